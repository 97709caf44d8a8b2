cd /root/repo
bash tools/gpu.sh tests_noexit b4
grep -o '"roi_pool": {[^}]*}' gpurun_out/bench_b4.log
bash tools/ab.sh "2 2 1" default
bash tools/ab.sh "4 4 1" default
bash tools/ab.sh "4 4 2" default
