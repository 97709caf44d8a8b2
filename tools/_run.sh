cd /root/repo
bash tools/gpu.sh tests_noexit b1 b4 racecheck
DC_LAYERS=1 bash tools/ab.sh "2 2 1" default
