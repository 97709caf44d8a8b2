cd /root/repo
DC_LAYERS=1 bash tools/ab.sh "2 2 1" default fakestore
DC_LAYERS=1 bash tools/ab.sh "1000 1000 1000" default
bash tools/gpu.sh tests_noexit
