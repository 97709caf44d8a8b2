cd /root/repo
bash tools/gpu.sh tests_noexit smoke bench ref mixed8 train launches:1
PB=4 bash tools/gpu.sh launches:4
