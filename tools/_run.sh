cd /root/repo
DENSECAP_B200_DEBUG=1 timeout 900 python -m pytest tests -m gpu -q --tb=line -x -s 2>&1 | grep -B2 -A2 -m 6 "DENSECAP_B200_DEBUG" | cut -c1-300
bash tools/gpu.sh b1 b4
