#!/bin/bash
# A/B of variant libraries on the GPU box: bash tools/ab.sh "<groups or ->" name1 name2 ...   (name "default" = the product library)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
G="$1"; shift
[ "$G" = "-" ] && G=""
for v in "$@"; do
  if [ "$v" = default ]; then unset DENSECAP_B200_LIB; else export DENSECAP_B200_LIB=$PWD/densecap_b200/libdensecap_b200_$v.so; fi
  timeout 300 python tools/bench_trunk.py $G 2>&1 | tail -n 1 | tee -a gpurun_out/ab.log
done
