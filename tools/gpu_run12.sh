#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log | cut -c1-400; }
run 300 tests python -m pytest tests -m gpu -q --tb=short -x
run 200 bench_b1 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline
timeout 300 ncu --set full --clock-control none -k regex:"roi_pool_kernel|roi_taps_kernel" -s 4 -c 2 -o /tmp/prof_roi -f python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_roi.log 2>&1
ncu -i /tmp/prof_roi.ncu-rep --page raw --csv > gpurun_out/prof_roi_r06_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
