#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log | cut -c1-300; }
run 200 ops python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py -m gpu -q --tb=short -x
run 200 bench_b1 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline
grep -o '"roi_pool": {[^}]*}' gpurun_out/bench_b1.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"roi_pool_kernel" -s 2 -c 1 -o /tmp/prof_roi -f python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_roi.log 2>&1
ncu -i /tmp/prof_roi.ncu-rep --page raw --csv > gpurun_out/prof_roi_r06_raw.csv 2>/dev/null
ncu -i /tmp/prof_roi.ncu-rep --page source --csv > gpurun_out/prof_roi_r06_source.csv 2>/dev/null
ls -la gpurun_out | tail -4
