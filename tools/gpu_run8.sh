#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log | cut -c1-300; }
run 200 pipe_dur python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_ops.py -m gpu -q --tb=short --durations=8
run 300 reference python bench.py --impl reference --steps 3 --warmup 1
# full-set captures of the kernels the roofline quotes: fc6 + fc7 (first two plain-GEMM cg2 launches of a step), one conv3 layer, ROI pooling, NMS scan
timeout 400 ncu --set full --clock-control none -k regex:"roi_pool_kernel|nms_scan_kernel|nms_mask_kernel|conv_first_kernel" -s 12 -c 6 -o /tmp/prof_misc -f python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_misc.log 2>&1
ncu -i /tmp/prof_misc.ncu-rep --page raw --csv > gpurun_out/prof_misc_r05_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none -k regex:"tc_gemm_kernel" -s 141 -c 18 -o /tmp/prof_tc -f python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1
ncu -i /tmp/prof_tc.ncu-rep --page raw --csv > gpurun_out/prof_tc_r05_raw.csv 2>/dev/null
ls -la gpurun_out
