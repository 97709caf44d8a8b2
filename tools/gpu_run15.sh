#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log | cut -c1-300; }
run 240 pipe python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_ops.py tests/test_gpu_train.py -m gpu -q --tb=short -x
timeout 200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"tc_|conv_first" -s 150 -c 60 --csv --log-file gpurun_out/launches_r06_halo.csv python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1
DC_TC_HALO=0 timeout 200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"tc_|conv_first" -s 150 -c 60 --csv --log-file gpurun_out/launches_r06_nohalo.csv python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1
ls -la gpurun_out | tail -4
