#!/bin/bash
# final-state evidence of round 1: default bench line, full launch list of one step, --set full of the main kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -2 gpurun_out/$name.log | cut -c1-400; }
run 300 bench python bench.py
run 120 bench_train python tools/bench_train.py
timeout 240 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 306 -c 104 --csv --log-file gpurun_out/launches_r06.csv python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:"tc_conv_halo|tc_gemm|roi_pool_kernel|nms_scan|conv_first64" -s 159 -c 22 -o /tmp/prof_r06 -f python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ncu -i /tmp/prof_r06.ncu-rep --page raw --csv > gpurun_out/prof_r06_raw.csv 2>/dev/null
ls -la gpurun_out | tail -6
