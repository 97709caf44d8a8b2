#!/bin/bash
# end-of-round sanity + source-level profiles of the two kernels round 2 starts from (conv2_1 halo<128>, LSTM gate GEMM)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log | cut -c1-300; }
run 300 tests python -m pytest tests -m gpu -q --tb=short
run 120 smoke python __graft_entry__.py smoke
run 200 bench_b1 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline
# conv2_1 = 2nd tc_conv_halo launch of a step (13 per step): skip 3*13+1, take 1.  gate GEMM = tc_gemm launch #6 of a step (35 per step: head, fc6, fc7, enc, gates0, vocab... ) -> take the 3rd gates launch
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"tc_conv_halo" -s 40 -c 1 -o /tmp/p_conv21 -f python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_c21.log 2>&1
ncu -i /tmp/p_conv21.ncu-rep --page source --csv > gpurun_out/prof_conv2_1_r06_source.csv 2>/dev/null
ncu -i /tmp/p_conv21.ncu-rep --page raw --csv > gpurun_out/prof_conv2_1_r06_raw.csv 2>/dev/null
timeout 200 ncu --set full --import-source on --clock-control none -k regex:"tc_gemm_kernel" -s 111 -c 2 -o /tmp/p_lm -f python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_lm.log 2>&1
ncu -i /tmp/p_lm.ncu-rep --page source --csv > gpurun_out/prof_lm_r06_source.csv 2>/dev/null
ncu -i /tmp/p_lm.ncu-rep --page raw --csv > gpurun_out/prof_lm_r06_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
