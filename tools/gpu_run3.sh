#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log | cut -c1-400; }
run tests python -m pytest tests -m gpu -q --tb=short -x
run smoke python __graft_entry__.py smoke
run bench python bench.py --steps 10 --warmup 3
run bench_b1 python bench.py --steps 10 --warmup 3 --batch 1 --no-cpu-baseline
ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 300 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 144 -c 48 -o /tmp/prof_tc -f python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_tc.log 2>&1
echo "ncu tc exit $?"
ncu -i /tmp/prof_tc.ncu-rep --page raw --csv > gpurun_out/prof_tc_r02_raw.csv 2>/dev/null
ncu -i /tmp/prof_tc.ncu-rep --page details --csv > gpurun_out/prof_tc_r02_details.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:"roi_pool_kernel|nms_chunk_kernel|nms_rank_kernel|conv_first_kernel|lstm_pointwise" -s 15 -c 12 -o /tmp/prof_misc -f python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_misc.log 2>&1
echo "ncu misc exit $?"
ncu -i /tmp/prof_misc.ncu-rep --page raw --csv > gpurun_out/prof_misc_r02_raw.csv 2>/dev/null
ls -la gpurun_out; du -sh gpurun_out
