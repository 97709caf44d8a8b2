#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log | cut -c1-600; }
run 300 tests python -m pytest tests -m gpu -q --tb=short -x
run 120 smoke python __graft_entry__.py smoke
run 300 bench python bench.py
run 200 bench_b1 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline
run 120 bench_train python tools/bench_train.py
timeout 240 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 261 -c 100 --csv --log-file gpurun_out/launches_r06.csv python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1
tail -2 gpurun_out/ncu_l1.log | cut -c1-200
