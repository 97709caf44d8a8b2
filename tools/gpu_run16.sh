#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log | cut -c1-300; }
run 300 tests python -m pytest tests -m gpu -q --tb=short -x
run 200 bench_b1 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline
grep -o '"phases_ms": {[^}]*}' gpurun_out/bench_b1.log
DC_TC_HALO=0 timeout 200 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/bench_b1_nohalo.log 2>&1
grep -o '"phases_ms": {[^}]*}' gpurun_out/bench_b1_nohalo.log
run 200 bench_b4 python bench.py --steps 20 --warmup 3 --batch 4 --no-cpu-baseline
grep -o '"phases_ms": {[^}]*}' gpurun_out/bench_b4.log
timeout 200 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"tc_|conv_first" -s 144 -c 60 --csv --log-file gpurun_out/launches_r06_halo.csv python bench.py --steps 1 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1
