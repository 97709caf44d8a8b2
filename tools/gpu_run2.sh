#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -4 gpurun_out/$name.log; }
run pipe python -m pytest tests/test_gpu_pipeline.py -m gpu -q --tb=short
run smoke python __graft_entry__.py smoke
run bench python bench.py --steps 5 --warmup 3
run bench_b1 python bench.py --steps 5 --warmup 3 --batch 1 --no-cpu-baseline
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --batch 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu exit $?"
