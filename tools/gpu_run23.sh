#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -3 gpurun_out/$name.log | cut -c1-300; }
run 300 tests python -m pytest tests -m gpu -q --tb=short
run 120 smoke python __graft_entry__.py smoke
