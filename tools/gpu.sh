#!/bin/bash
# One parameterised GPU-box script (run as: gpurun --timeout N -- 'bash tools/gpu.sh <step> [<step> ...]').
# Every command is wrapped in its own `timeout`; logs land in gpurun_out/<name>.log and the tail is echoed.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { t=$1; name=$2; shift; shift; echo "=== $name"; timeout $t "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -${TAIL:-4} gpurun_out/$name.log | cut -c1-400; }
phases() { grep -o '"phases_ms": {[^}]*}' gpurun_out/$1.log; grep -o '"e2e[_u8]*": {[^}]*}' gpurun_out/$1.log; grep -o '"parity": {.*}' gpurun_out/$1.log | cut -c1-1500; }
for step in "$@"; do
  case $step in
    tests)      TAIL=15 run 900 tests python -m pytest tests -m gpu -q --tb=short -x ;;
    tests_noexit) TAIL=60 run 900 tests python -m pytest tests -m gpu -q --tb=short --maxfail=12 ;;
    safe)       # the same suite against the variant build without the newest hot-path kernels (bisecting a failure on the box)
                TAIL=40 DENSECAP_B200_LIB=$PWD/densecap_b200/libdensecap_b200_safe.so run 600 tests_safe python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_fullsize.py tests/test_gpu_dense.py -m gpu -q --tb=short --maxfail=12 ;;
    test:*)     TAIL=30 run 600 test_${step#test:} python -m pytest tests -m gpu -q --tb=short -x -k "${step#test:}" ;;
    file:*)     TAIL=30 run 600 tfile python -m pytest "tests/${step#file:}" -m gpu -q --tb=short -x ;;
    smoke)      run 200 smoke python __graft_entry__.py smoke ;;
    b1)         run 200 bench_b1 python bench.py --steps 20 --warmup 3 --batch 1 --no-cpu-baseline; phases bench_b1 ;;
    b4)         run 200 bench_b4 python bench.py --steps 20 --warmup 3 --batch 4 --no-cpu-baseline; phases bench_b4 ;;
    bench)      run 400 bench python bench.py; phases bench ;;
    ref)        run 400 reference python bench.py --impl reference --steps 3 ;;
    mixed8)     run 300 bench_mixed8 python bench.py --workload mixed8 --steps 10 --warmup 3 --no-cpu-baseline; phases bench_mixed8 ;;
    train)      run 200 bench_train python tools/bench_train.py ;;
    launches:*) # launch list of one batch-N step: launches:1 / launches:4
                B=${step#launches:}
                run 600 ncu_launches_b$B ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
                    --clock-control none --csv --log-file gpurun_out/launches_b$B.csv python bench.py --steps 1 --warmup 3 --batch $B --no-cpu-baseline --no-extras $BENCH_EXTRA
                mv gpurun_out/launches_b$B.csv gpurun_out/launches_b$B$TAG.csv ;;
    full:*)     # ncu --set full of kernels matching a regex: full:<regex>  (batch 1, last step only)
                R=${step#full:}
                run 900 ncu_full ncu --set full --import-source on --clock-control none -k "regex:$R" --launch-skip-before-match 0 -c ${CNT:-40} -o gpurun_out/prof_full \
                    python bench.py --steps 1 --warmup 0 --batch ${PB:-1} --no-cpu-baseline --no-extras $BENCH_EXTRA
                ncu -i gpurun_out/prof_full.ncu-rep --page raw --csv > gpurun_out/prof_full_raw.csv 2>/dev/null
                ncu -i gpurun_out/prof_full.ncu-rep --page source --csv > gpurun_out/prof_full_source.csv 2>/dev/null
                rm -f gpurun_out/prof_full.ncu-rep ;;
    sanitizer)  # compute-sanitizer over the small-model end-to-end tests (every kernel of the forward path incl. the persistent decoder, TMA-store epilogues,
                # cluster kernels) and the dense-layer tests; summaries are copied to profiles/ by hand
                SEL="full_forward_matches_oracle or batched_forward_equals_single or beamsearch_matches_oracle or test_tc_conv or test_tc_linear"
                TAIL=25 run 1500 sanitizer_memcheck compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_dense.py -m gpu -q --tb=line -x -k "$SEL"
                TAIL=25 run 1500 sanitizer_racecheck compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_dense.py -m gpu -q --tb=line -x -k "full_forward_matches_oracle or test_tc_conv"
                TAIL=25 run 1500 sanitizer_synccheck compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --tb=line -x -k "full_forward_matches_oracle" ;;
    racecheck)  TAIL=25 run 1500 sanitizer_racecheck compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_dense.py -m gpu -q --tb=line -x -k "full_forward_matches_oracle or test_tc_conv" ;;
    py:*)       TAIL=60 run 600 py python "${step#py:}" ;;
    *) echo "unknown step $step" ;;
  esac
done
