#!/bin/bash
# Performance iteration: parity smoke + steady-state bench (4 full waves of subject tiles) + one ncu capture.
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/perf_summary.txt
run() { local name=$1; local t=$2; shift 2; echo "=== $name" >> $OUT/perf_summary.txt; timeout $t "$@" > $OUT/$name.log 2>&1; echo "exit=$? $(tail -n 2 $OUT/$name.log | cut -c1-1500)" >> $OUT/perf_summary.txt; }
run p_tests 600 python -m pytest tests -m gpu -q -x -k "random_vs_oracle or edge_cases or (golden_rankers and tc)"
run p_smoke 300 python __graft_entry__.py smoke
B200_TC_SPLITS=1 run p_bench 600 python bench.py --users 151552 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --parity-users 128
if [ "$1" = "ncu" ]; then
  B200_TC_SPLITS=1 run p_ncu 900 ncu --set full --clock-control none --import-source on -k regex:tc_topk -s 1 -c 1 -f -o $OUT/prof_tc python bench.py --users 37888 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --parity-users 0
fi
cat $OUT/perf_summary.txt
