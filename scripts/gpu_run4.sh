#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/summary.txt
run() { local name=$1; local t=$2; shift 2; echo "=== $name" | tee -a $OUT/summary.txt; local t0=$(date +%s); timeout $t "$@" > $OUT/$name.log 2>&1; echo "exit=$? $(( $(date +%s) - t0 ))s $(tail -n 3 $OUT/$name.log | tr '\n' ' ' | cut -c1-500)" | tee -a $OUT/summary.txt; }
run t_all 1500 python -m pytest tests -m gpu -x -q --durations=6
run bench_c2 900 python bench.py --steps 5 --warmup 3
grep '^{' $OUT/bench_c2.log | tail -n 1 > $OUT/r02_bench_c2_n1.json
run bench_c3 900 python bench.py --config c3 --steps 5 --warmup 3
grep '^{' $OUT/bench_c3.log | tail -n 1 > $OUT/r02_bench_c3_n1.json
run bench_ref 600 python bench.py --impl reference --steps 3 --warmup 1
cat $OUT/summary.txt
