#!/bin/bash
# Round-end style check on one B200: full GPU test suite, smoke, the default bench (both arms), ncu launch list + one
# full capture of the dominant kernel.  Usage: gpurun --timeout 2400 -- 'bash scripts/gpu_final.sh'
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/final_summary.txt
run() { local name=$1; local t=$2; shift 2; echo "=== $name" >> $OUT/final_summary.txt; timeout $t "$@" > $OUT/$name.log 2>&1; echo "exit=$? $(tail -n 2 $OUT/$name.log | cut -c1-3500)" >> $OUT/final_summary.txt; }
run f_tests 1500 python -m pytest tests -m gpu -q
run f_smoke 300 python __graft_entry__.py smoke
run f_bench 900 python bench.py --steps 5 --warmup 3
run f_bench_ref 900 python bench.py --impl reference --steps 3 --warmup 1
run f_ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --parity-users 0
run f_ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:topk_kernel -s 2 -c 1 -f -o $OUT/r01_prof_tc python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --parity-users 0
cat $OUT/final_summary.txt
