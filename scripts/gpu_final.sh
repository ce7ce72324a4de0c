#!/bin/bash
# Final single-GPU pass of a round: the whole GPU suite, the default bench line, the ncu launch list and full capture.
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/summary.txt
run() { local name=$1; local t=$2; shift 2; echo "=== $name" | tee -a $OUT/summary.txt; local t0=$(date +%s); timeout $t "$@" > $OUT/$name.log 2>&1; echo "exit=$? $(( $(date +%s) - t0 ))s $(tail -n 3 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)" | tee -a $OUT/summary.txt; }
run t_all 900 python -m pytest tests -m gpu -x -q
run smoke 200 python __graft_entry__.py smoke
run bench_c2 600 python bench.py --steps 5 --warmup 3
grep '^{' $OUT/bench_c2.log | tail -n 1 > $OUT/r02_bench_c2_n1.json
bash scripts/gpu_profile.sh > $OUT/profile.log 2>&1
run bench_c5 600 python bench.py --config c5 --steps 3 --warmup 3 --parity-users 256 --no-model
grep '^{' $OUT/bench_c5.log | tail -n 1 > $OUT/r02_bench_c5_n1.json
cat $OUT/summary.txt; cat $OUT/profile.log
