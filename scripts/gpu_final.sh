#!/bin/bash
# Final single-GPU pass of a round: the whole GPU suite, smoke, the default bench line (+ config 3), the ncu launch list and
# one full capture of the fused kernel.   Usage: gpurun --timeout 1200 -- 'bash scripts/gpu_final.sh [noprof]'
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/summary.txt
run() { local name=$1; local t=$2; shift 2; echo "=== $name" | tee -a $OUT/summary.txt; local t0=$(date +%s); timeout $t "$@" > $OUT/$name.log 2>&1; echo "exit=$? $(( $(date +%s) - t0 ))s $(tail -n 3 $OUT/$name.log | tr '\n' ' ' | cut -c1-400)" | tee -a $OUT/summary.txt; }
run t_all 900 python -m pytest tests -m gpu -x -q
run smoke 200 python __graft_entry__.py smoke
run bench_c3 600 python bench.py --config c3 --steps 5 --warmup 3
grep '^{' $OUT/bench_c3.log | tail -n 1 > $OUT/r02_bench_c3_n1.json
if [ "$1" != "noprof" ]; then
  run bench_c2 600 python bench.py --steps 5 --warmup 3
  grep '^{' $OUT/bench_c2.log | tail -n 1 > $OUT/r02_bench_c2_n1.json
  bash scripts/gpu_profile.sh > $OUT/profile.log 2>&1
fi
cat $OUT/summary.txt; cat $OUT/profile.log 2>/dev/null
