#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/summary.txt
run() { local name=$1; local t=$2; shift 2; echo "=== $name" | tee -a $OUT/summary.txt; local t0=$(date +%s); timeout $t "$@" > $OUT/$name.log 2>&1; echo "exit=$? $(( $(date +%s) - t0 ))s $(tail -n 3 $OUT/$name.log | tr '\n' ' ' | cut -c1-500)" | tee -a $OUT/summary.txt; }
run t_parity 900 python -m pytest tests/test_gpu_parity.py -q --durations=5
run t_recommend 300 python -m pytest tests/test_recommend_gpu.py -q
run t_scale2 600 python -m pytest tests/test_gpu_scale.py -q -k "second_chance"
run t_sharded 300 python -m pytest tests/test_gpu_sharded.py -q
bash scripts/gpu_ab.sh 303104 > $OUT/ab.log 2>&1
cat $OUT/summary.txt; cat $OUT/ab_summary.txt
