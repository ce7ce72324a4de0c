#!/bin/bash
# One gpurun call: parity tests in isolated processes (a trap in one kernel must not poison the others), smoke, bench.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh [quick|full]'
MODE=${1:-quick}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> $OUT/gpu.txt
run() { # name, timeout, cmd...
  local name=$1; local t=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  timeout $t "$@" > $OUT/$name.log 2>&1
  echo "exit=$? $(tail -n 3 $OUT/$name.log | tr '\n' ' ' | cut -c1-600)" | tee -a $OUT/summary.txt
}
: > $OUT/summary.txt
run t_exact 600 python -m pytest tests -m gpu -q -x -k "known_answers or raises or puresvd or (golden_rankers and not tc)"
run t_tc_golden 600 python -m pytest tests -m gpu -q -x -k "golden_rankers and tc"
run t_random 900 python -m pytest tests -m gpu -q -k "random_vs_oracle"
run t_edge 600 python -m pytest tests -m gpu -q -k "edge_cases or merge"
run smoke 300 python __graft_entry__.py smoke
run bench_small 600 python bench.py --users 131072 --items 1000000 --steps 3 --warmup 3 --ref-users 256
if [ "$MODE" = "full" ]; then
  run bench_full 900 python bench.py --steps 5 --warmup 3
  run ncu_list 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv python bench.py --users 262144 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --parity-users 0
  run ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:tc_topk -s 1 -c 1 -f -o $OUT/prof_tc python bench.py --users 65536 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --parity-users 0
fi
cat $OUT/summary.txt
