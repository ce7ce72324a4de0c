#!/bin/bash
# One gpurun call: the whole GPU test suite (groups in separate processes: a trap in one kernel must not poison the others),
# smoke, A/B cases, the default bench.
# Usage: gpurun --timeout 1700 -- 'bash scripts/gpu_check.sh [tests|all]'
MODE=${1:-all}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> $OUT/gpu.txt
run() { # name, timeout, cmd...
  local name=$1; local t=$2; shift 2
  echo "=== $name" | tee -a $OUT/summary.txt
  local t0=$(date +%s)
  timeout $t "$@" > $OUT/$name.log 2>&1
  echo "exit=$? $(( $(date +%s) - t0 ))s $(tail -n 3 $OUT/$name.log | tr '\n' ' ' | cut -c1-700)" | tee -a $OUT/summary.txt
}
: > $OUT/summary.txt
run smoke 300 python __graft_entry__.py smoke
run t_parity 900 python -m pytest tests/test_gpu_parity.py -q -x --durations=8
run t_recommend 300 python -m pytest tests/test_recommend_gpu.py -q
run t_models 600 python -m pytest tests/test_gpu_models.py -q --durations=6
run t_sharded 300 python -m pytest tests/test_gpu_sharded.py -q
run t_scale 1200 python -m pytest tests/test_gpu_scale.py -q --durations=6
if [ "$MODE" = "all" ]; then
  bash scripts/gpu_ab.sh 303104 > $OUT/ab.log 2>&1
  run bench_c2 900 python bench.py --steps 5 --warmup 3
fi
cat $OUT/summary.txt; cat $OUT/ab_summary.txt 2>/dev/null
