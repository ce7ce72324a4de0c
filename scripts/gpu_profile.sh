#!/bin/bash
# One gpurun call (ONE GPU): launch list of the default bench command + one `ncu --set full` capture of the fused kernel on
# the default workload (and, optionally, of other configs).  Numbers printed under ncu are never bench values.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_profile.sh [c3 ...]'
OUT=gpurun_out; mkdir -p $OUT
B="--no-e2e --no-cpu-baseline --no-model --parity-users 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/r02_launches.csv python bench.py --steps 2 --warmup 1 $B > $OUT/prof_list.log 2>&1
echo "launch list exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_topk -s 2 -c 1 -f -o $OUT/r02_prof_fused python bench.py --steps 1 --warmup 1 $B > $OUT/prof_full.log 2>&1
echo "full capture exit=$?"
for cfg in "$@"; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_topk -s 2 -c 1 -f -o $OUT/r02_prof_fused_$cfg python bench.py --config $cfg --users 303104 --steps 1 --warmup 1 $B > $OUT/prof_full_$cfg.log 2>&1
  echo "full capture $cfg exit=$?"
done
ls -la $OUT | tail -n 8
