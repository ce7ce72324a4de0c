#!/bin/bash
# Kernel-variant comparison at the full-catalogue shape (N = 1M) and at the 8-way item-sharded shape (N = 125K).
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/exp_summary.txt
brief() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']; p=d.get('parity') or {}
    print('value=%.0f ms_step=%.2f ms_main=%.2f tflops=%.0f fb=%s splits=%s mism=%s'%(d['value'],d['ms_per_step'],r['ms_per_launch'],r['achieved'],d['config']['engine'].get('n_fallback_rows'),d['config']['engine'].get('n_splits'),p.get('id_mismatches')))
except Exception as e:
    print('ERR',e, open(sys.argv[1]).read()[-300:])
PY
}
for items in 1000000 125000; do
 for cfg in "B200_TC_KERNEL=1" "B200_TC_TILE=256" "B200_TC_TILE=256 B200_TC_STAGE=0" "B200_TC_TILE=128"; do
  env $cfg timeout 300 python bench.py --users 303104 --items $items --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --parity-users 64 > $OUT/exp.log 2>&1
  echo "items=$items [$cfg]: $(brief $OUT/exp.log)" >> $OUT/exp_summary.txt
 done
done
cat $OUT/exp_summary.txt
