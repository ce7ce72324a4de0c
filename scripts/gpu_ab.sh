#!/bin/bash
# A/B of kernel geometries / modes on one B200 (cases: scripts/ab_cases.sh).
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_ab.sh [users]'
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/ab_summary.txt
USERS=${1:-303104}
brief() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']; p=d.get('parity') or {}; e=d['config']['engine']; c=d.get('clocks') or {}
    print('value=%.0f ms_step=%.2f ms_main=%.2f ms_select=%.2f tflops=%.0f frac=%.3f launches=%s fb=%s exact=%s mism=%s/%s sm_mhz=%s power=%s'%(d['value'],d['ms_per_step'],r['ms_per_launch'],r.get('ms_select_per_step') or 0,r['achieved'],r['frac'] or 0,e.get('n_tc_launches'),e.get('n_fallback_rows'),e.get('n_exact_rows'),p.get('id_mismatches'),p.get('users_checked'),c.get('sm_mhz'),c.get('power_w_max')))
except Exception as e:
    print('ERR',e, open(sys.argv[1]).read()[-800:])
PY
}
bench() { # label, extra bench args (quoted), env...
  local label=$1; local args=$2; shift 2
  env "$@" timeout 400 python bench.py --users $USERS --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-model --parity-users 128 $args > $OUT/ab_$label.log 2>&1
  echo "$label: $(brief $OUT/ab_$label.log)" | tee -a $OUT/ab_summary.txt
}
source scripts/ab_cases.sh
cat $OUT/ab_summary.txt
