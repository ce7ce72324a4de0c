#!/bin/bash
# A/B of library variants (python -m rectools_b200.build --variant NAME -D...) and tuning hooks on one B200.
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_ab.sh [users]'
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/ab_summary.txt
USERS=${1:-303104}
brief() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']; p=d.get('parity') or {}; e=d['config']['engine']; c=d.get('clocks') or {}
    print('value=%.0f ms_step=%.2f ms_main=%.2f tflops=%.0f frac=%.3f fb=%s exact=%s mism=%s sm_mhz=%s power=%s'%(d['value'],d['ms_per_step'],r['ms_per_launch'],r['achieved'],r['frac'] or 0,e.get('n_fallback_rows'),e.get('n_exact_rows'),p.get('id_mismatches'),c.get('sm_mhz'),c.get('power_w_max')))
except Exception as e:
    print('ERR',e, open(sys.argv[1]).read()[-600:])
PY
}
bench() { # label, extra bench args (quoted), env...
  local label=$1; local args=$2; shift 2
  env "$@" timeout 400 python bench.py --users $USERS --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --parity-users 64 $args > $OUT/ab_$label.log 2>&1
  echo "$label: $(brief $OUT/ab_$label.log)" | tee -a $OUT/ab_summary.txt
}
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/ab_tests.log 2>&1; echo "tests exit=$? $(tail -n 1 $OUT/ab_tests.log)" | tee -a $OUT/ab_summary.txt
L=$PWD/rectools_b200
source scripts/ab_cases.sh
cat $OUT/ab_summary.txt
