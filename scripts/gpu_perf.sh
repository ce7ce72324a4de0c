#!/bin/bash
# Performance iteration: parity tests + steady-state bench (4 full waves of subject tiles) for the kernel variants.
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/perf_summary.txt
run() { local name=$1; local t=$2; shift 2; echo "=== $name" >> $OUT/perf_summary.txt; timeout $t "$@" > $OUT/$name.log 2>&1; echo "exit=$? $(tail -n 2 $OUT/$name.log | cut -c1-300)" >> $OUT/perf_summary.txt; }
brief() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']; p=d.get('parity') or {}
    print('value=%.0f ms_step=%.2f ms_main=%.2f tflops=%.0f frac=%.3f fb=%s mism=%s power=%s reasons=%s'%(d['value'],d['ms_per_step'],r['ms_per_launch'],r['achieved'],r['frac'] or 0,d['config']['engine'].get('n_fallback_rows'),p.get('id_mismatches'),(d.get('clocks') or {}).get('power_w_max'),(d.get('clocks') or {}).get('reasons')))
except Exception as e:
    print('ERR',e, open(sys.argv[1]).read()[-400:])
PY
}
bench() { # label, env...
  local label=$1; shift
  env "$@" B200_TC_SPLITS=1 timeout 400 python bench.py --users 151552 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --parity-users 64 > $OUT/b_$label.log 2>&1
  echo "$label: $(brief $OUT/b_$label.log)" >> $OUT/perf_summary.txt
}
run p_tests 900 python -m pytest tests -m gpu -q -x -k "random_vs_oracle or edge_cases or (golden_rankers and tc) or merge or many_work"
B200_TC_STAGE=0 run p_tests_nostage 900 python -m pytest tests -m gpu -q -x -k "random_vs_oracle or (golden_rankers and tc)"
bench 2sm_t256_stage B200_TC_TILE=256
bench 2sm_t256_stage_dbg1 B200_TC_TILE=256 B200_TC_DEBUG=1
bench 2sm_t256_nostage B200_TC_TILE=256 B200_TC_STAGE=0
bench 2sm_t128_stage B200_TC_TILE=128
bench 2sm_t128_nostage B200_TC_TILE=128 B200_TC_STAGE=0
bench 1sm B200_TC_KERNEL=1
if [ "$1" = "ncu" ]; then
  B200_TC_SPLITS=1 run p_ncu 900 ncu --set full --clock-control none --import-source on -k regex:topk_kernel -s 1 -c 1 -f -o $OUT/prof_tc python bench.py --users 37888 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --parity-users 0
fi
cat $OUT/perf_summary.txt
