#!/bin/bash
# Performance iteration: parity tests + steady-state bench (4 full waves of subject tiles) + optional ncu capture.
OUT=gpurun_out
mkdir -p $OUT
: > $OUT/perf_summary.txt
run() { local name=$1; local t=$2; shift 2; echo "=== $name" >> $OUT/perf_summary.txt; timeout $t "$@" > $OUT/$name.log 2>&1; echo "exit=$? $(tail -n 2 $OUT/$name.log | cut -c1-1500)" >> $OUT/perf_summary.txt; }
brief() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']; print('value=%.0f ms_step=%.2f ms_main=%.2f tflops=%.0f frac=%.3f fb=%s parity=%s clocks=%s'%(d['value'],d['ms_per_step'],r['ms_per_launch'],r['achieved'],r['frac'] or 0,d['config']['engine'].get('n_fallback_rows'),d.get('parity'),d.get('clocks')))
except Exception as e:
    print('ERR',e)
PY
}
run p_tests 900 python -m pytest tests -m gpu -q -x -k "random_vs_oracle or edge_cases or (golden_rankers and tc) or merge"
run p_smoke 300 python __graft_entry__.py smoke
B200_TC_SPLITS=1 run p_bench 600 python bench.py --users 151552 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --parity-users 128
echo "2sm: $(brief $OUT/p_bench.log)" >> $OUT/perf_summary.txt
for mode in 2 1; do
  B200_TC_DEBUG=$mode B200_TC_SPLITS=1 timeout 300 python bench.py --users 151552 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --parity-users 0 > $OUT/exp_$mode.log 2>&1
  echo "2sm debug=$mode: $(brief $OUT/exp_$mode.log)" >> $OUT/perf_summary.txt
done
if [ "$1" = "ncu" ]; then
  B200_TC_SPLITS=1 run p_ncu 900 ncu --set full --clock-control none --import-source on -k regex:topk_kernel -s 1 -c 1 -f -o $OUT/prof_tc python bench.py --users 37888 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --parity-users 0
fi
B200_TC_KERNEL=1 run p_tests_1sm 900 python -m pytest tests -m gpu -q -x -k "random_vs_oracle or (golden_rankers and tc)"
B200_TC_KERNEL=1 B200_TC_SPLITS=1 run p_bench_1sm 600 python bench.py --users 151552 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --parity-users 128
echo "1sm: $(brief $OUT/p_bench_1sm.log)" >> $OUT/perf_summary.txt
cat $OUT/perf_summary.txt
