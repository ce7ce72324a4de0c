#!/bin/bash
# Pipeline experiments: MMA-only and fast-path-only rates of the tensor-core kernel (results invalid in debug modes).
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/exp_summary.txt
for mode in 2 1 0; do
  B200_TC_DEBUG=$mode B200_TC_SPLITS=1 timeout 300 python bench.py --users 151552 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --parity-users 0 > $OUT/exp_$mode.log 2>&1
  echo "mode=$mode $(python - <<PY
import json
try:
    d=json.loads(open('$OUT/exp_$mode.log').read().strip().splitlines()[-1])
    print('ms_main=%.2f tflops=%.0f value=%.0f power=%s'%(d['roofline']['ms_per_launch'], d['roofline']['achieved'], d['value'], d['clocks']))
except Exception as e:
    print('ERR', e)
PY
)" >> $OUT/exp_summary.txt
done
cat $OUT/exp_summary.txt
