#!/bin/bash
# e2e chunk-size sweep (B200_CHUNK_ROWS) on the default workload + a full ncu capture of the fused kernel on the per-GPU shard of
# an 8-way item split.  Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_misc.sh'
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/misc_summary.txt
for c in 75776 151552 303104 1000000; do
  B200_CHUNK_ROWS=$c timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --parity-users 16 > $OUT/chunk_$c.log 2>&1
  echo "chunk_rows=$c $(tail -n 1 $OUT/chunk_$c.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); e=d['e2e']
    print('value=%.0f e2e=%.0f e2e_ms_total=%.2f ms_main=%.2f h2d=%.2f d2h=%.2f'%(d['value'], e['value'], e['engine_ms_last_step']['ms_total'], e['engine_ms_last_step']['ms_main'], e['engine_ms_last_step']['ms_h2d'], e['engine_ms_last_step']['ms_d2h']))
except Exception as ex: print('ERR', ex)
")" | tee -a $OUT/misc_summary.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_kernel -s 2 -c 1 -f -o $OUT/r01_prof_tc_n125k python bench.py --users 303104 --items 125000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --parity-users 0 > $OUT/ncu_n125k.log 2>&1; echo "ncu exit=$?" | tee -a $OUT/misc_summary.txt
