#!/bin/bash
# final 8-GPU measurements of the default workload: item-sharded (north star, thresholds shared), user-sharded, 2 x 4 grid
N=8; OUT=gpurun_out; mkdir -p $OUT; : > $OUT/multi_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
brief() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    r=d['roofline']; p=d.get('parity') or {}; e=d['config']['engine']; c=d.get('clocks') or {}; x=d.get('e2e') or {}
    print('value=%.0f e2e=%.0f ms_step=%.2f ms_main=%.2f ms_select=%.2f tflops=%.0f frac=%.3f uncert=%s fb=%s mism=%s/%s sm_mhz=%s steps=%s'%(d['value'],x.get('value') or 0,d['ms_per_step'],r['ms_per_launch'],r.get('ms_select_per_step') or 0,r['achieved'],r['frac'] or 0,e.get('n_uncertified_rows'),e.get('n_fallback_rows'),p.get('id_mismatches'),p.get('users_checked'),c.get('sm_mhz'),d.get('ms_steps_rank0')))
except Exception as e:
    print('ERR',e, open(sys.argv[1]).read()[-1500:])
PY
}
port=29550
for tag in items8 users8 grid2x4; do
  extra=""; [ "$tag" = "users8" ] && extra="--item-shards 1"; [ "$tag" = "grid2x4" ] && extra="--item-shards 2"
  port=$((port+1))
  timeout 300 $TR --master-port $port bench.py --gpus $N --steps 8 --warmup 3 --parity-users 256 $extra > $OUT/bench_c2_n8_$tag.log 2>&1
  echo "c2 n=8 $tag: $(brief $OUT/bench_c2_n8_$tag.log)" | tee -a $OUT/multi_summary.txt
  grep '^{' $OUT/bench_c2_n8_$tag.log | tail -n 1 > $OUT/r02_bench_c2_n8_$tag.json
done
