#!/bin/bash
# Multi-GPU check on N GPUs of one box: correctness of the sharded path under torchrun (NCCL, threshold sharing over
# NVLink peer memory), then bench lines with and without sharing.
# Usage: gpurun --gpus N --timeout 1200 -- 'bash scripts/gpu_multi.sh N [c2 c2:noshare c4 ...]'
N=${1:-2}; shift
CONFIGS=${@:-c2 c2:noshare}
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/multi_summary.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "=== dist_gpu_check" | tee -a $OUT/multi_summary.txt
timeout 600 $TR --master-port 29511 scripts/dist_gpu_check.py > $OUT/dist_check.log 2>&1
echo "exit=$? $(grep -c OK $OUT/dist_check.log) OK / $(grep -c MISMATCH $OUT/dist_check.log) MISMATCH" | tee -a $OUT/multi_summary.txt
grep -E "OK|MISMATCH|Error|error" $OUT/dist_check.log | tail -n 40 >> $OUT/multi_summary.txt
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
port=29520
for item in $CONFIGS; do   # "c2" = with threshold sharing, "c2:noshare" = without
  cfg=${item%%:*}; share=share; extra=""
  if [ "$item" != "$cfg" ]; then share=noshare; extra="--no-share"; fi
  port=$((port+1))
  timeout 900 $TR --master-port $port bench.py --gpus $N --config $cfg --steps 5 --warmup 3 --parity-users 256 $extra > $OUT/bench_${cfg}_n${N}_${share}.log 2>&1
  echo "$cfg n=$N $share: $(brief $OUT/bench_${cfg}_n${N}_${share}.log)" | tee -a $OUT/multi_summary.txt
  grep '^{' $OUT/bench_${cfg}_n${N}_${share}.log | tail -n 1 > $OUT/r02_bench_${cfg}_n${N}_${share}.json
done
cat $OUT/multi_summary.txt
