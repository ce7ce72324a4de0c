#!/bin/bash
# quick 8-GPU measurement of the default workload (value only)
N=${1:-8}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 bench.py --gpus $N --steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-model --parity-users 64 > $OUT/quick_n$N.log 2>&1
python - <<'PY'
import json,sys
try:
    d=json.loads([l for l in open('gpurun_out/quick_n%s.log' % sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/quick_n8.log') if l.startswith('{')][-1])
    r=d['roofline']; e=d['config']['engine']
    print('value=%.0f ms_step=%.2f ms_main=%.2f ms_select=%.2f tflops=%.0f uncert=%s mism=%s steps=%s' % (d['value'], d['ms_per_step'], r['ms_per_launch'], r['ms_select_per_step'], r['achieved'], e.get('n_uncertified_rows'), d['parity']['id_mismatches'], d['ms_steps_rank0']))
except Exception as ex:
    print('ERR', ex, open('gpurun_out/quick_n8.log').read()[-1500:])
PY
