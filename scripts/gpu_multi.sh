#!/bin/bash
# Multi-GPU check (run with `gpurun --gpus N`): sharded parity under NCCL + the bench at N ranks: item sharding (default,
# north star) and, with a second argument, another --item-shards value (1 = user sharding).
N=${1:-2}
ALT=${2:-}
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 scripts/dist_gpu_check.py > $OUT/dist_check.log 2>&1; echo "dist_check exit=$? $(grep -E 'OK|MISMATCH|Error' $OUT/dist_check.log | tr '\n' ' ' | cut -c1-600)"
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_n$N.log 2>&1; echo "bench exit=$? $(tail -n 1 $OUT/bench_n$N.log | cut -c1-400)"
if [ -n "$ALT" ]; then
  timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --item-shards $ALT > $OUT/bench_n${N}_i$ALT.log 2>&1; echo "bench(item-shards=$ALT) exit=$? $(tail -n 1 $OUT/bench_n${N}_i$ALT.log | cut -c1-400)"
fi
