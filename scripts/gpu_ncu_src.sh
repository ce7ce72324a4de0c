#!/bin/bash
# Source-level ncu captures of the fused kernel (per-SASS-line instruction counts + stall samples).
# Usage: gpurun --timeout 1200 -- 'bash scripts/gpu_ncu_src.sh'
OUT=gpurun_out; mkdir -p $OUT
cap() { # name, bench args...
  local name=$1; shift
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:topk_kernel -s 2 -c 1 -f -o $OUT/$name \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --parity-users 0 "$@" > $OUT/$name.log 2>&1
  echo "$name exit=$?"; tail -n 2 $OUT/$name.log | cut -c1-300
}
cap src_n1m --users 75776 --items 1000000
cap src_n125k --users 151552 --items 125000
ls -la $OUT/*.ncu-rep
