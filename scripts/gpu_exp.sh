#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/exp_summary.txt
for cfg in "B200_TC_DEBUG=3" "B200_TC_DEBUG=3 B200_TC_TILE=128" ; do
  env $cfg B200_TC_SPLITS=1 timeout 300 python bench.py --users 151552 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --parity-users 0 > $OUT/exp.log 2>&1
  echo "$cfg: $(grep 'b200 tc debug' $OUT/exp.log | tail -1)" >> $OUT/exp_summary.txt
done
cat $OUT/exp_summary.txt
