#!/bin/bash
# DRAM traffic of the fused kernel vs number of waves / carousel variant.
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/exp_summary.txt
cap() { # label, users, env...
  local label=$1; local users=$2; shift 2
  env "$@" timeout 900 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum --clock-control none -k regex:topk_kernel -s 2 -c 1 --csv --log-file $OUT/tr.csv python bench.py --users $users --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --parity-users 0 > $OUT/exp_ncu.log 2>&1
  echo "$label users=$users: $(grep -E 'dram__bytes_read|gpu__time|lts__' $OUT/tr.csv | awk -F'","' '{printf "%s=%s ", $(NF-2), $NF}' | tr -d '"')" >> $OUT/exp_summary.txt
}
cap refpair 151552
cap refpair 303104
cap refpair 1000000
cap lastwriter 1000000 B200_TC_DEBUG=4
cap off 303104 B200_TC_CAROUSEL=0
cat $OUT/exp_summary.txt
