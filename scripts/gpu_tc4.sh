#!/bin/bash
# First hardware run of the experimental tc4_topk kernel (round-2 starting point):
#   gpurun --timeout 1500 -- 'bash scripts/gpu_tc4.sh'
# 1. its parity test alone (a protocol bug traps through the mbarrier watchdog instead of hanging; still under `timeout`),
# 2. the A/B of scripts/ab_cases.sh (gen3 vs gen4 at N = 1M and on a 125 K-item shard).
OUT=gpurun_out; mkdir -p $OUT
B200_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests -m gpu -q -x -k "gen4" > $OUT/tc4_tests.log 2>&1; echo "tc4 parity exit=$? $(tail -n 1 $OUT/tc4_tests.log)"
B200_TC_KERNEL=4 timeout 300 python -m pytest tests -m gpu -q -x -k "random_vs_oracle or edge_cases or golden" > $OUT/tc4_tests_all.log 2>&1; echo "suite under B200_TC_KERNEL=4 exit=$? $(tail -n 1 $OUT/tc4_tests_all.log)"
bash scripts/gpu_ab.sh
