#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/summary.txt
run() { local name=$1; local t=$2; shift 2; echo "=== $name" | tee -a $OUT/summary.txt; local t0=$(date +%s); timeout $t "$@" > $OUT/$name.log 2>&1; echo "exit=$? $(( $(date +%s) - t0 ))s $(tail -n 3 $OUT/$name.log | tr '\n' ' ' | cut -c1-500)" | tee -a $OUT/summary.txt; }
run t_all 900 python -m pytest tests -m gpu -x -q --durations=4
cat > scripts/ab_cases.sh <<'EOC'
bench epi8 ""
bench epi8_n125k "--items 125000"
bench epi16_n125k "--items 125000" B200_EPI_WARPS=16
bench c5_shard "--config c5 --items 625000 --users 151552"
EOC
bash scripts/gpu_ab.sh 303104 > $OUT/ab.log 2>&1
bash scripts/gpu_profile.sh > $OUT/profile.log 2>&1
cat $OUT/summary.txt; cat $OUT/ab_summary.txt; cat $OUT/profile.log
