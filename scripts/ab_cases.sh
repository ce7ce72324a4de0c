# A/B cases of scripts/gpu_ab.sh (sourced): label, extra bench args, environment
bench epi8 ""
bench epi16 "" B200_EPI_WARPS=16
bench epi16_kc10 "" B200_EPI_WARPS=16 B200_TC_KCAND=10
bench epi16_kc6 "" B200_EPI_WARPS=16 B200_TC_KCAND=6
bench epi8_n125k "--items 125000"
bench epi16_n125k "--items 125000" B200_EPI_WARPS=16
bench c3_wide "--config c3 --users 151552"
bench c3_multipass "--config c3 --users 75776" B200_WIDE=0
bench c5_shard "--config c5 --items 625000 --users 151552"
bench c5_shard16 "--config c5 --items 625000 --users 151552" B200_EPI_WARPS=16
