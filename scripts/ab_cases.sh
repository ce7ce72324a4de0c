bench gen3 ""
bench gen3_noview "--viewed 0"
bench gen3_n125k "--items 125000"
bench gen3_n125k_noview "--items 125000 --viewed 0"
bench gen3_kc16 "" B200_TC_KCAND=16
bench gen3_k100_cos "--k 100 --distance cosine --users 75776"
bench gen3_d256_bf16_k20 "--dim 256 --tc bf16 --k 20 --items 625000 --users 151552"
bench gen3_1M "--users 1000000"
