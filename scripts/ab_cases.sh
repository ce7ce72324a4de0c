# round-2 starting point: the experimental 16-epilogue-warp kernel against the default
bench gen3 ""
bench gen4 "" B200_TC_KERNEL=4
bench gen4_kc10 "" B200_TC_KERNEL=4 B200_TC_KCAND=10
bench gen4_kc6 "" B200_TC_KERNEL=4 B200_TC_KCAND=6
bench gen3_n125k "--items 125000"
bench gen4_n125k "--items 125000" B200_TC_KERNEL=4
bench gen4_1M "--users 1000000" B200_TC_KERNEL=4
