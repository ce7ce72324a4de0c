bench gen3 ""
bench gen2 "" B200_TC_KERNEL=2
bench gen3_q2 "" B200_RANK_LIB=$L/libb200rank_q2.so
bench gen3_q8 "" B200_RANK_LIB=$L/libb200rank_q8.so
bench gen3_noview "--viewed 0"
bench gen3_q2_noview "--viewed 0" B200_RANK_LIB=$L/libb200rank_q2.so
bench gen3_q8_noview "--viewed 0" B200_RANK_LIB=$L/libb200rank_q8.so
bench gen3_n125k "--items 125000"
bench gen3_q8_n125k "--items 125000" B200_RANK_LIB=$L/libb200rank_q8.so
bench gen3_q2_n125k "--items 125000" B200_RANK_LIB=$L/libb200rank_q2.so
bench gen3_n125k_noview "--items 125000 --viewed 0"
bench gen3_n125k_dbg1 "--items 125000" B200_TC_DEBUG=1
bench gen3_kc11 "" B200_TC_KCAND=11
bench gen3_1M "--users 1000000"
