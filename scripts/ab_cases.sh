bench base "" B200_RANK_LIB=$L/libb200rank_base.so
bench coop ""
bench base_noview "--viewed 0" B200_RANK_LIB=$L/libb200rank_base.so
bench coop_noview "--viewed 0"
bench coop_kc10 "" B200_TC_KCAND=10
bench coop_kc11 "" B200_TC_KCAND=11
bench coop_1sm "" B200_TC_KERNEL=1
bench coop_dbg1 "" B200_TC_DEBUG=1
bench coop_n125k "--items 125000"
bench base_n125k "--items 125000" B200_RANK_LIB=$L/libb200rank_base.so
bench coop_1M "--users 1000000"
