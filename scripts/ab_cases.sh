bench p4b2 ""
bench p2b2 "" B200_RANK_LIB=$L/libb200rank_p2.so
bench p8b2 "" B200_RANK_LIB=$L/libb200rank_p8.so
bench p8b3 "" B200_RANK_LIB=$L/libb200rank_p8b3.so
bench p16q8b4 "" B200_RANK_LIB=$L/libb200rank_p16q8b4.so
bench p4b2_n125k "--items 125000"
bench p8b2_n125k "--items 125000" B200_RANK_LIB=$L/libb200rank_p8.so
bench p8b3_n125k "--items 125000" B200_RANK_LIB=$L/libb200rank_p8b3.so
bench p16q8b4_n125k "--items 125000" B200_RANK_LIB=$L/libb200rank_p16q8b4.so
