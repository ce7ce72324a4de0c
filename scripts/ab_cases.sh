bench redux ""
bench votes "" B200_RANK_LIB=$L/libb200rank_votes.so
bench redux_n125k "--items 125000"
bench votes_n125k "--items 125000" B200_RANK_LIB=$L/libb200rank_votes.so
bench redux_1M "--users 1000000"
