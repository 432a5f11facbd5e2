/* empty: TBB is not used (RCPP_PARALLEL_USE_TBB == 0) */
