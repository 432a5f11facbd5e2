/* stand-in for RcppParallel: no TBB -> the reference uses its serial PARALLEL_FOR macros
 * (src/LibHLA.cpp:151-159) */
#define RCPP_PARALLEL_USE_TBB 0
