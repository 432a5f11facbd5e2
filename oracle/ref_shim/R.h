/* Minimal stand-in for <R.h> so the UNMODIFIED reference sources (src/LibHLA.cpp and
 * src/LibHLA_ext_*.cpp) compile without an R installation. Test infrastructure only:
 * nothing in the product path (hibag_b200/) includes this. The functions are defined in
 * oracle/ref_driver.cpp. */
#ifndef HIBAG_B200_ORACLE_R_SHIM_H
#define HIBAG_B200_ORACLE_R_SHIM_H

#include <cfloat>
#include <climits>
#include <cmath>
#include <cstddef>

#ifdef __cplusplus
extern "C" {
#endif

double unif_rand(void);                    /* R's Mersenne-Twister stream (ref_driver.cpp) */
void Rprintf(const char *fmt, ...);
void Rf_error(const char *fmt, ...);       /* throws std::runtime_error in the shim */
void R_CheckUserInterrupt(void);

#ifdef __cplusplus
}
#endif

typedef enum { FALSE = 0, TRUE = 1 } Rboolean;
#define NA_INTEGER   INT_MIN
#define R_FINITE(x)  std::isfinite(x)

static inline Rboolean R_ToplevelExec(void (*fun)(void *), void *data)
{
	fun(data);
	return TRUE;
}

#endif
