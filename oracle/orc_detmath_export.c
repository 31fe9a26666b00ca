/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Exposes the shared deterministic elementary functions
 * (galacticus_b200/csrc/glc_detmath.h) to the tests so their accuracy can be checked against libm. */
#include <math.h>
#include "../galacticus_b200/csrc/glc_detmath.h"
#include "../galacticus_b200/csrc/glc_specfun.h"
void orc_dm_eval(int which, long n, const double *x, const double *y, double *out) {
    long i;
    for (i = 0; i < n; i++) {
        switch (which) {
        case 0: out[i] = dm_exp(x[i]); break;
        case 1: out[i] = dm_log(x[i]); break;
        case 2: out[i] = dm_pow(x[i], y[i]); break;
        case 3: out[i] = dm_atan(x[i]); break;
        default: out[i] = dm_cbrt(x[i]); break;
        }
    }
}

/* I_m(x; beta) of glc_specfun.h for tests/test_specfun.py */
void orc_dm_beta_moment(int m, long n, const double *x, double beta, double *out) {
    long i;
    for (i = 0; i < n; i++) out[i] = dm_beta_moment(m, x[i], beta);
}
