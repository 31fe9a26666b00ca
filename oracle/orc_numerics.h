/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 * Restatements of the numerical utilities the RHS calls:
 *   rootFinder%find           source/numerical/root_finder.F90:587-1075 (Brent branch)
 *   gsl_root_fsolver_brent    libgsl 2.6 roots/brent.c        (not vendored)
 *   gsl_root_test_interval    libgsl 2.6 roots/convergence.c
 *   gsl_integration_qag (GK15) libgsl 2.6 integration/{qag.c,qk.c,qk15.c} (QUADPACK)
 *   table1DLinearLinear       source/objects/tables/_module.F90:1237-1405
 *   fastExponentiator         source/math/exponentiation.F90:57-104
 */
#ifndef ORC_NUMERICS_H
#define ORC_NUMERICS_H

#ifdef __cplusplus
extern "C" {
#endif

typedef double (*orc_fn1)(double x, void *ctx);

enum { ORC_EXPAND_NONE = 0, ORC_EXPAND_ADDITIVE = 1, ORC_EXPAND_MULTIPLICATIVE = 2 };
enum { ORC_SIGN_NONE = 0, ORC_SIGN_NEGATIVE = -1, ORC_SIGN_POSITIVE = 1 };

typedef struct orc_root_finder {
    orc_fn1 f;
    void *ctx;
    double tol_abs, tol_rel;
    /* rangeExpand() state */
    int expand_type;
    double expand_upward, expand_downward;
    int sign_expect_upward, sign_expect_downward;
    int upward_limit_set, downward_limit_set;
    double upward_limit, downward_limit;
    /* statistics */
    int n_eval, n_iter;
} orc_root_finder;

void orc_root_init(orc_root_finder *r, orc_fn1 f, void *ctx, double tol_abs, double tol_rel);
/* find with rootRange and (optionally) known function values at its ends; status 0 = ok */
double orc_root_find(orc_root_finder *r, double x_low, double x_high, int have_values, double f_low,
                     double f_high, int *status);

/* gsl_integration_qag with key = GSL_INTEG_GAUSS15; returns status (0 ok) */
int orc_qag15(orc_fn1 f, void *ctx, double a, double b, double epsabs, double epsrel, int limit,
              double *result, double *abserr, int *n_intervals);
void orc_qk15(orc_fn1 f, void *ctx, double a, double b, double *result, double *abserr,
              double *resabs, double *resasc);

/* value of a table1DLinearLinear with n points on [xmin,xmax] populated by g(x), extrapolation "fix"
 * or "extrapolate" (extrap_fix=1 clamps x into range first) */
double orc_linear_table_eval(orc_fn1 g, void *ctx, double xmin, double xmax, int n, double x,
                             int extrap_fix);
/* fastExponentiator(rangeMin,rangeMax,exponent,density,abortOutsideRange=.false.)%exponentiate(x) */
double orc_fast_exponentiate(double range_min, double range_max, double exponent, double density,
                             double x);

#ifdef __cplusplus
}
#endif
#endif
