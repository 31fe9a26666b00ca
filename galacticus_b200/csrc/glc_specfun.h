/*
 * glc_specfun.h -- special functions of the path that have no closed form, shared by the CUDA code and the CPU checker like
 * glc_detmath.h (same source on both sides => same bits; only correctly rounded operations and the dm_* elementary functions).
 *
 * dm_beta_moment(m, x, beta) = I_m(x) = int_0^x t^m (1 + t^2)^(-3 beta / 2) dt
 *                            = x^(m+1) / (m+1) * 2F1((m+1)/2, 3 beta/2; (m+3)/2; -x^2)
 * is what every general-beta branch of the reference's beta profile evaluates through Hypergeometric_2F1
 * (source/mass_distributions/spherical/beta_profile.F90: density normalisation :258, enclosed mass :430, radial moments
 * :606-612; the vendored GSL approximation source/external/gslSpecFuncApprox/hyperg_2F1.c needs GSL headers and cannot be
 * built here).  The integrand is analytic with its nearest singularities at t = +-i, so composite 16-point Gauss-Legendre
 * quadrature on panels no wider than 1 is exact to rounding (checked against scipy.special.hyp2f1 in
 * tests/test_specfun.py); no series switching, no transformation formulas, no gamma functions.
 */
#ifndef GLC_SPECFUN_H
#define GLC_SPECFUN_H

#include "glc_detmath.h"

GLC_HD_BIG double dm_beta_moment(int m, double x, double beta) {
    /* abscissae / weights of the 16-point Gauss-Legendre rule on [-1, 1] (positive half; the rule is symmetric) */
    const double gx[8] = {9.50125098376374405e-02, 2.81603550779258915e-01, 4.58016777657227370e-01, 6.17876244402643771e-01, 7.55404408355002999e-01, 8.65631202387831755e-01, 9.44575023073232600e-01, 9.89400934991649939e-01};
    const double gw[8] = {1.89450610455068641e-01, 1.82603415044923639e-01, 1.69156519395002647e-01, 1.49595988816576708e-01, 1.24628971255534071e-01, 9.51585116824926053e-02, 6.22535239386474565e-02, 2.71524594117541762e-02};
    double sum = 0.0, width, a;
    int np, p, k;
    if (!(x > 0.0)) return 0.0;
    a = 1.5 * beta;
    if (x < 1.0e-3) { /* leading terms of the series: t^m (1 - a t^2 + ...) */
        const double lead = (m == 2) ? x * x * x / 3.0 : ((m == 3) ? x * x * x * x / 4.0 : dm_pow(x, (double)(m + 1)) / (double)(m + 1));
        return lead * (1.0 - a * ((double)(m + 1) / (double)(m + 3)) * x * x);
    }
    np = (int)x + 1;
    if (np > 64) np = 64;
    width = x / (double)np;
    for (p = 0; p < np; p++) {
        const double mid = ((double)p + 0.5) * width, half = 0.5 * width;
        double panel = 0.0;
        for (k = 0; k < 8; k++) {
            const double t0 = mid - half * gx[k], t1 = mid + half * gx[k];
            const double p0 = (m == 2) ? t0 * t0 : ((m == 3) ? t0 * t0 * t0 : dm_pow(t0, (double)m));
            const double p1 = (m == 2) ? t1 * t1 : ((m == 3) ? t1 * t1 * t1 : dm_pow(t1, (double)m));
            const double f0 = p0 * dm_exp(-a * dm_log(1.0 + t0 * t0));
            const double f1 = p1 * dm_exp(-a * dm_log(1.0 + t1 * t1));
            panel += gw[k] * (f0 + f1);
        }
        sum += half * panel;
    }
    return sum;
}

#endif /* GLC_SPECFUN_H */
