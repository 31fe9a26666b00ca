/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_numerics.h).
 */
#include "orc_numerics.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>

#include "../galacticus_b200/csrc/glc_detmath.h"
#include <string.h>

/* ------------------------------------------------------------------ Brent (GSL 2.6 roots/brent.c) */
typedef struct {
    double a, b, c, d, e, fa, fb, fc;
} brent_state;

static double sign1(double x) { return signbit(x) ? -1.0 : 1.0; } /* Fortran sign(1.0d0,x) */

static double eval(orc_root_finder *r, double x) {
    r->n_eval++;
    return r->f(x, r->ctx);
}

static int brent_init(orc_root_finder *r, brent_state *s, double *root, double x_lower, double x_upper,
                      double f_lower, double f_upper) {
    *root = 0.5 * (x_lower + x_upper);
    s->a = x_lower;
    s->fa = f_lower;
    s->b = x_upper;
    s->fb = f_upper;
    s->c = x_upper;
    s->fc = f_upper;
    s->d = x_upper - x_lower;
    s->e = x_upper - x_lower;
    (void)r;
    if ((f_lower < 0.0 && f_upper < 0.0) || (f_lower > 0.0 && f_upper > 0.0)) return -1;
    return 0;
}

static int brent_iterate(orc_root_finder *r, brent_state *s, double *root, double *x_lower,
                         double *x_upper) {
    double tol, m;
    int ac_equal = 0;
    double a = s->a, b = s->b, c = s->c, fa = s->fa, fb = s->fb, fc = s->fc, d = s->d, e = s->e;
    if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) {
        ac_equal = 1;
        c = a;
        fc = fa;
        d = b - a;
        e = b - a;
    }
    if (fabs(fc) < fabs(fb)) {
        ac_equal = 1;
        a = b;
        b = c;
        c = a;
        fa = fb;
        fb = fc;
        fc = fa;
    }
    tol = 0.5 * DBL_EPSILON * fabs(b);
    m = 0.5 * (c - b);
    if (fb == 0) {
        *root = b;
        *x_lower = b;
        *x_upper = b;
        return 0;
    }
    if (fabs(m) <= tol) {
        *root = b;
        if (b < c) {
            *x_lower = b;
            *x_upper = c;
        } else {
            *x_lower = c;
            *x_upper = b;
        }
        return 0;
    }
    if (fabs(e) < tol || fabs(fa) <= fabs(fb)) {
        d = m; /* bisection */
        e = m;
    } else {
        double p, q, rr;
        double sx = fb / fa;
        if (ac_equal) {
            p = 2 * m * sx;
            q = 1 - sx;
        } else {
            q = fa / fc;
            rr = fb / fc;
            p = sx * (2 * m * q * (q - rr) - (b - a) * (rr - 1));
            q = (q - 1) * (rr - 1) * (sx - 1);
        }
        if (p > 0)
            q = -q;
        else
            p = -p;
        {
            double t1 = 3 * m * q - fabs(tol * q), t2 = fabs(e * q);
            if (2 * p < (t1 < t2 ? t1 : t2)) {
                e = d;
                d = p / q;
            } else {
                d = m;
                e = m;
            }
        }
    }
    a = b;
    fa = fb;
    if (fabs(d) > tol)
        b += d;
    else
        b += (m > 0 ? +tol : -tol);
    fb = eval(r, b);
    s->a = a;
    s->b = b;
    s->c = c;
    s->d = d;
    s->e = e;
    s->fa = fa;
    s->fb = fb;
    s->fc = fc;
    *root = b;
    if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) c = a;
    if (b < c) {
        *x_lower = b;
        *x_upper = c;
    } else {
        *x_lower = c;
        *x_upper = b;
    }
    return 0;
}

/* gsl_root_test_interval */
static int test_interval(double x_lower, double x_upper, double epsabs, double epsrel) {
    const double abs_lower = fabs(x_lower), abs_upper = fabs(x_upper);
    double min_abs, tolerance;
    if ((x_lower > 0.0 && x_upper > 0.0) || (x_lower < 0.0 && x_upper < 0.0))
        min_abs = abs_lower < abs_upper ? abs_lower : abs_upper;
    else
        min_abs = 0;
    tolerance = epsabs + epsrel * min_abs;
    return fabs(x_upper - x_lower) < tolerance ? 0 : 1;
}

void orc_root_init(orc_root_finder *r, orc_fn1 f, void *ctx, double tol_abs, double tol_rel) {
    memset(r, 0, sizeof(*r));
    r->f = f;
    r->ctx = ctx;
    r->tol_abs = tol_abs;
    r->tol_rel = tol_rel;
}

/* rootFinderFind, numerical/root_finder.F90:587-1075, non-derivative branch, default
 * stopping criterion (interval), no testLimits */
double orc_root_find(orc_root_finder *r, double x_low, double x_high, int have_values, double f_low,
                     double f_high, int *status) {
    const int iteration_maximum = 1000;
    brent_state s;
    double x_root = 0.0, xl, xh;
    int iteration, st;
    *status = 0;
    if (!have_values) {
        f_low = eval(r, x_low);
        f_high = eval(r, x_high);
    }
    if (x_high == x_low) f_high = f_low;
    /* range expansion, :818-996 */
    while (sign1(f_low) * sign1(f_high) > 0.0 && f_low != 0.0 && f_high != 0.0) {
        int range_changed = 0, lower_ok, upper_ok;
        lower_ok = r->sign_expect_downward == ORC_SIGN_NEGATIVE ? (f_low < 0.0)
                 : r->sign_expect_downward == ORC_SIGN_POSITIVE ? (f_low > 0.0) : 0;
        upper_ok = r->sign_expect_upward == ORC_SIGN_NEGATIVE ? (f_high < 0.0)
                 : r->sign_expect_upward == ORC_SIGN_POSITIVE ? (f_high > 0.0) : 0;
        if (r->expand_type == ORC_EXPAND_ADDITIVE) {
            if (r->expand_upward > 0.0 && !upper_ok && (x_high < r->upward_limit || !r->upward_limit_set)) {
                if (lower_ok) {
                    x_low = x_high;
                    f_low = f_high;
                }
                x_high += r->expand_upward;
                if (r->upward_limit_set && x_high > r->upward_limit) x_high = r->upward_limit;
                f_high = eval(r, x_high);
                range_changed = 1;
            }
            if (r->expand_downward < 0.0 && !lower_ok && (x_low > r->downward_limit || !r->downward_limit_set)) {
                if (upper_ok) {
                    x_high = x_low;
                    f_high = f_low;
                }
                x_low += r->expand_downward;
                if (r->downward_limit_set && x_low < r->downward_limit) x_low = r->downward_limit;
                f_low = eval(r, x_low);
                range_changed = 1;
            }
        } else if (r->expand_type == ORC_EXPAND_MULTIPLICATIVE) {
            if (((r->expand_upward > 1.0 && x_high > 0.0) || (r->expand_upward < 1.0 && x_high < 0.0)) &&
                !upper_ok && (x_high < r->upward_limit || !r->upward_limit_set)) {
                if (lower_ok) {
                    x_low = x_high;
                    f_low = f_high;
                }
                x_high *= r->expand_upward;
                if (r->upward_limit_set && x_high > r->upward_limit) x_high = r->upward_limit;
                f_high = eval(r, x_high);
                range_changed = 1;
            }
            if (((r->expand_downward < 1.0 && x_low > 0.0) || (r->expand_downward > 1.0 && x_low < 0.0)) &&
                !lower_ok && (x_low > r->downward_limit || !r->downward_limit_set)) {
                if (upper_ok) {
                    x_high = x_low;
                    f_high = f_low;
                }
                x_low *= r->expand_downward;
                if (r->downward_limit_set && x_low < r->downward_limit) x_low = r->downward_limit;
                f_low = eval(r, x_low);
                range_changed = 1;
            }
        }
        if (!range_changed) {
            *status = 2; /* errorStatusOutOfRange: failed to bracket */
            return 0.0;
        }
    }
    if (brent_init(r, &s, &x_root, x_low, x_high, f_low, f_high) != 0) {
        *status = 3;
        return 0.0;
    }
    iteration = 0;
    x_root = 0.0;
    st = 0;
    for (;;) {
        iteration++;
        st = brent_iterate(r, &s, &x_root, &xl, &xh);
        r->n_iter++;
        if (st != 0 || iteration > iteration_maximum) break;
        if (iteration > 1) {
            st = test_interval(xl, xh, r->tol_abs, r->tol_rel);
            if (st == 0) break;
        }
    }
    if (st != 0) {
        *status = 4;
        return 0.0;
    }
    return x_root;
}

/* ------------------------------------------------ QAG, GK15 (GSL 2.6 integration/qk.c, qk15.c, qag.c) */
static const double xgk[8] = {0.991455371120812639206854697526329, 0.949107912342758524526189684047851,
                              0.864864423359769072789712788640926, 0.741531185599394439863864773280788,
                              0.586087235467691130294144838258730, 0.405845151377397166906606412076961,
                              0.207784955007898467600689403773245, 0.000000000000000000000000000000000};
static const double wg[4] = {0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
                             0.381830050505118944950369775488975, 0.417959183673469387755102040816327};
static const double wgk[8] = {0.022935322010529224963732008058970, 0.063092092629978553290700663189204,
                              0.104790010322250183839876322541518, 0.140653259715525918745189590510238,
                              0.169004726639267902826583426598550, 0.190350578064785409913256402421014,
                              0.204432940075298892414161999234649, 0.209482141084727828012999174891714};

static double rescale_error(double err, const double result_abs, const double result_asc) {
    err = fabs(err);
    if (result_asc != 0 && err != 0) {
        double scale = dm_pow((200 * err / result_asc), 1.5);
        if (scale < 1)
            err = result_asc * scale;
        else
            err = result_asc;
    }
    if (result_abs > DBL_MIN / (50 * DBL_EPSILON)) {
        double min_err = 50 * DBL_EPSILON * result_abs;
        if (min_err > err) err = min_err;
    }
    return err;
}

void orc_qk15(orc_fn1 f, void *ctx, double a, double b, double *result, double *abserr, double *resabs,
              double *resasc) {
    const int n = 8;
    double fv1[8], fv2[8];
    const double center = 0.5 * (a + b);
    const double half_length = 0.5 * (b - a);
    const double abs_half_length = fabs(half_length);
    const double f_center = f(center, ctx);
    double result_gauss = 0;
    double result_kronrod = f_center * wgk[n - 1];
    double result_abs = fabs(result_kronrod);
    double result_asc = 0;
    double mean = 0, err = 0;
    int j;
    if (n % 2 == 0) result_gauss = f_center * wg[n / 2 - 1];
    for (j = 0; j < (n - 1) / 2; j++) {
        const int jtw = j * 2 + 1;
        const double abscissa = half_length * xgk[jtw];
        const double fval1 = f(center - abscissa, ctx);
        const double fval2 = f(center + abscissa, ctx);
        const double fsum = fval1 + fval2;
        fv1[jtw] = fval1;
        fv2[jtw] = fval2;
        result_gauss += wg[j] * fsum;
        result_kronrod += wgk[jtw] * fsum;
        result_abs += wgk[jtw] * (fabs(fval1) + fabs(fval2));
    }
    for (j = 0; j < n / 2; j++) {
        int jtwm1 = j * 2;
        const double abscissa = half_length * xgk[jtwm1];
        const double fval1 = f(center - abscissa, ctx);
        const double fval2 = f(center + abscissa, ctx);
        fv1[jtwm1] = fval1;
        fv2[jtwm1] = fval2;
        result_kronrod += wgk[jtwm1] * (fval1 + fval2);
        result_abs += wgk[jtwm1] * (fabs(fval1) + fabs(fval2));
    }
    mean = result_kronrod * 0.5;
    result_asc = wgk[n - 1] * fabs(f_center - mean);
    for (j = 0; j < n - 1; j++) result_asc += wgk[j] * (fabs(fv1[j] - mean) + fabs(fv2[j] - mean));
    err = (result_kronrod - result_gauss) * half_length;
    result_kronrod *= half_length;
    result_abs *= abs_half_length;
    result_asc *= abs_half_length;
    *result = result_kronrod;
    *resabs = result_abs;
    *resasc = result_asc;
    *abserr = rescale_error(err, result_abs, result_asc);
}

static int subinterval_too_small(double a1, double a2, double b2) {
    const double e = DBL_EPSILON;
    const double u = DBL_MIN;
    double tmp = (1 + 100 * e) * (fabs(a2) + 1000 * u);
    return fabs(a1) <= tmp && fabs(b2) <= tmp;
}

#define ORC_QAG_LIMIT_MAX 1000

int orc_qag15(orc_fn1 f, void *ctx, double a, double b, double epsabs, double epsrel, int limit,
              double *result, double *abserr, int *n_intervals) {
    double alist[ORC_QAG_LIMIT_MAX], blist[ORC_QAG_LIMIT_MAX], rlist[ORC_QAG_LIMIT_MAX],
        elist[ORC_QAG_LIMIT_MAX];
    int size = 0, i_max = 0, iteration = 0, error_type = 0, roundoff_type1 = 0, roundoff_type2 = 0, i;
    double area, errsum, result0, abserr0, resabs0, resasc0, tolerance, round_off;
    if (limit > ORC_QAG_LIMIT_MAX) limit = ORC_QAG_LIMIT_MAX;
    *result = 0;
    *abserr = 0;
    if (n_intervals) *n_intervals = 1;
    orc_qk15(f, ctx, a, b, &result0, &abserr0, &resabs0, &resasc0);
    alist[0] = a;
    blist[0] = b;
    rlist[0] = result0;
    elist[0] = abserr0;
    size = 1;
    tolerance = fmax(epsabs, epsrel * fabs(result0));
    round_off = 50 * DBL_EPSILON * resabs0;
    if (abserr0 <= round_off && abserr0 > tolerance) {
        *result = result0;
        *abserr = abserr0;
        return 18; /* GSL_EROUND */
    } else if ((abserr0 <= tolerance && abserr0 != resasc0) || abserr0 == 0.0) {
        *result = result0;
        *abserr = abserr0;
        return 0;
    } else if (limit == 1) {
        *result = result0;
        *abserr = abserr0;
        return 11; /* GSL_EMAXITER */
    }
    area = result0;
    errsum = abserr0;
    iteration = 1;
    do {
        double a1, b1, a2, b2, a_i, b_i, r_i, e_i, area1 = 0, area2 = 0, area12 = 0, error1 = 0, error2 = 0,
            error12 = 0, resasc1, resasc2, resabs1, resabs2;
        /* retrieve: the interval with the largest error estimate */
        i_max = 0;
        for (i = 1; i < size; i++)
            if (elist[i] > elist[i_max]) i_max = i;
        a_i = alist[i_max];
        b_i = blist[i_max];
        r_i = rlist[i_max];
        e_i = elist[i_max];
        a1 = a_i;
        b1 = 0.5 * (a_i + b_i);
        a2 = b1;
        b2 = b_i;
        orc_qk15(f, ctx, a1, b1, &area1, &error1, &resabs1, &resasc1);
        orc_qk15(f, ctx, a2, b2, &area2, &error2, &resabs2, &resasc2);
        area12 = area1 + area2;
        error12 = error1 + error2;
        errsum += (error12 - e_i);
        area += area12 - r_i;
        if (resasc1 != error1 && resasc2 != error2) {
            double delta = r_i - area12;
            if (fabs(delta) <= 1.0e-5 * fabs(area12) && error12 >= 0.99 * e_i) roundoff_type1++;
            if (iteration >= 10 && error12 > e_i) roundoff_type2++;
        }
        tolerance = fmax(epsabs, epsrel * fabs(area));
        if (errsum > tolerance) {
            if (roundoff_type1 >= 6 || roundoff_type2 >= 20) error_type = 2;
            if (subinterval_too_small(a1, a2, b2)) error_type = 3;
        }
        /* update */
        if (error2 > error1) {
            alist[i_max] = a2;
            rlist[i_max] = area2;
            elist[i_max] = error2;
            alist[size] = a1;
            blist[size] = b1;
            rlist[size] = area1;
            elist[size] = error1;
        } else {
            blist[i_max] = b1;
            rlist[i_max] = area1;
            elist[i_max] = error1;
            alist[size] = a2;
            blist[size] = b2;
            rlist[size] = area2;
            elist[size] = error2;
        }
        size++;
        iteration++;
    } while (iteration < limit && !error_type && errsum > tolerance);
    {
        double sum = 0;
        for (i = 0; i < size; i++) sum += rlist[i];
        *result = sum;
    }
    *abserr = errsum;
    if (n_intervals) *n_intervals = size;
    if (errsum <= tolerance) return 0;
    if (error_type == 2) return 18;
    if (error_type == 3) return 21; /* GSL_ESING */
    if (iteration == limit) return 11;
    return -1;
}

/* ------------------------------------------------------------------ tables */
double orc_linear_table_eval(orc_fn1 g, void *ctx, double xmin, double xmax, int n, double x, int extrap_fix) {
    /* Table_Linear_1D_Create / _Interpolate, objects/tables/_module.F90:1237-1405 */
    const double dx = (xmax - xmin) / (double)(n - 1);
    const double inverse_dx = 1.0 / ((xmin + dx) - xmin);
    double xe = x, h, xi, xi1;
    int i;
    if (extrap_fix) {
        if (xe < xmin) xe = xmin;
        if (xe > xmax) xe = xmax;
    }
    if (xe < xmin)
        i = 1;
    else if (xe >= xmax)
        i = n - 1;
    else {
        i = (int)((xe - xmin) * inverse_dx) + 1;
        if (i > n - 1) i = n - 1;
        if (i < 1) i = 1;
    }
    xi = xmin + dx * (double)(i - 1);
    xi1 = (i == n - 1) ? xmax : xmin + dx * (double)i;
    h = (xe - xi) * inverse_dx;
    return g(xi, ctx) * (1.0 - h) + g(xi1, ctx) * h;
}

static double powfn(double x, void *ctx) { return dm_pow(x, *(double *)ctx); }

/* The reference tabulates x^exponent once per fastExponentiator object (math/exponentiation.F90:57-104).  The
 * values are kept in a small per-process cache keyed by the table's definition; a cached entry is exactly the
 * dm_pow value the on-the-fly evaluation would produce, so results do not change -- only the CPU baseline gets
 * the reference's own cost profile (table look-ups instead of pow calls). */
#define ORC_POW_TABLES 4
static struct {
    double range_min, range_max, exponent, density;
    int n;
    double *v;
} orc_pow_cache[ORC_POW_TABLES];
static int orc_pow_cache_n = 0;

static const double *orc_pow_table(double range_min, double range_max, double exponent, double density, int *n_out) {
    int t, k;
    for (t = 0; t < orc_pow_cache_n; t++)
        if (orc_pow_cache[t].range_min == range_min && orc_pow_cache[t].range_max == range_max &&
            orc_pow_cache[t].exponent == exponent && orc_pow_cache[t].density == density) {
            *n_out = orc_pow_cache[t].n;
            return orc_pow_cache[t].v;
        }
    {
        const double *result = 0;
#pragma omp critical(orc_pow_cache_build)
        {
            for (t = 0; t < orc_pow_cache_n; t++)
                if (orc_pow_cache[t].range_min == range_min && orc_pow_cache[t].range_max == range_max &&
                    orc_pow_cache[t].exponent == exponent && orc_pow_cache[t].density == density)
                    break;
            if (t == orc_pow_cache_n && orc_pow_cache_n < ORC_POW_TABLES) {
                const int n = (int)((range_max - range_min) * density) + 1;
                const double dx = (range_max - range_min) / (double)(n - 1);
                double *v = (double *)malloc(sizeof(double) * (size_t)n);
                for (k = 0; k < n; k++) v[k] = dm_pow((k == n - 1) ? range_max : range_min + dx * (double)k, exponent);
                orc_pow_cache[t].range_min = range_min;
                orc_pow_cache[t].range_max = range_max;
                orc_pow_cache[t].exponent = exponent;
                orc_pow_cache[t].density = density;
                orc_pow_cache[t].n = n;
                orc_pow_cache[t].v = v;
#pragma omp flush
                orc_pow_cache_n = t + 1;
            }
            if (t < orc_pow_cache_n) {
                *n_out = orc_pow_cache[t].n;
                result = orc_pow_cache[t].v;
            }
        }
        return result;
    }
}

double orc_fast_exponentiate(double range_min, double range_max, double exponent, double density, double x) {
    /* math/exponentiation.F90:57-104 */
    int n, i;
    const double *v;
    if (x < range_min || x > range_max) return dm_pow(x, exponent);
    v = orc_pow_table(range_min, range_max, exponent, density, &n);
    if (!v) return orc_linear_table_eval(powfn, &exponent, range_min, range_max, (int)((range_max - range_min) * density) + 1, x, 0);
    {
        const double dx = (range_max - range_min) / (double)(n - 1);
        const double inverse_dx = 1.0 / ((range_min + dx) - range_min);
        double xi, h;
        if (x >= range_max)
            i = n - 1;
        else {
            i = (int)((x - range_min) * inverse_dx) + 1;
            if (i > n - 1) i = n - 1;
            if (i < 1) i = 1;
        }
        xi = range_min + dx * (double)(i - 1);
        h = (x - xi) * inverse_dx;
        return v[i - 1] * (1.0 - h) + v[i] * h;
    }
}
