/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_ode.h).
 * CPU restatement of the GSL-2.6 RKCK stepper / evolve loop, Galacticus' scaled2
 * error controller, its modified driver and odeSolverSolve.
 */
#include "orc_ode.h"

#include <float.h>
#include <math.h>

#include "../galacticus_b200/csrc/glc_detmath.h"
#include <string.h>
#ifdef ORC_TRACE
#include <stdio.h>
#define TR(...) fprintf(stderr, __VA_ARGS__)
#else
#define TR(...)
#endif

/* ---- Cash-Karp tableau: libgsl 2.6 ode-initval2/rkck.c (published coefficients,
 *      Cash & Karp 1990, ACM TOMS 16, 201) -------------------------------------- */
static const double ah[] = {1.0 / 5.0, 0.3, 3.0 / 5.0, 1.0, 7.0 / 8.0};
static const double b21 = 1.0 / 5.0;
static const double b3[] = {3.0 / 40.0, 9.0 / 40.0};
static const double b4[] = {0.3, -0.9, 1.2};
static const double b5[] = {-11.0 / 54.0, 2.5, -70.0 / 27.0, 35.0 / 27.0};
static const double b6[] = {1631.0 / 55296.0, 175.0 / 512.0, 575.0 / 13824.0,
                            44275.0 / 110592.0, 253.0 / 4096.0};
static const double c1 = 37.0 / 378.0;
static const double c3 = 250.0 / 621.0;
static const double c4 = 125.0 / 594.0;
static const double c6 = 512.0 / 1771.0;
/* fifth minus fourth order weights */
static const double ec[] = {0.0,
                            37.0 / 378.0 - 2825.0 / 27648.0,
                            0.0,
                            250.0 / 621.0 - 18575.0 / 48384.0,
                            125.0 / 594.0 - 13525.0 / 55296.0,
                            -277.0 / 14336.0,
                            512.0 / 1771.0 - 0.25};
#define RKCK_ORDER 5u /* rkck_order() returns 5 in libgsl */

void orc_ode_init(orc_ode_solver *s, size_t dim, orc_rhs_fn f, void *ctx, double eps_abs,
                  double eps_rel, const double *scale, const int *is_non_negative,
                  orc_poststep_fn post_step) {
    memset(s, 0, sizeof(*s));
    s->dim = dim;
    s->f = f;
    s->ctx = ctx;
    s->eps_abs = eps_abs;
    s->eps_rel = eps_rel;
    s->a_y = 1.0;    /* yScale    default, solver.F90:365 */
    s->a_dydt = 0.0; /* dydtScale default, solver.F90:366 */
    for (size_t i = 0; i < dim; i++) {
        s->scale_abs[i] = scale ? scale[i] : 1.0;
        s->is_non_negative[i] = is_non_negative ? is_non_negative[i] : 0;
    }
    s->post_step = post_step;
    s->h = 1.0; /* hStart default, solver.F90:367 */
}

static int fn_eval(orc_ode_solver *s, double t, const double *y, double *dydt) {
    s->n_rhs++;
    return s->f(t, y, dydt, s->ctx);
}

/* source/external/gslODEInitVal2/cscal2.c:93-169 */
int orc_sc2_hadjust(const orc_ode_solver *s, unsigned int ord, const double *y,
                    const double *yerr, const double *yp, double *h) {
    const double S = 0.9;
    const double h_old = *h;
    double rmax = DBL_MIN;
    int forbidden_negatives = 0;
    for (size_t i = 0; i < s->dim; i++) {
        const double D0 = s->eps_rel * (s->a_y * fabs(y[i]) + s->a_dydt * fabs(h_old * yp[i])) +
                          s->eps_abs * s->scale_abs[i];
        const double r = fabs(yerr[i]) / fabs(D0);
        rmax = (r > rmax) ? r : rmax; /* GSL_MAX_DBL(r, rmax) */
        if (s->is_non_negative[i] && y[i] < 0.0) forbidden_negatives = 1;
    }
    if (rmax > 1.1) {
        double r = S / dm_pow(rmax, 1.0 / ord);
        if (r < 0.2) r = 0.2;
        *h = r * h_old;
        return ORC_HADJ_DEC;
    } else if (forbidden_negatives == 1) {
        *h = 0.5 * h_old;
        return ORC_HADJ_DEC;
    } else if (rmax < 0.5) {
        double r = S / dm_pow(rmax, 1.0 / (ord + 1.0));
        if (r > 4.9) r = 4.9;
        if (r < 1.0) r = 1.0;
        *h = r * h_old;
        return ORC_HADJ_INC;
    }
    return ORC_HADJ_NIL;
}

/* libgsl 2.6 rkck.c: rkck_apply */
int orc_rkck_apply(orc_ode_solver *s, double t, double h, double *y, double *yerr,
                   const double *dydt_in, double *dydt_out) {
    const size_t dim = s->dim;
    double *k1 = s->k1, *k2 = s->k2, *k3 = s->k3, *k4 = s->k4, *k5 = s->k5, *k6 = s->k6;
    double *ytmp = s->ytmp, *y0 = s->ystep0;
    size_t i;
    int st;
    memcpy(y0, y, dim * sizeof(double));
    if (dydt_in != NULL) {
        memcpy(k1, dydt_in, dim * sizeof(double));
    } else {
        st = fn_eval(s, t, y, k1);
        if (st != ORC_GSL_SUCCESS) return st;
    }
    for (i = 0; i < dim; i++) ytmp[i] = y[i] + b21 * h * k1[i];
    st = fn_eval(s, t + ah[0] * h, ytmp, k2);
    if (st != ORC_GSL_SUCCESS) return st;
    for (i = 0; i < dim; i++) ytmp[i] = y[i] + h * (b3[0] * k1[i] + b3[1] * k2[i]);
    st = fn_eval(s, t + ah[1] * h, ytmp, k3);
    if (st != ORC_GSL_SUCCESS) return st;
    for (i = 0; i < dim; i++) ytmp[i] = y[i] + h * (b4[0] * k1[i] + b4[1] * k2[i] + b4[2] * k3[i]);
    st = fn_eval(s, t + ah[2] * h, ytmp, k4);
    if (st != ORC_GSL_SUCCESS) return st;
    for (i = 0; i < dim; i++)
        ytmp[i] = y[i] + h * (b5[0] * k1[i] + b5[1] * k2[i] + b5[2] * k3[i] + b5[3] * k4[i]);
    st = fn_eval(s, t + ah[3] * h, ytmp, k5);
    if (st != ORC_GSL_SUCCESS) return st;
    for (i = 0; i < dim; i++)
        ytmp[i] = y[i] + h * (b6[0] * k1[i] + b6[1] * k2[i] + b6[2] * k3[i] + b6[3] * k4[i] +
                              b6[4] * k5[i]);
    st = fn_eval(s, t + ah[4] * h, ytmp, k6);
    if (st != ORC_GSL_SUCCESS) return st;
    for (i = 0; i < dim; i++) {
        const double d_i = c1 * k1[i] + c3 * k3[i] + c4 * k4[i] + c6 * k6[i];
        y[i] += h * d_i;
    }
    if (dydt_out != NULL) {
        st = fn_eval(s, t + h, y, dydt_out);
        if (st != ORC_GSL_SUCCESS) {
            memcpy(y, y0, dim * sizeof(double));
            return st;
        }
    }
    for (i = 0; i < dim; i++)
        yerr[i] = h * (ec[1] * k1[i] + ec[3] * k3[i] + ec[4] * k4[i] + ec[5] * k5[i] + ec[6] * k6[i]);
    return ORC_GSL_SUCCESS;
}

static void evolve_reset(orc_ode_solver *s) {
    s->count = 0;
    s->failed_steps = 0;
    s->last_step = 0.0;
}

/* libgsl 2.6 evolve.c: gsl_odeiv2_evolve_apply (rkck: can_use_dydt_in = 1) */
int orc_evolve_apply(orc_ode_solver *s, double *t, double t1, double *h, double *y) {
    const double t0 = *t;
    double h0 = *h;
    int step_status;
    int final_step = 0;
    const double dt = t1 - t0;
    const size_t dim = s->dim;

    if ((dt < 0.0 && h0 > 0.0) || (dt > 0.0 && h0 < 0.0)) return ORC_GSL_EINVAL;
    memcpy(s->y0, y, dim * sizeof(double));
    if (s->count == 0) {
        int status = fn_eval(s, t0, y, s->dydt_in);
        if (status) return status;
    } else {
        memcpy(s->dydt_in, s->dydt_out, dim * sizeof(double));
    }

    for (;;) { /* try_step: */
        if ((dt >= 0.0 && h0 > dt) || (dt < 0.0 && h0 < dt)) {
            h0 = dt;
            final_step = 1;
        } else {
            final_step = 0;
        }
        step_status = orc_rkck_apply(s, t0, h0, y, s->yerr, s->dydt_in, s->dydt_out);
        TR("attempt t0=%.17g h0=%.17g t1=%.17g final=%d count=%lu status=%d\n", t0, h0, t1, final_step, s->count, step_status);
        if (step_status == ORC_GSL_EFAULT || step_status == ORC_GSL_EBADFUNC) return step_status;
        if (step_status != ORC_GSL_SUCCESS) {
            const double h_old = h0;
            h0 *= 0.5;
            {
                volatile double t_curr = *t;
                volatile double t_next = (*t) + h0;
                if (fabs(h0) < fabs(h_old) && t_next != t_curr) {
                    memcpy(y, s->y0, dim * sizeof(double));
                    s->failed_steps++;
                    continue;
                } else {
                    *h = h0;
                    *t = t0;
                    return step_status;
                }
            }
        }
        s->count++;
        s->last_step = h0;
        if (final_step)
            *t = t1;
        else
            *t = t0 + h0;
        {
            const double h_old = h0;
            const int hadjust_status = orc_sc2_hadjust(s, RKCK_ORDER, y, s->yerr, s->dydt_out, &h0);
            TR("  hadjust=%d h_old=%.17g h_new=%.17g\n", hadjust_status, h_old, h0);
            if (hadjust_status == ORC_HADJ_DEC) {
                volatile double t_curr = *t;
                volatile double t_next = (*t) + h0;
                if (fabs(h0) < fabs(h_old) && t_next != t_curr) {
                    memcpy(y, s->y0, dim * sizeof(double));
                    s->failed_steps++;
                    s->n_steps_rejected++;
                    continue;
                } else {
                    *h = h0;
                    return ORC_GSL_FAILURE;
                }
            }
        }
        break;
    }
    if (final_step == 0) *h = h0;
    s->n_steps_accepted++;
    return step_status;
}

/* source/external/gslODEInitVal2/driver2.c:148-250 (latentIntegrator, stepAnalyzer = NULL;
 * hmin=0, hmax=DBL_MAX, nmax=0 so those guards are inert, driver2.c:94-97) */
int orc_driver2_apply(orc_ode_solver *s, double *t, double t1, double *y) {
    int sign;
    s->n = 0;
    sign = (s->h > 0.0) ? 1 : -1;
    if (sign * (t1 - *t) < 0.0) return ORC_GSL_EINVAL;
    while (sign * (t1 - *t) > 0.0) {
        int st = orc_evolve_apply(s, t, t1, &s->h, y);
        if (s->analyzer != NULL) s->analyzer(*t, t1, y, s->yerr, s->last_step, st, s->ctx); /* driver2.c:195-198 */
        if (st != ORC_GSL_SUCCESS) return st;
        if (s->post_step != NULL) {
            int ps = ORC_GSL_SUCCESS;
            s->post_step(*t, y, &ps, s->ctx);
            if (ps != 0) evolve_reset(s);
        }
        s->n++;
    }
    return ORC_GSL_SUCCESS;
}

/* source/numerical/ODE_solver/solver.F90:492-636 (no latent variables) */
int orc_ode_solve(orc_ode_solver *s, double *x0, double x1, double *y, double *x_step) {
    double y0[ORC_ODE_DIM_MAX];
    double x_step_ = x1 - *x0;
    double x1_ = x1, x;
    int status_ = ORC_GSL_SUCCESS;
    const int evolve_forward = x1 > *x0;
    if (x_step && *x_step > 0.0) x_step_ = (*x_step < x_step_) ? *x_step : x_step_;
    memcpy(y0, y, s->dim * sizeof(double));
    x = *x0;
    evolve_reset(s); /* GSL_ODEIV2_Driver_Reset = evolve reset + step reset (rkck reset is a no-op on results) */
    if (x_step_ != 0.0) s->h = x_step_;
    while ((evolve_forward && x < x1_) || (!evolve_forward && x > x1_)) {
        status_ = orc_driver2_apply(s, &x, x1_, y);
        if (status_ == ORC_GSL_SUCCESS) {
            if (x_step) *x_step = s->h;
        } else if (status_ == ORC_GSL_EBADFUNC) {
            x1_ = s->interrupted_at_x;
            if (x > x1_) {
                memcpy(y, y0, s->dim * sizeof(double));
                x = *x0;
                evolve_reset(s);
            }
        } else {
            /* GSL_FAILURE and everything else: hand status back (status present on the node path) */
            *x0 = x;
            return status_;
        }
    }
    *x0 = x;
    return status_;
}

/* ------------------------------- KATs ------------------------------------------ */
static int kat_sin_rhs(double t, const double *y, double *dydt, void *ctx) {
    (void)y;
    (void)ctx;
    dydt[0] = sin(t);
    return ORC_GSL_SUCCESS;
}

double orc_kat_sin(double x0, double x1, double y0, unsigned long *n_steps) {
    orc_ode_solver s;
    const double scale[1] = {1.0};
    double y[1] = {y0};
    double xs = x0;
    orc_ode_init(&s, 1, kat_sin_rhs, NULL, 1.0e-9, 1.0e-9, scale, NULL, NULL);
    orc_ode_solve(&s, &xs, x1, y, NULL);
    if (n_steps) *n_steps = s.n_steps_accepted;
    return y[0];
}

static int kat_harm_rhs(double t, const double *y, double *dydt, void *ctx) {
    (void)t;
    (void)ctx;
    dydt[0] = y[1];
    dydt[1] = -1.0 * y[0];
    return ORC_GSL_SUCCESS;
}

void orc_kat_harmonic(double x1, double *y_out) {
    orc_ode_solver s;
    const double scale[2] = {1.0, 1.0};
    double y[2] = {1.0, 0.0};
    double xs = 0.0;
    orc_ode_init(&s, 2, kat_harm_rhs, NULL, 1.0e-9, 1.0e-9, scale, NULL, NULL);
    orc_ode_solve(&s, &xs, x1, y, NULL);
    y_out[0] = y[0];
    y_out[1] = y[1];
}
