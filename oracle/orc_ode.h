/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * CPU restatement of the ODE machinery on Galacticus' node-evolution hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may build, load or call anything in oracle/.
 *
 * Restates (reference paths relative to /root/reference):
 *   - gsl_odeiv2_step_rkck            libgsl 2.6 ode-initval2/rkck.c  (NOT vendored in
 *                                     the reference; pinned at 2.6 by
 *                                     docs/manuals/user-guide/installation/source-linux.rst:59)
 *   - gsl_odeiv2_evolve_apply/_reset  libgsl 2.6 ode-initval2/evolve.c
 *   - sc2_control_hadjust             source/external/gslODEInitVal2/cscal2.c:93-169
 *   - gsl_odeiv2_driver2_apply        source/external/gslODEInitVal2/driver2.c:148-250
 *   - odeSolverSolve                  source/numerical/ODE_solver/solver.F90:492-636
 *
 * Parity pinning: the accuracy-level KAT of source/tests/ODE_solver.F90:78-90
 * (y'=sin x, default RKCK + scaled2 control) is checked in tests/test_oracle_ode.py.
 * Step-sequence-level parity with libgsl is UNPINNED (no reference test holds
 * step sequences, and libgsl is absent from this image).
 */
#ifndef ORC_ODE_H
#define ORC_ODE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* GSL status codes used on the path (gsl_errno.h, GSL 2.6). */
enum {
    ORC_GSL_SUCCESS  = 0,
    ORC_GSL_FAILURE  = -1,
    ORC_GSL_CONTINUE = -2,
    ORC_GSL_EFAULT   = 3,
    ORC_GSL_EINVAL   = 4,
    ORC_GSL_EBADFUNC = 9   /* == odeSolverInterrupt, source/numerical/ODE_solver/error_codes.F90:31 */
};

#define ORC_HADJ_INC 1
#define ORC_HADJ_NIL 0
#define ORC_HADJ_DEC (-1)

#define ORC_ODE_DIM_MAX 64

/* RHS: returns ORC_GSL_SUCCESS, or ORC_GSL_EBADFUNC for an interrupt. */
typedef int (*orc_rhs_fn)(double t, const double *y, double *dydt, void *ctx);
/* postStep(t, y, &status): may edit y; status != 0 forces an evolve reset (driver2.c:206-217). */
typedef void (*orc_poststep_fn)(double t, double *y, int *status, void *ctx);

/* stepAnalyzer(t, t1, y, yerr, last_step, status): driver2.c:195-198 */
typedef void (*orc_analyzer_fn)(double t, double t1, const double *y, const double *yerr, double h, int status, void *ctx);

typedef struct {
    /* system */
    size_t dim;
    orc_rhs_fn f;
    void *ctx;
    /* control (cscal2.c state) */
    double eps_abs, eps_rel, a_y, a_dydt;
    double scale_abs[ORC_ODE_DIM_MAX];
    int is_non_negative[ORC_ODE_DIM_MAX];
    /* evolve (evolve.c state) */
    double y0[ORC_ODE_DIM_MAX], yerr[ORC_ODE_DIM_MAX];
    double dydt_in[ORC_ODE_DIM_MAX], dydt_out[ORC_ODE_DIM_MAX];
    double last_step;
    unsigned long count, failed_steps;
    /* rkck workspace */
    double k1[ORC_ODE_DIM_MAX], k2[ORC_ODE_DIM_MAX], k3[ORC_ODE_DIM_MAX];
    double k4[ORC_ODE_DIM_MAX], k5[ORC_ODE_DIM_MAX], k6[ORC_ODE_DIM_MAX];
    double ytmp[ORC_ODE_DIM_MAX], ystep0[ORC_ODE_DIM_MAX];
    /* driver */
    double h;
    unsigned long n;
    /* odeSolver-level */
    orc_poststep_fn post_step;
    orc_analyzer_fn analyzer; /* NULL unless profileOdeEvolver */
    double interrupted_at_x; /* module variable interruptedAtX (ODE_Solver_Error_Codes) */
    /* statistics (not in the reference; used for the metric "node-ODE steps") */
    unsigned long n_steps_accepted, n_steps_rejected, n_rhs;
} orc_ode_solver;

void orc_ode_init(orc_ode_solver *s, size_t dim, orc_rhs_fn f, void *ctx,
                  double eps_abs, double eps_rel, const double *scale,
                  const int *is_non_negative, orc_poststep_fn post_step);

/* odeSolverSolve: integrate from *x0 to x1; *x_step in = initial step guess
 * (<=0: use x1-x0), out = driver step size after success. Returns GSL status. */
int orc_ode_solve(orc_ode_solver *s, double *x0, double x1, double *y, double *x_step);

/* individual pieces exported for unit tests */
int orc_sc2_hadjust(const orc_ode_solver *s, unsigned int ord, const double *y,
                    const double *yerr, const double *yp, double *h);
int orc_rkck_apply(orc_ode_solver *s, double t, double h, double *y, double *yerr,
                   const double *dydt_in, double *dydt_out);
int orc_evolve_apply(orc_ode_solver *s, double *t, double t1, double *h, double *y);
int orc_driver2_apply(orc_ode_solver *s, double *t, double t1, double *y);

/* KAT entry (tests/ODE_solver.F90:52-90): y'=sin x from x0 to x1, tol 1e-9, scale 1. */
double orc_kat_sin(double x0, double x1, double y0, unsigned long *n_steps);
/* harmonic oscillator y''=-y from 0 to x1 (active part of tests/ODE_solver.F90:93-110) */
void orc_kat_harmonic(double x1, double *y_out);

#ifdef __cplusplus
}
#endif
#endif
