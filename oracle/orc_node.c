/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.
 * Restatement of standardEvolve / standardODEs / standardPostStepProcessing
 * (source/merger_trees/node_evolver/standard.F90).
 */
#include "orc_node.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "../galacticus_b200/csrc/glc_detmath.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TRIAL_COUNT_MAXIMUM 8 /* standard.F90:135 */

/* ---- standardODEs, standard.F90:831-946 --------------------------------------------
 * Differences from the reference, none of which alters results for this operator set:
 *  - the (timePrevious,yPrevious) memo (:871-881) is omitted: it can only hit when the
 *    solver re-evaluates the RHS at a bit-identical (t,y), which does not occur on the
 *    RKCK path (FSAL copies dydt_out without calling the RHS);
 *  - the isAccurate guard (:887-891) is omitted: nodeOperatorCosmicTime sets basic%time
 *    to the requested time in solveAnalytics, so the guard cannot trigger. */
static int standard_odes(double time, const double *y, double *dydt, void *vctx) {
    orc_evolve_ctx *c = (orc_evolve_ctx *)vctx;
    double rate[GLC_NY];
    int i, code;
    /* node%deserializeValues(y) */
    for (i = 0; i < c->n_active; i++) c->p[c->active[i]] = y[i];
    orc_model_solve_analytics(c, time);
    if (c->interrupt_first_found && time >= c->time_interrupt_first) {
        for (i = 0; i < c->n_active; i++) dydt[i] = 0.0;
        orc_model_solve_analytics(c, c->time_interrupt_first);
        return ORC_GSL_SUCCESS;
    }
    for (i = 0; i < GLC_NY; i++) rate[i] = 0.0; /* node%odeStepRatesInitialize() */
    code = orc_model_rates(c, time, rate);
    if (code == GLC_INT_NONE) {
        for (i = 0; i < c->n_active; i++) {
            dydt[i] = rate[c->active[i]];
            if (!isfinite(dydt[i])) {
                if (getenv("ORC_DEBUG_NAN") && !c->nonfinite) {
                    int k;
                    fprintf(stderr, "[orc] non-finite rate of property %d at t=%.17g\n", c->active[i], time);
                    for (k = 0; k < GLC_NPROP; k++) fprintf(stderr, "  p[%d]=%.17g rate=%g\n", k, c->p[k], k < GLC_NY ? rate[k] : 0.0);
                }
                c->nonfinite = 1;
            }
        }
        return ORC_GSL_SUCCESS;
    }
    for (i = 0; i < c->n_active; i++) dydt[i] = 0.0;
    if (time < c->time_interrupt_first || !c->interrupt_first_found) {
        c->interrupt_first_found = 1;
        c->time_interrupt_first = time;
        c->interrupt_first_code = code;
        c->solver->interrupted_at_x = time;
        return ORC_GSL_EBADFUNC; /* odeSolverInterrupt */
    }
    return ORC_GSL_SUCCESS;
}

/* standardPostStepProcessing, standard.F90:1160-1185 */
static void standard_post_step(double time, double *y, int *status, void *vctx) {
    orc_evolve_ctx *c = (orc_evolve_ctx *)vctx;
    int i;
    for (i = 0; i < c->n_active; i++) c->p[c->active[i]] = y[i];
    orc_model_solve_analytics(c, time);
    orc_model_post_step(c, status);
    if (*status != ORC_GSL_SUCCESS)
        for (i = 0; i < c->n_active; i++) y[i] = c->p[c->active[i]];
    if (*status == ORC_GSL_CONTINUE) *status = ORC_GSL_SUCCESS;
}

/* ---- profiling: standardStepErrorAnalyzer (standard.F90:1187-1239) -> mergerTreeEvolveProfilerSimple::profile
 * (merger_trees/evolve/profiler/simple.F90:250-304) */
static glc_profile g_profile;
void orc_profiler_reset(const glc_params *P) {
    int i, n;
    double l0, l1;
    memset(&g_profile, 0, sizeof(g_profile));
    /* simple.F90:128-147: Make_Range(min, max, n, logarithmic), n = int(log10(max / min) * pointsPerDecade) + 1 */
    n = (int)(log10(P->profilerTimeStepMaximum / P->profilerTimeStepMinimum) * (double)P->profilerTimeStepPointsPerDecade) + 1;
    if (n < 2) n = 2;
    if (n > GLC_PROFILE_BINS) n = GLC_PROFILE_BINS;
    l0 = dm_log(P->profilerTimeStepMinimum);
    l1 = dm_log(P->profilerTimeStepMaximum);
    for (i = 0; i < n; i++) g_profile.time_step[i] = dm_exp(l0 + (l1 - l0) * (double)i / (double)(n - 1));
    g_profile.n_bins = n;
    g_profile.time_step_smallest = DBL_MAX;
}
void orc_profiler_read(glc_profile *out) { *out = g_profile; }

static void standard_step_error_analyzer(double time, double time_end, const double *y, const double *yerr, double time_step,
                                         int step_status, void *vctx) {
    orc_evolve_ctx *c = (orc_evolve_ctx *)vctx;
    double scaled_error_maximum = 0.0;
    int i, limiting = -1, lo, hi;
    (void)time;
    (void)time_end;
    c->evals_to_success++;
    if (step_status != ORC_GSL_SUCCESS) return;
    for (i = 0; i < c->n_active; i++) {
        const double scale = c->P->odeToleranceAbsolute * c->scale[i] + c->P->odeToleranceRelative * fabs(y[i]);
        const double scaled_error = fabs(yerr[i]) / scale;
        if (scaled_error > scaled_error_maximum) {
            scaled_error_maximum = scaled_error;
            limiting = c->active[i];
        }
    }
    /* searchArray = gsl_interp_bsearch over the bin edges */
    lo = 0;
    hi = g_profile.n_bins - 1;
    while (hi > lo + 1) {
        const int mid = (hi + lo) >> 1;
        if (g_profile.time_step[mid] > time_step)
            hi = mid;
        else
            lo = mid;
    }
#ifdef _OPENMP
#pragma omp critical(orc_profile)
#endif
    {
        g_profile.time_step_count[lo] += 1;
        g_profile.evaluation_count[lo] += c->evals_to_success;
        if (c->interrupt_first_found) {
            g_profile.time_step_count_interrupted[lo] += 1;
            g_profile.evaluation_count_interrupted[lo] += c->evals_to_success;
        }
        if (limiting >= 0)
            g_profile.property_hits[limiting] += 1;
        else
            g_profile.property_hits_unknown += 1;
        if (time_step < g_profile.time_step_smallest) g_profile.time_step_smallest = time_step;
    }
    c->evals_to_success = 0;
}

static int is_non_negative_prop(int prop) {
    /* isNonNegative="true" attributes of the component definitions; only the satellite
       bound mass (satellite/standard.F90:46-49) is not flagged */
    return prop != GLC_P_SAT_BOUND_MASS;
}

int orc_evolve_node_segment(const glc_params *P, const orc_tables *T, double *props, int *flags,
                            double time_end, int *interrupt, glc_counters *C) {
    orc_evolve_ctx c;
    orc_ode_solver solver;
    double y[GLC_NY], y_saved[GLC_NY], scale[GLC_NY], scale_by_prop[GLC_NY];
    int nonneg[GLC_NY];
    double time_start, time_start_saved, step_size;
    int i, ode_status, trial_count, status = GLC_STATUS_SUCCESS;

    memset(&c, 0, sizeof(c));
    c.P = P;
    c.T = T;
    c.p = props;
    c.flags = *flags;
    c.solver = &solver;
    *interrupt = GLC_INT_NONE;
    if (C) C->segments++;

    orc_model_pre_evolve(&c); /* differentialEvolutionPre + preEvolveTask, :434-441 */
    time_start = props[GLC_P_TIME];
    time_start_saved = time_start;
    c.n_active = orc_model_active_list(&c, c.active);
    for (i = 0; i < c.n_active; i++) {
        y[i] = props[c.active[i]];
        y_saved[i] = y[i];
        nonneg[i] = P->enforceNonNegativity ? is_non_negative_prop(c.active[i]) : 0;
    }
    for (i = 0; i < GLC_NY; i++) scale_by_prop[i] = 0.0;
    orc_model_scales(&c, scale_by_prop);
    for (i = 0; i < c.n_active; i++) scale[i] = scale_by_prop[c.active[i]];

    c.interrupt_first_found = 0;
    c.time_interrupt_first = 0.0;
    c.interrupt_first_code = GLC_INT_NONE;
    step_size = -1.0;
    if (time_start != time_end && c.n_active > 0) {
        trial_count = 0;
        ode_status = ORC_GSL_FAILURE;
        orc_ode_init(&solver, (size_t)c.n_active, standard_odes, &c, P->odeToleranceAbsolute,
                     P->odeToleranceRelative, scale, nonneg, standard_post_step);
        c.scale = scale;
        c.evals_to_success = 0;
        if (P->profileOdeEvolver) solver.analyzer = standard_step_error_analyzer; /* errorAnalyzer, :633 */
        while (trial_count < TRIAL_COUNT_MAXIMUM &&
               !(ode_status == ORC_GSL_SUCCESS || ode_status == ORC_GSL_EBADFUNC)) {
            if (P->reuseODEStepSize)
                step_size = props[GLC_P_TIME_STEP] / dm_scale2(1.0, trial_count);
            else
                step_size = -1.0;
            time_start = time_start_saved;
            ode_status = orc_ode_solve(&solver, &time_start, time_end, y, &step_size);
            if (P->enforceNonNegativity &&
                !(ode_status == ORC_GSL_SUCCESS || ode_status == ORC_GSL_EBADFUNC)) {
                int any = 0;
                for (i = 0; i < c.n_active; i++)
                    if (y[i] < 0.0 && nonneg[i]) any = 1;
                if (any) {
                    for (i = 0; i < c.n_active; i++)
                        if (y[i] < 0.0 && nonneg[i]) y[i] = 0.0;
                    step_size = props[GLC_P_TIME_STEP] / dm_scale2(1.0, trial_count);
                    ode_status = ORC_GSL_SUCCESS;
                }
            }
            if (!(ode_status == ORC_GSL_SUCCESS || ode_status == ORC_GSL_EBADFUNC)) {
                trial_count++;
                if (C) C->trials_failed++;
                for (i = 0; i < c.n_active; i++) y[i] = y_saved[i];
            }
        }
        if (C) {
            C->steps_accepted += solver.n_steps_accepted;
            C->steps_rejected += solver.n_steps_rejected;
            C->rhs_evaluations += solver.n_rhs;
        }
        if (!(ode_status == ORC_GSL_SUCCESS || ode_status == ORC_GSL_EBADFUNC)) {
            /* status=errorStatusUnderflow, :717-719 (node left at its saved values) */
            for (i = 0; i < c.n_active; i++) props[c.active[i]] = y_saved[i];
            orc_model_solve_analytics(&c, time_start_saved);
            return GLC_STATUS_UNDERFLOW;
        }
    }
    /* :726-740 */
    for (i = 0; i < c.n_active; i++) props[c.active[i]] = y[i];
    orc_model_solve_analytics(&c, time_end);
    if (c.time_interrupt_first != 0.0) {
        *interrupt = c.interrupt_first_code;
        /* NB: as in the reference, analytic properties stay at their time_end values
           (solveAnalytics(node,timeEnd) :728) while basic%time is set to the interrupt time */
        props[GLC_P_TIME] = c.time_interrupt_first;
        props[GLC_P_TIME_STEP] = -1.0;
    } else {
        props[GLC_P_TIME] = time_end;
        props[GLC_P_TIME_STEP] = step_size;
    }
    orc_model_post_evolve(&c); /* differentialEvolutionPost + postEvolve hooks, :744-753 */
    if (c.nonfinite) status = GLC_STATUS_NONFINITE;
    *flags = c.flags;
    return status;
}

static int evolve_one(const glc_params *P, const orc_tables *T, double *props, int *flags,
                      double time_end, int *interrupt, glc_counters *C) {
    /* the host loop of evolver/standard.F90:425-476 for component-creation interrupts */
    int status;
    int guard = 0;
    for (;;) {
        status = orc_evolve_node_segment(P, T, props, flags, time_end, interrupt, C);
        if (status != GLC_STATUS_SUCCESS) return status;
        if (*interrupt == GLC_INT_NONE) return status;
        if (!P->resolveInterruptsOnDevice) return status;
        orc_apply_interrupt(P, props, flags, *interrupt);
        *interrupt = GLC_INT_NONE;
        if (!(props[GLC_P_TIME] < time_end)) return status;
        if (++guard > 64) return GLC_STATUS_FAIL;
    }
}

int orc_evolve_batch(const glc_params *P, const orc_tables *T, long n, double *props, int *flags,
                     const double *time_end, int *status, int *interrupt, glc_counters *C,
                     int n_threads) {
    glc_counters total;
    memset(&total, 0, sizeof(total));
#ifdef _OPENMP
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel num_threads(n_threads)
#endif
    {
        glc_counters local;
        long i;
        memset(&local, 0, sizeof(local));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 16)
#endif
        for (i = 0; i < n; i++) {
            status[i] = evolve_one(P, T, props + i * GLC_NPROP, flags + i, time_end[i],
                                   interrupt + i, &local);
            local.nodes++;
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        {
            total.steps_accepted += local.steps_accepted;
            total.steps_rejected += local.steps_rejected;
            total.rhs_evaluations += local.rhs_evaluations;
            total.segments += local.segments;
            total.trials_failed += local.trials_failed;
            total.nodes += local.nodes;
        }
    }
    (void)n_threads;
    if (C) *C = total;
    return 0;
}

int orc_rhs_node(const glc_params *P, const orc_tables *T, double *props, int flags, double *dydt,
                 int *interrupt) {
    orc_evolve_ctx c;
    int i;
    memset(&c, 0, sizeof(c));
    c.P = P;
    c.T = T;
    c.p = props;
    c.flags = flags;
    c.n_active = orc_model_active_list(&c, c.active);
    orc_model_solve_analytics(&c, props[GLC_P_TIME]);
    for (i = 0; i < GLC_NY; i++) dydt[i] = 0.0;
    *interrupt = orc_model_rates(&c, props[GLC_P_TIME], dydt);
    return 0;
}

/* standardErrorHandler (node_evolver/standard.F90:1063-1140) + standardODEStepTolerances (:1142-1158): the "ODE system
   parameters" table of a node, restated for glc_error_report_node: y, dy/dt at the node's time, yScale, yTolerance, and as
   yError the embedded error of one Cash-Karp step of size h from the node's state after the pre-evolve hooks (the reference
   reads the failed solver's own estimate; a batch interface has no such state).  out7: [7][GLC_NY] = y, dydt, scale,
   tolerance, error, error_scaled, active. */
int orc_error_report(const glc_params *P, const orc_tables *T, const double *record, int flags, double h, double *out7,
                     int *interrupt) {
    orc_evolve_ctx c;
    orc_ode_solver solver;
    double props[GLC_NPROP], y[GLC_NY], y0[GLC_NY], scale[GLC_NY], scale_by_prop[GLC_NY], yerr[GLC_NY], k1[GLC_NY], out[GLC_NY];
    int nonneg[GLC_NY];
    int i, st;
    memcpy(props, record, sizeof props);
    memset(&c, 0, sizeof(c));
    c.P = P;
    c.T = T;
    c.p = props;
    c.flags = flags;
    c.solver = &solver;
    *interrupt = GLC_INT_NONE;
    orc_model_pre_evolve(&c);
    c.n_active = orc_model_active_list(&c, c.active);
    for (i = 0; i < c.n_active; i++) {
        y[i] = props[c.active[i]];
        y0[i] = y[i];
        nonneg[i] = 0;
    }
    for (i = 0; i < GLC_NY; i++) scale_by_prop[i] = 0.0;
    orc_model_scales(&c, scale_by_prop);
    for (i = 0; i < c.n_active; i++) scale[i] = scale_by_prop[c.active[i]];
    orc_ode_init(&solver, (size_t)c.n_active, standard_odes, &c, P->odeToleranceAbsolute, P->odeToleranceRelative, scale, nonneg,
                 standard_post_step);
    c.scale = scale;
    for (i = 0; i < 7 * GLC_NY; i++) out7[i] = 0.0;
    st = standard_odes(props[GLC_P_TIME], y, k1, &c);
    if (st == ORC_GSL_EBADFUNC) *interrupt = c.interrupt_first_code;
    /* one embedded step; a frozen system past an interrupt gives zero rates, as on the device */
    for (i = 0; i < c.n_active; i++) {
        y[i] = y0[i];
        yerr[i] = 0.0;
    }
    (void)out;
    orc_rkck_apply(&solver, record[GLC_P_TIME], h, y, yerr, k1, NULL);
    for (i = 0; i < c.n_active; i++) {
        const int pidx = c.active[i];
        const double tol = P->odeToleranceRelative * fabs(y0[i]) + P->odeToleranceAbsolute * scale[i];
        out7[0 * GLC_NY + pidx] = y0[i];
        out7[1 * GLC_NY + pidx] = k1[i];
        out7[2 * GLC_NY + pidx] = scale[i];
        out7[3 * GLC_NY + pidx] = tol;
        out7[4 * GLC_NY + pidx] = yerr[i];
        out7[5 * GLC_NY + pidx] = fabs(yerr[i]) / tol;
        out7[6 * GLC_NY + pidx] = 1.0;
    }
    return 0;
}
