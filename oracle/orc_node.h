/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.
 * CPU restatement of mergerTreeNodeEvolverStandard::standardEvolve and its RHS
 * (source/merger_trees/node_evolver/standard.F90:385-755, 831-946, 1019-1061, 1160-1185).
 * Node-record layout and parameter struct are the public ones of include/glc_b200.h
 * (the boundary both sides share); nothing else of the product is used.
 */
#ifndef ORC_NODE_H
#define ORC_NODE_H

#include "../include/glc_b200.h"
#include "orc_ode.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_table2d {
    int n0, n1;
    double *x0, *x1, *v; /* v[n0][n1] */
} orc_table2d;

typedef struct orc_tables {
    orc_table2d t[GLC_NTABLES];
    /* derived (set by orc_tables_finalize) for the CIE tables, CIE_file.F90:627-659 */
    int cooling_log, cooling_first_z_zero;
    double cooling_first_nonzero_z;
    double *cooling_lnZ, *cooling_lnT, *cooling_lnL;
    int electron_log, electron_first_z_zero;
    double electron_first_nonzero_z;
    double *electron_lnZ, *electron_lnT, *electron_lnV;
} orc_tables;

orc_tables *orc_tables_create(void);
void orc_tables_destroy(orc_tables *T);
int orc_tables_set(orc_tables *T, int table_id, int n0, int n1, const double *x0,
                   const double *x1, const double *values);

/* working state of one standardEvolve call */
typedef struct orc_evolve_ctx {
    const glc_params *P;
    const orc_tables *T;
    double *p;  /* node record, GLC_NPROP */
    int flags;
    int n_active;
    int active[GLC_NY]; /* y index -> property index */
    /* interrupt bookkeeping, standard.F90:573-576,916-928 */
    int interrupt_first_found;
    double time_interrupt_first;
    int interrupt_first_code;
    orc_ode_solver *solver;
    int nonfinite;
    const double *scale;          /* by y index: propertyScalesActive */
    unsigned long evals_to_success; /* countEvaluationsToSuccess (per call here; an object member in the reference) */
} orc_evolve_ctx;

/* one standardEvolve call: evolves the record to time_end or the first interrupt.
 * returns status (enum glc_status); *interrupt = code if interrupted. */
int orc_evolve_node_segment(const glc_params *P, const orc_tables *T, double *props, int *flags,
                            double time_end, int *interrupt, glc_counters *C);

/* batched form with the semantics of glc_evolve_batch (same argument meaning);
 * n_threads > 1 mirrors the reference's OpenMP-over-trees by OpenMP-over-nodes. */
int orc_evolve_batch(const glc_params *P, const orc_tables *T, long n, double *props, int *flags,
                     const double *time_end, int *status, int *interrupt, glc_counters *C,
                     int n_threads);

/* apply an interrupt the way the host's functionInterrupt would (component creation) */
void orc_apply_interrupt(const glc_params *P, double *props, int *flags, int code);

/* one RHS evaluation (standardODEs) at the record's own time; dydt[GLC_NY] indexed by property */
int orc_error_report(const glc_params *P, const orc_tables *T, const double *record, int flags, double h, double *out7,
                     int *interrupt);
int orc_rhs_node(const glc_params *P, const orc_tables *T, double *props, int flags, double *dydt,
                 int *interrupt);

void orc_params_default(glc_params *P, int model);

/* mergerTreeEvolveProfilerSimple fed by standardStepErrorAnalyzer when P->profileOdeEvolver (process-wide accumulators) */
void orc_profiler_reset(const glc_params *P);
void orc_profiler_read(glc_profile *out);

/* model hooks (orc_model_box.c, orc_model_standard.c) */
int orc_model_active_list(const orc_evolve_ctx *c, int *active);
void orc_model_scales(orc_evolve_ctx *c, double *scale_by_prop);
void orc_model_solve_analytics(orc_evolve_ctx *c, double time);
/* returns 0 or an interrupt code; rates indexed by property (GLC_NY) */
int orc_model_rates(orc_evolve_ctx *c, double time, double *rate);
void orc_model_post_step(orc_evolve_ctx *c, int *status);
void orc_model_post_evolve(orc_evolve_ctx *c);
void orc_model_pre_evolve(orc_evolve_ctx *c);

#ifdef __cplusplus
}
#endif
#endif
