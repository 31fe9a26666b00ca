/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.
 * CPU restatement of the tree-level caller of the hot path: mergerTreeEvolverStandard::evolve
 * (source/merger_trees/evolver/standard.F90:291-635) walking ONE tree at a time, depth first, and handing ONE node at a
 * time to the node evolver (:452), as the reference does; OpenMP over trees mirrors tasks/evolve_forests/_class.F90:622.
 *
 *   evolvability            standardNodeIsEvolvable            evolver/standard.F90:723-760
 *   time limits             standardTimeEvolveTo               evolver/standard.F90:762-1035
 *                           mergerTreeEvolveTimestepSimple     evolve/timesteps/simple.F90
 *   promotion / node merger standardPromote / standardMerge    node_evolver/standard.F90:1241-1356
 *                           mergerTreeNodeMergerSingleLevelHierarchy
 *   hooks                   nodeOperatorDMOInterpolate         dark_matter_only_mass/interpolate.F90:84-291
 *                           nodeOperatorCGMAccretion           circumgalactic_medium/accretion.F90:144-426
 *                           accretionHaloSimple                accretion/halo/simple.F90
 *
 * Scope (identical on the product side): satellites are visited before their host; no satellite
 * merging times / galaxy mergers (satellites live to the end of the tree); no tree or node events; the cap that keeps a
 * primary progenitor from leading its siblings (:984-1000) is not applied.  The walk order differs from the product's
 * bulk-synchronous rounds; the per-node sequences of (state, end time) -- and therefore the results -- do not.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../galacticus_b200/csrc/glc_detmath.h"
#include "orc_constants.h"
#include "orc_node.h"

enum { T_PENDING = 0, T_ISOLATED = 1, T_SATELLITE = 2, T_PROMOTED = 3 };

typedef struct tree_ctx {
    const glc_params *P;
    const orc_tables *T;
    long n;
    const int *parent;
    const double *mass, *time, *scale, *angmom;
    double *rec;
    int *flags, *state;
    int *first_child, *sibling, *host, *children_left, *first_sat, *next_sat; /* satellite lists kept in ascending index */
    double *time_end;
    /* what cgmAccretionNodesMerge needs of a merged progenitor: unaccreted mass / metals and halo mass at merger */
    double *merged_unaccreted, *merged_unaccreted_abund, *merged_mass;
    unsigned char *merged_hot;
} tree_ctx;

static double *R(const tree_ctx *c, long i) { return c->rec + i * GLC_NPROP; }

static double expansion_timescale(const tree_ctx *c, double t) {
    /* 1/H(t) of a flat matter + Lambda universe in closed form (cosmologyFunctionsMatterLambda) */
    const double OL = 1.0 - c->P->OmegaMatter;
    const double H0 = c->P->HubbleConstant / ORC_MPC_PER_KMS_TO_GYR;
    const double e = dm_exp(2.0 * (1.5 * sqrt(OL) * H0 * t));
    return 1.0 / (H0 * sqrt(OL) * ((e + 1.0) / (e - 1.0)));
}
static double timestep(const tree_ctx *c, double t) {
    /* mergerTreeEvolveTimestepSimple, evolve/timesteps/simple.F90: min(timeStepRelative / H, timeStepAbsolute) */
    return c->P->timestepSimpleRelative > 0.0 ? fmin(c->P->timestepSimpleRelative * expansion_timescale(c, t), c->P->timestepSimpleAbsolute)
                                              : c->P->timestepSimpleAbsolute;
}
static double timestep_host(const tree_ctx *c, double t) {
    /* evolver/standard.F90:942-968: min(timestepHostRelative / H, timestepHostAbsolute); the absolute value alone when the
       relative one is not positive (:951-954) */
    return c->P->timestepHostRelative > 0.0 ? fmin(c->P->timestepHostRelative * expansion_timescale(c, t), c->P->timestepHostAbsolute)
                                            : c->P->timestepHostAbsolute;
}
static double virial_velocity(const tree_ctx *c, double m, double t) {
    /* virial_density_contrast.F90:195-417 through the tabulated mean halo density */
    const orc_table2d *tb = &c->T->t[GLC_TABLE_HALO_MEAN_DENSITY];
    const double l0 = dm_log(tb->x0[0]), ln = dm_log(tb->x0[tb->n0 - 1]);
    const double inv = (double)(tb->n0 - 1) / (ln - l0), lt = dm_log(t);
    const double x = (lt - l0) * inv;
    int i = (int)x;
    double h, rho, rvir;
    if (lt < l0) i = 0;
    if (i > tb->n0 - 2) i = tb->n0 - 2;
    if (i < 0) i = 0;
    h = x - (double)i;
    rho = tb->v[2 * i] * (1.0 - h) + tb->v[2 * (i + 1)] * h;
    rvir = dm_cbrt(3.0 * m / 4.0 / ORC_PI / rho);
    return sqrt(ORC_G_INTERNAL * m / rvir);
}
static double failed_fraction(const tree_ctx *c, double m, double t) {
    return (t > c->P->timeReionization && virial_velocity(c, m, t) < c->P->velocitySuppressionReionization) ? 1.0 : 0.0;
}
static int is_primary(const tree_ctx *c, int i) { return c->parent[i] >= 0 && c->first_child[c->parent[i]] == i; }
static double node_time(const tree_ctx *c, int i) { return c->state[i] == T_PENDING ? c->time[i] : R(c, i)[GLC_P_TIME]; }

static void set_targets(const tree_ctx *c, int i, double *r) {
    /* dmoInterpolateNodeInitialize :84-200 (+ the scale-radius and angular-momentum interpolators) */
    const int p = c->parent[i];
    double unresolved, dt;
    int k;
    r[GLC_P_MASS_TARGET] = c->mass[i];
    r[GLC_P_MASS_RATE] = 0.0;
    r[GLC_P_TIME_TARGET] = c->time[i];
    r[GLC_P_DMSCALE_TARGET] = c->scale[i];
    r[GLC_P_DMSCALE_RATE] = 0.0;
    r[GLC_P_SPIN_TARGET] = c->angmom[i];
    r[GLC_P_SPIN_RATE] = 0.0;
    if (p < 0) return;
    unresolved = c->mass[p];
    for (k = c->first_child[p]; k >= 0; k = c->sibling[k]) unresolved = unresolved - c->mass[k];
    dt = c->time[p] - c->time[i];
    r[GLC_P_TIME_TARGET] = c->time[p];
    if (unresolved > 0.0) {
        if (is_primary(c, i)) {
            if (dt > 0.0) r[GLC_P_MASS_RATE] = unresolved / dt;
            r[GLC_P_MASS_TARGET] = c->mass[i] + unresolved;
        }
    } else {
        const double total = c->mass[p] - unresolved;
        if (dt > 0.0) r[GLC_P_MASS_RATE] = (unresolved / dt) * (c->mass[i] / total);
        r[GLC_P_MASS_TARGET] = c->mass[i] + unresolved * c->mass[i] / total;
    }
    if (is_primary(c, i) && dt > 0.0) {
        r[GLC_P_DMSCALE_TARGET] = c->scale[p];
        r[GLC_P_DMSCALE_RATE] = (c->scale[p] - c->scale[i]) / dt;
        r[GLC_P_SPIN_TARGET] = c->angmom[p];
        r[GLC_P_SPIN_RATE] = (c->angmom[p] - c->angmom[i]) / dt;
    }
}

static void sat_insert(tree_ctx *c, int h, int s) {
    int *link = &c->first_sat[h];
    c->host[s] = h;
    while (*link >= 0 && *link < s) link = &c->next_sat[*link];
    c->next_sat[s] = *link;
    *link = s;
}
static void sats_move(tree_ctx *c, int from, int to) {
    int s = c->first_sat[from];
    c->first_sat[from] = -1;
    while (s >= 0) {
        const int nx = c->next_sat[s];
        sat_insert(c, to, s);
        s = nx;
    }
}
static double baryons(const tree_ctx *c, int i) {
    const double *r = R(c, i);
    double m = 0.0;
    if (c->flags[i] & GLC_F_HAS_HOTHALO) m += r[GLC_P_HH_MASS] + r[GLC_P_HH_OUTFLOWED_MASS];
    if (c->flags[i] & GLC_F_HAS_DISK) m += r[GLC_P_DISK_MASS_GAS] + r[GLC_P_DISK_MASS_STELLAR];
    if (c->flags[i] & GLC_F_HAS_SPHEROID) m += r[GLC_P_SPH_MASS_GAS] + r[GLC_P_SPH_MASS_STELLAR];
    if (c->flags[i] & GLC_F_HAS_BH) m += r[GLC_P_BH_MASS];
    return m;
}

static void apply_merged_progenitors(tree_ctx *c, int p) {
    /* the parent's side of cgmAccretionNodesMerge :281-362 for every merged progenitor, in progenitor order.  That side is
       order dependent; the reference applies it in arrival order, a property of its walk.  Checker and product apply it at
       the parent's own promotion in progenitor order: schedule independent, and identical to arrival order whenever a node
       has at most one non-primary progenitor. */
    double *rp = R(c, p);
    const double fb = c->P->OmegaBaryon / c->P->OmegaMatter;
    int k;
    for (k = c->first_child[p]; k >= 0; k = c->sibling[k]) {
        double failed, acc_hot, unacc;
        if (!c->merged_hot[k]) continue;
        c->flags[p] |= GLC_F_HAS_HOTHALO;
        rp[GLC_P_HH_UNACCRETED_MASS] = rp[GLC_P_HH_UNACCRETED_MASS] + c->merged_unaccreted[k];
        rp[GLC_P_HH_UNACCRETED_ABUND] = rp[GLC_P_HH_UNACCRETED_ABUND] + c->merged_unaccreted_abund[k];
        failed = failed_fraction(c, c->mass[p], c->time[p]);
        acc_hot = fb * c->mass[p] * (1.0 - failed);
        unacc = fb * c->mass[p] * failed;
        if (acc_hot > 0.0) {
            const double fraction = acc_hot / (acc_hot + unacc);
            const double re = rp[GLC_P_HH_UNACCRETED_MASS] * fraction * c->merged_mass[k] / c->mass[p];
            rp[GLC_P_HH_UNACCRETED_MASS] = rp[GLC_P_HH_UNACCRETED_MASS] - re;
            rp[GLC_P_HH_MASS] = rp[GLC_P_HH_MASS] + re;
            rp[GLC_P_HH_ANGMOM] = rp[GLC_P_HH_ANGMOM] + re * c->angmom[p] / c->mass[p];
        }
    }
}

static void node_merge(tree_ctx *c, int i, glc_forest_counters *fc) {
    /* standardMerge :1329-1356 + the node's side of cgmAccretionNodesMerge :265-280 + dmoInterpolateNodesMerge :277-291 */
    const int p = c->parent[i];
    double *r = R(c, i);
    if (c->flags[i] & GLC_F_HAS_HOTHALO) {
        c->merged_hot[i] = 1;
        c->merged_unaccreted[i] = r[GLC_P_HH_UNACCRETED_MASS];
        c->merged_unaccreted_abund[i] = r[GLC_P_HH_UNACCRETED_ABUND];
        c->merged_mass[i] = r[GLC_P_BASIC_MASS];
        r[GLC_P_HH_UNACCRETED_MASS] = 0.0;
        r[GLC_P_HH_UNACCRETED_ABUND] = 0.0;
    }
    r[GLC_P_MASS_RATE] = 0.0;
    r[GLC_P_MASS_TARGET] = r[GLC_P_BASIC_MASS];
    r[GLC_P_DMSCALE_RATE] = 0.0;
    r[GLC_P_DMSCALE_TARGET] = r[GLC_P_DMSCALE];
    r[GLC_P_SPIN_RATE] = 0.0;
    r[GLC_P_SPIN_TARGET] = r[GLC_P_SPIN];
    r[GLC_P_TIME_LAST_ISOLATED] = r[GLC_P_TIME];
    r[GLC_P_SAT_BOUND_MASS] = r[GLC_P_BASIC_MASS];
    r[GLC_P_MASS_BARYONIC_SUBHALOS] = 0.0;
    c->flags[i] |= GLC_F_IS_SATELLITE;
    c->state[i] = T_SATELLITE;
    sat_insert(c, p, i);
    sats_move(c, i, p);
    c->children_left[p]--;
    fc->node_mergers++;
}

static void node_promote(tree_ctx *c, int i, glc_forest_counters *fc) {
    /* standardPromote :1241-1327 + cgmAccretionNodePromote :202-263 + dmoInterpolateNodePromote :241-275 */
    const int p = c->parent[i];
    double *r = R(c, i), *rp = R(c, p);
    apply_merged_progenitors(c, p);
    if (c->flags[p] & GLC_F_HAS_HOTHALO) {
        c->flags[i] |= GLC_F_HAS_HOTHALO;
        if (r[GLC_P_HH_MASS] <= 0.0) r[GLC_P_HH_MASS] = r[GLC_P_HH_ANGMOM] = r[GLC_P_HH_ABUND] = 0.0;
        r[GLC_P_HH_UNACCRETED_MASS] = r[GLC_P_HH_UNACCRETED_MASS] + rp[GLC_P_HH_UNACCRETED_MASS];
        r[GLC_P_HH_MASS] = r[GLC_P_HH_MASS] + rp[GLC_P_HH_MASS];
        r[GLC_P_HH_ANGMOM] = r[GLC_P_HH_ANGMOM] + rp[GLC_P_HH_ANGMOM];
        r[GLC_P_HH_UNACCRETED_ABUND] = r[GLC_P_HH_UNACCRETED_ABUND] + rp[GLC_P_HH_UNACCRETED_ABUND];
        r[GLC_P_HH_ABUND] = r[GLC_P_HH_ABUND] + rp[GLC_P_HH_ABUND];
    }
    memcpy(rp, r, sizeof(double) * GLC_NPROP);
    c->flags[p] = c->flags[i];
    rp[GLC_P_BASIC_MASS] = c->mass[p];
    rp[GLC_P_DMSCALE] = c->scale[p];
    rp[GLC_P_SPIN] = c->angmom[p];
    rp[GLC_P_SAT_BOUND_MASS] = c->mass[p];
    set_targets(c, p, rp);
    sats_move(c, i, p);
    c->state[i] = T_PROMOTED;
    c->state[p] = T_ISOLATED;
    c->children_left[p]--;
    fc->promotions++;
}

static int evolve_to(tree_ctx *c, int i, double to, glc_forest_counters *fc, glc_counters *C) {
    int status[1], interrupt[1], fl = c->flags[i];
    glc_counters local;
    memset(&local, 0, sizeof(local));
    orc_evolve_batch(c->P, c->T, 1, R(c, i), &fl, &to, status, interrupt, &local, 1);
    c->flags[i] = fl;
    C->steps_accepted += local.steps_accepted;
    C->steps_rejected += local.steps_rejected;
    C->rhs_evaluations += local.rhs_evaluations;
    C->segments += local.segments;
    C->trials_failed += local.trials_failed;
    C->nodes += local.nodes;
    fc->evolve_calls++;
    if (local.rhs_evaluations > 20000 && getenv("ORC_DEBUG_LONG_CALLS"))
        fprintf(stderr, "[orc_tree] node %d: %llu RHS evaluations in one evolve call (to t=%.6g, flags %d, M=%.4g)\n", i,
                (unsigned long long)local.rhs_evaluations, to, c->flags[i], R(c, i)[GLC_P_BASIC_MASS]);
    if (!(status[0] == GLC_STATUS_SUCCESS && interrupt[0] == GLC_INT_NONE)) {
        /* the reference aborts here (standard.F90:697-722); the checker and the product move the node on and count it */
        fc->failed_evolves++;
        R(c, i)[GLC_P_TIME] = to;
    }
    return 0;
}

/* visit node i of the walk: its satellites first, then the node itself; returns 1 if anything moved, <0 on error */
static int visit(tree_ctx *c, int i, glc_forest_counters *fc, glc_counters *C) {
    int moved = 0, s;
    if (c->state[i] == T_PROMOTED || c->state[i] == T_SATELLITE) return 0;
    /* hosted satellites (:916-982) */
    for (s = c->first_sat[i]; s >= 0; s = c->next_sat[s]) {
        const double tn = R(c, s)[GLC_P_TIME];
        double to = fmin(c->time_end[s], tn + timestep(c, tn)), th, limit;
        if (to == tn) continue;
        th = c->parent[i] >= 0 ? node_time(c, i) : fmax(node_time(c, i), tn);
        limit = c->children_left[i] > 0 ? fmax(th, tn) : fmax(th + timestep_host(c, th), tn);
        to = fmin(to, limit);
        if (to > tn) {
            if (evolve_to(c, s, to, fc, C)) return -10;
            moved = 1;
        }
    }
    if (c->state[i] != T_ISOLATED || c->parent[i] < 0) return moved;
    {
        /* the node itself (:905-914, :1003-1030) */
        double *r = R(c, i);
        const double tn = r[GLC_P_TIME];
        double to = fmin(c->time_end[i], tn + timestep(c, tn));
        to = fmin(to, c->time[c->parent[i]]);
        for (s = c->first_sat[i]; s >= 0; s = c->next_sat[s]) {
            const double ts = R(c, s)[GLC_P_TIME];
            if (ts < to) to = fmax(ts, tn);
        }
        if (to > tn) {
            double sub = 0.0;
            for (s = c->first_sat[i]; s >= 0; s = c->next_sat[s]) sub += baryons(c, s);
            r[GLC_P_MASS_BARYONIC_SUBHALOS] = sub;
            if (evolve_to(c, i, to, fc, C)) return -10;
            moved = 1;
        }
        /* arrival at the parent (evolver/standard.F90:478-540): merge, or promote once all siblings have merged */
        if (r[GLC_P_TIME] == c->time[c->parent[i]]) {
            if (!is_primary(c, i)) {
                node_merge(c, i, fc);
                moved = 1;
            } else if (c->children_left[c->parent[i]] == 1) {
                node_promote(c, i, fc);
                moved = 1;
            }
        }
    }
    return moved;
}

int orc_forest_evolve(const glc_params *P, const orc_tables *T, long n, const int *parent, const double *mass, const double *time,
                      const double *scale, const double *angmom, double *rec, int *flags, int *state,
                      glc_forest_counters *fc_out, glc_counters *C_out, int n_threads) {
    tree_ctx c;
    long i;
    int *order, *root_of, *post, n_roots = 0, rc = 0;
    long *root_first; /* post-order segment of each root */
    int *roots;
    glc_forest_counters fc;
    glc_counters C;
    const double fb = P->OmegaBaryon / P->OmegaMatter;
    memset(&fc, 0, sizeof(fc));
    memset(&C, 0, sizeof(C));
    c.P = P; c.T = T; c.n = n; c.parent = parent; c.mass = mass; c.time = time; c.scale = scale; c.angmom = angmom;
    c.rec = rec; c.flags = flags; c.state = state;
    c.first_child = malloc(sizeof(int) * n); c.sibling = malloc(sizeof(int) * n); c.host = malloc(sizeof(int) * n);
    c.children_left = calloc(n, sizeof(int)); c.first_sat = malloc(sizeof(int) * n); c.next_sat = malloc(sizeof(int) * n);
    c.time_end = malloc(sizeof(double) * n);
    c.merged_unaccreted = calloc(n, sizeof(double)); c.merged_unaccreted_abund = calloc(n, sizeof(double));
    c.merged_mass = calloc(n, sizeof(double)); c.merged_hot = calloc(n, 1);
    order = malloc(sizeof(int) * n); root_of = malloc(sizeof(int) * n); post = malloc(sizeof(int) * n);
    for (i = 0; i < n; i++) c.first_child[i] = c.sibling[i] = c.host[i] = c.first_sat[i] = c.next_sat[i] = -1;
    /* progenitors ordered by descending mass: insertion sort of each node into its parent's list (ties: lower index first) */
    for (i = 0; i < n; i++) {
        const int p = parent[i];
        int *link;
        if (p < 0) { n_roots++; continue; }
        link = &c.first_child[p];
        while (*link >= 0 && (mass[*link] > mass[i] || (mass[*link] == mass[i] && *link < i))) link = &c.sibling[*link];
        c.sibling[i] = *link;
        *link = (int)i;
        c.children_left[p]++;
    }
    roots = malloc(sizeof(int) * (n_roots + 1));
    root_first = malloc(sizeof(long) * (n_roots + 2));
    n_roots = 0;
    for (i = 0; i < n; i++) if (parent[i] < 0) roots[n_roots++] = (int)i;
    /* depth-first post-order of every tree (children before their parent), iterative */
    {
        long np = 0;
        int t, *stack = malloc(sizeof(int) * n), *cursor = malloc(sizeof(int) * n);
        for (t = 0; t < n_roots; t++) {
            int sp = 0;
            root_first[t] = np;
            stack[sp] = roots[t]; cursor[sp] = c.first_child[roots[t]]; sp++;
            while (sp > 0) {
                const int k = cursor[sp - 1];
                if (k >= 0) {
                    cursor[sp - 1] = c.sibling[k];
                    stack[sp] = k; cursor[sp] = c.first_child[k]; sp++;
                } else {
                    post[np] = stack[sp - 1];
                    root_of[stack[sp - 1]] = roots[t];
                    np++; sp--;
                }
            }
        }
        root_first[n_roots] = np;
        free(stack); free(cursor);
    }
    memset(rec, 0, sizeof(double) * (size_t)n * GLC_NPROP);
    for (i = 0; i < n; i++) {
        double *r = R(&c, i), failed, m_hot, m_failed;
        c.time_end[i] = time[root_of[i]];
        flags[i] = 0;
        state[i] = c.children_left[i] > 0 ? T_PENDING : T_ISOLATED;
        if (state[i] != T_ISOLATED) continue;
        r[GLC_P_TIME] = time[i];
        r[GLC_P_TIME_STEP] = -1.0;
        r[GLC_P_BASIC_MASS] = mass[i];
        r[GLC_P_DMSCALE] = scale[i];
        r[GLC_P_SPIN] = angmom[i];
        r[GLC_P_SAT_BOUND_MASS] = mass[i];
        set_targets(&c, (int)i, r);
        /* cgmAccretionNodeInitialize :144-200 */
        failed = failed_fraction(&c, mass[i], time[i]);
        m_hot = fb * mass[i] * (1.0 - failed);
        m_failed = fb * mass[i] * failed;
        if (m_hot > 0.0 || m_failed > 0.0) {
            flags[i] |= GLC_F_HAS_HOTHALO;
            r[GLC_P_HH_MASS] = m_hot;
            r[GLC_P_HH_UNACCRETED_MASS] = m_failed;
            r[GLC_P_HH_ANGMOM] = angmom[i] * m_hot / mass[i];
        }
    }
    fc.trees = (uint64_t)n_roots;
    fc.nodes = (uint64_t)n;
    {
        int t;
#ifdef _OPENMP
        if (n_threads < 1) n_threads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
#endif
        for (t = 0; t < n_roots; t++) {
            glc_forest_counters lfc;
            glc_counters lC;
            int moved = 1, err = 0;
            long k;
            memset(&lfc, 0, sizeof(lfc));
            memset(&lC, 0, sizeof(lC));
            while (moved && !err) { /* repeated walks of the tree until nothing can move (:398-577) */
                moved = 0;
                for (k = root_first[t]; k < root_first[t + 1]; k++) {
                    const int m = visit(&c, post[k], &lfc, &lC);
                    if (m < 0) { err = 1; break; }
                    moved |= m;
                }
                lfc.rounds++;
            }
#ifdef _OPENMP
#pragma omp critical
#endif
            {
                if (err) rc = -10;
                fc.evolve_calls += lfc.evolve_calls; fc.promotions += lfc.promotions; fc.node_mergers += lfc.node_mergers;
                fc.failed_evolves += lfc.failed_evolves;
                if (lfc.rounds > fc.rounds) fc.rounds = lfc.rounds;
                C.steps_accepted += lC.steps_accepted; C.steps_rejected += lC.steps_rejected;
                C.rhs_evaluations += lC.rhs_evaluations; C.segments += lC.segments; C.trials_failed += lC.trials_failed;
                C.nodes += lC.nodes;
            }
        }
    }
    (void)n_threads;
    if (fc_out) *fc_out = fc;
    if (C_out) *C_out = C;
    free(c.first_child); free(c.sibling); free(c.host); free(c.children_left); free(c.first_sat); free(c.next_sat);
    free(c.merged_unaccreted); free(c.merged_unaccreted_abund); free(c.merged_mass); free(c.merged_hot);
    free(c.time_end); free(order); free(root_of); free(post); free(roots); free(root_first);
    return rc;
}
