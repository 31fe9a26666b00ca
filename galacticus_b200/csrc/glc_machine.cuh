// glc_machine.cuh -- the micro-task machine: the evolve path for the standard (quickTest) model.
//
// Why.  One evaluation of the model's rate function contains nested adaptive solvers (galactic-structure
// fixed point -> adiabatic-contraction Brent root find; star-formation-rate root finds and QAG; cooling-radius
// root find) whose trip counts are heavy-tailed.  Run as ordinary nested loops with one node per lane, a warp
// pays the MAXIMUM trip count of its 32 lanes in every loop: measured 4.7 active lanes per warp instruction
// (ncu, profiles/), i.e. 15 % SIMT efficiency, whatever reconvergence discipline is used.
//
// What.  The rate function is therefore cut into small UNITS (one Brent step + one function evaluation, one
// 15-point Gauss-Kronrod pass, one structure-solver visit, the straight-line set-up pieces, the RK bookkeeping),
// and every node slot carries an explicit continuation (SlotState, resident in HBM/L2).  One iteration of
// the kernel's main loop executes ONE unit per lane.  Before each iteration the block re-groups its slots by
// pending unit with a shared-memory counting sort, so that the lanes of a warp execute the same unit: lanes
// are bound to slots only for the duration of one unit.  A slot that needs 100 structure iterations simply
// takes part in more iterations while its neighbours move on to their next Runge-Kutta stage; nobody waits.
// Modelled SIMT efficiency with 256-slot blocks: 0.80 (tests/emu statistics).
//
// The arithmetic of every unit is the shared helper code of glc_model_standard.cuh / glc_numerics.cuh /
// glc_evolve_kernel.cuh, so the result of a node does not depend on how its units are scheduled; the host
// emulation (tests/emu) runs the same machine and is bit-identical to the CPU checker.
#pragma once

#include "glc_common.cuh"
#include "glc_evolve_kernel.cuh"
#include "glc_model_standard.cuh"

namespace glc {

// Sort key = unit id.  U_IDLE must be the largest so that idle slots collect in the last warps of the block.
enum Unit : int {
    U_RK = 0,       // rate accumulation of the finished evaluation + RK bookkeeping up to the next evaluation
    U_RHS_BEGIN,    // halo scales, hot-halo profile, plausibility, NFW normalisation
    U_STRUCT,       // structure solver: one (iteration, component) visit up to its root find
    U_ROOT_AC,      // Brent step + adiabatic-contraction function
    U_STRUCT_FIN,   // structure solver: digest the root, fixed-point update
    U_SFR_BEGIN,    // Krumholz-McKee-Tumlinson set-up
    U_ROOT_TRUNC,   // Brent step + surface-density truncation function
    U_SFR_MID,      // after the truncation radius: second root find or the integration intervals
    U_ROOT_CRIT,    // Brent step + critical-surface-density function
    U_SFR_MID2,     // after the critical radius: the integration intervals
    U_QAG,          // one 15-point Gauss-Kronrod pass of the star-formation-rate integral
    U_COOL_BEGIN,   // CIE table look-ups + cooling-radius shortcuts
    U_ROOT_COOL,    // Brent step + cooling-time function
    U_IDLE,
    U_COUNT
};

static_assert(U_RHS_BEGIN == kUnitRhsBegin, "kUnitRhsBegin (glc_common.cuh) must equal U_RHS_BEGIN");

typedef ModelStandard MS;

#ifndef GLC_ROOT_STEPS
#define GLC_ROOT_STEPS 2  // Brent steps per execution of a root-find unit (measured: 1 -> 2 = -2...-4 % on the 10^6-node pass, 4 = +1 %)
#endif

// The x^omega table of the adiabatic-contraction function (fastExponentiator, 9 991 doubles = 80 KB) is looked up once per
// Brent step of the most frequent unit.  Through the read-only global path those look-ups were 11 % of the long-scoreboard
// samples of a bulk slice (profiles/r02i: the continuation records stream through L1 and evict the table, so most look-ups
// go to L2).  With -DGLC_STAGED_POW machine_kernel stages the table in shared memory behind its unit queues (104 KB + 80 KB
// of 227 KB).  Measured (profiles/r02j_machine_variants.txt): no gain -- the carve-out takes the same 80 KB away from L1, and
// with two warps per scheduler a warp that no longer waits here waits at its next dependent instruction -- so it is OFF.
#if defined(__CUDACC__)
extern __shared__ unsigned int s_qdyn[];  // machine_kernel's dynamic shared memory: [U_IDLE][SLOTS] ring cells, then the table
#endif
#if defined(__CUDACC__) && defined(GLC_STAGED_POW)
#define GLC_MACHINE_STAGED_POW 1
GLC_DEVICE_INLINE const double *machine_pow_ac() {
    return reinterpret_cast<const double *>(s_qdyn + (size_t)U_IDLE * GLC_MSLOTS);  // (13 x 2048 x 4 bytes: 8-byte aligned)
}
#else
#define GLC_MACHINE_STAGED_POW 0
GLC_DEVICE_INLINE const double *machine_pow_ac() { return GLC_TABLES.powAc; }
#endif

struct RhsState {
    Work w;
    double hist[4], fit;
    double j, radius;
    MS::AcProblem ac;
    MS::SfrProblem sfr;
    double lo[2], hi[2], total, psiDisk, rinfall, logSlopeT;
    int count, comp, active, bad, structureOnly, go, dOn, coolOn, radiusOn, two, nIv, iv;
};

// Everything a root-find unit touches: the Brent state and the parameters of the function being solved.
//   AC    p = {dmoNorm, dmoScale, rvir, fi, fd, radius, bterm}
//   TRUNC/CRIT p = {sigma0, rdisk, sigmaTrunc, xh}
//   COOL  p = {coolXH, coolFHn, coolEfrac, coolLambda, tvir, coolTavail, hhRho0, hhRcore, hhRouter, hhValid}
struct RootState {
    BrentState B;
    double p[10];
    double pad;  // 256 bytes = two 128-byte lines
};
static_assert(sizeof(RootState) == 256, "RootState must stay two cache lines");

// The per-slot continuation, split by access group into separate HBM arrays so that the most frequent units
// (root-find steps: ~2/3 of all unit executions) touch one compact 256-byte record that stays L2-resident.
struct SlotArrays {
    LaneState *L;
    RhsState *R;
    RootState *root;
    double *yt;  // [nslots][NY]
    QagState *Q;
    int *unit;
};
struct SlotRef {
    LaneState &L;
    RhsState &R;
    BrentState &B;
    double (&p)[10];
    double (&yt)[NY];
    QagState &Q;
    int &unit;
};
GLC_DEVICE_INLINE SlotRef slot_ref(const SlotArrays &a, int64_t s) {
    return SlotRef{a.L[s], a.R[s], a.root[s].B, a.root[s].p, *reinterpret_cast<double (*)[NY]>(a.yt + s * NY), a.Q[s], a.unit[s]};
}

GLC_DEVICE_INLINE void slot_reset(const SlotRef &S) {
    lane_reset(S.L);
    S.unit = U_RK;
}

// ---------------------------------------------------------------- cheap transitions (a few instructions)
GLC_DEVICE_INLINE void m_cool_decide(const SlotRef &S) {
    RhsState &R = S.R;
    R.coolOn = MS::cooling_on(S.L.ctx, S.yt, R.w, R.go != 0) ? 1 : 0;
    R.radiusOn = MS::cooling_radius_on(S.L.ctx, S.yt, R.w, R.coolOn != 0, R.go != 0) ? 1 : 0;
    R.rinfall = 0.0;
    R.logSlopeT = 0.0;
    S.unit = R.radiusOn ? U_COOL_BEGIN : U_RK;
}

GLC_DEVICE_INLINE void m_after_struct(const SlotRef &S) {
    RhsState &R = S.R;
    R.psiDisk = 0.0;
    R.rinfall = 0.0;
    R.logSlopeT = 0.0;
    R.dOn = R.coolOn = R.radiusOn = 0;
    if (R.structureOnly || !R.w.solvable) {
        R.go = 0;
        S.unit = U_RK;
        return;
    }
    R.go = 1;
    R.dOn = MS::disk_sfr_on(S.L.ctx, S.yt, true) ? 1 : 0;
    if (R.dOn)
        S.unit = U_SFR_BEGIN;
    else
        m_cool_decide(S);
}

// loop control of galacticStructureSolverEquilibrium::solve (equilibrium.F90:278-292): next component to
// visit, next iteration, or convergence
GLC_DEVICE_INLINE void m_struct_next(const SlotRef &S) {
    RhsState &R = S.R;
    const double tolerance = GLC_PARAMS.structureSolutionTolerance;
    for (;;) {
        while (R.comp < 2 && !MS::has(S.L.ctx, R.comp == 0 ? GLC_F_HAS_DISK : GLC_F_HAS_SPHEROID)) R.comp++;
        if (R.comp < 2) {
            S.unit = U_STRUCT;
            return;
        }
        if (R.active == 0) {
            R.fit = 0.0;
            break;
        }
        R.fit /= (double)R.active;
        if (!(R.count <= 1 || (R.fit > tolerance && R.count < 100))) break;
        R.active = 0;
        R.count++;
        GLC_COUNT(1);
        if (R.count > 1) R.fit = 0.0;
        R.comp = 0;
    }
    m_after_struct(S);
}

template <int UNIT>
GLC_DEVICE_INLINE RootOptions m_root_options() {
    return UNIT == U_ROOT_AC     ? MS::ac_root_options()
           : UNIT == U_ROOT_COOL ? MS::cooling_root_options()
                                 : MS::sfr_root_options();
}
// unit that digests the finished root find of type UNIT
template <int UNIT>
GLC_DEVICE_INLINE int m_root_done_unit() {
    return (UNIT == U_ROOT_AC) ? U_STRUCT_FIN
           : UNIT == U_ROOT_TRUNC                  ? U_SFR_MID
           : UNIT == U_ROOT_CRIT                   ? U_SFR_MID2
                                                   : U_RK;
}
template <int UNIT>
GLC_DEVICE_INLINE void m_root_finish(const SlotRef &S) {
    if (UNIT == U_ROOT_COOL) {
        S.R.rinfall = S.B.result;
        if (S.B.status != 0) S.R.bad = 1;
    }
    S.unit = m_root_done_unit<UNIT>();
}
// initialise a root find and run the state machine to its first abscissa
template <int UNIT>
GLC_DEVICE_INLINE void m_root_start(const SlotRef &S, double xLow, double xHigh, bool haveValues, double fLow,
                                    double fHigh) {
    BrentState B;
    brent_begin(B, true, xLow, xHigh, haveValues, fLow, fHigh);
    brent_advance(B, m_root_options<UNIT>());
    S.B = B;
    if (B.busy)
        S.unit = UNIT;
    else
        m_root_finish<UNIT>(S);
}

GLC_DEVICE_INLINE void m_qag_start(const SlotRef &S) {
    RhsState &R = S.R;
    qag_begin(S.Q, true, R.lo[R.iv], R.hi[R.iv], 1.0e-12, GLC_PARAMS.sfrIntegrationTolerance);
    S.unit = U_QAG;
}

GLC_DEVICE_INLINE void m_sfr_intervals(const SlotRef &S, double rCrit) {
    RhsState &R = S.R;
    R.lo[0] = 0.0;
    R.lo[1] = rCrit;
    R.hi[0] = R.two ? rCrit : R.sfr.rMax;
    R.hi[1] = R.sfr.rMax;
    R.nIv = R.two ? 2 : 1;
    R.total = 0.0;
    R.iv = 0;
    m_qag_start(S);
}

GLC_DEVICE_INLINE void m_sfr_root_params(const SlotRef &S) {
    const MS::Kmt &k = S.R.sfr.k;
    S.p[0] = k.sigma0;
    S.p[1] = k.rdisk;
    S.p[2] = k.sigmaTrunc;
    S.p[3] = k.xh;
}

GLC_DEVICE_INLINE void m_sfr_after_trunc(const SlotRef &S, double rTrunc, int st) {
    RhsState &R = S.R;
    R.two = MS::sfr_after_trunc(R.sfr, rTrunc, st, R.bad) ? 1 : 0;
    if (R.two) {
        m_sfr_root_params(S);
        m_root_start<U_ROOT_CRIT>(S, 0.0, R.sfr.rMax, false, 0.0, 0.0);
    } else
        m_sfr_intervals(S, 0.0);
}

// ---------------------------------------------------------------- the units
// Every unit returns the slot's NEXT unit.  Inside a unit the pending-unit word is a local (GLC_UNIT_ENTER rebinds the slot
// view to it), so the transitions above never touch the word in HBM: the kernel read it back right after the unit had stored
// it, and that load-after-store was the largest single memory stall of a bulk slice (8 % of the long-scoreboard samples,
// profiles/r02i).  The caller stores the returned value.
#define GLC_UNIT_ENTER(S0, SELF) \
    int unit__ = (SELF);         \
    const SlotRef S { (S0).L, (S0).R, (S0).B, (S0).p, (S0).yt, (S0).Q, unit__ }
// U_RK: rates_accumulate of the evaluation whose nested solvers have just finished, lane_consume (store the
// stage derivative; at the end of an attempt: controller, accept/reject, post-step), lane_prepare (next stage
// input, or epilogue/fetch/prologue).
GLC_DEVICE_NOINLINE int unit_rk(const SlotRef S0, const LaneMem M) {
    GLC_UNIT_ENTER(S0, U_RK);
    LaneState L = S.L;
    double yt[NY], rate[NY];
    int code = GLC_INT_NONE;
#pragma unroll 1
    for (int i = 0; i < NY; i++) {
        yt[i] = S.yt[i];
        rate[i] = 0.0;
    }
    if (L.heavy == HV_RHS || L.heavy == HV_POST_EVOLVE) {
        const RhsState &R = S.R;
        code = MS::rates_accumulate(L.ctx, L.ts, yt, rate, R.w, R.go != 0, R.dOn != 0, R.psiDisk, R.coolOn != 0,
                                    R.radiusOn != 0, R.rinfall, R.logSlopeT, R.bad);
    }
    for (;;) {
        lane_consume<MS>(L, M, yt, rate, code);
        lane_prepare<MS>(L, M, yt);
        if (L.heavy != HV_FROZEN) break;
        code = GLC_INT_NONE;
#pragma unroll 1
        for (int i = 0; i < NY; i++) rate[i] = 0.0;
    }
    S.L = L;
    if (L.heavy == HV_NONE) {
        S.unit = U_IDLE;
        return unit__;
    }
#pragma unroll 1
    for (int i = 0; i < NY; i++) S.yt[i] = yt[i];
    S.unit = U_RHS_BEGIN;
    return unit__;
}

GLC_DEVICE_NOINLINE int unit_rhs_begin(const SlotRef S0) {
    GLC_UNIT_ENTER(S0, U_RHS_BEGIN);
    RhsState &R = S.R;
    NodeCtx &c = S.L.ctx;
    Work w;
    GLC_COUNT(6);
    MS::work_clear(w);
    MS::halo_scales(c, S.L.ts, w);
    MS::hh_profile(c, S.yt, w);
    MS::plausibility(c, S.yt, S.L.ts, w);
    R.structureOnly = (S.L.heavy == HV_POST_EVOLVE) ? 1 : 0;
    R.bad = 0;
    R.hist[0] = R.hist[1] = R.hist[2] = R.hist[3] = -1.0;
    R.fit = 2.0 * GLC_PARAMS.structureSolutionTolerance;
    if (w.plausible) MS::dmo_prepare(c, w);
    R.w = w;
    if (!w.plausible) {
        m_after_struct(S);
        return unit__;
    }
    R.count = 1;
    GLC_COUNT(1);
    R.active = 0;
    R.comp = 0;
    m_struct_next(S);
    return unit__;
}

// one (iteration, component) visit of the structure solver, up to the point where a root is needed
GLC_DEVICE_NOINLINE int unit_struct(const SlotRef S0) {
    GLC_UNIT_ENTER(S0, U_STRUCT);
    RhsState &R = S.R;
    NodeCtx &c = S.L.ctx;
    const int comp = R.comp;
    const double j = MS::component_j(S.yt, comp);
    double radius = 0.0, velocity = 0.0;
    R.active++;
    R.j = j;
    if (R.count == 1) {
        MS::structure_first_pass(c, R.w, comp, j, radius, velocity);
        MS::structure_store(c, comp, radius, velocity);
        R.comp++;
        m_struct_next(S);
        return unit__;
    }
    if (j <= 0.0) {
        R.comp++;
        m_struct_next(S);
        return unit__;
    }
    radius = comp == 0 ? c.diskRadius : c.sphRadius;
    R.radius = radius;
    MS::AcProblem P;
    P.fd = P.fi = P.bterm = 0.0;
    P.rup = P.rInit = radius;
    P.need = 0;
    if (GLC_PARAMS.adiabaticContraction && !(radius <= 0.0))
        MS::ac_setup<GLC_MACHINE_STAGED_POW != 0>(c, S.yt, R.w, radius, P, machine_pow_ac());
    R.ac = P;
    if (P.need) {
        S.p[0] = R.w.dmoNorm;
        S.p[1] = R.w.dmoScale;
        S.p[2] = R.w.rvir;
        S.p[3] = P.fi;
        S.p[4] = P.fd;
        S.p[5] = radius;
        S.p[6] = P.bterm;
        m_root_start<U_ROOT_AC>(S, radius, P.rup, false, 0.0, 0.0);
    } else {
        S.B.busy = 0;
        S.B.status = 0;
        S.unit = U_STRUCT_FIN;
    }
    return unit__;
}

// digest the root of a structure visit: the contracted dark-matter mass and the fixed-point update
GLC_DEVICE_NOINLINE int unit_struct_fin(const SlotRef S0) {
    GLC_UNIT_ENTER(S0, U_STRUCT_FIN);
    RhsState &R = S.R;
    NodeCtx &c = S.L.ctx;
    const int comp = R.comp;
    double radius = R.radius, velocity;
    {
        const double fDm = 1.0 - GLC_PARAMS.OmegaBaryon / GLC_PARAMS.OmegaMatter;
        double mdm;
        if (!GLC_PARAMS.adiabaticContraction)
            mdm = MS::dmo_mass(R.w.dmoNorm, R.w.dmoScale, radius);
        else if (radius <= 0.0)
            mdm = 0.0;
        else {
            double rInit = R.ac.rInit;
            if (R.ac.need) {
                rInit = S.B.result;
                if (S.B.status != 0) R.bad = 1;
            }
            mdm = fDm * MS::dmo_mass(R.w.dmoNorm, R.w.dmoScale, rInit);
        }
        MS::structure_update(c, S.yt, R.w, R.j, mdm, R.count, R.hist[2 * comp], R.hist[2 * comp + 1], R.fit, R.bad, radius,
                             velocity);
    }
    MS::structure_store(c, comp, radius, velocity);
    R.comp++;
    m_struct_next(S);
    return unit__;
}

// One Brent step: evaluate the function at the pending abscissa, digest it, advance to the next abscissa.
// Touches only the slot's RootState record.
template <int UNIT>
GLC_DEVICE_NOINLINE int unit_root(const SlotRef S0) {
    GLC_UNIT_ENTER(S0, UNIT);
    BrentState B = S.B;
    const RootOptions o = m_root_options<UNIT>();
#pragma unroll 1
    for (int step = 0;; step++) {
    const double x = B.x;
    double fx;
    if (UNIT == U_ROOT_AC) {
        GLC_COUNT(0);
        Work w;
        w.rvir = S.p[2];
        MS::AcProblem P;
        P.fi = S.p[3];
        P.fd = S.p[4];
        P.bterm = S.p[6];
        fx = MS::ac_function<GLC_MACHINE_STAGED_POW != 0>(S.p[0], S.p[1], w, P, S.p[5], x, machine_pow_ac());
    } else if (UNIT == U_ROOT_TRUNC || UNIT == U_ROOT_CRIT) {
        MS::Kmt k;
        k.sigma0 = S.p[0];
        k.rdisk = S.p[1];
        k.sigmaTrunc = S.p[2];
        k.xh = S.p[3];
        GLC_COUNT(4);
        fx = (UNIT == U_ROOT_TRUNC) ? MS::sfr_trunc_function(k, x) : MS::sfr_crit_function(k, x);
    } else {
        GLC_COUNT(3);
        Work w;
        w.coolXH = S.p[0];
        w.coolFHn = S.p[1];
        w.coolEfrac = S.p[2];
        w.coolLambda = S.p[3];
        w.tvir = S.p[4];
        w.coolTavail = S.p[5];
        w.hhRho0 = S.p[6];
        w.hhRcore = S.p[7];
        w.hhRouter = S.p[8];
        w.hhValid = S.p[9] != 0.0;
        fx = MS::cooling_function(w, x);
    }
    brent_feed(B, o, fx);
    brent_advance(B, o);
    // GLC_ROOT_STEPS Brent steps per unit execution while the lane's root find is still busy: the continuation stays in
    // registers between them (a unit execution costs a queue round trip plus the load and store of the RootState record)
    if (!B.busy || step + 1 >= GLC_ROOT_STEPS) break;
    }
    S.B = B;
    if (!B.busy) m_root_finish<UNIT>(S);
    return unit__;
}

GLC_DEVICE_NOINLINE int unit_sfr_begin(const SlotRef S0) {
    GLC_UNIT_ENTER(S0, U_SFR_BEGIN);
    RhsState &R = S.R;
    MS::sfr_setup(S.L.ctx, S.yt, true, R.sfr);
    if (!R.sfr.live) {
        R.psiDisk = 0.0;
        m_cool_decide(S);
        return unit__;
    }
    if (R.sfr.needRmax) {
        m_sfr_root_params(S);
        m_root_start<U_ROOT_TRUNC>(S, 0.0, R.sfr.rOut, false, 0.0, 0.0);
    } else
        m_sfr_after_trunc(S, 0.0, 0);
    return unit__;
}
GLC_DEVICE_NOINLINE int unit_sfr_mid(const SlotRef S0) {
    GLC_UNIT_ENTER(S0, U_SFR_MID);
    m_sfr_after_trunc(S, S.B.result, S.B.status);
    return unit__;
}
GLC_DEVICE_NOINLINE int unit_sfr_mid2(const SlotRef S0) {
    GLC_UNIT_ENTER(S0, U_SFR_MID2);
    if (S.B.status != 0) S.R.bad = 1;
    m_sfr_intervals(S, S.B.result);
    return unit__;
}

GLC_DEVICE_NOINLINE int unit_qag(const SlotRef S0) {
    GLC_UNIT_ENTER(S0, U_QAG);
    RhsState &R = S.R;
    const MS::Kmt k = R.sfr.k;
    GLC_COUNT(7);
    qag_pass(S.Q, [&](double r) {
        GLC_COUNT(2);
        return MS::sfr_integrand(k, r);
    });
    if (S.Q.busy) return unit__;
    const double v = qag_finish(S.Q);
    R.total += v;
    if (S.Q.status == 11) R.bad = 1;
    R.iv++;
    if (R.iv < R.nIv) {
        m_qag_start(S);
        return unit__;
    }
    R.psiDisk = 2.0 * kPi * R.total;
    m_cool_decide(S);
    return unit__;
}

GLC_DEVICE_NOINLINE int unit_cool_begin(const SlotRef S0) {
    GLC_UNIT_ENTER(S0, U_COOL_BEGIN);
    RhsState &R = S.R;
    Work w = R.w;
    double logSlopeT = 0.0, rootOuter, rootZero, result;
    bool need;
    MS::cooling_prepare(S.yt, w, logSlopeT);
    MS::cooling_setup(w, rootOuter, rootZero, result, need);
    R.w = w;
    R.logSlopeT = logSlopeT;
    if (need) {
        S.p[0] = w.coolXH;
        S.p[1] = w.coolFHn;
        S.p[2] = w.coolEfrac;
        S.p[3] = w.coolLambda;
        S.p[4] = w.tvir;
        S.p[5] = w.coolTavail;
        S.p[6] = w.hhRho0;
        S.p[7] = w.hhRcore;
        S.p[8] = w.hhRouter;
        S.p[9] = w.hhValid ? 1.0 : 0.0;
        m_root_start<U_ROOT_COOL>(S, 0.0, w.hhRouter, true, rootZero, rootOuter);
    } else {
        R.rinfall = result;
        S.unit = U_RK;
    }
    return unit__;
}

// One unit (`u`) of one slot.  Returns the slot's next unit, -1 when `u` is not an executable unit (idle slot).
GLC_DEVICE_INLINE int machine_dispatch(const SlotRef &S, const LaneMem &M, int u) {
    switch (u) {
        case U_RK: return unit_rk(S, M);
        case U_RHS_BEGIN: return unit_rhs_begin(S);
        case U_STRUCT: return unit_struct(S);
        case U_STRUCT_FIN: return unit_struct_fin(S);
        case U_ROOT_AC: return unit_root<U_ROOT_AC>(S);
        case U_ROOT_TRUNC: return unit_root<U_ROOT_TRUNC>(S);
        case U_ROOT_CRIT: return unit_root<U_ROOT_CRIT>(S);
        case U_ROOT_COOL: return unit_root<U_ROOT_COOL>(S);
        case U_SFR_BEGIN: return unit_sfr_begin(S);
        case U_SFR_MID: return unit_sfr_mid(S);
        case U_SFR_MID2: return unit_sfr_mid2(S);
        case U_QAG: return unit_qag(S);
        case U_COOL_BEGIN: return unit_cool_begin(S);
        default: return -1;
    }
}
// ... of the unit the slot's pending-unit word names (host-driven machine, tests/emu).  Returns false when the slot is idle.
GLC_DEVICE_INLINE bool machine_step(const SlotRef &S, const LaneMem &M) {
    const int nu = machine_dispatch(S, M, S.unit);
    if (nu < 0) return false;
    S.unit = nu;
    return true;
}

// Start of a time slice: the pending unit of a slot, after re-arming what must be re-armed.  Shared by machine_kernel and
// the host-driven machine of tests/emu (the stale-slot defect of round 2 lived in exactly this code).
//   * first slice of a batch / session: every slot is reset;
//   * idle slot: back to the node queue, which may have grown since the last slice;
//   * not a unit (-1): the slot was finished by drain_kernel in an earlier hand-over and is free.  The lane state in memory is
//     the STALE copy the drain kernel loaded (heavy == HV_RHS: unit_rk would digest an evaluation that never ran and evolve
//     the old node a second time), so the slot is reset completely, not just re-armed like an idle one.
GLC_DEVICE_INLINE int machine_rearm(const SlotRef &own, bool resume) {
    if (!resume) slot_reset(own);
    if (own.unit == U_IDLE) {
        own.L.phase = PH_FETCH;
        own.unit = U_RK;
    }
    int u = own.unit;
    if (u < 0 || u > U_IDLE) {
        slot_reset(own);
        u = U_RK;
    }
    return u;
}

#if defined(__CUDACC__)
// Persistent time-sliced kernel.  One block per SM owns SLOTS slots.  Scheduling is barrier-free: the block keeps
// one ring-buffer queue per unit type in shared memory; a warp pops up to 32 slots that all wait for the SAME unit
// (preferring the unit it executed last, whose code is hot in its instruction cache, else the fullest queue),
// executes that unit once per lane, and pushes every slot onto the queue of its next unit.  Lanes are bound to
// slots only for the duration of one unit, warps are pure by construction, and a slow unit (the RK bookkeeping,
// a cold-started structure solve) delays only its own slots.  All per-slot state lives in HBM/L2 (SlotArrays +
// the RK stage vectors); the queues are rebuilt from the slots' pending-unit words at the start of every time
// slice, so parking costs nothing.
//
// Queue protocol (bounded multi-producer / multi-consumer ring, one per unit).  A cell is one 32-bit word
//     (lap << 12) | (full << 11) | slot            lap = (position / SLOTS) mod 2^20
// and goes  empty(lap) -> full(lap) -> empty(lap + 1).  A producer takes a position with atomicAdd on the tail, WAITS until
// the cell is empty for its lap (i.e. until the consumer of the position one lap earlier has read it) and writes it; a
// consumer (lane 0 reserves up to 32 positions with a CAS on the head) waits until its cell is full for its lap, reads
// the slot id and immediately marks the cell empty for the next lap.  Round 1 used 16-bit cells with a 5-bit lap tag and
// producers that did not wait: a consumer that was slow between reserving and reading could be lapped (its cell
// overwritten: the slot was lost), and because the tag repeats every 32 laps it then took another slot's entry (the slot
// ran on two lanes at once).  Seen on the device as nodes that never finish and as corrupted lane states in batches with
// very fast turnover (profiles/r02a_ledger_4000_trees_round1_queues.txt); the GLC_LEDGER build (node-ownership ledger + per-slot busy flags)
// proves the absence of both with this protocol.  All spins are bounded: a time-out sets the block's abort flag and the
// host returns GLC_ERR_STALLED.
constexpr unsigned int kCellFull = 1u << 11, kCellSlotMask = 0x7ffu, kLapMask = 0xfffffu;
constexpr unsigned int kSpinLimit = 1u << 24;  // x ~64 ns

template <int THREADS, int SLOTS>
__global__ void __launch_bounds__(THREADS, 1) machine_kernel(KernelArgs A, SlotArrays slots) {
    static_assert((SLOTS & (SLOTS - 1)) == 0 && SLOTS <= 2048, "SLOTS must be a power of two <= 2048 (11-bit ids)");
    constexpr int PER = SLOTS / THREADS;
    static_assert(SLOTS == GLC_MSLOTS, "machine_pow_ac() places the staged table behind U_IDLE x GLC_MSLOTS ring cells");
    // s_qdyn: [U_IDLE][SLOTS] ring cells (dynamic: > 48 KB), then the staged x^omega table
    unsigned int(*s_q)[SLOTS] = reinterpret_cast<unsigned int(*)[SLOTS]>(s_qdyn);
    __shared__ unsigned int s_head[U_IDLE], s_tail[U_IDLE];
    __shared__ int s_idle, s_cur, s_abort;
    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t base = (int64_t)blockIdx.x * SLOTS;
    for (int i = tid; i < U_IDLE * SLOTS; i += THREADS) s_qdyn[i] = 0u;  // empty, lap 0
#if GLC_MACHINE_STAGED_POW
    {
        double *tab = reinterpret_cast<double *>(s_qdyn + (size_t)U_IDLE * SLOTS);
        for (int i = tid; i < GLC_TABLES.powAcN; i += THREADS) tab[i] = GLC_LDG(GLC_TABLES.powAc + i);
    }
#endif
    if (tid < U_IDLE) s_head[tid] = s_tail[tid] = 0u;
    if (tid == 0) {
        s_idle = 0;
        s_cur = U_RK;
        s_abort = 0;
    }
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < PER; k++) {
        const int s = tid + k * THREADS;
        const SlotRef own = slot_ref(slots, base + s);
        const int u = machine_rearm(own, A.resume != 0);
        if (A.hold && u == U_RHS_BEGIN)
            atomicAdd(&s_idle, 1);  // held at an RK boundary for the drain kernel: out of work as far as this block goes
        else
            s_q[u][atomicAdd(&s_tail[u], 1u) & (SLOTS - 1)] = kCellFull | (unsigned int)s;  // lap 0, full
    }
    __syncthreads();

    volatile unsigned int *vhead = s_head, *vtail = s_tail;
    volatile int *vidle = &s_idle, *vcur = &s_cur, *vabort = &s_abort;
#ifdef GLC_DEBUG_HANG
    volatile int *dbg = A.debug ? A.debug + ((int64_t)blockIdx.x * (THREADS / 32) + (tid >> 5)) * 8 : nullptr;
#define GLC_DBG(k, v) do { if (dbg && lane == 0) dbg[k] = (v); } while (0)
#define GLC_DBG_LANE(k, v) do { if (dbg) dbg[k] = (v); } while (0)
#else
#define GLC_DBG(k, v) ((void)0)
#define GLC_DBG_LANE(k, v) ((void)0)
#endif
    int warpLast = -1;
    (void)warpLast;
    unsigned int unitsDone = 0;  // unit executions of this warp in this slice (lane 0's copy is reported)
#pragma unroll 1
    for (int it = 0; it < A.budget; ++it) {
        // ---- lane 0 reserves up to 32 entries of the block's CURRENT unit.  The whole block works on one unit
        // at a time (its code, 10-30 KB, then lives in the SM's instruction cache) and moves on to the fullest
        // queue when the current one cannot fill a warp any more.  No barrier: a warp that still holds a stale
        // choice just pops from that queue once more.
        int u = -1, take = 0;
        unsigned int start = 0;
        GLC_DBG(0, 1);
        GLC_DBG(1, it);
        if (lane == 0) {
#pragma unroll 1
            for (int attempt = 0; attempt < 4 && u < 0; attempt++) {
                int cur = *vcur;
                unsigned int n = vtail[cur] - vhead[cur];
#ifdef GLC_POLICY_WARP
                if (warpLast >= 0 && vtail[warpLast] - vhead[warpLast] >= 32u && vtail[warpLast] - vhead[warpLast] <= (unsigned int)SLOTS) {
                    cur = warpLast;
                    n = 32u;
                }
#endif
                if (n < 32u) {
                    int best = cur;
                    unsigned int bestN = n;
                    for (int q = 0; q < U_IDLE; q++) {
                        const unsigned int m = vtail[q] - vhead[q];
                        if (m > bestN && m <= (unsigned int)SLOTS) {
                            bestN = m;
                            best = q;
                        }
                    }
                    if (best != cur) {
                        *vcur = best;
                        cur = best;
                    }
                }
                const unsigned int old = vhead[cur];
                n = vtail[cur] - old;
                if (n == 0u || n > (unsigned int)SLOTS) continue;
                const unsigned int t = n < 32u ? n : 32u;
                if (atomicCAS(&s_head[cur], old, old + t) == old) {
                    u = cur;
                    start = old;
                    take = (int)t;
                }
            }
        }
        // NB: every decision that steers the warp's control flow is taken by lane 0 and broadcast.  Letting each
        // lane read the volatile idle counter itself dead-locked the block: the lanes of a warp are not guaranteed
        // to execute that load at the same instant, so when the counter reached SLOTS between two of them part of
        // the warp left the loop while the rest waited for it in the next __shfl_sync.
        int allIdle = 0;
        if (lane == 0 && u < 0) allIdle = (*vidle >= SLOTS) ? 1 : 0;
        if (lane == 0 && *vabort) allIdle = 2;  // a queue time-out somewhere in the block: leave
        u = __shfl_sync(0xffffffffu, u, 0);
        start = __shfl_sync(0xffffffffu, start, 0);
        take = __shfl_sync(0xffffffffu, take, 0);
        allIdle = __shfl_sync(0xffffffffu, allIdle, 0);
        GLC_DBG(2, u);
        GLC_DBG(3, (int)start);
        GLC_DBG(4, take);
        if (allIdle == 2) break;
        if (u < 0) {
            GLC_DBG(0, 6);
            GLC_DBG(5, allIdle);
#ifdef GLC_WATCHDOG
            if (lane == 0 && (it % 4000) == 3999)
                printf("[watchdog] block %d warp %d it %d: nothing to pop, idle=%d cur=%d\n", blockIdx.x, tid >> 5, it, *vidle, *vcur);
#endif
            if (allIdle) break;  // every slot of the block is out of work
            __nanosleep(256);
            continue;
        }
        // ---- one unit per lane
        if (lane < take) {
            const unsigned int pos = start + (unsigned int)lane;
            const unsigned int lap = (pos / SLOTS) & kLapMask;
            volatile unsigned int *entry = &s_q[u][pos & (SLOTS - 1)];
            const unsigned int want = (lap << 1) | 1u;  // full, my lap
            unsigned int e, spins = 0;
            GLC_DBG_LANE(0, 2);
            GLC_DBG_LANE(6, lane);
            while (((e = *entry) >> 11) != want) {  // the producer has taken the position and is about to write it
                if (++spins > kSpinLimit) break;
                __nanosleep(32);
            }
            if (spins > kSpinLimit) {
                *vabort = 1;
                atomicAdd(&A.counters[10], 1ull);
            } else {
            *entry = ((lap + 1u) & kLapMask) << 12;  // empty for the next lap: the producer one lap on may write
            GLC_DBG_LANE(0, 3);
            GLC_DBG_LANE(6, lane);
            GLC_DBG_LANE(7, (int)(e & kCellSlotMask));
            const int s = (int)(e & kCellSlotMask);
            __threadfence_block();  // acquire: the producer's stores to the slot's continuation are visible
            const int64_t slot = base + s;
            const SlotRef S = slot_ref(slots, slot);
            LaneMem M{&A, A.ws + slot * (WS_NVEC * NY), 1};
#ifdef GLC_LEDGER
            M.dbgSlot = (int)slot;
            if (A.slotBusy && atomicExch(&A.slotBusy[slot], 1) != 0) atomicAdd(&A.ledgerErr[2], 1ull);  // two lanes in one slot
            if (A.slotBusy && S.unit != u && !(u == U_RK && S.unit == U_IDLE)) atomicAdd(&A.ledgerErr[3], 1ull);  // wrong queue
#endif
            const int nu = machine_dispatch(S, M, u);  // the queue a slot sits in IS its pending unit
            S.unit = nu;
#ifdef GLC_LEDGER
            if (A.slotBusy) atomicExch(&A.slotBusy[slot], 0);
#endif
            __threadfence_block();  // release: continuation stores before the queue entry
            if (nu < 0 || nu > U_IDLE) {  // not a unit: the continuation is corrupt -- never index a queue with it
                *vabort = 1;
                atomicAdd(&A.counters[10], 1ull);
            } else if (nu == U_IDLE || (A.hold && nu == U_RHS_BEGIN))
                atomicAdd(&s_idle, 1);
            else {
                const unsigned int np = atomicAdd(&s_tail[nu], 1u);
                const unsigned int plap = (np / SLOTS) & kLapMask;
                volatile unsigned int *cell = &s_q[nu][np & (SLOTS - 1)];
                unsigned int pspins = 0;
                while (*cell != (plap << 12)) {  // the consumer one lap back has not read its entry yet
                    if (++pspins > kSpinLimit) break;
                    __nanosleep(32);
                }
                if (pspins > kSpinLimit) {
                    *vabort = 1;
                    atomicAdd(&A.counters[10], 1ull);
                }
                *cell = (plap << 12) | kCellFull | (unsigned int)s;
            }
            }
        }
        GLC_DBG_LANE(0, 4);
        __syncwarp();
        GLC_DBG(0, 5);
        warpLast = u;
        unitsDone += (unsigned int)take;
    }
    GLC_DBG(0, 9);
    __syncthreads();
    // ---- counters of this block's slots: warp-reduce then one atomic per warp per counter
    unsigned int vals[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (int k = 0; k < PER; k++) {
        const SlotRef own = slot_ref(slots, base + tid + k * THREADS);
        LaneState &L = own.L;
        vals[0] += L.nAcc;
        vals[1] += L.nRej;
        vals[2] += L.nRhs;
        vals[3] += L.nSeg;
        vals[4] += L.nTrialFail;
        vals[5] += L.nNodes;
        vals[6] += L.nDone;
        vals[7] += own.unit != U_IDLE ? 1u : 0u;
        vals[8] += (own.unit != U_IDLE && own.unit != U_RHS_BEGIN) ? 1u : 0u;
        L.nAcc = L.nRej = L.nRhs = L.nSeg = L.nTrialFail = L.nNodes = L.nDone = 0;
    }
#pragma unroll
    for (int k = 0; k < 9; k++) {
        unsigned int v = vals[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&A.counters[k], (unsigned long long)v);
    }
    if (lane == 0 && unitsDone) atomicAdd(&A.counters[9], (unsigned long long)unitsDone);
}

// compacts the ids of the slots held at an RK boundary into a list for drain_kernel; score = predicted number of
// remaining steps of the node, (t_end - t) / h with the controller's current step size.  Streaming sessions also list up to
// `wantFree` free slots (entries tagged kHeldFresh, counted in count[3]) into which the drain kernel fetches queued nodes.
__global__ void held_list_kernel(const int *__restrict__ unit, const LaneState *__restrict__ L, int nslots, int32_t *held,
                                 float *score, int *count, int wantFree, float ageWeight = 0.0f, const float *priority = nullptr) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nslots; i += gridDim.x * blockDim.x) {
        const int u = unit[i];
        if (u == U_RHS_BEGIN) {
            const int k = atomicAdd(count, 1);
            held[k] = i;
            if (score) {
                const double h = L[i].h, left = L[i].x1 - L[i].x;
                const float steps = (L[i].heavy == HV_RHS && h > 0.0 && left > 0.0) ? (float)fmin(left / h, 1.0e30) : 0.0f;
                // ageWeight > 0 (streaming ticks): evaluations already spent on the node count too -- step sizes collapse late,
                // and a node that has been running for long is the best predictor of a node that will go on running
                score[k] = ageWeight > 0.0f ? 6.0f * steps + ageWeight * (float)L[i].age : steps;
                // ageWeight < 0: the component-set bucket of the node (queue_bucket), for lists sorted by kind of node
                if (ageWeight < 0.0f) score[k] = (float)queue_bucket(L[i].ctx.flags);
                // priority given by the submitter (the tree scheduler knows which nodes its trees wait for): per ticket = arena row
                if (priority) score[k] = (L[i].node >= 0) ? priority[L[i].node] : 0.0f;
            }
        } else if (wantFree > 0 && (u == U_IDLE || u < 0)) {
            if (atomicAdd(count + 3, 1) < wantFree) {
                const int k = atomicAdd(count, 1);
                held[k] = i | kHeldFresh;
                if (score) score[k] = ageWeight < 0.0f ? 64.0f : 0.0f;
            }
        }
    }
}
#endif

}  // namespace glc
