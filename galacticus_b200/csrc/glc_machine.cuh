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
    U_STRUCT_FIN,   // structure solver: digest the root, fixed-point update
    U_ROOT_AC,      // Brent step + adiabatic-contraction function
    U_ROOT_J,       // Brent step + specific-angular-momentum function (first-guess radius)
    U_ROOT_TRUNC,   // Brent step + surface-density truncation function
    U_ROOT_CRIT,    // Brent step + critical-surface-density function
    U_ROOT_COOL,    // Brent step + cooling-time function
    U_SFR_BEGIN,    // Krumholz-McKee-Tumlinson set-up
    U_QAG,          // one 15-point Gauss-Kronrod pass of the star-formation-rate integral
    U_COOL_BEGIN,   // CIE table look-ups + cooling-radius shortcuts
    U_IDLE,
    U_COUNT
};

typedef ModelStandard MS;

struct RhsState {
    Work w;
    double nfwNorm, hist[4], fit;
    double j, radius, lnj;
    MS::AcProblem ac;
    MS::SfrProblem sfr;
    double lo[2], hi[2], total, psiDisk, rinfall, logSlopeT;
    int count, comp, active, bad, structureOnly, go, dOn, coolOn, radiusOn, two, nIv, iv, guess;
};

struct SlotState {
    LaneState L;
    RhsState R;
    BrentState B;
    double yt[NY];
    int unit, pad;
    QagState Q;
};

GLC_DEVICE_INLINE void slot_reset(SlotState &S) {
    lane_reset(S.L);
    S.unit = U_RK;
}

// ---------------------------------------------------------------- cheap transitions (a few instructions)
GLC_DEVICE_INLINE void m_cool_decide(SlotState &S) {
    RhsState &R = S.R;
    R.coolOn = MS::cooling_on(S.L.ctx, S.yt, R.w, R.go != 0) ? 1 : 0;
    R.radiusOn = MS::cooling_radius_on(R.w, R.coolOn != 0) ? 1 : 0;
    R.rinfall = 0.0;
    R.logSlopeT = 0.0;
    S.unit = R.radiusOn ? U_COOL_BEGIN : U_RK;
}

GLC_DEVICE_INLINE void m_after_struct(SlotState &S) {
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
GLC_DEVICE_INLINE void m_struct_next(SlotState &S) {
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

GLC_DEVICE_INLINE void m_root_complete(SlotState &S, int unit);

// first advance of a freshly initialised root find: usually yields the first abscissa
GLC_DEVICE_INLINE void m_root_start(SlotState &S, int unit) {
    brent_advance(S.B);
    if (S.B.busy)
        S.unit = unit;
    else
        m_root_complete(S, unit);
}

GLC_DEVICE_INLINE void m_qag_start(SlotState &S) {
    RhsState &R = S.R;
    qag_begin(S.Q, true, R.lo[R.iv], R.hi[R.iv], 1.0e-12, GLC_PARAMS.sfrIntegrationTolerance);
    S.unit = U_QAG;
}

GLC_DEVICE_INLINE void m_sfr_intervals(SlotState &S, double rCrit) {
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

GLC_DEVICE_INLINE void m_sfr_after_trunc(SlotState &S, double rTrunc, int st) {
    RhsState &R = S.R;
    R.two = MS::sfr_after_trunc(R.sfr, rTrunc, st, R.bad) ? 1 : 0;
    if (R.two) {
        brent_begin(S.B, true, MS::sfr_root_options(), 0.0, R.sfr.rMax, false, 0.0, 0.0);
        m_root_start(S, U_ROOT_CRIT);
    } else
        m_sfr_intervals(S, 0.0);
}

GLC_DEVICE_INLINE void m_root_complete(SlotState &S, int unit) {
    RhsState &R = S.R;
    if (unit == U_ROOT_AC || unit == U_ROOT_J) {
        S.unit = U_STRUCT_FIN;
    } else if (unit == U_ROOT_TRUNC) {
        m_sfr_after_trunc(S, S.B.result, S.B.status);
    } else if (unit == U_ROOT_CRIT) {
        if (S.B.status != 0) R.bad = 1;
        m_sfr_intervals(S, S.B.result);
    } else {  // U_ROOT_COOL
        R.rinfall = S.B.result;
        if (S.B.status != 0) R.bad = 1;
        S.unit = U_RK;
    }
}

// ---------------------------------------------------------------- the units
// U_RK: rates_accumulate of the evaluation whose nested solvers have just finished, lane_consume (store the
// stage derivative; at the end of an attempt: controller, accept/reject, post-step), lane_prepare (next stage
// input, or epilogue/fetch/prologue).
GLC_DEVICE_INLINE void unit_rk(SlotState &S, const LaneMem &M) {
    LaneState L = S.L;
    double yt[NY], rate[NY];
    int code = GLC_INT_NONE;
#pragma unroll
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
#pragma unroll
        for (int i = 0; i < NY; i++) rate[i] = 0.0;
    }
    S.L = L;
    if (L.heavy == HV_NONE) {
        S.unit = U_IDLE;
        return;
    }
#pragma unroll
    for (int i = 0; i < NY; i++) S.yt[i] = yt[i];
    S.unit = U_RHS_BEGIN;
}

GLC_DEVICE_INLINE void unit_rhs_begin(SlotState &S) {
    RhsState &R = S.R;
    NodeCtx &c = S.L.ctx;
    Work w;
    MS::work_clear(w);
    MS::halo_scales(c, S.L.ts, w);
    MS::hh_profile(c, S.yt, w);
    MS::plausibility(c, S.yt, S.L.ts, w);
    R.structureOnly = (S.L.heavy == HV_POST_EVOLVE) ? 1 : 0;
    R.bad = 0;
    R.hist[0] = R.hist[1] = R.hist[2] = R.hist[3] = -1.0;
    R.fit = 2.0 * GLC_PARAMS.structureSolutionTolerance;
    R.nfwNorm = w.plausible ? MS::nfw_norm(c, w) : 0.0;
    R.w = w;
    if (!w.plausible) {
        m_after_struct(S);
        return;
    }
    R.count = 1;
    GLC_COUNT(1);
    R.active = 0;
    R.comp = 0;
    m_struct_next(S);
}

// one (iteration, component) visit of the structure solver, up to the point where a root is needed
GLC_DEVICE_INLINE void unit_struct(SlotState &S) {
    RhsState &R = S.R;
    NodeCtx &c = S.L.ctx;
    const int comp = R.comp;
    const double j = MS::component_j(S.yt, comp);
    double radius = 0.0, velocity = 0.0;
    R.active++;
    R.j = j;
    R.guess = 0;
    if (R.count == 1) {
        bool guess, needRoot;
        MS::structure_first_pass(c, R.w, R.nfwNorm, comp, j, radius, velocity, guess, needRoot);
        if (needRoot) {
            if (j > 0.0) {
                R.lnj = dm_log(j);
                const double lnrv = dm_log(R.w.rvir);
                R.guess = 1;
                brent_begin(S.B, true, MS::jroot_options(), lnrv - 4.0, lnrv, false, 0.0, 0.0);
                m_root_start(S, U_ROOT_J);
                return;
            }
            radius = 0.0;  // nfw_radius_from_j of a non-positive j
        }
        if (guess) velocity = MS::structure_guess_velocity(c, R.nfwNorm, radius);
        MS::structure_store(c, comp, radius, velocity);
        R.comp++;
        m_struct_next(S);
        return;
    }
    if (j <= 0.0) {
        R.comp++;
        m_struct_next(S);
        return;
    }
    radius = comp == 0 ? c.diskRadius : c.sphRadius;
    R.radius = radius;
    MS::AcProblem P;
    P.fd = P.fi = P.bterm = 0.0;
    P.rup = P.rInit = radius;
    P.need = 0;
    if (GLC_PARAMS.adiabaticContraction && !(radius <= 0.0)) MS::ac_setup(c, S.yt, R.w, R.nfwNorm, radius, P);
    R.ac = P;
    if (P.need) {
        brent_begin(S.B, true, MS::ac_root_options(), radius, P.rup, false, 0.0, 0.0);
        m_root_start(S, U_ROOT_AC);
    } else {
        S.B.busy = 0;
        S.B.status = 0;
        S.unit = U_STRUCT_FIN;
    }
}

// digest the root of a structure visit: first-guess radius, or the contracted dark-matter mass and the
// fixed-point update
GLC_DEVICE_INLINE void unit_struct_fin(SlotState &S) {
    RhsState &R = S.R;
    NodeCtx &c = S.L.ctx;
    const int comp = R.comp;
    double radius, velocity;
    if (R.guess) {
        radius = (S.B.status != 0) ? R.w.rvir : dm_exp(S.B.result);
        velocity = MS::structure_guess_velocity(c, R.nfwNorm, radius);
    } else {
        const double rs = c.dmScale;
        const double fDm = 1.0 - GLC_PARAMS.OmegaBaryon / GLC_PARAMS.OmegaMatter;
        radius = R.radius;
        double mdm;
        if (!GLC_PARAMS.adiabaticContraction)
            mdm = MS::nfw_mass(R.nfwNorm, rs, radius);
        else if (radius <= 0.0)
            mdm = 0.0;
        else {
            double rInit = R.ac.rInit;
            if (R.ac.need) {
                rInit = S.B.result;
                if (S.B.status != 0) R.bad = 1;
            }
            mdm = fDm * MS::nfw_mass(R.nfwNorm, rs, rInit);
        }
        MS::structure_update(c, S.yt, R.w, R.j, mdm, R.count, R.hist[2 * comp], R.hist[2 * comp + 1], R.fit, R.bad, radius,
                             velocity);
    }
    MS::structure_store(c, comp, radius, velocity);
    R.comp++;
    m_struct_next(S);
}

GLC_DEVICE_INLINE void unit_root(SlotState &S, int unit) {
    RhsState &R = S.R;
    BrentState B = S.B;
    const double x = B.x;
    double fx;
    if (unit == U_ROOT_AC) {
        GLC_COUNT(0);
        fx = MS::ac_function(R.nfwNorm, S.L.ctx.dmScale, R.w, R.ac, R.radius, x);
    } else if (unit == U_ROOT_J) {
        fx = MS::jroot_function(R.nfwNorm, S.L.ctx.dmScale, R.lnj, x);
    } else if (unit == U_ROOT_TRUNC) {
        fx = MS::sfr_trunc_function(R.sfr.k, x);
    } else if (unit == U_ROOT_CRIT) {
        fx = MS::sfr_crit_function(R.sfr.k, x);
    } else {
        GLC_COUNT(3);
        fx = MS::cooling_function(R.w, x);
    }
    brent_feed(B, fx);
    brent_advance(B);
    S.B = B;
    if (!B.busy) m_root_complete(S, unit);
}

GLC_DEVICE_INLINE void unit_sfr_begin(SlotState &S) {
    RhsState &R = S.R;
    MS::sfr_setup(S.L.ctx, S.yt, true, R.sfr);
    if (!R.sfr.live) {
        R.psiDisk = 0.0;
        m_cool_decide(S);
        return;
    }
    if (R.sfr.needRmax) {
        brent_begin(S.B, true, MS::sfr_root_options(), 0.0, R.sfr.rOut, false, 0.0, 0.0);
        m_root_start(S, U_ROOT_TRUNC);
    } else
        m_sfr_after_trunc(S, 0.0, 0);
}

GLC_DEVICE_INLINE void unit_qag(SlotState &S) {
    RhsState &R = S.R;
    const MS::Kmt k = R.sfr.k;
    qag_pass(S.Q, [&](double r) {
        GLC_COUNT(2);
        return MS::sfr_integrand(k, r);
    });
    if (S.Q.busy) return;
    const double v = qag_finish(S.Q);
    R.total += v;
    if (S.Q.status == 11) R.bad = 1;
    R.iv++;
    if (R.iv < R.nIv) {
        m_qag_start(S);
        return;
    }
    R.psiDisk = 2.0 * kPi * R.total;
    m_cool_decide(S);
}

GLC_DEVICE_INLINE void unit_cool_begin(SlotState &S) {
    RhsState &R = S.R;
    Work w = R.w;
    double logSlopeT = 0.0, rootOuter, rootZero, result;
    bool need;
    MS::cooling_prepare(S.yt, w, logSlopeT);
    MS::cooling_setup(w, rootOuter, rootZero, result, need);
    R.w = w;
    R.logSlopeT = logSlopeT;
    if (need) {
        brent_begin(S.B, true, MS::cooling_root_options(), 0.0, w.hhRouter, true, rootZero, rootOuter);
        m_root_start(S, U_ROOT_COOL);
    } else {
        R.rinfall = result;
        S.unit = U_RK;
    }
}

// One unit of one slot.  Returns false when the slot is idle.
GLC_DEVICE_INLINE bool machine_step(SlotState &S, const LaneMem &M) {
    const int unit = S.unit;
    switch (unit) {
        case U_RK: unit_rk(S, M); break;
        case U_RHS_BEGIN: unit_rhs_begin(S); break;
        case U_STRUCT: unit_struct(S); break;
        case U_STRUCT_FIN: unit_struct_fin(S); break;
        case U_ROOT_AC:
        case U_ROOT_J:
        case U_ROOT_TRUNC:
        case U_ROOT_CRIT:
        case U_ROOT_COOL: unit_root(S, unit); break;
        case U_SFR_BEGIN: unit_sfr_begin(S); break;
        case U_QAG: unit_qag(S); break;
        case U_COOL_BEGIN: unit_cool_begin(S); break;
        default: return false;
    }
    return true;
}

#if defined(__CUDACC__)
// Persistent time-sliced kernel.  One block per SM owns SLOTS slots (SLOTS a multiple of THREADS).  Every
// iteration it (1) counting-sorts its slots by pending unit in shared memory, (2) sweeps the sorted list: warp w
// executes chunks w, w+W, w+2W, ... of 32 consecutive sorted slots, one unit per slot.  Because the warps of the
// SM walk through the sorted list side by side, they execute the same one or two units at any instant: the
// instruction working set of the SM is one unit's code, not the whole machine's, and a warp is pure except at the
// few unit boundaries of the list.  All per-slot state lives in HBM/L2 (SlotState + the RK stage vectors), so
// parking at the end of a time slice costs nothing.
template <int THREADS, int SLOTS>
__global__ void __launch_bounds__(THREADS, 1) machine_kernel(KernelArgs A, SlotState *slots) {
    constexpr int PER = SLOTS / THREADS;
    constexpr int WARPS = THREADS / 32;
    __shared__ int s_hist[U_COUNT];
    __shared__ int s_off[U_COUNT];
    __shared__ unsigned short s_perm[SLOTS];
    __shared__ unsigned char s_unit[SLOTS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * SLOTS;
#pragma unroll 1
    for (int k = 0; k < PER; k++) {
        const int s = tid + k * THREADS;
        SlotState &own = slots[base + s];
        if (!A.resume) slot_reset(own);
        if (own.unit == U_IDLE) {  // the queue may have grown since the last slice
            own.L.phase = PH_FETCH;
            own.unit = U_RK;
        }
        s_unit[s] = (unsigned char)own.unit;
    }
    __syncthreads();
    for (int it = 0; it < A.budget; ++it) {
        // ---- regroup the block's slots by pending unit
        if (tid < U_COUNT) s_hist[tid] = 0;
        __syncthreads();
        int myUnit[PER], myRank[PER];
#pragma unroll
        for (int k = 0; k < PER; k++) {
            myUnit[k] = s_unit[tid + k * THREADS];
            myRank[k] = atomicAdd(&s_hist[myUnit[k]], 1);
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (int u = 0; u < U_COUNT; u++) {
                s_off[u] = acc;
                acc += s_hist[u];
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PER; k++) s_perm[s_off[myUnit[k]] + myRank[k]] = (unsigned short)(tid + k * THREADS);
        const int nActive = SLOTS - s_hist[U_IDLE];
        __syncthreads();
        if (nActive == 0) break;
        // ---- sweep: one unit per active slot
#pragma unroll 1
        for (int c = warp; c * 32 < nActive; c += WARPS) {
            const int p = c * 32 + lane;
            if (p < nActive) {
                const int s = s_perm[p];
                const int64_t slot = base + s;
                LaneMem M{&A, A.ws + slot * (WS_NVEC * NY), 1};
                machine_step(slots[slot], M);
                s_unit[s] = (unsigned char)slots[slot].unit;
            }
            __syncwarp();
        }
        __syncthreads();
    }
    __syncthreads();
    // ---- counters of this block's slots: warp-reduce then one atomic per warp per counter
    unsigned int vals[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 1
    for (int k = 0; k < PER; k++) {
        SlotState &own = slots[base + tid + k * THREADS];
        LaneState &L = own.L;
        vals[0] += L.nAcc;
        vals[1] += L.nRej;
        vals[2] += L.nRhs;
        vals[3] += L.nSeg;
        vals[4] += L.nTrialFail;
        vals[5] += L.nNodes;
        vals[6] += L.nDone;
        vals[7] += own.unit != U_IDLE ? 1u : 0u;
        L.nAcc = L.nRej = L.nRhs = L.nSeg = L.nTrialFail = L.nNodes = L.nDone = 0;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        unsigned int v = vals[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&A.counters[k], (unsigned long long)v);
    }
}
#endif

}  // namespace glc
