// glc_evolve_kernel.cuh -- the batched adaptive Cash-Karp solver.
//
// Execution model (B200): one lane integrates one node; a lane that finishes its node pulls the next
// one from a global queue (atomic counter over a component-sorted order), so divergence in step counts
// between nodes is absorbed at RHS granularity ("compaction by refill").  The per-lane solver is an
// explicit state machine whose main loop performs EXACTLY ONE heavy call per iteration -- one evaluation
// of the model's rate function (a Cash-Karp stage, the dydt_out evaluation, or the post-evolve structure
// solve, all through the same call site) -- so the 32 lanes of a warp re-converge at the top of the rate
// function every iteration whatever stage/attempt/segment each of them is in.  The bookkeeping between
// heavy calls (tableau combinations, error norm, step-size controller, post-step clamps, node fetch and
// write-back) is short.  A launch is a TIME SLICE: every lane performs at most `budget` heavy calls and
// then parks its state in HBM (LaneState, one record per resident lane); the next launch resumes it.
//
// Semantics restated (reference paths relative to /root/reference):
//   standardEvolve          source/merger_trees/node_evolver/standard.F90:385-755
//   standardODEs            :831-946   (interrupt bookkeeping :901-928)
//   standardPostStepProcessing :1160-1185
//   odeSolverSolve          source/numerical/ODE_solver/solver.F90:492-636
//   gsl_odeiv2_driver2_apply source/external/gslODEInitVal2/driver2.c:148-250
//   sc2_control_hadjust     source/external/gslODEInitVal2/cscal2.c:93-169
//   rkck_apply / gsl_odeiv2_evolve_apply   libgsl 2.6 ode-initval2/{rkck.c,evolve.c}
#pragma once

#include <float.h>

#include "glc_common.cuh"

#ifdef GLC_TRACE
#define GTR(...) printf(__VA_ARGS__)
#else
#define GTR(...)
#endif

// NB: the 24-component loops of the RK bookkeeping are kept rolled (#pragma unroll 1): the kernels are
// instruction-fetch bound, not issue bound, so a compact loop that runs out of the L0 instruction cache beats
// 24 unrolled copies streamed from L2.

#ifdef GLC_RK_FULL_UNROLL
#define GLC_UNROLL_RK _Pragma("unroll")
#else
#define GLC_UNROLL_RK _Pragma("unroll 1")
#endif

namespace glc {

// GLC_LEDGER (debug builds of the machine): every node records which slot holds it from fetch to write-back; a second
// fetch, a write-back by a slot that does not own the node, or a node left owned at the end of a batch are violations.
#if defined(GLC_LEDGER) && defined(__CUDACC__)
#define GLC_LEDGER_FETCH(A, M, node)                                                       \
    do {                                                                                   \
        if ((A).ledger) {                                                                  \
            const int old__ = atomicExch(&(A).ledger[node], (M).dbgSlot + 1);              \
            if (old__ != -1) atomicAdd(&(A).ledgerErr[0], 1ull);                           \
        }                                                                                  \
    } while (0)
#define GLC_LEDGER_DONE(A, M, node)                                                        \
    do {                                                                                   \
        if ((A).ledger) {                                                                  \
            const int old__ = atomicExch(&(A).ledger[node], -2);                           \
            if (old__ != (M).dbgSlot + 1) atomicAdd(&(A).ledgerErr[1], 1ull);              \
        }                                                                                  \
    } while (0)
#else
#define GLC_LEDGER_FETCH(A, M, node) ((void)0)
#define GLC_LEDGER_DONE(A, M, node) ((void)0)
#endif

// Cash-Karp tableau (Cash & Karp 1990). Row s = weights of k1..k6 used to build the input of
// stage s (s=1..5 -> k2..k6; s=6 -> 5th-order solution); row 0 = error weights (5th-4th order).
#define GLC_RK_B_INIT                                                                                    \
    {                                                                                                    \
        {37.0 / 378.0 - 2825.0 / 27648.0, 0.0, 250.0 / 621.0 - 18575.0 / 48384.0,                        \
         125.0 / 594.0 - 13525.0 / 55296.0, -277.0 / 14336.0, 512.0 / 1771.0 - 0.25},                    \
            {1.0 / 5.0, 0, 0, 0, 0, 0}, {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0, 0}, {0.3, -0.9, 1.2, 0, 0, 0}, \
            {-11.0 / 54.0, 2.5, -70.0 / 27.0, 35.0 / 27.0, 0, 0},                                        \
            {1631.0 / 55296.0, 175.0 / 512.0, 575.0 / 13824.0, 44275.0 / 110592.0, 253.0 / 4096.0, 0},   \
        {                                                                                                \
            37.0 / 378.0, 0.0, 250.0 / 621.0, 125.0 / 594.0, 0.0, 512.0 / 1771.0                         \
        }                                                                                                \
    }
__constant__ double c_rk_b[7][6] = GLC_RK_B_INIT;
__constant__ double c_rk_a[7] = {0.0, 1.0 / 5.0, 0.3, 3.0 / 5.0, 1.0, 7.0 / 8.0, 1.0};

enum Phase : int { PH_FETCH = 0, PH_SEGMENT, PH_TRIAL, PH_ATTEMPT, PH_STAGE, PH_SOLVE_DONE, PH_WRITEBACK, PH_IDLE };
enum Heavy : int { HV_NONE = 0, HV_RHS, HV_FROZEN, HV_POST_EVOLVE };

constexpr int kTrialCountMaximum = 8;  // standard.F90:135
constexpr int kSegmentGuard = 64;

// Everything a lane carries between heavy calls.  POD: parked in HBM between time slices.
struct LaneState {
    NodeCtx ctx;
    double tEnd, x, x1, h, t0, h0, timeStartSaved, timeStepIn, timeInterruptFirst, rmax, ts;
    int phase, node, stage, heavy;
    uint32_t mask;
    int interruptFound, interruptCode, count, inApply, outWritten, yslot, kslot, trial, finalStep;
    int segmentsThisNode, nodeStatus, solveFailed, forbiddenNegatives;
    unsigned int nAcc, nRej, nRhs, nSeg, nTrialFail, nNodes, nDone;
    unsigned int age;  // rate-function evaluations spent on the current node so far (scheduling hint only: oldest nodes first)
    // profileOdeEvolver: largest scaled error of the attempt and the property that has it (standardStepErrorAnalyzer
    // :1210-1219), evolve_apply calls since the last successful one (countEvaluationsToSuccess)
    double profErrMax;
    int limiting, evalsToSuccess;
};

GLC_DEVICE_INLINE void lane_reset(LaneState &L) {
    L = LaneState();
    L.phase = PH_FETCH;
    L.node = -1;
    L.h = 1.0;
}

GLC_DEVICE_INLINE bool prop_is_non_negative(int prop) {
    return prop != GLC_P_SAT_BOUND_MASS;  // isNonNegative attributes of the component definitions
}

// Queue order key (see glc_api.cu, "queue order"): ascending bucket id = position in the queue.
GLC_DEVICE_INLINE int queue_bucket(int flags) {
    const int rich = ((flags & GLC_F_HAS_SPHEROID) ? 8 : 0) | ((flags & GLC_F_HAS_DISK) ? 4 : 0) |
                     ((flags & GLC_F_HAS_BH) ? 2 : 0) | ((flags & GLC_F_IS_SATELLITE) ? 1 : 0);
    const int fresh = (flags & GLC_F_HAS_HOTHALO) ? ((flags & GLC_F_HH_INITIALIZED) ? 0 : 1) : 2;
    return (2 - fresh) * 16 + (15 - rich);  // ascending bucket id = queue order; < 64
}

// Accessors of one lane's slice of the arena / workspace (both SoA: consecutive lanes touch consecutive
// addresses, so every access is one coalesced 256-B line per warp).
struct LaneMem {
    const KernelArgs *A;
    double *ws;       // this lane's / slot's first workspace word
    int64_t wstride;  // nslots: SoA over resident lanes (evolve_kernel); 1: one contiguous record per slot (machine)
    int dbgSlot = -1; // GLC_LEDGER builds: id of the slot that runs this lane state
    GLC_DEVICE_METHOD double &W(int vec, int comp) const { return ws[((int64_t)vec * NY + comp) * wstride]; }
    GLC_DEVICE_METHOD double &AR(int prop, int node) const { return A->props[(int64_t)prop * A->cap + node]; }
};

// standardEvolve epilogue, part 2 (:657-753): interrupt hand-off and write-back of the final state yt.
template <class Model>
GLC_DEVICE_INLINE void lane_writeback(LaneState &L, const LaneMem &M, double (&yt)[NY]) {
    const KernelArgs &A = *M.A;
    NodeCtx &ctx = L.ctx;
    const int node = L.node;
    double timeOut, timeStepOut;
    int interrupted = 0;
    if (L.timeInterruptFirst != 0.0) {
        interrupted = 1;
        timeOut = L.timeInterruptFirst;
        timeStepOut = -1.0;
    } else {
        timeOut = L.tEnd;
        timeStepOut = (L.timeStartSaved != L.tEnd && L.mask != 0u) ? L.h : -1.0;
    }
    if (ctx.numericsFailed) L.nodeStatus = GLC_STATUS_NONFINITE;
    int code = interrupted ? L.interruptCode : GLC_INT_NONE;
    if (interrupted && GLC_PARAMS.resolveInterruptsOnDevice) {
        // functionInterrupt: <class>CreateByInterrupt / blackHoleCreate
        if (code == GLC_INT_HOTHALO_CREATE) ctx.flags |= GLC_F_HAS_HOTHALO;
        if (code == GLC_INT_DISK_CREATE) ctx.flags |= GLC_F_HAS_DISK;
        if (code == GLC_INT_SPHEROID_CREATE) ctx.flags |= GLC_F_HAS_SPHEROID;
        if (code == GLC_INT_BH_CREATE) {
            ctx.flags |= GLC_F_HAS_BH;
            yt[GLC_P_BH_MASS] = GLC_PARAMS.bhSeedMass;
            yt[GLC_P_BH_SPIN] = GLC_PARAMS.bhSeedSpin;
        }
        code = GLC_INT_NONE;
    }
GLC_UNROLL_RK
    for (int i = 0; i < NY; i++) M.AR(i, node) = yt[i];
    M.AR(GLC_P_TIME, node) = timeOut;
    M.AR(GLC_P_TIME_STEP, node) = timeStepOut;
    M.AR(GLC_P_DISK_RADIUS, node) = ctx.diskRadius;
    M.AR(GLC_P_DISK_VELOCITY, node) = ctx.diskVelocity;
    M.AR(GLC_P_SPH_RADIUS, node) = ctx.sphRadius;
    M.AR(GLC_P_SPH_VELOCITY, node) = ctx.sphVelocity;
    M.AR(GLC_P_BASIC_MASS, node) = ctx.basicMass;
    M.AR(GLC_P_DMSCALE, node) = ctx.dmScale;
    M.AR(GLC_P_SPIN, node) = ctx.spinJ;
    A.flags[node] = ctx.flags;
    if (interrupted && code == GLC_INT_NONE && timeOut < L.tEnd) {
        if (L.segmentsThisNode < kSegmentGuard) {
            L.phase = PH_SEGMENT;  // host loop evolver/standard.F90:425-476, resolved in place
            return;
        }
        L.nodeStatus = GLC_STATUS_FAIL;
    }
    A.status[node] = L.nodeStatus;
    A.interrupt[node] = code;
    GLC_LEDGER_DONE(A, M, node);
    L.nDone++;
    L.phase = PH_FETCH;
}

// Runs the lane's bookkeeping up to the next heavy call.  On return L.heavy says what the caller must
// evaluate (HV_RHS / HV_POST_EVOLVE: Model::rates on yt at time L.ts; HV_FROZEN: nothing) before calling
// lane_consume; HV_NONE means the lane is out of work.
template <class Model>
GLC_DEVICE_INLINE void lane_prepare(LaneState &L, const LaneMem &M, double (&yt)[NY]) {
    const KernelArgs &A = *M.A;
    for (;;) {
        // ------------------------------------------------------------------ fetch a node
        if (L.phase == PH_FETCH) {
            const int q = glc_atomic_add(A.work_counter, 1);
            if (q >= A.n) {
                glc_atomic_add(A.work_counter, -1);  // keep the cursor at n: the queue may grow between slices
                L.phase = PH_IDLE;
                L.heavy = HV_NONE;
                return;
            }
            L.node = A.order ? A.order[q] : q;
            GLC_LEDGER_FETCH(A, M, L.node);
            L.nNodes++;
            L.age = 0;
            L.ctx.flags = A.flags[L.node];
            L.tEnd = A.time_end[L.node];
            L.segmentsThisNode = 0;
            L.nodeStatus = GLC_STATUS_SUCCESS;
            L.phase = PH_SEGMENT;
        }
        // ------------------------------------------------- standardEvolve prologue (:434-576)
        if (L.phase == PH_SEGMENT) {
            const int node = L.node;
            NodeCtx &ctx = L.ctx;
            double s[NY];
GLC_UNROLL_RK
            for (int i = 0; i < NY; i++) {
                yt[i] = M.AR(i, node);
                s[i] = 0.0;
            }
            ctx.massTarget = M.AR(GLC_P_MASS_TARGET, node);
            ctx.massRate = M.AR(GLC_P_MASS_RATE, node);
            ctx.timeTarget = M.AR(GLC_P_TIME_TARGET, node);
            ctx.scaleTarget = M.AR(GLC_P_DMSCALE_TARGET, node);
            ctx.scaleRate = M.AR(GLC_P_DMSCALE_RATE, node);
            ctx.spinTarget = M.AR(GLC_P_SPIN_TARGET, node);
            ctx.spinRate = M.AR(GLC_P_SPIN_RATE, node);
            ctx.timeLastIsolated = M.AR(GLC_P_TIME_LAST_ISOLATED, node);
            ctx.diskRadius = M.AR(GLC_P_DISK_RADIUS, node);
            ctx.diskVelocity = M.AR(GLC_P_DISK_VELOCITY, node);
            ctx.sphRadius = M.AR(GLC_P_SPH_RADIUS, node);
            ctx.sphVelocity = M.AR(GLC_P_SPH_VELOCITY, node);
            ctx.basicMass = M.AR(GLC_P_BASIC_MASS, node);
            ctx.dmScale = M.AR(GLC_P_DMSCALE, node);
            ctx.spinJ = M.AR(GLC_P_SPIN, node);
            ctx.massBaryonicSubhalos = M.AR(GLC_P_MASS_BARYONIC_SUBHALOS, node);
            ctx.numericsFailed = 0;
            L.timeStartSaved = M.AR(GLC_P_TIME, node);
            ctx.timeNode = L.timeStartSaved;
            L.timeStepIn = M.AR(GLC_P_TIME_STEP, node);
            L.nSeg++;
            L.segmentsThisNode++;
            {
                // pre-evolve hooks edit the node itself (they run before the solver's saved copy is taken,
                // standard.F90:434-441 vs :518-527), so their edits must survive a solver restart
                const int flagsBefore = ctx.flags;
                Model::pre_evolve(ctx, yt);
                if (ctx.flags != flagsBefore) {
GLC_UNROLL_RK
                    for (int i = 0; i < NY; i++) M.AR(i, node) = yt[i];
                    M.AR(GLC_P_BASIC_MASS, node) = ctx.basicMass;
                    A.flags[node] = ctx.flags;
                }
            }
            L.mask = Model::active_mask(ctx.flags);
            Model::scales(ctx, yt, s);
GLC_UNROLL_RK
            for (int i = 0; i < NY; i++) {
                M.W(WS_YA, i) = yt[i];
                M.W(WS_SCALE, i) = s[i];
            }
            L.yslot = 0;
            L.interruptFound = 0;
            L.timeInterruptFirst = 0.0;
            L.interruptCode = GLC_INT_NONE;
            L.trial = 0;
            L.solveFailed = 0;
            L.evalsToSuccess = 0;
            L.x = L.timeStartSaved;
            L.phase = (L.timeStartSaved != L.tEnd && L.mask != 0u) ? PH_TRIAL : PH_SOLVE_DONE;
        }
        // ------------------------------ trial start (:587-652) + odeSolverSolve prologue
        if (L.phase == PH_TRIAL) {
            if (L.trial > 0) {
GLC_UNROLL_RK
                for (int i = 0; i < NY; i++) M.W(WS_YA, i) = M.AR(i, L.node);
            }
            L.yslot = 0;
            const double stepSize = GLC_PARAMS.reuseODEStepSize ? L.timeStepIn / dm_scale2(1.0, L.trial) : -1.0;
            L.x = L.timeStartSaved;
            L.x1 = L.tEnd;
            double xStep = L.x1 - L.x;
            if (stepSize > 0.0) xStep = fmin(stepSize, xStep);
            L.count = 0;  // GSL_ODEIV2_Driver_Reset
            if (xStep != 0.0) L.h = xStep;
            L.inApply = 0;
            L.outWritten = 0;
            L.phase = PH_ATTEMPT;
        }
        // --------------------------------------------- start of one try_step of evolve_apply
        if (L.phase == PH_ATTEMPT) {
            int needK1 = 0;
            if (!L.inApply) {
                L.t0 = L.x;
                L.h0 = L.h;
                if (L.count > 0 && L.outWritten) L.kslot ^= 1;  // dydt_in := dydt_out
                L.outWritten = 0;
                needK1 = (L.count == 0);
            }
            const double dt = L.x1 - L.t0;
            if (L.h0 > dt) {
                L.h0 = dt;
                L.finalStep = 1;
            } else {
                L.finalStep = 0;
            }
            L.rmax = DBL_MIN;
            L.profErrMax = 0.0;
            L.limiting = -1;
            L.forbiddenNegatives = 0;
            L.stage = needK1 ? 0 : 1;
            L.phase = PH_STAGE;
        }
        // --------------------------------------------- input of one Cash-Karp stage
        if (L.phase == PH_STAGE) {
            const int stage = L.stage;
            const int ySrc = WS_YA + L.yslot, yDst = WS_YA + (L.yslot ^ 1);
            const int k1v = WS_KA + L.kslot;
            const double h0 = L.h0;
            L.ts = L.t0 + c_rk_a[stage] * h0;
            if (stage == 0) {
GLC_UNROLL_RK
                for (int i = 0; i < NY; i++) yt[i] = M.W(ySrc, i);
            } else if (stage == 1) {
                const double b10 = c_rk_b[1][0];
GLC_UNROLL_RK
                for (int i = 0; i < NY; i++) yt[i] = M.W(ySrc, i) + b10 * h0 * M.W(k1v, i);  // rkck.c: y + b21*h*k1
            } else if (stage < 6) {
                double b[5];
GLC_UNROLL_RK
                for (int j = 0; j < 5; j++) b[j] = c_rk_b[stage][j];
GLC_UNROLL_RK
                for (int i = 0; i < NY; i++) {
                    double acc = b[0] * M.W(k1v, i);
                    for (int j = 1; j < stage; j++) acc += b[j] * M.W(WS_K2 + j - 1, i);
                    yt[i] = M.W(ySrc, i) + h0 * acc;
                }
            } else {
                // 5th-order solution, error estimate and the controller's rmax (cscal2.c:109-126;
                // a_dydt = 0 so rmax does not depend on dydt_out)
                const double epsAbs = GLC_PARAMS.odeToleranceAbsolute, epsRel = GLC_PARAMS.odeToleranceRelative;
                const bool nonNeg = GLC_PARAMS.enforceNonNegativity != 0;
                double rmax = L.rmax;
                int forbidden = L.forbiddenNegatives;
GLC_UNROLL_RK
                for (int i = 0; i < NY; i++) {
                    const double k1 = M.W(k1v, i), k3 = M.W(WS_K3, i), k4 = M.W(WS_K4, i), k5 = M.W(WS_K5, i),
                                 k6 = M.W(WS_K6, i);
                    const double d = c_rk_b[6][0] * k1 + c_rk_b[6][2] * k3 + c_rk_b[6][3] * k4 + c_rk_b[6][5] * k6;
                    const double ynew = M.W(ySrc, i) + h0 * d;
                    const double yerr = h0 * (c_rk_b[0][0] * k1 + c_rk_b[0][2] * k3 + c_rk_b[0][3] * k4 +
                                              c_rk_b[0][4] * k5 + c_rk_b[0][5] * k6);
                    yt[i] = ynew;
                    M.W(yDst, i) = ynew;
                    if (L.mask & (1u << i)) {
                        const double D0 = epsRel * fabs(ynew) + epsAbs * M.W(WS_SCALE, i);
                        const double r = fabs(yerr) / fabs(D0);
                        rmax = fmax(r, rmax);
                        if (GLC_TABLES.profile && r > L.profErrMax) {  // scaledError > scaledErrorMaximum: the first maximum wins
                            L.profErrMax = r;
                            L.limiting = i;
                        }
                        if (nonNeg && prop_is_non_negative(i) && ynew < 0.0) forbidden = 1;
                    }
                }
                L.rmax = rmax;
                L.forbiddenNegatives = forbidden;
            }
            // ---- standardODEs, part 1
            Model::solve_analytics(L.ctx, L.ts);
            L.nRhs++;  // every call of the derivatives function counts (also the frozen ones past an interrupt)
            L.age++;
            if (L.interruptFound && L.ts >= L.timeInterruptFirst) {
                Model::solve_analytics(L.ctx, L.timeInterruptFirst);
                L.heavy = HV_FROZEN;
            } else {
                L.heavy = HV_RHS;
            }
            return;
        }
        // -------------------------------- trial epilogue + standardEvolve epilogue (:657-753), part 1
        if (L.phase == PH_SOLVE_DONE) {
            const int yv = WS_YA + L.yslot;
GLC_UNROLL_RK
            for (int i = 0; i < NY; i++) yt[i] = M.W(yv, i);
            if (L.solveFailed) {
                int rescued = 0;
                if (GLC_PARAMS.enforceNonNegativity) {
GLC_UNROLL_RK
                    for (int i = 0; i < NY; i++)
                        if ((L.mask & (1u << i)) && prop_is_non_negative(i) && yt[i] < 0.0) {
                            yt[i] = 0.0;
                            rescued = 1;
                        }
                }
                if (rescued) {
                    L.h = L.timeStepIn / dm_scale2(1.0, L.trial);
                    L.solveFailed = 0;
                } else {
                    L.trial++;
                    L.nTrialFail++;
                    L.solveFailed = 0;
                    if (L.trial < kTrialCountMaximum) {
                        L.phase = PH_TRIAL;
                        continue;
                    }
                    // errorStatusUnderflow: node left at its saved values
                    A.status[L.node] = GLC_STATUS_UNDERFLOW;
                    A.interrupt[L.node] = GLC_INT_NONE;
                    GLC_LEDGER_DONE(A, M, L.node);
                    L.nDone++;
                    L.phase = PH_FETCH;
                    continue;
                }
            }
            Model::solve_analytics(L.ctx, L.tEnd);
            L.ctx.timeNode = (L.timeInterruptFirst != 0.0) ? L.timeInterruptFirst : L.tEnd;
            L.ts = L.ctx.timeNode;
            L.phase = PH_WRITEBACK;
            if (Model::kHasPostEvolve) {
                L.heavy = HV_POST_EVOLVE;  // <eventHook postEvolve>: structure solve at the final state
                return;                    // lane_consume writes the node back
            }
        }
        if (L.phase == PH_WRITEBACK) {
            lane_writeback<Model>(L, M, yt);
            continue;
        }
        if (L.phase == PH_IDLE) {
            L.heavy = HV_NONE;
            return;
        }
    }
}

// Digests the result of the heavy call: stores the stage derivative, and at the end of an attempt runs the
// step-size controller, the accept/reject logic of gsl_odeiv2_evolve_apply and the post-step hook.
template <class Model>
GLC_DEVICE_INLINE void lane_consume(LaneState &L, const LaneMem &M, double (&yt)[NY], double (&rate)[NY], int code) {
    if (L.heavy == HV_NONE) return;
    if (L.heavy == HV_POST_EVOLVE) {
        if (Model::kHasPostEvolve) lane_writeback<Model>(L, M, yt);
        return;
    }
    // ---- standardODEs, part 2 (interrupt bookkeeping :901-928)
    int ebadfunc = 0;
    if (L.heavy == HV_RHS && code != GLC_INT_NONE) {
GLC_UNROLL_RK
        for (int i = 0; i < NY; i++) rate[i] = 0.0;
        if (L.ts < L.timeInterruptFirst || !L.interruptFound) {
            L.interruptFound = 1;
            L.timeInterruptFirst = L.ts;
            L.interruptCode = code;
            ebadfunc = 1;
        }
    }
    {
        const int stage = L.stage;
        const int kv = (stage == 0) ? WS_KA + L.kslot : ((stage == 6) ? WS_KA + (L.kslot ^ 1) : WS_K2 + stage - 1);
        int nonfinite = 0;
GLC_UNROLL_RK
        for (int i = 0; i < NY; i++) {
            const double r = (L.mask & (1u << i)) ? rate[i] : 0.0;
            if (!isfinite(r)) nonfinite = 1;
            M.W(kv, i) = r;
        }
        if (nonfinite) L.nodeStatus = GLC_STATUS_NONFINITE;
        if (stage == 6) L.outWritten = 1;
    }
    if (ebadfunc) {
        // odeSolverInterrupt, solver.F90:608-618.  NB (GSL quirk, reproduced): gsl_odeiv2_evolve_apply
        // leaves *t at the end of a REJECTED attempt when the retry returns early, so x may be ahead
        // of t0 here; if it is beyond the interrupt time the solver restarts from the initial state.
        L.x1 = L.timeInterruptFirst;
        L.inApply = 0;
        L.evalsToSuccess++;  // the analyzer sees the interrupted evolve_apply call too (status /= success: counted, :1205-1207)
        if (L.x > L.x1) {
GLC_UNROLL_RK
            for (int i = 0; i < NY; i++) M.W(WS_YA, i) = M.AR(i, L.node);
            L.yslot = 0;
            L.x = L.timeStartSaved;
            L.count = 0;  // GSL_ODEIV2_Driver_Reset
            L.outWritten = 0;
        }
        L.phase = (L.x < L.x1) ? PH_ATTEMPT : PH_SOLVE_DONE;
        return;
    }
    if (L.stage < 6) {
        L.stage++;
        return;  // phase stays PH_STAGE
    }
    // ---------------------------------------------------------------- end of the attempt
    L.count++;
    const double tNew = L.finalStep ? L.x1 : L.t0 + L.h0;
    L.x = tNew;  // evolve.c sets *t before the controller runs and does not restore it on a retry
    // ---- sc2_control_hadjust (ord = 5)
    const double hOld = L.h0;
    const double rmax = L.rmax;
    int dec = 0;
    if (rmax > 1.1) {
        double r = 0.9 / dm_pow(rmax, 1.0 / 5.0);
        if (r < 0.2) r = 0.2;
        L.h0 = r * hOld;
        dec = 1;
    } else if (L.forbiddenNegatives) {
        L.h0 = 0.5 * hOld;
        dec = 1;
    } else if (rmax < 0.5) {
        double r = 0.9 / dm_pow(rmax, 1.0 / 6.0);
        if (r > 4.9) r = 4.9;
        if (r < 1.0) r = 1.0;
        L.h0 = r * hOld;
    }
    if (dec) {
        const double tNext = tNew + L.h0;
        if (fabs(L.h0) < fabs(hOld) && tNext != tNew) {
            L.nRej++;
            L.inApply = 1;  // y := y0 is implicit (yslot unchanged)
            L.phase = PH_ATTEMPT;
            return;
        }
        // GSL_FAILURE: step-size underflow (the analyzer counts the call and returns, :1205-1207)
        L.evalsToSuccess++;
        L.h = L.h0;
        L.inApply = 0;
        L.solveFailed = 1;
        L.yslot ^= 1;  // as in GSL, y holds the failed step's result
        L.phase = PH_SOLVE_DONE;
        return;
    }
    // ---- accepted
    if (GLC_TABLES.profile) {
        // standardStepErrorAnalyzer (:1187-1239) -> mergerTreeEvolveProfilerSimple::profile (simple.F90:250-304); the step
        // handed to it is evolve's last_step (driver2.c:195-198)
        unsigned long long *P = GLC_TABLES.profile;
        const int nb = GLC_TABLES.profBins;
        int lo = 0, hi = nb - 1;  // gsl_interp_bsearch over the bin edges
        while (hi > lo + 1) {
            const int mid = (hi + lo) >> 1;
            if (GLC_TABLES.profEdges[mid] > hOld)
                hi = mid;
            else
                lo = mid;
        }
        const unsigned long long evals = (unsigned long long)L.evalsToSuccess + 1ull;
        glc_atomic_add(&P[0 * GLC_PROFILE_BINS + lo], 1ull);
        glc_atomic_add(&P[1 * GLC_PROFILE_BINS + lo], evals);
        if (L.interruptFound) {
            glc_atomic_add(&P[2 * GLC_PROFILE_BINS + lo], 1ull);
            glc_atomic_add(&P[3 * GLC_PROFILE_BINS + lo], evals);
        }
        glc_atomic_add(&P[(L.limiting >= 0) ? kProfHits + L.limiting : kProfUnknown], 1ull);
        glc_atomic_min_positive_double(&P[kProfSmallest], hOld);
        L.evalsToSuccess = 0;
    }
    if (!L.finalStep) L.h = L.h0;
    L.yslot ^= 1;
    L.inApply = 0;
    L.nAcc++;
    {
        // standardPostStepProcessing (rate[] is dead here and is re-used as the state buffer)
        const int yv = WS_YA + L.yslot;
GLC_UNROLL_RK
        for (int i = 0; i < NY; i++) rate[i] = M.W(yv, i);
        Model::solve_analytics(L.ctx, L.x);
        const int st = Model::post_step(L.ctx, rate);
        if (st != kGslSuccess) {
GLC_UNROLL_RK
            for (int i = 0; i < NY; i++) M.W(yv, i) = rate[i];
        }
        if (st != kGslSuccess && st != kGslContinue) L.count = 0;  // gsl_odeiv2_evolve_reset
    }
    L.phase = (L.x < L.x1) ? PH_ATTEMPT : PH_SOLVE_DONE;
}

// One iteration of a lane: bookkeeping, ONE heavy call, digestion.  Returns false when the lane is idle.
template <class Model>
GLC_DEVICE_INLINE bool lane_iterate(LaneState &L, const LaneMem &M) {
    double yt[NY], rate[NY];
    lane_prepare<Model>(L, M, yt);
    int code = GLC_INT_NONE;
GLC_UNROLL_RK
    for (int i = 0; i < NY; i++) rate[i] = 0.0;
    // The single heavy call site.  Warp-synchronous: every lane of the warp calls it in every iteration (idle and
    // frozen lanes with on = false), so the votes inside the rate function see the whole warp.
    GLC_SYNCWARP();
    code = Model::rates(L.ctx, L.ts, yt, rate, L.heavy == HV_POST_EVOLVE, L.heavy == HV_RHS || L.heavy == HV_POST_EVOLVE);
    lane_consume<Model>(L, M, yt, rate, code);
    return L.heavy != HV_NONE;
}

// Drain path.  Takes over the slots the micro-task machine (glc_machine.cuh) has parked at an RK boundary -- lane
// state after lane_prepare, stage input yt ready, nothing of the evaluation started -- and finishes their nodes with
// whole evaluations: one lane per node, the warp-synchronous rate function, 255 registers.  When few nodes are left
// there is nothing to re-group, and running an evaluation straight through is ~5x cheaper than unit by unit.
// One call = take the next held slot if this lane has none, then ONE iteration of it.  Returns false when the lane
// is out of work.  `fresh` = the lane's state was just loaded (skip lane_prepare: it has already run).
GLC_DEVICE_INLINE void drain_tally(LaneState &L, unsigned int (&tot)[7]) {
    tot[0] += L.nAcc;
    tot[1] += L.nRej;
    tot[2] += L.nRhs;
    tot[3] += L.nSeg;
    tot[4] += L.nTrialFail;
    tot[5] += L.nNodes;
    tot[6] += L.nDone;
    L.nAcc = L.nRej = L.nRhs = L.nSeg = L.nTrialFail = L.nNodes = L.nDone = 0;
}

template <class Model>
GLC_DEVICE_INLINE bool drain_iterate(LaneState &L, LaneMem &M, const KernelArgs &A, double (&yt)[NY], bool &fresh,
                                     int &slotHeld, unsigned int (&tot)[7], bool mayTake) {
    double rate[NY];
    if (slotHeld == -1 && mayTake) {
        for (;;) {
            const int h = glc_atomic_add(A.held_counter, 1);
            if (h >= A.nheld) {
                slotHeld = -2;  // list exhausted
                break;
            }
            const int entry = A.held[h];
            const int s = entry & ~kHeldFresh;
            M.ws = A.ws + (int64_t)s * (WS_NVEC * NY);
            M.wstride = 1;
            M.dbgSlot = s;
            if (entry & kHeldFresh) {  // a free slot: start with a fetch from the node queue (streaming sessions)
                slotHeld = s;
                lane_reset(L);
                fresh = false;
                break;
            }
            if (A.slotUnit && A.slotUnit[s] < 0) continue;  // finished in an earlier pass over the same list
            slotHeld = s;
            L = A.slotL[slotHeld];
            GLC_UNROLL_RK
            for (int i = 0; i < NY; i++) yt[i] = A.slotYt[(int64_t)slotHeld * NY + i];
            fresh = true;
            break;
        }
    }
    bool have = slotHeld >= 0;
    if (have && !fresh) {
        lane_prepare<Model>(L, M, yt);
        if (L.heavy == HV_NONE) {  // the fetch found the node queue empty: release the slot
            drain_tally(L, tot);
            if (A.slotUnit) A.slotUnit[slotHeld] = -1;
            slotHeld = -1;
            have = false;
        }
    }
    fresh = false;
    GLC_UNROLL_RK
    for (int i = 0; i < NY; i++) rate[i] = 0.0;
    const bool on = have && (L.heavy == HV_RHS || L.heavy == HV_POST_EVOLVE);
    GLC_SYNCWARP();
    const int code = Model::rates(L.ctx, L.ts, yt, rate, L.heavy == HV_POST_EVOLVE, on);
    if (have) {
        lane_consume<Model>(L, M, yt, rate, code);
        if (L.phase == PH_FETCH && !A.drainRefill) {  // node written back; no refill in the drain of a batch: release the slot
            drain_tally(L, tot);
            if (A.slotUnit) A.slotUnit[slotHeld] = -1;
            slotHeld = -1;
        }
        // (with refill the next lane_prepare fetches the next queued node into this slot, or releases it above)
    }
    return slotHeld >= 0 || (slotHeld == -1 && mayTake);
}

// Parks a node the drain kernel could not finish within its budget: brings the lane to the next RK boundary (state
// after lane_prepare, exactly what the machine parks) and stores it back into the slot.
template <class Model>
GLC_DEVICE_INLINE void drain_park(LaneState &L, LaneMem &M, const KernelArgs &A, double (&yt)[NY], bool fresh, int slotHeld,
                                  unsigned int (&tot)[7]) {
    if (slotHeld < 0) return;
    double rate[NY];
    if (!fresh) {
        for (;;) {
            lane_prepare<Model>(L, M, yt);
            if (L.heavy != HV_FROZEN) break;
            GLC_UNROLL_RK
            for (int i = 0; i < NY; i++) rate[i] = 0.0;
            lane_consume<Model>(L, M, yt, rate, GLC_INT_NONE);
        }
    }
    drain_tally(L, tot);
    if (L.heavy == HV_NONE) {
        if (A.slotUnit) A.slotUnit[slotHeld] = -1;
        return;
    }
    A.slotL[slotHeld] = L;
    GLC_UNROLL_RK
    for (int i = 0; i < NY; i++) A.slotYt[(int64_t)slotHeld * NY + i] = yt[i];
    if (A.slotUnit) A.slotUnit[slotHeld] = kUnitRhsBegin;  // (a free slot that fetched a node is now an occupied one)
}

#if defined(__CUDACC__)
template <class Model>
__global__ void __launch_bounds__(GLC_BLOCK, GLC_MIN_BLOCKS) drain_kernel(KernelArgs A) {
    LaneMem M{&A, A.ws, 1};
    LaneState L;
    lane_reset(L);
    double yt[NY];
    bool fresh = false;
    int slotHeld = -1;
    unsigned int tot[7] = {0, 0, 0, 0, 0, 0, 0};
    const bool mayTake = A.drainLanes <= 0 || (int)(threadIdx.x & 31) < A.drainLanes;
    glc_vote_init(A.drainBlockSync);
    for (int it = 0; it < A.budget; ++it) {
        const bool active = drain_iterate<Model>(L, M, A, yt, fresh, slotHeld, tot, mayTake);
        if (A.drainBlockSync) {
            if (!__syncthreads_or(active ? 1 : 0)) break;  // block-uniform: every thread takes part in every barrier
        } else if (!__any_sync(0xffffffffu, active))
            break;
    }
    drain_park<Model>(L, M, A, yt, fresh, slotHeld, tot);
#pragma unroll
    for (int k = 0; k < 7; k++) {
        unsigned int v = tot[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&A.counters[k], (unsigned long long)v);
    }
}

template <class Model>
__global__ void __launch_bounds__(GLC_BLOCK, GLC_MIN_BLOCKS) evolve_kernel(KernelArgs A) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    LaneMem M{&A, A.ws + slot, A.nslots};
    LaneState L;
    if (A.resume)
        L = A.lanes[slot];
    else
        lane_reset(L);
    L.nAcc = L.nRej = L.nRhs = L.nSeg = L.nTrialFail = L.nNodes = L.nDone = 0;
    if (L.phase == PH_IDLE) L.phase = PH_FETCH;  // the queue may have grown since the last slice
    glc_vote_init(0);

    for (int it = 0; it < A.budget; ++it) {
        // All 32 lanes stay in this loop until the whole warp is out of work; the vote is a reconvergence
        // point in front of the heavy call.
        const bool active = lane_iterate<Model>(L, M);
        if (!__any_sync(0xffffffffu, active)) break;
    }
    A.lanes[slot] = L;

    // ---- counters: warp-reduce then one atomic per warp per counter
    unsigned int vals[8] = {L.nAcc, L.nRej, L.nRhs, L.nSeg, L.nTrialFail, L.nNodes, L.nDone,
                            (L.phase != PH_IDLE && L.phase != PH_FETCH) ? 1u : 0u};
#pragma unroll
    for (int k = 0; k < 8; k++) {
        unsigned int v = vals[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&A.counters[k], (unsigned long long)v);
    }
}
#endif

}  // namespace glc
