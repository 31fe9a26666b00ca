// glc_evolve_kernel.cuh -- the batched adaptive Cash-Karp solver.
//
// One thread integrates one node; a thread that finishes its node pulls the next one from a
// global queue (atomic counter), so divergence in step counts between nodes is absorbed at
// step granularity ("compaction by refill") instead of idling lanes until the slowest node of
// a warp is done.  One iteration of the main loop = one attempt of gsl_odeiv2_evolve_apply's
// try_step: the 5 intermediate Cash-Karp stages + the dydt_out evaluation, all through ONE
// call site of the model's rate function (the RHS is by far the largest piece of code).
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

namespace glc {

// Cash-Karp tableau (Cash & Karp 1990). Row s = weights of k1..k6 used to build the input of
// stage s (s=1..5 -> k2..k6; s=6 -> 5th-order solution); row 0 = error weights (5th-4th order).
__constant__ double c_rk_b[7][6] = {
    {37.0 / 378.0 - 2825.0 / 27648.0, 0.0, 250.0 / 621.0 - 18575.0 / 48384.0,
     125.0 / 594.0 - 13525.0 / 55296.0, -277.0 / 14336.0, 512.0 / 1771.0 - 0.25},
    {1.0 / 5.0, 0, 0, 0, 0, 0},
    {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0, 0},
    {0.3, -0.9, 1.2, 0, 0, 0},
    {-11.0 / 54.0, 2.5, -70.0 / 27.0, 35.0 / 27.0, 0, 0},
    {1631.0 / 55296.0, 175.0 / 512.0, 575.0 / 13824.0, 44275.0 / 110592.0, 253.0 / 4096.0, 0},
    {37.0 / 378.0, 0.0, 250.0 / 621.0, 125.0 / 594.0, 0.0, 512.0 / 1771.0}};
__constant__ double c_rk_a[7] = {0.0, 1.0 / 5.0, 0.3, 3.0 / 5.0, 1.0, 7.0 / 8.0, 1.0};

enum Phase : int { PH_FETCH = 0, PH_SEGMENT, PH_TRIAL, PH_STEP, PH_SOLVE_DONE, PH_IDLE };

constexpr int kTrialCountMaximum = 8;  // standard.F90:135
constexpr int kSegmentGuard = 64;

__device__ __forceinline__ bool prop_is_non_negative(int prop) {
    return prop != GLC_P_SAT_BOUND_MASS;  // isNonNegative attributes of the component definitions
}

template <class Model>
__global__ void __launch_bounds__(128) evolve_kernel(KernelArgs A) {
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t S = A.nslots;
    double *ws = A.ws + slot;
    auto W = [&](int vec, int comp) -> double & { return ws[((int64_t)vec * NY + comp) * S]; };
    auto AR = [&](int prop, int node) -> double & { return A.props[(int64_t)prop * A.cap + node]; };

    const double epsAbs = c_params.odeToleranceAbsolute;
    const double epsRel = c_params.odeToleranceRelative;

    int phase = PH_FETCH;
    int node = -1;
    NodeCtx ctx;
    uint32_t mask = 0;
    double tEnd = 0, x = 0, x1 = 0, h = 1.0, t0 = 0, h0 = 0, timeStartSaved = 0, timeStepIn = -1;
    double timeInterruptFirst = 0;
    int interruptFound = 0, interruptCode = 0;
    int count = 0, inApply = 0, outWritten = 0, yslot = 0, kslot = 0, trial = 0, finalStep = 0;
    int segmentsThisNode = 0, nodeStatus = GLC_STATUS_SUCCESS, solveFailed = 0;
    unsigned int nAcc = 0, nRej = 0, nRhs = 0, nSeg = 0, nTrialFail = 0, nNodes = 0;

    for (;;) {
        // All 32 lanes stay in this loop until the whole warp is out of work, and the vote below is a
        // reconvergence point: with independent thread scheduling the lanes would otherwise drift apart
        // after the first divergent `continue` and execute serially for the rest of the kernel
        // (measured: 2.0 active threads per warp instruction without it).
        if (__all_sync(0xffffffffu, phase == PH_IDLE)) break;
        // ------------------------------------------------------------------ fetch a node
        if (phase == PH_FETCH) {
            node = atomicAdd(A.work_counter, 1);
            if (node >= A.n) {
                phase = PH_IDLE;
                continue;
            }
            nNodes++;
            ctx.flags = A.flags[node];
            tEnd = A.time_end[node];
            segmentsThisNode = 0;
            nodeStatus = GLC_STATUS_SUCCESS;
            phase = PH_SEGMENT;
        }
        // ------------------------------------------------- standardEvolve prologue (:434-576)
        if (phase == PH_SEGMENT) {
            double y[NY], s[NY];
#pragma unroll
            for (int i = 0; i < NY; i++) {
                y[i] = AR(i, node);
                s[i] = 0.0;
            }
            ctx.massTarget = AR(GLC_P_MASS_TARGET, node);
            ctx.massRate = AR(GLC_P_MASS_RATE, node);
            ctx.timeTarget = AR(GLC_P_TIME_TARGET, node);
            ctx.scaleTarget = AR(GLC_P_DMSCALE_TARGET, node);
            ctx.scaleRate = AR(GLC_P_DMSCALE_RATE, node);
            ctx.spinTarget = AR(GLC_P_SPIN_TARGET, node);
            ctx.spinRate = AR(GLC_P_SPIN_RATE, node);
            ctx.timeLastIsolated = AR(GLC_P_TIME_LAST_ISOLATED, node);
            ctx.diskRadius = AR(GLC_P_DISK_RADIUS, node);
            ctx.diskVelocity = AR(GLC_P_DISK_VELOCITY, node);
            ctx.sphRadius = AR(GLC_P_SPH_RADIUS, node);
            ctx.sphVelocity = AR(GLC_P_SPH_VELOCITY, node);
            ctx.basicMass = AR(GLC_P_BASIC_MASS, node);
            ctx.dmScale = AR(GLC_P_DMSCALE, node);
            ctx.spinJ = AR(GLC_P_SPIN, node);
            ctx.massBaryonicSubhalos = AR(GLC_P_MASS_BARYONIC_SUBHALOS, node);
            ctx.numericsFailed = 0;
            timeStartSaved = AR(GLC_P_TIME, node);
            ctx.timeNode = timeStartSaved;
            timeStepIn = AR(GLC_P_TIME_STEP, node);
            nSeg++;
            segmentsThisNode++;
            {
                // pre-evolve hooks edit the node itself (they run before the solver's saved copy is taken,
                // standard.F90:434-441 vs :518-527), so their edits must survive a solver restart
                const int flagsBefore = ctx.flags;
                Model::pre_evolve(ctx, y);
                if (ctx.flags != flagsBefore) {
#pragma unroll
                    for (int i = 0; i < NY; i++) AR(i, node) = y[i];
                    AR(GLC_P_BASIC_MASS, node) = ctx.basicMass;
                    A.flags[node] = ctx.flags;
                }
            }
            mask = Model::active_mask(ctx.flags);
            Model::scales(ctx, y, s);
#pragma unroll
            for (int i = 0; i < NY; i++) {
                W(WS_YA, i) = y[i];
                W(WS_SCALE, i) = s[i];
            }
            yslot = 0;
            interruptFound = 0;
            timeInterruptFirst = 0.0;
            interruptCode = GLC_INT_NONE;
            trial = 0;
            solveFailed = 0;
            x = timeStartSaved;
            if (timeStartSaved != tEnd && mask != 0u)
                phase = PH_TRIAL;
            else
                phase = PH_SOLVE_DONE;
        }
        // ------------------------------ trial start (:587-652) + odeSolverSolve prologue
        if (phase == PH_TRIAL) {
            if (trial > 0) {
#pragma unroll
                for (int i = 0; i < NY; i++) W(WS_YA, i) = AR(i, node);
            }
            yslot = 0;
            double stepSize = c_params.reuseODEStepSize ? timeStepIn / dm_scale2(1.0, trial) : -1.0;
            x = timeStartSaved;
            x1 = tEnd;
            double xStep = x1 - x;
            if (stepSize > 0.0) xStep = fmin(stepSize, xStep);
            count = 0;  // GSL_ODEIV2_Driver_Reset
            if (xStep != 0.0) h = xStep;
            inApply = 0;
            outWritten = 0;
            phase = PH_STEP;
        }
        // --------------------------------------------- one try_step of evolve_apply
        if (phase == PH_STEP) {
            int needK1 = 0;
            if (!inApply) {
                t0 = x;
                h0 = h;
                if (count > 0 && outWritten) kslot ^= 1;  // dydt_in := dydt_out
                outWritten = 0;
                needK1 = (count == 0);
            }
            {
                const double dt = x1 - t0;
                if (h0 > dt) {
                    h0 = dt;
                    finalStep = 1;
                } else {
                    finalStep = 0;
                }
            }
            const int ySrc = WS_YA + yslot, yDst = WS_YA + (yslot ^ 1);
            const int k1v = WS_KA + kslot, kOutv = WS_KA + (kslot ^ 1);
            double rmax = DBL_MIN;
            int forbiddenNegatives = 0;
            int aborted = 0;
#pragma unroll 1
            for (int stage = needK1 ? 0 : 1; stage <= 6; ++stage) {
                double yt[NY], rate[NY];
                const double ts = t0 + c_rk_a[stage] * h0;
                // ---- stage input
                if (stage == 0) {
#pragma unroll
                    for (int i = 0; i < NY; i++) yt[i] = W(ySrc, i);
                } else if (stage < 6) {
#pragma unroll
                    for (int i = 0; i < NY; i++) {
                        if (stage == 1) {
                            yt[i] = W(ySrc, i) + c_rk_b[1][0] * h0 * W(k1v, i);  // rkck.c: y + b21*h*k1
                        } else {
                            double acc = c_rk_b[stage][0] * W(k1v, i);
                            for (int j = 1; j < stage; j++) acc += c_rk_b[stage][j] * W(WS_K2 + j - 1, i);
                            yt[i] = W(ySrc, i) + h0 * acc;
                        }
                    }
                } else {
                    // 5th-order solution, error estimate and the controller's rmax (cscal2.c:109-126;
                    // a_dydt = 0 so rmax does not depend on dydt_out)
#pragma unroll
                    for (int i = 0; i < NY; i++) {
                        const double k1 = W(k1v, i), k3 = W(WS_K3, i), k4 = W(WS_K4, i),
                                     k5 = W(WS_K5, i), k6 = W(WS_K6, i);
                        const double d = c_rk_b[6][0] * k1 + c_rk_b[6][2] * k3 + c_rk_b[6][3] * k4 +
                                         c_rk_b[6][5] * k6;
                        const double ynew = W(ySrc, i) + h0 * d;
                        const double yerr = h0 * (c_rk_b[0][0] * k1 + c_rk_b[0][2] * k3 +
                                                  c_rk_b[0][3] * k4 + c_rk_b[0][4] * k5 +
                                                  c_rk_b[0][5] * k6);
                        yt[i] = ynew;
                        W(yDst, i) = ynew;
                        if (mask & (1u << i)) {
                            const double D0 = epsRel * fabs(ynew) + epsAbs * W(WS_SCALE, i);
                            const double r = fabs(yerr) / fabs(D0);
                            rmax = fmax(r, rmax);
                            if (c_params.enforceNonNegativity && prop_is_non_negative(i) && ynew < 0.0)
                                forbiddenNegatives = 1;
                        }
                    }
                }
                // ---- standardODEs
#pragma unroll
                for (int i = 0; i < NY; i++) rate[i] = 0.0;
                Model::solve_analytics(ctx, ts);
                int code = GLC_INT_NONE;
                int ebadfunc = 0;
                nRhs++;  // every call of the derivatives function counts (also the frozen ones past an interrupt)
                if (interruptFound && ts >= timeInterruptFirst) {
                    Model::solve_analytics(ctx, timeInterruptFirst);
                } else {
                    code = Model::rates(ctx, ts, yt, rate);
                    if (code != GLC_INT_NONE) {
#pragma unroll
                        for (int i = 0; i < NY; i++) rate[i] = 0.0;
                        if (ts < timeInterruptFirst || !interruptFound) {
                            interruptFound = 1;
                            timeInterruptFirst = ts;
                            interruptCode = code;
                            ebadfunc = 1;
                        }
                    }
                }
                {
                    const int kv = (stage == 0) ? k1v : ((stage == 6) ? kOutv : WS_K2 + stage - 1);
#pragma unroll
                    for (int i = 0; i < NY; i++) {
                        const double r = (mask & (1u << i)) ? rate[i] : 0.0;
                        if (!isfinite(r)) nodeStatus = GLC_STATUS_NONFINITE;
                        W(kv, i) = r;
                    }
                    if (stage == 6) outWritten = 1;
                }
                if (ebadfunc) {
                    aborted = 1;
                    break;
                }
            }
            GTR("attempt t0=%.17g h0=%.17g t1=%.17g final=%d count=%d aborted=%d rmax=%.17g\n", t0, h0, x1, finalStep, count, aborted, rmax);
            if (aborted) {
                // odeSolverInterrupt, solver.F90:608-618.  NB (GSL quirk, reproduced): gsl_odeiv2_evolve_apply
                // leaves *t at the end of a REJECTED attempt when the retry returns early, so x may be ahead
                // of t0 here; if it is beyond the interrupt time the solver restarts from the initial state.
                x1 = timeInterruptFirst;
                inApply = 0;
                if (x > x1) {
#pragma unroll
                    for (int i = 0; i < NY; i++) W(WS_YA, i) = AR(i, node);
                    yslot = 0;
                    x = timeStartSaved;
                    count = 0;  // GSL_ODEIV2_Driver_Reset
                    outWritten = 0;
                }
                if (!(x < x1)) phase = PH_SOLVE_DONE;
                continue;
            }
            count++;
            const double tNew = finalStep ? x1 : t0 + h0;
            x = tNew;  // evolve.c sets *t before the controller runs and does not restore it on a retry
            // ---- sc2_control_hadjust (ord = 5)
            const double hOld = h0;
            int dec = 0;
            if (rmax > 1.1) {
                double r = 0.9 / dm_pow(rmax, 1.0 / 5.0);
                if (r < 0.2) r = 0.2;
                h0 = r * hOld;
                dec = 1;
            } else if (forbiddenNegatives) {
                h0 = 0.5 * hOld;
                dec = 1;
            } else if (rmax < 0.5) {
                double r = 0.9 / dm_pow(rmax, 1.0 / 6.0);
                if (r > 4.9) r = 4.9;
                if (r < 1.0) r = 1.0;
                h0 = r * hOld;
            }
            GTR("  dec=%d h_old=%.17g h_new=%.17g\n", dec, hOld, h0);
            if (dec) {
                const double tNext = tNew + h0;
                if (fabs(h0) < fabs(hOld) && tNext != tNew) {
                    nRej++;
                    inApply = 1;  // y := y0 is implicit (yslot unchanged)
                    continue;
                }
                // GSL_FAILURE: step-size underflow
                h = h0;
                inApply = 0;
                solveFailed = 1;
                yslot ^= 1;  // as in GSL, y holds the failed step's result
                phase = PH_SOLVE_DONE;
                continue;
            }
            // ---- accepted
            if (!finalStep) h = h0;
            yslot ^= 1;
            inApply = 0;
            nAcc++;
            {
                // standardPostStepProcessing
                double y[NY];
                const int yv = WS_YA + yslot;
#pragma unroll
                for (int i = 0; i < NY; i++) y[i] = W(yv, i);
                Model::solve_analytics(ctx, x);
                const int st = Model::post_step(ctx, y);
                if (st != kGslSuccess) {
#pragma unroll
                    for (int i = 0; i < NY; i++) W(yv, i) = y[i];
                }
                if (st != kGslSuccess && st != kGslContinue) count = 0;  // gsl_odeiv2_evolve_reset
            }
            if (!(x < x1)) phase = PH_SOLVE_DONE;
        }
        // -------------------------------- trial epilogue + standardEvolve epilogue (:657-753)
        if (phase == PH_SOLVE_DONE) {
            double y[NY];
            const int yv = WS_YA + yslot;
#pragma unroll
            for (int i = 0; i < NY; i++) y[i] = W(yv, i);
            if (solveFailed) {
                int rescued = 0;
                if (c_params.enforceNonNegativity) {
#pragma unroll
                    for (int i = 0; i < NY; i++)
                        if ((mask & (1u << i)) && prop_is_non_negative(i) && y[i] < 0.0) {
                            y[i] = 0.0;
                            rescued = 1;
                        }
                }
                if (rescued) {
                    h = timeStepIn / dm_scale2(1.0, trial);
                    solveFailed = 0;
                } else {
                    trial++;
                    nTrialFail++;
                    solveFailed = 0;
                    if (trial < kTrialCountMaximum) {
                        phase = PH_TRIAL;
                        continue;
                    }
                    // errorStatusUnderflow: node left at its saved values
                    A.status[node] = GLC_STATUS_UNDERFLOW;
                    A.interrupt[node] = GLC_INT_NONE;
                    phase = PH_FETCH;
                    continue;
                }
            }
            Model::solve_analytics(ctx, tEnd);
            double timeOut, timeStepOut;
            int interrupted = 0;
            if (timeInterruptFirst != 0.0) {
                interrupted = 1;
                timeOut = timeInterruptFirst;
                timeStepOut = -1.0;
            } else {
                timeOut = tEnd;
                timeStepOut = (timeStartSaved != tEnd && mask != 0u) ? h : -1.0;
            }
            ctx.timeNode = timeOut;
            Model::post_evolve(ctx, y);
            if (ctx.numericsFailed) nodeStatus = GLC_STATUS_NONFINITE;
            int code = interrupted ? interruptCode : GLC_INT_NONE;
            if (interrupted && c_params.resolveInterruptsOnDevice) {
                // functionInterrupt: <class>CreateByInterrupt / blackHoleCreate
                if (code == GLC_INT_HOTHALO_CREATE) ctx.flags |= GLC_F_HAS_HOTHALO;
                if (code == GLC_INT_DISK_CREATE) ctx.flags |= GLC_F_HAS_DISK;
                if (code == GLC_INT_SPHEROID_CREATE) ctx.flags |= GLC_F_HAS_SPHEROID;
                if (code == GLC_INT_BH_CREATE) {
                    ctx.flags |= GLC_F_HAS_BH;
                    y[GLC_P_BH_MASS] = c_params.bhSeedMass;
                    y[GLC_P_BH_SPIN] = c_params.bhSeedSpin;
                }
                code = GLC_INT_NONE;
            }
#pragma unroll
            for (int i = 0; i < NY; i++) AR(i, node) = y[i];
            AR(GLC_P_TIME, node) = timeOut;
            AR(GLC_P_TIME_STEP, node) = timeStepOut;
            AR(GLC_P_DISK_RADIUS, node) = ctx.diskRadius;
            AR(GLC_P_DISK_VELOCITY, node) = ctx.diskVelocity;
            AR(GLC_P_SPH_RADIUS, node) = ctx.sphRadius;
            AR(GLC_P_SPH_VELOCITY, node) = ctx.sphVelocity;
            AR(GLC_P_BASIC_MASS, node) = ctx.basicMass;
            AR(GLC_P_DMSCALE, node) = ctx.dmScale;
            AR(GLC_P_SPIN, node) = ctx.spinJ;
            A.flags[node] = ctx.flags;
            if (interrupted && code == GLC_INT_NONE && timeOut < tEnd) {
                if (segmentsThisNode < kSegmentGuard) {
                    phase = PH_SEGMENT;  // host loop evolver/standard.F90:425-476, resolved in place
                    continue;
                }
                nodeStatus = GLC_STATUS_FAIL;
            }
            A.status[node] = nodeStatus;
            A.interrupt[node] = code;
            phase = PH_FETCH;
        }
    }

    // ---- counters: warp-reduce then one atomic per warp per counter
    unsigned int vals[6] = {nAcc, nRej, nRhs, nSeg, nTrialFail, nNodes};
#pragma unroll
    for (int k = 0; k < 6; k++) {
        unsigned int v = vals[k];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&A.counters[k], (unsigned long long)v);
    }
}

}  // namespace glc
