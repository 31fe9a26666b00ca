// glc_common.cuh -- device-side types shared by the kernels of libglcb200.
//
// Data layout in HBM:
//   arena props : double [GLC_NPROP][capacity]   SoA, one column per node
//   arena flags : int32  [capacity]
//   workspace   : double [WS_NVEC][GLC_NY][nslots]  per-resident-thread RK stage vectors;
//                 slot = global thread id, so every access is a coalesced 256-B line per warp
#pragma once

#include <stdint.h>

// ---- platform layer.  The product is compiled by nvcc for sm_100a.  The SAME headers can also be compiled by
// g++ into tests/emu (TEST INFRASTRUCTURE: lets the CPU-only test suite execute the kernels' per-lane logic
// bit-for-bit against the reference restatement without a GPU); nothing in the product path ever loads that build.
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define GLC_DEVICE_INLINE static __device__ __forceinline__
#define GLC_DEVICE_METHOD __device__ __forceinline__
#define GLC_DEVICE_NOINLINE __device__ __noinline__
#define GLC_LDG(p) __ldg(p)
#define GLC_PARAMS c_params
#define GLC_TABLES c_tables
#define glc_atomic_add(p, v) atomicAdd((p), (v))
// Vote level of the running kernel (block-uniform, set by glc_vote_init at the top of every kernel that calls the
// warp-synchronous rate function).  0/1: the data-dependent loops vote per warp.  2: the rate function also puts a block
// barrier between its phases.  3: the fixed-point loop of the structure solve votes per BLOCK (one barrier per pass).
// The point: the warps of a block walk through an evaluation in step and share the instruction stream (drain_kernel
// only -- it is instruction-fetch bound, profiles/r02ac; every thread of the block must then call the rate function the
// same number of times).  (Every loop voting per block, Brent and QAG included, was slower than no block vote at all.)
__shared__ int s_glcVoteLevel;
static __device__ __forceinline__ void glc_vote_init(int level) {
    if (threadIdx.x == 0) s_glcVoteLevel = level;
    __syncthreads();
}
static __device__ __forceinline__ bool glc_any_outer(bool p) {
    return s_glcVoteLevel >= 3 ? __syncthreads_or(p ? 1 : 0) != 0 : __any_sync(0xffffffffu, p) != 0;
}
#define GLC_ANY(p) __any_sync(0xffffffffu, (p))
#define GLC_ANY_OUTER(p) glc_any_outer(p)
#define GLC_PHASE_SYNC() do { if (s_glcVoteLevel >= 2) __syncthreads(); } while (0)
#define GLC_COUNT(k) ((void)0)
#define GLC_SYNCWARP() __syncwarp()
// smallest positive double: positive IEEE doubles order like their bit patterns
#define glc_atomic_min_positive_double(p, v) atomicMin((p), (unsigned long long)__double_as_longlong(v))
#ifndef GLC_BLOCK
#define GLC_BLOCK 128
#endif
#ifndef GLC_MIN_BLOCKS
#define GLC_MIN_BLOCKS 2
#endif
#ifndef GLC_MTHREADS
#define GLC_MTHREADS 256  // threads per block of the micro-task machine (one block per SM): 8 warps x 255 registers beat
                          // 16 warps x 128 registers by 8 % on the bench workload
#endif
#ifndef GLC_MSLOTS
#define GLC_MSLOTS 2048   // slots per block = regrouping domain
#endif
#else
#include <math.h>
#include <string.h>
#include <algorithm>
#define GLC_DEVICE_INLINE static inline
#define GLC_DEVICE_METHOD inline
#define GLC_DEVICE_NOINLINE static __attribute__((noinline))
#define GLC_LDG(p) (*(p))
#define GLC_PARAMS c_params
#define GLC_TABLES c_tables
#define __constant__ static
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define GLC_ANY(p) (p)
#define GLC_ANY_OUTER(p) (p)
#define GLC_PHASE_SYNC() ((void)0)
#ifdef GLC_EMU_COUNTERS
static long long g_emu_count[8];
#define GLC_COUNT(k) (g_emu_count[k]++)
#else
#define GLC_COUNT(k) ((void)0)
#endif
#define GLC_SYNCWARP() ((void)0)
template <class T>
static inline T glc_atomic_add(T *p, T v) {
    T o = *p;
    *p = o + v;
    return o;
}
static inline void glc_atomic_min_positive_double(unsigned long long *p, double v) {
    unsigned long long b;
    memcpy(&b, &v, sizeof b);
    if (b < *p) *p = b;
}
using std::max;
using std::min;
#endif

#include "../../include/glc_b200.h"
#include "glc_detmath.h"
#include "glc_specfun.h"

namespace glc {

// ---- physical constants: source/numerical/constants/*.F90 with GSL-2.6 MKSA values ----
constexpr double kPi = 3.14159265358979323846;
constexpr double kGravitationalConstant = 6.673e-11;  // GSL_CONST_MKSA_GRAVITATIONAL_CONSTANT
constexpr double kParsec = 3.08567758135e16;          // GSL_CONST_MKSA_PARSEC
constexpr double kMassSolar = 1.98892e30;             // GSL_CONST_MKSA_SOLAR_MASS
constexpr double kBoltzmann = 1.3806504e-23;          // GSL_CONST_MKSA_BOLTZMANN
constexpr double kAtomicMassUnit = 1.660538782e-27;   // GSL_CONST_MKSA_UNIFIED_ATOMIC_MASS
constexpr double kSpeedLight = 2.99792458e8;          // GSL_CONST_MKSA_SPEED_OF_LIGHT
constexpr double kThomsonCrossSection = 6.65245893699e-29;
constexpr double kKilo = 1.0e3, kMega = 1.0e6, kGiga = 1.0e9, kHecto = 1.0e2, kErgs = 1.0e-7;
constexpr double kMegaParsec = kMega * kParsec;
constexpr double kYear = 3.15581497635456e7;
constexpr double kGigaYear = kGiga * kYear;
constexpr double kGInternal = kGravitationalConstant * kMassSolar / (kKilo * kKilo) / kMegaParsec;
constexpr double kMpcPerKmPerSToGyr = kMegaParsec / kKilo / kGigaYear;
constexpr double kHydrogenByMassSolar = 0.7070, kHeliumByMassSolar = 0.2740;
constexpr double kMetallicitySolar = 0.0188;
constexpr double kHydrogenByMassPrimordial = 0.7514, kHeliumByMassPrimordial = 0.2486;
constexpr double kAtomicMassHydrogen = 1.0078250322, kAtomicMassHelium = 4.0026032545;
constexpr double kMassHydrogenAtom = kAtomicMassHydrogen * kAtomicMassUnit;
constexpr double kMeanAtomicMassPrimordial =
    1.0 / (2.0 * kHydrogenByMassPrimordial / kAtomicMassHydrogen +
           3.0 * kHeliumByMassPrimordial / kAtomicMassHelium);
constexpr double kFeedbackEnergyInputAtInfinityCanonical = 4.517e5;

// GSL status codes on the path
constexpr int kGslSuccess = 0, kGslFailure = -1, kGslContinue = -2;

constexpr int NY = GLC_NY;
constexpr int NPROP = GLC_NPROP;

// workspace vector ids
enum : int { WS_YA = 0, WS_YB, WS_KA, WS_KB, WS_K2, WS_K3, WS_K4, WS_K5, WS_K6, WS_SCALE, WS_NVEC };

struct DeviceTable2D {
    int n0, n1;
    const double *x0, *x1, *v;  // device pointers; v[n0][n1]
};

struct DeviceTables {
    // CIE tables pre-processed as cieFileReadFile does (CIE_file.F90:627-659)
    DeviceTable2D cooling;   // x0 = ln Z (or -999), x1 = ln T, v = ln Lambda when cooling_log
    int cooling_log, cooling_first_z_zero;
    double cooling_first_nonzero_z, cooling_z_min, cooling_z_max, cooling_t_min, cooling_t_max;
    DeviceTable2D electron;
    int electron_log, electron_first_z_zero;
    double electron_first_nonzero_z, electron_z_min, electron_z_max, electron_t_min, electron_t_max;
    // halo mean density, uniform in ln t
    DeviceTable2D density;   // x0 = ln t, v[n0][2] = {rho_mean, d rho_mean/dt}
    double density_lnt0, density_inv_dlnt;
    // exponential-disk rotation-curve factor, uniform in ln(half-radius)
    DeviceTable2D diskrc;
    double diskrc_lnx0, diskrc_inv_dlnx;
    // ADAF tabulations (accretion_disks/ADAF.F90:394-447), uniform in ln(1-j): x0 = ln(1-j), v[n0][2] = {jet power per
    // unit accretion rate, spin-up ratio}
    DeviceTable2D adaf;
    double adaf_inv_dlnx;
    // fastExponentiator tables (math/exponentiation.F90:57-104), tabulated once like the reference does:
    // x^adiabaticOmega on [1e-3,1] (adiabatic_Gnedin2004.F90:664-687) and x^0.33 on [1,1000] (Krumholz2009.F90)
    const double *powAc, *powKmt;
    int powAcN, powKmtN;
    double powAcDx, powAcInvDx, powKmtDx, powKmtInvDx;  // lattice spacing and its reciprocal, computed once on the host
    // inverse tabulation of the scale-free NFW specific angular momentum (NFW.F90:589-639; numerical/tabulations_inverse.F90):
    // nfwJx[k] = 2^(k/30) on the octave lattice, nfwJv[k] = sqrt(4 pi m(x_k) x_k), built once on the host
    const double *nfwJx, *nfwJv;
    int nfwJN;
    // constructor-time constants of accretionDisksSwitched (switched.F90:259-297 takes these logarithms at every call; they
    // are pure functions of the parameters): ln(accretionRateThinDiskMinimum), ln(accretionRateThinDiskMaximum)
    double lnThinDiskMin, lnThinDiskMax;
    // mergerTreeEvolveProfilerSimple: bin edges of the step-size histogram (Make_Range logarithmic, simple.F90:147) and the
    // device accumulators [time_step_count | evaluation_count | ..._interrupted | ..._interrupted][bins], then
    // property_hits[NY], hits "unknown", smallest step (bits of a positive double); null = profiling off
    double profEdges[GLC_PROFILE_BINS];
    int profBins;
    unsigned long long *profile;
};
constexpr int kProfHits = 4 * GLC_PROFILE_BINS, kProfUnknown = kProfHits + GLC_NY, kProfSmallest = kProfUnknown + 1,
              kProfWords = kProfSmallest + 1;

struct LaneState;

struct KernelArgs {
    double *props;        // [NPROP][cap]
    int32_t *flags;       // [cap]
    const double *time_end;  // [cap]
    int32_t *status;      // [cap]
    int32_t *interrupt;   // [cap]
    int64_t cap;
    int n;
    double *ws;           // [WS_NVEC][NY][nslots]
    int64_t nslots;
    int *work_counter;             // queue cursor (persists across time slices)
    unsigned long long *counters;  // 8 x u64: glc_counters fields, nodes done, lanes parked mid-node
    const int32_t *order;          // queue order (node ids sorted by component set) or null = identity
    struct LaneState *lanes;       // [nslots] parked lane states
    int resume;                    // 0: first slice of a batch (lanes start empty), 1: resume parked lanes
    int budget;                    // heavy calls per lane in this slice
    volatile int *debug;           // GLC_DEBUG_HANG builds: host-mapped per-warp progress words (else null)
    // drain hand-over (machine -> drain_kernel): see glc_api.cu launch_machine
    int hold;                      // machine: slots that reach an RK boundary (U_RHS_BEGIN) are held, not re-queued
    const int32_t *held;           // drain: list of held slot ids
    int nheld;
    int *held_counter;             // drain: cursor into `held`
    struct LaneState *slotL;       // drain: the machine's per-slot lane states and stage inputs
    double *slotYt;                // [nslots][NY]
    int *slotUnit;                 // drain: pending-unit words (a finished slot is marked -1)
    int drainLanes;                // drain: lanes per warp that take nodes (0 or >= 32: all of them).  A warp serialises the divergent
                                   // evaluations of its lanes, so a node advances fastest alone in its warp (1: lone-lane speed, ~0.2 ms
                                   // per evaluation against ~0.7 ms with 32 nodes per warp); the host spreads the nodes of a pass
                                   // over all resident warps: lanes = ceil(nodes / resident warps)
    int drainBlockSync;            // drain: 1 = the warps of a block start every evaluation together (__syncthreads_or at the top of
                                   // the loop).  The dense pass is bound by instruction fetch (profiles/r02ac: no_instruction 12.8
                                   // stall cycles per issue, 340 KB of SASS); warps that walk the rate function side by side share
                                   // what one of them has fetched
    int drainRefill;               // drain, streaming sessions: a lane whose node is done fetches the next one from the node
                                   // queue into the same slot (list entries with kHeldFresh set are free slots that start
                                   // with a fetch)
    // GLC_LEDGER builds (debug): node-ownership ledger and per-slot execution flags, see glc_machine.cuh
    int *ledger;                   // [n] -1 = never fetched, s+1 = held by slot s, -2 = written back
    int *slotBusy;                 // [nslots] 1 while a lane executes a unit of the slot
    unsigned long long *ledgerErr; // [8] violation counters
};

// Per-node context kept in registers while a node is resident in a thread.
struct NodeCtx {
    int flags;
    // linear-in-time interpolants of the analytically solved properties
    double massTarget, massRate, timeTarget;
    double scaleTarget, scaleRate, spinTarget, spinRate;
    double timeLastIsolated;
    // their current values (refreshed by solve_analytics at every RHS call)
    double basicMass, dmScale, spinJ;
    // structure-solver state (warm start in, solution out)
    double diskRadius, diskVelocity, sphRadius, sphVelocity;
    double timeNode;              // basic%time() of the node outside the ODE solve (scale-set, pre/post-evolve hooks)
    double massBaryonicSubhalos;  // frozen input, see GLC_P_MASS_BARYONIC_SUBHALOS
    int numericsFailed;           // a nested solver (Brent/QAG) failed where the reference would abort
};

// drain hand-over list: bit 30 of an entry marks a FREE slot handed to the drain kernel to fetch queued nodes into
constexpr int32_t kHeldFresh = 1 << 30;
// pending-unit word of a slot parked at an RK boundary (== U_RHS_BEGIN of glc_machine.cuh, checked there)
constexpr int kUnitRhsBegin = 1;

// single translation unit (glc_api.cu): defined here
__constant__ glc_params c_params;
__constant__ DeviceTables c_tables;

}  // namespace glc
