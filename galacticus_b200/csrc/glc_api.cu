// glc_api.cu -- C-ABI of libglcb200 (include/glc_b200.h): context, tables, arena, launches.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <time.h>
#include <unistd.h>

#include "glc_common.cuh"
#include "glc_tables_host.h"
#include "host/glc_forest.hpp"

#include "glc_evolve_kernel.cuh"
#include "glc_model_box.cuh"
#include "glc_model_standard.cuh"
#include "glc_machine.cuh"

using namespace glc;

namespace {

constexpr int kBlock = GLC_BLOCK;

struct HostTable {
    int n0 = 0, n1 = 0;
    double *d_x0 = nullptr, *d_x1 = nullptr, *d_v = nullptr;
};

}  // namespace

struct glc_evolver {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    glc_params params{};
    bool params_set = false;
    DeviceTables tables{};
    HostTable host_tables[GLC_NTABLES];
    glcf::HaloTable halo_host;  // host copy of GLC_TABLE_HALO_MEAN_DENSITY for the forest scheduler
    // arena
    int64_t cap = 0;
    double *d_props = nullptr;
    int32_t *d_flags = nullptr, *d_status = nullptr, *d_interrupt = nullptr;
    double *d_time_end = nullptr;
    double *d_stage = nullptr;  // node-major staging [cap][NPROP] for transposes
    double *d_dydt = nullptr;   // [cap][NY] for glc_rhs_batch
    double *d_snap_props = nullptr;  // snapshot of the arena (glc_arena_snapshot)
    int32_t *d_snap_flags = nullptr;
    int64_t snap_cap = 0;
    int64_t launches = 0;
    // workspace
    double *d_ws = nullptr;
    int64_t nslots = 0;
    int *d_work = nullptr;
    unsigned long long *d_counters = nullptr;
    double *d_pow_ac = nullptr, *d_pow_kmt = nullptr;  // fastExponentiator tables
    double *d_nfw_jx = nullptr, *d_nfw_jv = nullptr;   // inverse tabulation of the NFW specific angular momentum
    unsigned long long *d_profile = nullptr;           // mergerTreeEvolveProfilerSimple accumulators (profileOdeEvolver)
    double pow_ac_exponent = 0.0;
    LaneState *d_lanes = nullptr;   // parked lane states, one per resident lane
    SlotArrays d_slots{};           // micro-task machine: per-slot continuations, split by access group
    int64_t nslots_machine = 0;
    int32_t use_machine = 2;        // standard model: 1 = micro-task machine, 0 = warp-synchronous evolve_kernel, 2 = by batch size
    int64_t machine_min_nodes = -1; // use_machine == 2: batches smaller than this (default: drain_threshold) go to the
                                    // warp-synchronous evolve_kernel -- too few nodes to fill the machine's per-unit queues, it would
                                    // hand over to the drain after one slice on a few SMs anyway; results are identical either way
    int32_t drain_handover = 1;     // run-to-completion mode: finish the last nodes with drain_kernel
    int64_t drain_threshold = 120000; // hand over when fewer slots than this are still in flight (measured: profiles/r01f_knobs.txt)
    int32_t drain_dense_budget = 0x3fffffff; // evaluations per lane in a dense drain pass.  Unbounded since round 2: once the list is
                                       // exhausted a long node is left alone in its warp and runs at lone-lane speed, which beats parking it
                                       // after 1024 evaluations and re-listing it (10^6-node pass 1941 -> 1808 ms, profiles/r02w)
    int32_t hybrid_budget = 4096;      // pops per warp of one internal machine slice in run-to-completion mode
    int32_t *d_held = nullptr;
    float *d_held_score = nullptr;
    int64_t held_cap = 0;
    int32_t drain_express = 0;      // 1 = first drain pass: the predicted-longest nodes one per warp on stream2 beside a dense launch on the
                                    // other block per SM.  Off since the GK15 pass is shared by lanes: the express launch is over after 121 ms
                                    // (profiles/r02aq: the nodes that end the pass are not the predicted ones) and meanwhile halves the lanes of
                                    // the dense launch; 10^6-node pass 1588-1642 ms without it, 1623-1715 with it (gpu_r2bc.sh, alternating)
    float drain_age_weight = 0.0f;  // express selection: score = predicted remaining steps (0) or 6 x that + weight x evaluations so far
    int32_t drain_block_sync = 2;   // vote level of dense drain / lane passes (glc_common.cuh): the warps of a block start every evaluation
                                    // together and meet again between the phases of the rate function -- the kernel is instruction-fetch
                                    // bound (profiles/r02ac): dense drain 1105 -> 1000 ms at level 2 (profiles/r02ag); 1 and 3 gain less
    int32_t drain_lanes_max = 32;   // most nodes per warp in a drain pass (fewer: less divergence per warp, more passes over the list)
    int32_t drain_spread = 1;       // drain / lane passes: spread the nodes over all resident warps (KernelArgs::drainLanes); 0 = one per
                                    // warp when they fit, else 32 per warp (round-2 behaviour before the measurement in profiles/r02k)
    cudaStream_t stream2 = nullptr;
    // streaming session (glc_stream_*)
    bool stream_active = false, stream_started = false;
    bool stream_lane_mode = false;  // streaming, adaptive ticks: every occupied slot stands at an RK boundary (lane passes)
    int64_t stream_live = 0;        // occupied slots after the last tick
    int64_t stream_n = 0;           // tickets handed out so far = length of the node queue
    unsigned char *d_collected = nullptr;
    int64_t *d_collect_list = nullptr;
    int32_t *d_collect_meta = nullptr;   // [cap][3] flags, status, interrupt of the collected nodes
    int32_t *h_collect_meta = nullptr;   // pinned
    int64_t collect_meta_cap = 0;
    int32_t stream_sparse_budget = 32, stream_dense_budget = 12;  // evaluations per lane in one lane pass of a streaming tick
    int32_t stream_spread = 2;           // lane passes with more nodes than warps: 0 = 32 nodes per warp on as few blocks as needed, 1 = spread
                                         // over all resident warps, 2 = over one block per SM (a second block per SM only when 32 per warp do not
                                         // suffice).  2 since the GK15 pass is shared by lanes (thin warps got cheaper): 1000-tree forest 20.6-21.3 ->
                                         // 19.9 s, volume forest 16.2 -> 14.7 s (gpu_r2aw.sh); 1 is as slow as 0 (8 warps per SM on 8 code paths)
    int32_t stream_sort = 0;             // lane passes with more nodes than warps: list sorted by kind of node (component set)
    int32_t stream_express = 0;          // lane passes with more nodes than warps: nodes whose score (evaluations spent so far +
                                         // 6 x predicted remaining steps) reaches this get a warp each; 0 = off
    int32_t stream_express_budget = 48;  // evaluations of an express warp per tick
    int32_t stream_priority_express = 0; // > 0: that many of the highest-PRIORITY nodes (glc_forest_evolve: halo mass, i.e. the main
                                         // branches every tree waits for) get a warp each in lane passes with more nodes than warps.
                                         // Measured (profiles/r02k): 200 = no change, 592 / 1000 = 18 % slower; off
    float *d_stream_priority = nullptr;  // [cap] per ticket
    int64_t stream_priority_cap = 0;
    std::vector<int32_t> h_held, h_ordered;
    std::vector<float> h_score;
    std::vector<int> h_idx;
    int32_t stream_machine_budget = 4096; // pops per warp of a machine slice used as a streaming tick
    int64_t stream_machine_above = -1;   // adaptive ticks: machine slice when queued + occupied >= this (default: drain_threshold)
    // tick statistics of the session (forest log): machine slices / lane passes, their device-side wall time, nodes in flight
    int64_t tick_machine = 0, tick_lane = 0, tick_hold = 0;
    double tick_machine_s = 0.0, tick_lane_s = 0.0, tick_live_sum = 0.0, tick_lanes_sum = 0.0, tick_express_sum = 0.0;
    int32_t l2_persist = 0;              // GLC_L2_PERSIST=1: pin the machine's RootState array in L2 (experiment)
    bool l2_window_set = false;
    int32_t forest_schedule = 1;         // glc_forest_evolve: 1 = asynchronous groups over the streaming machine, 0 = bulk-synchronous rounds
    int64_t collected_cap = 0;
    int64_t stream_collected = 0;
    int32_t *d_order = nullptr;     // queue order (component-sorted node ids)
    int *d_sort = nullptr;          // 2 x 64 bucket counters
    int64_t order_cap = 0;
    int32_t slice_budget = 0;       // heavy calls per lane per launch; 0 = run to completion in one launch
    int32_t sort_queue = 1;
    int32_t slice_log = 0;
    int32_t max_slices = 0;
    int64_t slices = 0;
    float last_ms = 0.f;
    // phases of the last hybrid machine batch (glc_last_phase_stats): device ms and rate-function evaluations of the machine slices and
    // of the drain passes
    cudaEvent_t ev_mid = nullptr;
    double phase_ms[2] = {0.0, 0.0}, phase_rhs[2] = {0.0, 0.0}, phase_steps[2] = {0.0, 0.0}, phase_nodes[2] = {0.0, 0.0};
    bool phase_split = false;
    std::string err;
#ifdef GLC_LEDGER
    // debug build: node-ownership ledger (see glc_evolve_kernel.cuh GLC_LEDGER_*)
    int *d_ledger = nullptr, *d_slot_busy = nullptr;
    unsigned long long *d_ledger_err = nullptr;
    int64_t ledger_cap = 0;
#endif
};

// page-locked host staging (std::vector interface): H2D / D2H copies of the forest batches run at PCIe/NVLink-C2C speed
// instead of going through the driver's pageable bounce buffers
template <class T>
struct PinnedAllocator {
    typedef T value_type;
    PinnedAllocator() = default;
    template <class U>
    PinnedAllocator(const PinnedAllocator<U> &) {}
    T *allocate(size_t n) {
        void *p = nullptr;
        if (cudaHostAlloc(&p, n * sizeof(T), cudaHostAllocDefault) != cudaSuccess) throw std::bad_alloc();
        return static_cast<T *>(p);
    }
    void deallocate(T *p, size_t) { cudaFreeHost(p); }
    template <class U>
    bool operator==(const PinnedAllocator<U> &) const { return true; }
    template <class U>
    bool operator!=(const PinnedAllocator<U> &) const { return false; }
};

static double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

#define GLC_CHECK(ev, call)                                                                  \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            (ev)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                 \
            return -(int)e__ - 1000;                                                         \
        }                                                                                    \
    } while (0)

// ------------------------------------------------------------------ layout transposes
// host/staging layout is node-major [n][NPROP]; the arena is SoA [NPROP][cap].
__global__ void aos_to_soa_kernel(const double *__restrict__ aos, double *__restrict__ soa, int n,
                                  int64_t cap) {
    // (callers pass aos/soa already offset to the first node of the range)
    __shared__ double tile[64 * NPROP];
    const int node0 = blockIdx.x * 64;
    const int nn = min(64, n - node0);
    for (int i = threadIdx.x; i < nn * NPROP; i += blockDim.x) tile[i] = aos[(int64_t)node0 * NPROP + i];
    __syncthreads();
    for (int i = threadIdx.x; i < nn * NPROP; i += blockDim.x) {
        const int p = i / nn, k = i % nn;
        soa[(int64_t)p * cap + node0 + k] = tile[k * NPROP + p];
    }
}

__global__ void soa_to_aos_kernel(const double *__restrict__ soa, double *__restrict__ aos, int n,
                                  int64_t cap) {
    __shared__ double tile[64 * NPROP];
    const int node0 = blockIdx.x * 64;
    const int nn = min(64, n - node0);
    for (int i = threadIdx.x; i < nn * NPROP; i += blockDim.x) {
        const int p = i / nn, k = i % nn;
        tile[k * NPROP + p] = soa[(int64_t)p * cap + node0 + k];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nn * NPROP; i += blockDim.x) aos[(int64_t)node0 * NPROP + i] = tile[i];
}

// one RHS evaluation per node (unit-level parity tests)
template <class Model>
__global__ void rhs_kernel(KernelArgs A, double *dydt) {
    // the rate function is warp-synchronous: out-of-range lanes go through it with on = false
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = gid < A.n;
    const int node = on ? gid : A.n - 1;
    glc_vote_init(0);
    auto AR = [&](int prop) -> double & { return A.props[(int64_t)prop * A.cap + node]; };
    NodeCtx ctx;
    double y[NY], rate[NY];
    for (int i = 0; i < NY; i++) {
        y[i] = AR(i);
        rate[i] = 0.0;
    }
    ctx.flags = A.flags[node];
    ctx.massTarget = AR(GLC_P_MASS_TARGET);
    ctx.massRate = AR(GLC_P_MASS_RATE);
    ctx.timeTarget = AR(GLC_P_TIME_TARGET);
    ctx.scaleTarget = AR(GLC_P_DMSCALE_TARGET);
    ctx.scaleRate = AR(GLC_P_DMSCALE_RATE);
    ctx.spinTarget = AR(GLC_P_SPIN_TARGET);
    ctx.spinRate = AR(GLC_P_SPIN_RATE);
    ctx.timeLastIsolated = AR(GLC_P_TIME_LAST_ISOLATED);
    ctx.diskRadius = AR(GLC_P_DISK_RADIUS);
    ctx.diskVelocity = AR(GLC_P_DISK_VELOCITY);
    ctx.sphRadius = AR(GLC_P_SPH_RADIUS);
    ctx.sphVelocity = AR(GLC_P_SPH_VELOCITY);
    ctx.basicMass = AR(GLC_P_BASIC_MASS);
    ctx.dmScale = AR(GLC_P_DMSCALE);
    ctx.spinJ = AR(GLC_P_SPIN);
    ctx.massBaryonicSubhalos = AR(GLC_P_MASS_BARYONIC_SUBHALOS);
    ctx.numericsFailed = 0;
    const double time = AR(GLC_P_TIME);
    ctx.timeNode = time;
    Model::solve_analytics(ctx, time);
    const int code = Model::rates(ctx, time, y, rate, false, on);
    if (!on) return;
    for (int i = 0; i < NY; i++) dydt[(int64_t)node * NY + i] = rate[i];
    A.interrupt[node] = code;
    AR(GLC_P_DISK_RADIUS) = ctx.diskRadius;
    AR(GLC_P_DISK_VELOCITY) = ctx.diskVelocity;
    AR(GLC_P_SPH_RADIUS) = ctx.sphRadius;
    AR(GLC_P_SPH_VELOCITY) = ctx.sphVelocity;
    AR(GLC_P_BASIC_MASS) = ctx.basicMass;
}

// standardErrorHandler's table for ONE node (glc_error_report_node): one warp, lane 0 carries the node (the rate function is
// warp-synchronous).  State = the record after the pre-evolve hooks; dy/dt = the evaluation at the node's time (k1); yError =
// the embedded error of one Cash-Karp step of size h (rkck.c: h * sum ec_i k_i), with the stage inputs built as
// lane_prepare builds them.  out: [7][NY] = y, dydt, scale, tolerance, error, error_scaled, active; then the interrupt code.
template <class Model>
__global__ void error_report_kernel(KernelArgs A, double h, double *out) {
    const bool on = threadIdx.x == 0;
    glc_vote_init(0);
    auto AR = [&](int prop) -> double & { return A.props[(int64_t)prop * A.cap]; };
    NodeCtx ctx;
    double y0[NY], yt[NY], rate[NY], s[NY], k[6][NY];
    for (int i = 0; i < NY; i++) {
        y0[i] = AR(i);
        s[i] = 0.0;
    }
    ctx.flags = A.flags[0];
    ctx.massTarget = AR(GLC_P_MASS_TARGET);
    ctx.massRate = AR(GLC_P_MASS_RATE);
    ctx.timeTarget = AR(GLC_P_TIME_TARGET);
    ctx.scaleTarget = AR(GLC_P_DMSCALE_TARGET);
    ctx.scaleRate = AR(GLC_P_DMSCALE_RATE);
    ctx.spinTarget = AR(GLC_P_SPIN_TARGET);
    ctx.spinRate = AR(GLC_P_SPIN_RATE);
    ctx.timeLastIsolated = AR(GLC_P_TIME_LAST_ISOLATED);
    ctx.diskRadius = AR(GLC_P_DISK_RADIUS);
    ctx.diskVelocity = AR(GLC_P_DISK_VELOCITY);
    ctx.sphRadius = AR(GLC_P_SPH_RADIUS);
    ctx.sphVelocity = AR(GLC_P_SPH_VELOCITY);
    ctx.basicMass = AR(GLC_P_BASIC_MASS);
    ctx.dmScale = AR(GLC_P_DMSCALE);
    ctx.spinJ = AR(GLC_P_SPIN);
    ctx.massBaryonicSubhalos = AR(GLC_P_MASS_BARYONIC_SUBHALOS);
    ctx.numericsFailed = 0;
    const double t0 = AR(GLC_P_TIME);
    ctx.timeNode = t0;
    Model::pre_evolve(ctx, y0);
    const uint32_t mask = Model::active_mask(ctx.flags);
    Model::scales(ctx, y0, s);
    int codeFirst = GLC_INT_NONE;
#pragma unroll 1
    for (int stage = 0; stage < 6; stage++) {
        const double ts = t0 + c_rk_a[stage] * h;
        if (stage == 0) {
            for (int i = 0; i < NY; i++) yt[i] = y0[i];
        } else if (stage == 1) {
            const double b10 = c_rk_b[1][0];
            for (int i = 0; i < NY; i++) yt[i] = y0[i] + b10 * h * k[0][i];
        } else {
            for (int i = 0; i < NY; i++) {
                double acc = c_rk_b[stage][0] * k[0][i];
                for (int j = 1; j < stage; j++) acc += c_rk_b[stage][j] * k[j][i];
                yt[i] = y0[i] + h * acc;
            }
        }
        for (int i = 0; i < NY; i++) rate[i] = 0.0;
        Model::solve_analytics(ctx, ts);
        const int code = Model::rates(ctx, ts, yt, rate, false, on);
        if (stage == 0) codeFirst = code;
        for (int i = 0; i < NY; i++) k[stage][i] = ((mask & (1u << i)) && code == GLC_INT_NONE) ? rate[i] : 0.0;
    }
    if (!on) return;
    const double epsAbs = GLC_PARAMS.odeToleranceAbsolute, epsRel = GLC_PARAMS.odeToleranceRelative;
    for (int i = 0; i < NY; i++) {
        const bool active = (mask & (1u << i)) != 0;
        const double yerr = h * (c_rk_b[0][0] * k[0][i] + c_rk_b[0][2] * k[2][i] + c_rk_b[0][3] * k[3][i] + c_rk_b[0][4] * k[4][i] +
                                 c_rk_b[0][5] * k[5][i]);
        const double tol = epsRel * fabs(y0[i]) + epsAbs * s[i];
        out[0 * NY + i] = active ? y0[i] : 0.0;
        out[1 * NY + i] = active ? k[0][i] : 0.0;
        out[2 * NY + i] = active ? s[i] : 0.0;
        out[3 * NY + i] = active ? tol : 0.0;
        out[4 * NY + i] = active ? yerr : 0.0;
        out[5 * NY + i] = active ? fabs(yerr) / tol : 0.0;
        out[6 * NY + i] = active ? 1.0 : 0.0;
    }
    out[7 * NY] = (double)codeFirst;
}

// FP64 FMA-chain microbenchmark (16 independent chains per thread): the measured FP64 roofline denominator
__global__ void fp64_peak_kernel(double *out, int iters) {
    double a[16];
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = threadIdx.x * 1.0e-9 + k;
    const double b = 1.0000001, c = 1.0e-7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) a[k] = fma(a[k], b, c);
    }
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < 16; k++) sum += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

__global__ void histogram_kernel(const double *__restrict__ v, int n, double lo, double hi, int nb,
                                 double *hist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = v[i];
    if (!(x > 0.0)) return;
    const double lx = log10(x);
    const int b = (int)floor((lx - lo) / (hi - lo) * nb);
    if (b >= 0 && b < nb) atomicAdd(&hist[b], 1.0);
}

// systemClockMaximum (node_evolver/standard.F90:694-705,861-867): nodes that were not finished when the wall-clock budget of
// the batched call ran out come back with errorStatusXCPU
__global__ void mark_xcpu_kernel(int32_t *status, int32_t *interrupt, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && status[i] == GLC_STATUS_PENDING) {
        status[i] = GLC_STATUS_XCPU;
        interrupt[i] = GLC_INT_NONE;
    }
}

// ---------------------------------------------------------------------------- queue order
// Nodes are handed to lanes in an order sorted by component set, so that the lanes of a warp -- which
// fetch consecutive queue entries -- mostly run the same branches of the rate function.  Bucket order:
// nodes that still lack a hot halo first (they go through component-creation segments and tend to be
// the longest), then from the richest component set to the poorest.
__global__ void queue_hist_kernel(const int32_t *__restrict__ flags, int n, int *hist) {
    __shared__ int sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = 0;
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&sh[queue_bucket(flags[i])], 1);
    __syncthreads();
    if (threadIdx.x < 64 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}
__global__ void queue_scan_kernel(int *hist) {  // hist[0..63] counts -> hist[64..127] exclusive offsets
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = 0; b < 64; b++) {
            hist[64 + b] = acc;
            acc += hist[b];
        }
    }
}
__global__ void queue_scatter_kernel(const int32_t *__restrict__ flags, int n, int *hist, int32_t *order) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        order[atomicAdd(&hist[64 + queue_bucket(flags[i])], 1)] = i;
}

// ---------------------------------------------------------------------------- helpers
static int ensure_workspace(glc_evolver *ev, int grid) {
    const int64_t need = (int64_t)grid * kBlock;
    if (need <= ev->nslots) return 0;
    if (ev->d_ws) cudaFree(ev->d_ws);
    if (ev->d_lanes) cudaFree(ev->d_lanes);
    ev->d_ws = nullptr;
    ev->d_lanes = nullptr;
    GLC_CHECK(ev, cudaMalloc(&ev->d_ws, sizeof(double) * WS_NVEC * NY * need));
    GLC_CHECK(ev, cudaMalloc(&ev->d_lanes, sizeof(LaneState) * need));
    ev->nslots = need;
    return 0;
}

static int build_queue_order(glc_evolver *ev, int n) {
    if (!ev->sort_queue || n < 2 * kBlock) return 0;
    if (n > ev->order_cap) {
        cudaFree(ev->d_order);
        ev->d_order = nullptr;
        GLC_CHECK(ev, cudaMalloc(&ev->d_order, sizeof(int32_t) * (size_t)n));
        ev->order_cap = n;
    }
    if (!ev->d_sort) GLC_CHECK(ev, cudaMalloc(&ev->d_sort, sizeof(int) * 128));
    GLC_CHECK(ev, cudaMemsetAsync(ev->d_sort, 0, sizeof(int) * 128, ev->stream));
    const int grid = std::min((n + 255) / 256, ev->num_sms * 8);
    queue_hist_kernel<<<grid, 256, 0, ev->stream>>>(ev->d_flags, n, ev->d_sort);
    queue_scan_kernel<<<1, 32, 0, ev->stream>>>(ev->d_sort);
    queue_scatter_kernel<<<grid, 256, 0, ev->stream>>>(ev->d_flags, n, ev->d_sort, ev->d_order);
    ev->launches += 3;
    GLC_CHECK(ev, cudaGetLastError());
    return 1;
}

// One batch = one or more time slices of the persistent evolve kernel.
template <class Model>
static int launch_evolve(glc_evolver *ev, int n, unsigned long long *hc) {
    int blocksPerSm = 0;
    GLC_CHECK(ev, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, evolve_kernel<Model>,
                                                                 kBlock, 0));
    if (blocksPerSm < 1) blocksPerSm = 1;
    int grid = ev->num_sms * blocksPerSm;
    grid = std::min(grid, (n + kBlock - 1) / kBlock);
    if (grid < 1) grid = 1;
    int rc = ensure_workspace(ev, ev->num_sms * blocksPerSm);
    if (rc) return rc;
    GLC_CHECK(ev, cudaEventRecord(ev->ev0, ev->stream));
    const int sorted = build_queue_order(ev, n);
    if (sorted < 0) return sorted;
    KernelArgs A{};
    A.props = ev->d_props;
    A.flags = ev->d_flags;
    A.time_end = ev->d_time_end;
    A.status = ev->d_status;
    A.interrupt = ev->d_interrupt;
    A.cap = ev->cap;
    A.n = n;
    A.ws = ev->d_ws;
    A.nslots = ev->nslots;
    A.work_counter = ev->d_work;
    A.counters = ev->d_counters;
    A.order = sorted ? ev->d_order : nullptr;
    A.lanes = ev->d_lanes;
    A.resume = 0;
    const double wallMax = ev->params.wallClockMaximumSeconds;
    const int budget = ev->slice_budget > 0 ? ev->slice_budget : (wallMax > 0.0 ? 256 : 0);  // the guard needs time slices
    A.budget = budget > 0 ? budget : 0x7fffffff;
    GLC_CHECK(ev, cudaMemsetAsync(ev->d_work, 0, sizeof(int), ev->stream));
    GLC_CHECK(ev, cudaMemsetAsync(ev->d_counters, 0, sizeof(unsigned long long) * 16, ev->stream));
    if (wallMax > 0.0) GLC_CHECK(ev, cudaMemsetAsync(ev->d_status, 0x80, sizeof(int32_t) * (size_t)n, ev->stream));
    const double t_start = now_s();
    int nslice = 0;
    for (;;) {
        evolve_kernel<Model><<<grid, kBlock, 0, ev->stream>>>(A);
        ev->launches++;
        ev->slices++;
        GLC_CHECK(ev, cudaGetLastError());
        GLC_CHECK(ev, cudaMemcpyAsync(hc, ev->d_counters, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost,
                                      ev->stream));
        if (budget <= 0) break;
        GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
        if (wallMax > 0.0 && now_s() - t_start > wallMax && hc[6] < (unsigned long long)n) {
            mark_xcpu_kernel<<<(n + 255) / 256, 256, 0, ev->stream>>>(ev->d_status, ev->d_interrupt, n);
            ev->launches++;
            break;
        }
        if (ev->slice_log)
            fprintf(stderr, "[glc slice %lld] t=%.3f ms done=%llu/%d rhs=%llu accepted=%llu parked=%llu\n",
                    (long long)ev->slices, 1e3 * (now_s() - t_start), hc[6], n, hc[2], hc[0], hc[7]);
        GLC_CHECK(ev, cudaMemsetAsync(ev->d_counters + 7, 0, sizeof(unsigned long long), ev->stream));
        if (hc[6] >= (unsigned long long)n) break;
        if (ev->max_slices > 0 && ++nslice >= ev->max_slices) break;  // profiling aid: leaves the batch unfinished
        A.resume = 1;
    }
    GLC_CHECK(ev, cudaEventRecord(ev->ev1, ev->stream));
    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
    return 0;
}

static void free_slots(glc_evolver *ev) {
    cudaFree(ev->d_slots.L);
    cudaFree(ev->d_slots.R);
    cudaFree(ev->d_slots.root);
    cudaFree(ev->d_slots.yt);
    cudaFree(ev->d_slots.Q);
    cudaFree(ev->d_slots.unit);
    ev->d_slots = SlotArrays{};
    ev->nslots_machine = 0;
}

#ifdef GLC_LEDGER
static int ledger_setup(glc_evolver *ev, int n, KernelArgs &A, bool fresh) {
    if (n > ev->ledger_cap) {
        cudaFree(ev->d_ledger);
        ev->d_ledger = nullptr;
        GLC_CHECK(ev, cudaMalloc(&ev->d_ledger, sizeof(int) * (size_t)n));
        ev->ledger_cap = n;
    }
    if (!ev->d_slot_busy) GLC_CHECK(ev, cudaMalloc(&ev->d_slot_busy, sizeof(int) * (size_t)ev->nslots_machine));
    if (!ev->d_ledger_err) GLC_CHECK(ev, cudaMalloc(&ev->d_ledger_err, sizeof(unsigned long long) * 8));
    if (fresh) {
        GLC_CHECK(ev, cudaMemsetAsync(ev->d_ledger, 0xff, sizeof(int) * (size_t)n, ev->stream));
        GLC_CHECK(ev, cudaMemsetAsync(ev->d_slot_busy, 0, sizeof(int) * (size_t)ev->nslots_machine, ev->stream));
        GLC_CHECK(ev, cudaMemsetAsync(ev->d_ledger_err, 0, sizeof(unsigned long long) * 8, ev->stream));
    }
    A.ledger = ev->d_ledger;
    A.slotBusy = ev->d_slot_busy;
    A.ledgerErr = ev->d_ledger_err;
    return 0;
}
// prints what the ledger knows: violation counters, nodes never fetched, nodes fetched but never written back and the
// state of the slots that took them
static int ledger_report(glc_evolver *ev, int n, const char *tag) {
    unsigned long long herr[8];
    std::vector<int> led((size_t)n);
    int work = 0;
    GLC_CHECK(ev, cudaMemcpy(herr, ev->d_ledger_err, sizeof herr, cudaMemcpyDeviceToHost));
    GLC_CHECK(ev, cudaMemcpy(led.data(), ev->d_ledger, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost));
    GLC_CHECK(ev, cudaMemcpy(&work, ev->d_work, sizeof(int), cudaMemcpyDeviceToHost));
    long long never = 0, held = 0, done = 0;
    std::vector<int> heldNodes;
    for (int i = 0; i < n; i++) {
        if (led[i] == -1) never++;
        else if (led[i] == -2) done++;
        else { held++; if (heldNodes.size() < 24) heldNodes.push_back(i); }
    }
    fprintf(stderr, "[glc ledger %s] n=%d work_counter=%d never-fetched=%lld held=%lld done=%lld | violations: double-fetch=%llu "
                    "foreign-writeback=%llu two-lanes-in-slot=%llu wrong-queue=%llu\n",
            tag, n, work, never, held, done, herr[0], herr[1], herr[2], herr[3]);
    FILE *dump = nullptr;
    if (const char *path = getenv("GLC_LEDGER_DUMP")) dump = fopen(path, "wb");
    if (dump) {
        const int sizes[8] = {(int)heldNodes.size(), (int)sizeof(LaneState), (int)sizeof(RhsState), (int)sizeof(RootState),
                              (int)sizeof(QagState), NY, WS_NVEC * NY, NPROP};
        fwrite(sizes, sizeof(int), 8, dump);
    }
    for (int node : heldNodes) {
        const int slot = led[node] - 1;
        LaneState L;
        RhsState R;
        RootState root;
        QagState Q;
        double yt[NY], ws[WS_NVEC * NY], rec[NPROP], tEnd = 0.0;
        int unit = -99, flags = 0;
        cudaMemcpy(&L, ev->d_slots.L + slot, sizeof(LaneState), cudaMemcpyDeviceToHost);
        cudaMemcpy(&R, ev->d_slots.R + slot, sizeof(RhsState), cudaMemcpyDeviceToHost);
        cudaMemcpy(&root, ev->d_slots.root + slot, sizeof(RootState), cudaMemcpyDeviceToHost);
        cudaMemcpy(&Q, ev->d_slots.Q + slot, sizeof(QagState), cudaMemcpyDeviceToHost);
        cudaMemcpy(yt, ev->d_slots.yt + (int64_t)slot * NY, sizeof yt, cudaMemcpyDeviceToHost);
        cudaMemcpy(ws, ev->d_ws + (int64_t)slot * (WS_NVEC * NY), sizeof ws, cudaMemcpyDeviceToHost);
        cudaMemcpy(&unit, ev->d_slots.unit + slot, sizeof(int), cudaMemcpyDeviceToHost);
        cudaMemcpy(&flags, ev->d_flags + node, sizeof(int), cudaMemcpyDeviceToHost);
        cudaMemcpy(&tEnd, ev->d_time_end + node, sizeof(double), cudaMemcpyDeviceToHost);
        for (int p = 0; p < NPROP; p++) cudaMemcpy(&rec[p], ev->d_props + (int64_t)p * ev->cap + node, sizeof(double), cudaMemcpyDeviceToHost);
        fprintf(stderr, "   node %d held by slot %d (block %d): slot.unit=%d L.node=%d phase=%d heavy=%d stage=%d trial=%d x=%.9g x1=%.9g h=%.3g "
                        "tEnd=%.9g flags=0x%x | R.count=%d comp=%d active=%d fit=%.3g bad=%d | B.state=%d it=%d busy=%d x=%.6g xLow=%.6g xHigh=%.6g "
                        "fLow=%.3g fHigh=%.3g | Q.busy=%d size=%d it=%d\n",
                node, slot, slot / GLC_MSLOTS, unit, L.node, L.phase, L.heavy, L.stage, L.trial, L.x, L.x1, L.h, tEnd, flags, R.count, R.comp,
                R.active, R.fit, R.bad, root.B.state, root.B.iteration, root.B.busy, root.B.x, root.B.xLow, root.B.xHigh, root.B.fLow,
                root.B.fHigh, Q.busy, Q.size, Q.iteration);
        if (dump) {
            fwrite(&node, sizeof(int), 1, dump);
            fwrite(&slot, sizeof(int), 1, dump);
            fwrite(&unit, sizeof(int), 1, dump);
            fwrite(&flags, sizeof(int), 1, dump);
            fwrite(&tEnd, sizeof(double), 1, dump);
            fwrite(&L, sizeof L, 1, dump);
            fwrite(&R, sizeof R, 1, dump);
            fwrite(&root, sizeof root, 1, dump);
            fwrite(&Q, sizeof Q, 1, dump);
            fwrite(yt, sizeof yt, 1, dump);
            fwrite(ws, sizeof ws, 1, dump);
            fwrite(rec, sizeof rec, 1, dump);
        }
    }
    if (dump) fclose(dump);
    int shown = 0;
    for (int i = 0; i < n && shown < 16; i++)
        if (led[i] == -1) {
            int flags = 0;
            double t = 0.0, tEnd = 0.0;
            cudaMemcpy(&flags, ev->d_flags + i, sizeof(int), cudaMemcpyDeviceToHost);
            cudaMemcpy(&tEnd, ev->d_time_end + i, sizeof(double), cudaMemcpyDeviceToHost);
            cudaMemcpy(&t, ev->d_props + (int64_t)GLC_P_TIME * ev->cap + i, sizeof(double), cudaMemcpyDeviceToHost);
            fprintf(stderr, "   node %d never fetched: flags=0x%x time=%.9g tEnd=%.9g\n", i, flags, t, tEnd);
            shown++;
        }
    return 0;
}
#endif

// One batch on the micro-task machine (standard model): same slice protocol as launch_evolve.
// mode 0: one batch, fresh queue, run to completion (hybrid) or in user time slices
// mode 1: streaming -- ONE time slice of `streamBudget` pops over the (possibly grown) node queue
// mode 2: streaming -- continue to completion (hybrid)
static int launch_machine(glc_evolver *ev, int n, unsigned long long *hc, int mode = 0, int streamBudget = 0, int streamHold = 0) {
    size_t kMachineSmem = sizeof(unsigned int) * (size_t)U_IDLE * GLC_MSLOTS;
#if GLC_MACHINE_STAGED_POW
    kMachineSmem += sizeof(double) * (size_t)ev->tables.powAcN;  // the x^omega table staged behind the queues
    if (kMachineSmem > 227u * 1024u) {
        ev->err = "micro-task machine: unit queues + staged exponentiation table exceed 227 KB of shared memory";
        return -9;
    }
#endif
    GLC_CHECK(ev, cudaFuncSetAttribute(machine_kernel<GLC_MTHREADS, GLC_MSLOTS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kMachineSmem));
    const int gridMax = ev->num_sms;  // one block per SM
    int grid = std::min(gridMax, (n + GLC_MSLOTS - 1) / GLC_MSLOTS);
    if (grid < 1) grid = 1;
    if (mode != 0) grid = gridMax;  // the queue grows: every block takes part from the first slice on
    const bool fresh = mode == 0 || !ev->stream_started;
    const int64_t need = (int64_t)gridMax * GLC_MSLOTS;
    if (need > ev->nslots_machine) {
        free_slots(ev);
        GLC_CHECK(ev, cudaMalloc(&ev->d_slots.L, sizeof(LaneState) * need));
        GLC_CHECK(ev, cudaMalloc(&ev->d_slots.R, sizeof(RhsState) * need));
        GLC_CHECK(ev, cudaMalloc(&ev->d_slots.root, sizeof(RootState) * need));
        GLC_CHECK(ev, cudaMalloc(&ev->d_slots.yt, sizeof(double) * NY * need));
        GLC_CHECK(ev, cudaMalloc(&ev->d_slots.Q, sizeof(QagState) * need));
        GLC_CHECK(ev, cudaMalloc(&ev->d_slots.unit, sizeof(int) * need));
        ev->nslots_machine = need;
    }
    int rc = ensure_workspace(ev, (int)((need + kBlock - 1) / kBlock));
    if (rc) return rc;
    if (ev->l2_persist && !ev->l2_window_set) {
        // The root-find units (a quarter of the warp-state samples, 43 % of them waiting for memory) touch one 256-byte
        // RootState per slot: 78 MB for 303 104 slots, which fits the 126 MB L2 if nothing else evicts it.  Pin it.
        cudaDeviceProp prop;
        cudaGetDeviceProperties(&prop, ev->device);
        const size_t bytes = sizeof(RootState) * (size_t)ev->nslots_machine;
        const size_t setAside = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, bytes);
        if (setAside > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, setAside) == cudaSuccess) {
            cudaStreamAttrValue attr{};
            attr.accessPolicyWindow.base_ptr = ev->d_slots.root;
            attr.accessPolicyWindow.num_bytes = std::min<size_t>(bytes, (size_t)prop.accessPolicyMaxWindowSize);
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)setAside / (double)attr.accessPolicyWindow.num_bytes);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaStreamSetAttribute(ev->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
        }
        cudaGetLastError();
        ev->l2_window_set = true;
    }
    GLC_CHECK(ev, cudaEventRecord(ev->ev0, ev->stream));
    const int sorted = mode == 0 ? build_queue_order(ev, n) : 0;  // a growing queue is served in submission order
    if (sorted < 0) return sorted;
    KernelArgs A{};
    A.props = ev->d_props;
    A.flags = ev->d_flags;
    A.time_end = ev->d_time_end;
    A.status = ev->d_status;
    A.interrupt = ev->d_interrupt;
    A.cap = ev->cap;
    A.n = n;
    A.ws = ev->d_ws;
    A.nslots = ev->nslots;
    A.work_counter = ev->d_work;
    A.counters = ev->d_counters;
    A.order = sorted ? ev->d_order : nullptr;
    A.lanes = nullptr;
    A.resume = fresh ? 0 : 1;
    A.budget = ev->slice_budget > 0 ? ev->slice_budget : 0x7fffffff;
    if (mode == 1) A.budget = streamBudget > 0 ? streamBudget : 4096;
    A.debug = nullptr;
    A.ledger = nullptr;
    A.slotBusy = nullptr;
    A.ledgerErr = nullptr;
#ifdef GLC_LEDGER
    if (int rcl = ledger_setup(ev, n, A, fresh)) return rcl;
#endif
#ifdef GLC_DEBUG_HANG
    static int *h_dbg = nullptr;
    const int nDbg = gridMax * (GLC_MTHREADS / 32) * 8;
    if (!h_dbg) cudaHostAlloc(&h_dbg, sizeof(int) * nDbg, cudaHostAllocMapped);
    memset(h_dbg, 0, sizeof(int) * nDbg);
    {
        int *d_dbg = nullptr;
        cudaHostGetDevicePointer(&d_dbg, h_dbg, 0);
        A.debug = d_dbg;
    }
#endif
    if (fresh) {
        GLC_CHECK(ev, cudaMemsetAsync(ev->d_work, 0, sizeof(int), ev->stream));
        GLC_CHECK(ev, cudaMemsetAsync(ev->d_counters, 0, sizeof(unsigned long long) * 16, ev->stream));
        // status = GLC_STATUS_PENDING (0x80808080) until a node is written back
        if (mode == 0) GLC_CHECK(ev, cudaMemsetAsync(ev->d_status, 0x80, sizeof(int32_t) * (size_t)n, ev->stream));
    } else
        GLC_CHECK(ev, cudaMemsetAsync(ev->d_counters + 7, 0, sizeof(unsigned long long) * 3, ev->stream));
    if (mode != 0) ev->stream_started = true;
    const double t_start = now_s();
    int nslice = 0;
    ev->phase_split = false;
    A.hold = (mode == 1) ? streamHold : 0;
    A.held = nullptr;
    A.nheld = 0;
    A.held_counter = nullptr;
    A.slotL = nullptr;
    A.slotYt = nullptr;
    A.slotUnit = nullptr;
    A.drainLanes = 0;
    // Run-to-completion mode (no user time slices) is a hybrid: the machine works in internal slices while the node
    // queue still refills its slots and for as long as enough slots stay in flight to fill warps; then the slots
    // are brought to an RK boundary (hold slices) and handed to drain_kernel, which finishes their nodes with
    // whole evaluations.  With a user slice budget the machine alone runs (resumable by construction).
    const bool hybrid = mode != 1 && (ev->slice_budget <= 0 || mode == 2) && ev->drain_handover;
    if (hybrid) A.budget = ev->hybrid_budget;
    const unsigned long long drainBelow = (unsigned long long)ev->drain_threshold;
    bool draining = false;
    int stalled = 0;
    bool xcpu = false;
    unsigned long long prevDone = ~0ull, prevRhs = ~0ull, prevParked = ~0ull, unitsSinceProgress = 0;
    for (;;) {
        if (ev->slice_log > 1) fprintf(stderr, "[glc host] launching machine_kernel grid=%d budget=%d resume=%d hold=%d n=%d\n", grid, A.budget, A.resume, A.hold, n);
        machine_kernel<GLC_MTHREADS, GLC_MSLOTS><<<grid, GLC_MTHREADS, kMachineSmem, ev->stream>>>(A, ev->d_slots);
        ev->launches++;
        ev->slices++;
        GLC_CHECK(ev, cudaGetLastError());
        GLC_CHECK(ev, cudaMemcpyAsync(hc, ev->d_counters, sizeof(unsigned long long) * 11, cudaMemcpyDeviceToHost,
                                      ev->stream));
        if (mode != 1 && ev->slice_budget <= 0 && !hybrid) break;
        GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
        if (hc[10] != 0) {
            ev->err = "micro-task machine: a queue wait timed out inside the kernel (protocol violation)";
#ifdef GLC_LEDGER
            ledger_report(ev, n, "queue time-out");
#endif
            return GLC_ERR_STALLED;
        }
        if (mode == 1) {  // streaming: exactly one slice per call
            if (ev->slice_log)
                fprintf(stderr, "[glc stream slice] budget=%d hold=%d n=%d done=%llu fetched=%llu rhs=%llu parked=%llu mid-evaluation=%llu units=%llu\n",
                        A.budget, A.hold, n, hc[6], hc[5], hc[2], hc[7], hc[8], hc[9]);
            break;
        }
        if (mode == 0 && ev->params.wallClockMaximumSeconds > 0.0 && now_s() - t_start > ev->params.wallClockMaximumSeconds &&
            hc[6] < (unsigned long long)n) {
            mark_xcpu_kernel<<<(n + 255) / 256, 256, 0, ev->stream>>>(ev->d_status, ev->d_interrupt, n);
            ev->launches++;
            xcpu = true;
            break;
        }
        if (ev->slice_log)
            fprintf(stderr, "[glc slice %lld] t=%.3f ms done=%llu/%d fetched=%llu rhs=%llu accepted=%llu parked=%llu mid-evaluation=%llu units=%llu%s\n",
                    (long long)ev->slices, 1e3 * (now_s() - t_start), hc[6], n, hc[5], hc[2], hc[0], hc[7], hc[8], hc[9], draining ? " (hold)" : "");
        const unsigned long long parked = hc[7], midEvaluation = hc[8];
        GLC_CHECK(ev, cudaMemsetAsync(ev->d_counters + 7, 0, sizeof(unsigned long long) * 3, ev->stream));
        if (hc[6] >= (unsigned long long)n) break;
        if (ev->max_slices > 0 && ++nslice >= ev->max_slices) break;  // profiling aid: leaves the batch unfinished
        A.resume = 1;
        // No-progress guard (every mode): a slice that executed no unit at all (hc[9]) and changed no counter cannot be
        // followed by a better one.  Three in a row end the call with an error instead of repeating empty slices for ever (the reference finishes every tree or
        // reports: tasks/evolve_forests/_class.F90:887-897).
        if (hc[9] == 0 && hc[6] == prevDone && hc[2] == prevRhs && parked == prevParked) {
            if (++stalled >= 3) {
                char msg[256];
                snprintf(msg, sizeof msg,
                         "micro-task machine made no progress in 3 consecutive slices: %llu of %d nodes done, %llu slots occupied "
                         "(%llu mid-evaluation)", hc[6], n, parked, midEvaluation);
                ev->err = msg;
#ifdef GLC_LEDGER
                ledger_report(ev, n, "stalled");
#endif
                return GLC_ERR_STALLED;
            }
        } else
            stalled = 0;
        // ... and so does a long run of slices that execute units without finishing a single evaluation or node: the nested
        // solvers of one evaluation are bounded (100 structure iterations, 1000 Brent steps, 24 quadrature intervals)
        if (hc[6] == prevDone && hc[2] == prevRhs) {
            unitsSinceProgress += hc[9];
            if (unitsSinceProgress > 400000ull * std::max<unsigned long long>(parked, 1ull) && mode != 1) {
                char msg[256];
                snprintf(msg, sizeof msg,
                         "micro-task machine: %llu slots executed %llu units without finishing an evaluation "
                         "(%llu of %d nodes done)", parked, unitsSinceProgress, hc[6], n);
                ev->err = msg;
#ifdef GLC_LEDGER
                ledger_report(ev, n, "livelock");
#endif
                return GLC_ERR_STALLED;
            }
        } else
            unitsSinceProgress = 0;
        prevDone = hc[6];
        prevRhs = hc[2];
        prevParked = parked;
        if (hybrid) {
            // every node has been handed out once done + parked covers the batch
            const bool queueDry = hc[6] + parked >= (unsigned long long)n;
            if (!draining && queueDry && parked < drainBelow) {
                draining = true;
                A.hold = 1;
                A.budget = 512;
            }
            if (draining && midEvaluation == 0) {
                GLC_CHECK(ev, cudaEventRecord(ev->ev_mid, ev->stream));  // end of the machine phase
                ev->phase_split = true;
                ev->phase_rhs[0] = (double)hc[2];
                ev->phase_steps[0] = (double)hc[0];
                ev->phase_nodes[0] = (double)hc[6];
                // ---- hand the held slots to the drain kernel: dense passes (one node per lane, bounded number of
                // evaluations) while there are more nodes than warps, then one node per warp to the end
                if (!ev->d_held || ev->held_cap < ev->nslots_machine) {
                    cudaFree(ev->d_held);
                    ev->d_held = nullptr;
                    GLC_CHECK(ev, cudaMalloc(&ev->d_held, sizeof(int32_t) * (ev->nslots_machine + 8)));
                    cudaFree(ev->d_held_score);
                    ev->d_held_score = nullptr;
                    GLC_CHECK(ev, cudaMalloc(&ev->d_held_score, sizeof(float) * ev->nslots_machine));
                    ev->held_cap = ev->nslots_machine;
                }
                int *d_count = reinterpret_cast<int *>(ev->d_held + ev->nslots_machine);  // [0] list length, [1] cursor
                int bps = 0;
                GLC_CHECK(ev, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, drain_kernel<ModelStandard>, kBlock, 0));
                if (bps < 1) bps = 1;
                const int nslotsActive = grid * GLC_MSLOTS;
                const int warpsResident = ev->num_sms * bps * (kBlock / 32);
                A.slotL = ev->d_slots.L;
                A.slotYt = ev->d_slots.yt;
                A.slotUnit = ev->d_slots.unit;
                for (int pass = 0;; pass++) {
                    GLC_CHECK(ev, cudaMemsetAsync(d_count, 0, sizeof(int) * 4, ev->stream));
                    held_list_kernel<<<std::min((nslotsActive + 255) / 256, ev->num_sms * 8), 256, 0, ev->stream>>>(
                        ev->d_slots.unit, ev->d_slots.L, nslotsActive, ev->d_held, pass == 0 ? ev->d_held_score : nullptr, d_count, 0,
                        ev->drain_age_weight);
                    int nheld = 0;
                    GLC_CHECK(ev, cudaMemcpyAsync(&nheld, d_count, sizeof(int), cudaMemcpyDeviceToHost, ev->stream));
                    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
                    if (nheld == 0) break;
                    const int express = ev->num_sms * (kBlock / 32);  // one block per SM, one node per warp
                    if (pass == 0 && ev->drain_express && nheld > 2 * express) {
                        // ---- first pass: the nodes predicted to need the most steps are the critical path of the
                        // whole batch.  They get a warp each (one block per SM, stream2) right away, while the other
                        // block per SM works through the rest one node per lane.
                        std::vector<int32_t> h_held(nheld);
                        std::vector<float> h_score(nheld);
                        GLC_CHECK(ev, cudaMemcpyAsync(h_held.data(), ev->d_held, sizeof(int32_t) * nheld, cudaMemcpyDeviceToHost, ev->stream));
                        GLC_CHECK(ev, cudaMemcpyAsync(h_score.data(), ev->d_held_score, sizeof(float) * nheld, cudaMemcpyDeviceToHost, ev->stream));
                        GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
                        std::vector<int> idx(nheld);
                        for (int k = 0; k < nheld; k++) idx[k] = k;
                        std::nth_element(idx.begin(), idx.begin() + express, idx.end(),
                                         [&](int a, int b) { return h_score[a] > h_score[b]; });
                        std::vector<int32_t> ordered(nheld);
                        for (int k = 0; k < nheld; k++) ordered[k] = h_held[idx[k]];
                        GLC_CHECK(ev, cudaMemcpyAsync(ev->d_held, ordered.data(), sizeof(int32_t) * nheld, cudaMemcpyHostToDevice, ev->stream));
                        GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
                        KernelArgs X = A;  // express: the first `express` entries, one per warp, to completion
                        X.held = ev->d_held;
                        X.nheld = express;
                        X.held_counter = d_count + 1;
                        X.drainLanes = 1;
                        X.budget = 0x7fffffff;
                        drain_kernel<ModelStandard><<<ev->num_sms, kBlock, 0, ev->stream2>>>(X);
                        KernelArgs D = A;  // dense: the rest, bounded passes until fewer nodes than warps are left
                        D.held = ev->d_held + express;
                        D.nheld = nheld - express;
                        D.held_counter = d_count + 2;
                        D.drainLanes = 0;
                        D.drainBlockSync = ev->drain_block_sync;
                        D.budget = ev->drain_dense_budget;
                        drain_kernel<ModelStandard><<<ev->num_sms, kBlock, 0, ev->stream>>>(D);
                        ev->launches += 3;
                        GLC_CHECK(ev, cudaGetLastError());
                        GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
                        if (ev->slice_log)
                            fprintf(stderr, "[glc drain pass 0 dense+express] dense part done t=%.3f ms held=%d express=%d\n",
                                    1e3 * (now_s() - t_start), nheld, express);
                        // further dense passes over what the dense part left, while the express kernel is still running
                        for (int sub = 0; sub < 8; sub++) {
                            GLC_CHECK(ev, cudaMemsetAsync(d_count + 2, 0, sizeof(int), ev->stream));
                            GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
                            if (cudaStreamQuery(ev->stream2) != cudaErrorNotReady) break;
                            // nodes parked by the dense part keep unit == U_RHS_BEGIN, finished ones are -1: re-run the
                            // same list (finished entries are skipped by the kernel through their unit word)
                            D.budget = ev->drain_dense_budget;
                            drain_kernel<ModelStandard><<<ev->num_sms, kBlock, 0, ev->stream>>>(D);
                            ev->launches++;
                            GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
                        }
                        GLC_CHECK(ev, cudaStreamSynchronize(ev->stream2));
                        GLC_CHECK(ev, cudaMemcpyAsync(hc, ev->d_counters, sizeof(unsigned long long) * 9, cudaMemcpyDeviceToHost, ev->stream));
                        GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
                        if (ev->slice_log)
                            fprintf(stderr, "[glc drain pass 0 dense+express] t=%.3f ms done=%llu/%d rhs=%llu\n",
                                    1e3 * (now_s() - t_start), hc[6], n, hc[2]);
                        continue;
                    }
                    // the nodes of a pass are spread over all resident warps (KernelArgs::drainLanes)
                    const int lanes = ev->drain_spread ? std::min(ev->drain_lanes_max, (nheld + warpsResident - 1) / warpsResident)
                                                       : (nheld <= warpsResident ? 1 : ev->drain_lanes_max);
                    const bool sparse = lanes == 1;
                    A.held = ev->d_held;
                    A.nheld = nheld;
                    A.held_counter = d_count + 1;
                    A.drainLanes = lanes >= 32 ? 0 : lanes;
                    A.drainBlockSync = (lanes > 1) ? ev->drain_block_sync : 0;
                    A.budget = sparse ? 0x7fffffff : ev->drain_dense_budget;
                    const int perBlock = lanes * (kBlock / 32);
                    int dgrid = (nheld + perBlock - 1) / perBlock;
                    dgrid = std::max(1, std::min(ev->num_sms * bps, dgrid));
                    drain_kernel<ModelStandard><<<dgrid, kBlock, 0, ev->stream>>>(A);
                    ev->launches += 2;
                    GLC_CHECK(ev, cudaGetLastError());
                    GLC_CHECK(ev, cudaMemcpyAsync(hc, ev->d_counters, sizeof(unsigned long long) * 9, cudaMemcpyDeviceToHost,
                                                  ev->stream));
                    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
                    if (ev->slice_log)
                        fprintf(stderr, "[glc drain pass %d, %d node(s) per warp] t=%.3f ms held=%d done=%llu/%d rhs=%llu\n", pass,
                                lanes, 1e3 * (now_s() - t_start), nheld, hc[6], n, hc[2]);
                }
                break;
            }
        }
    }
    GLC_CHECK(ev, cudaEventRecord(ev->ev1, ev->stream));
    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
    {
        float total = 0.f, first = 0.f;
        cudaEventElapsedTime(&total, ev->ev0, ev->ev1);
        if (ev->phase_split) {
            cudaEventElapsedTime(&first, ev->ev0, ev->ev_mid);
            ev->phase_ms[0] = first;
            ev->phase_ms[1] = total - first;
            ev->phase_rhs[1] = (double)hc[2] - ev->phase_rhs[0];
            ev->phase_steps[1] = (double)hc[0] - ev->phase_steps[0];
            ev->phase_nodes[1] = (double)hc[6] - ev->phase_nodes[0];
        } else {
            ev->phase_ms[0] = total;
            ev->phase_ms[1] = 0.0;
            ev->phase_rhs[0] = (double)hc[2];
            ev->phase_rhs[1] = 0.0;
            ev->phase_steps[0] = (double)hc[0];
            ev->phase_steps[1] = 0.0;
            ev->phase_nodes[0] = (double)hc[6];
            ev->phase_nodes[1] = 0.0;
        }
    }
#ifdef GLC_LEDGER
    {
        unsigned long long herr[8];
        GLC_CHECK(ev, cudaMemcpy(herr, ev->d_ledger_err, sizeof herr, cudaMemcpyDeviceToHost));
        if (herr[0] | herr[1] | herr[2] | herr[3] || (mode != 1 && hc[6] != (unsigned long long)n) || ev->slice_log) ledger_report(ev, n, "end of call");
    }
#endif
    if (hc[10] != 0) {
        ev->err = "micro-task machine: a queue wait timed out inside the kernel (protocol violation)";
        return GLC_ERR_STALLED;
    }
    // a call that runs to completion must have written back every node of the batch
    if (mode != 1 && ev->max_slices <= 0 && !xcpu && hc[6] != (unsigned long long)n) {
        char msg[160];
        snprintf(msg, sizeof msg, "micro-task machine returned with %llu of %d nodes written back", hc[6], n);
        ev->err = msg;
        return GLC_ERR_STALLED;
    }
    return 0;
}

// ---------------------------------------------------------------------------- streaming: adaptive time slices
// glc_stream_run(pops_per_warp = 0).  The machine is a throughput engine: a slot advances at (block throughput) / (slots in
// flight), and a lone slot at ~1 ms per evaluation (every unit is a cold piece of code plus a queue round trip).  A streaming
// host (the asynchronous tree scheduler) mostly has few nodes in flight and waits for the slowest of them, so a tick is
//   * a machine slice while the node queue plus the occupied slots would fill the machine (>= drain_threshold), else
//   * a LANE PASS: every occupied slot is brought to an RK boundary once (hold slices), then drain_kernel -- whole
//     evaluations, one node per lane, or one per WARP when there are fewer nodes than resident warps -- runs all of them for a
//     bounded number of evaluations, fetching queued nodes into free slots as it goes (drainRefill), and parks what is left.
// Both engines work on the same slots; a node's result does not depend on which engine advances it.
static int stream_tick(glc_evolver *ev, int n, unsigned long long *hc) {
    const int64_t need = (int64_t)ev->num_sms * GLC_MSLOTS;
    if (need > ev->nslots_machine || !ev->stream_started) {
        // first tick of a session: a one-pop machine slice allocates the slot arrays (first use) and resets every slot, the
        // queue cursor and the counters
        int rc = launch_machine(ev, n, hc, 1, 1, 0);
        if (rc) return rc;
        ev->stream_lane_mode = false;
        ev->stream_live = (int64_t)hc[7];
    }
    int work = 0;
    GLC_CHECK(ev, cudaMemcpyAsync(&work, ev->d_work, sizeof(int), cudaMemcpyDeviceToHost, ev->stream));
    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
    const int64_t queued = std::max<int64_t>(0, (int64_t)n - (int64_t)std::min(work, n));
    const int64_t big = ev->stream_machine_above >= 0 ? ev->stream_machine_above : ev->drain_threshold;
    const double t_tick = now_s();
    ev->tick_live_sum += (double)(queued + ev->stream_live);
    if (!ev->stream_lane_mode) {
        if (queued + ev->stream_live >= big) {
            int rc = launch_machine(ev, n, hc, 1, ev->stream_machine_budget, 0);
            ev->stream_live = (int64_t)hc[7];
            ev->tick_machine++;
            ev->tick_machine_s += now_s() - t_tick;
            return rc;
        }
        // few nodes: bring every occupied slot to an RK boundary (slots mid-evaluation finish their evaluation)
        for (int k = 0; k < 64; k++) {
            int rc = launch_machine(ev, n, hc, 1, 512, 1);
            ev->tick_hold++;
            if (rc) return rc;
            if (hc[8] == 0) break;
        }
        ev->stream_live = (int64_t)hc[7];
        ev->stream_lane_mode = true;
    } else if (queued + ev->stream_live >= big) {
        ev->stream_lane_mode = false;  // many nodes again: back to the machine (parked slots are re-queued by their unit words)
        int rc = launch_machine(ev, n, hc, 1, ev->stream_machine_budget, 0);
        ev->stream_live = (int64_t)hc[7];
        ev->tick_machine++;
        ev->tick_machine_s += now_s() - t_tick;
        return rc;
    }
    // ---- lane pass
    if (!ev->d_held || ev->held_cap < ev->nslots_machine) {
        cudaFree(ev->d_held);
        ev->d_held = nullptr;
        GLC_CHECK(ev, cudaMalloc(&ev->d_held, sizeof(int32_t) * (ev->nslots_machine + 8)));
        cudaFree(ev->d_held_score);
        ev->d_held_score = nullptr;
        GLC_CHECK(ev, cudaMalloc(&ev->d_held_score, sizeof(float) * ev->nslots_machine));
        ev->held_cap = ev->nslots_machine;
    }
    int *d_count = reinterpret_cast<int *>(ev->d_held + ev->nslots_machine);  // [0] list length, [1] cursor, [3] free slots seen
    int bps = 0;
    GLC_CHECK(ev, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, drain_kernel<ModelStandard>, kBlock, 0));
    if (bps < 1) bps = 1;
    const int nslotsAll = (int)ev->nslots_machine;
    const int warpsResident = ev->num_sms * bps * (kBlock / 32);
    GLC_CHECK(ev, cudaMemsetAsync(d_count, 0, sizeof(int) * 4, ev->stream));
    const bool wantPriority = ev->stream_priority_express > 0 && ev->d_stream_priority != nullptr;
    const bool wantExpress = ev->stream_express > 0 || wantPriority;
    const bool wantSort = !wantExpress && ev->stream_sort > 0;
    held_list_kernel<<<std::min((nslotsAll + 255) / 256, ev->num_sms * 8), 256, 0, ev->stream>>>(
        ev->d_slots.unit, ev->d_slots.L, nslotsAll, ev->d_held, (wantExpress || wantSort) ? ev->d_held_score : nullptr, d_count,
        (int)std::min<int64_t>(queued, nslotsAll), wantExpress ? 1.0f : (wantSort ? -1.0f : 0.0f), wantPriority ? ev->d_stream_priority : nullptr);
    int nlist = 0;
    GLC_CHECK(ev, cudaMemcpyAsync(&nlist, d_count, sizeof(int), cudaMemcpyDeviceToHost, ev->stream));
    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
    ev->launches++;
    if (nlist > 0) {
        KernelArgs A{};
        A.props = ev->d_props;
        A.flags = ev->d_flags;
        A.time_end = ev->d_time_end;
        A.status = ev->d_status;
        A.interrupt = ev->d_interrupt;
        A.cap = ev->cap;
        A.n = n;
        A.ws = ev->d_ws;
        A.nslots = ev->nslots;
        A.work_counter = ev->d_work;
        A.counters = ev->d_counters;
        A.order = nullptr;
        A.resume = 1;
        A.slotL = ev->d_slots.L;
        A.slotYt = ev->d_slots.yt;
        A.slotUnit = ev->d_slots.unit;
        A.held = ev->d_held;
        A.nheld = nlist;
        A.held_counter = d_count + 1;
        A.drainRefill = 1;
        if (wantSort && nlist > warpsResident) {
            // ---- more nodes than warps: the lanes of a warp share one instruction stream, so nodes of the same kind (same
            // component set: queue_bucket) are put next to each other -- a warp of hot-halo-only nodes is done after one or
            // two evaluations and leaves, a warp of disk + spheroid + black-hole nodes iterates its nested solvers together
            std::vector<int32_t> &h_held = ev->h_held;
            std::vector<float> &h_score = ev->h_score;
            h_held.resize(nlist);
            h_score.resize(nlist);
            GLC_CHECK(ev, cudaMemcpyAsync(h_held.data(), ev->d_held, sizeof(int32_t) * nlist, cudaMemcpyDeviceToHost, ev->stream));
            GLC_CHECK(ev, cudaMemcpyAsync(h_score.data(), ev->d_held_score, sizeof(float) * nlist, cudaMemcpyDeviceToHost, ev->stream));
            GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
            int start[66] = {0};
            for (int k = 0; k < nlist; k++) start[std::min(64, std::max(0, (int)h_score[k])) + 1]++;
            for (int b = 0; b < 65; b++) start[b + 1] += start[b];
            std::vector<int32_t> &ordered = ev->h_ordered;
            ordered.resize(nlist);
            for (int k = 0; k < nlist; k++) ordered[start[std::min(64, std::max(0, (int)h_score[k]))]++] = h_held[k];
            GLC_CHECK(ev, cudaMemcpyAsync(ev->d_held, ordered.data(), sizeof(int32_t) * nlist, cudaMemcpyHostToDevice, ev->stream));
            GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
        }
        int nexpress = 0;
        const int expressCap = ev->num_sms * (kBlock / 32) * std::max(1, bps - 1);  // all but one block per SM, one node per warp
        if (wantExpress && nlist > warpsResident) {
            // ---- more nodes than warps.  The tree scheduler waits for the SLOWEST node of a group, and a node alone in its warp
            // advances ~4x faster than one of 32 (a warp serialises the divergent evaluations of its lanes).  The nodes that
            // have run longest / have the most steps left (held_list_kernel's score) therefore get a warp each (express kernel,
            // stream2) while the rest shares the remaining block per SM one node per lane.
            std::vector<int32_t> &h_held = ev->h_held;
            std::vector<float> &h_score = ev->h_score;
            h_held.resize(nlist);
            h_score.resize(nlist);
            GLC_CHECK(ev, cudaMemcpyAsync(h_held.data(), ev->d_held, sizeof(int32_t) * nlist, cudaMemcpyDeviceToHost, ev->stream));
            GLC_CHECK(ev, cudaMemcpyAsync(h_score.data(), ev->d_held_score, sizeof(float) * nlist, cudaMemcpyDeviceToHost, ev->stream));
            GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
            std::vector<int> &idx = ev->h_idx;
            idx.resize(nlist);
            for (int k = 0; k < nlist; k++) idx[k] = k;
            nexpress = std::min(wantPriority ? std::min(expressCap, ev->stream_priority_express) : expressCap, nlist);
            const float expressMin = wantPriority ? 1.0e-30f : (float)ev->stream_express;  // (fresh slots carry priority 0)
            std::nth_element(idx.begin(), idx.begin() + nexpress, idx.end(), [&](int a, int b) { return h_score[a] > h_score[b]; });
            // only nodes that have actually been running qualify (fresh slots and just-fetched nodes score 0)
            int keep = 0;
            std::vector<int32_t> &ordered = ev->h_ordered;
            ordered.resize(nlist);
            for (int k = 0; k < nexpress; k++)
                if (h_score[idx[k]] >= expressMin) ordered[keep++] = h_held[idx[k]];
            int tail = keep;
            for (int k = 0; k < nexpress; k++)
                if (!(h_score[idx[k]] >= expressMin)) ordered[tail++] = h_held[idx[k]];
            for (int k = nexpress; k < nlist; k++) ordered[tail++] = h_held[idx[k]];
            nexpress = keep;
            if (nexpress > 0) {
                GLC_CHECK(ev, cudaMemcpyAsync(ev->d_held, ordered.data(), sizeof(int32_t) * nlist, cudaMemcpyHostToDevice, ev->stream));
                GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
                KernelArgs X = A;
                X.nheld = nexpress;
                X.held_counter = d_count + 2;
                X.drainLanes = 1;
                X.drainRefill = 0;  // an express warp is done when its node is
                X.budget = ev->stream_express_budget;
                drain_kernel<ModelStandard><<<(nexpress + kBlock / 32 - 1) / (kBlock / 32), kBlock, 0, ev->stream2>>>(X);
                ev->launches++;
                A.held = ev->d_held + nexpress;
                A.nheld = nlist - nexpress;
            }
        }
        // spread over the resident warps the express kernel leaves: a node advances faster the fewer nodes share its warp; the
        // evaluation budget of the pass shrinks with the number of nodes per warp so that a tick stays a few milliseconds long
        const int warpsFree = std::max(kBlock / 32, warpsResident - (nexpress + kBlock / 32 - 1) / (kBlock / 32) * (kBlock / 32));
        // stream_spread: 0 = one node per warp when they fit, else 32 per warp on as few blocks as needed; 1 = over all resident
        // warps; 2 = over one block per SM as soon as there are more nodes than that block has warps (a second block per SM only
        // when 32 nodes per warp do not suffice)
        const int warpsOne = ev->num_sms * (kBlock / 32);
        int lanes = A.nheld <= warpsFree ? 1 : 32;
        if (ev->stream_spread == 1) lanes = std::min(32, (A.nheld + warpsFree - 1) / warpsFree);
        if (ev->stream_spread == 2 && A.nheld > warpsOne) lanes = std::min(32, (A.nheld + warpsOne - 1) / warpsOne);
        A.drainLanes = lanes >= 32 ? 0 : lanes;
        A.drainBlockSync = (lanes > 1) ? ev->drain_block_sync : 0;
        A.budget = ev->stream_sparse_budget - (int)((long long)(ev->stream_sparse_budget - ev->stream_dense_budget) * (lanes - 1) / 31);
        const int perBlock = lanes * (kBlock / 32);
        int dgrid = (A.nheld + perBlock - 1) / perBlock;
        dgrid = std::max(1, std::min(warpsFree / (kBlock / 32), dgrid));
        ev->tick_lanes_sum += lanes;
        ev->tick_express_sum += nexpress;
        if (A.nheld > 0) {
            drain_kernel<ModelStandard><<<dgrid, kBlock, 0, ev->stream>>>(A);
            ev->launches++;
        }
        ev->slices++;
        GLC_CHECK(ev, cudaGetLastError());
        if (nexpress > 0) GLC_CHECK(ev, cudaStreamSynchronize(ev->stream2));
    }
    GLC_CHECK(ev, cudaMemcpyAsync(hc, ev->d_counters, sizeof(unsigned long long) * 11, cudaMemcpyDeviceToHost, ev->stream));
    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
    ev->stream_live = nlist;  // upper bound until the next listing
    ev->tick_lane++;
    ev->tick_lane_s += now_s() - t_tick;
    if (ev->slice_log)
        fprintf(stderr, "[glc stream tick: lane pass] n=%d queued=%lld listed=%d fetched=%llu done=%llu rhs=%llu\n", n, (long long)queued,
                nlist, hc[5], hc[6], hc[2]);
    return 0;
}

static int upload_constants(glc_evolver *ev) {
    GLC_CHECK(ev, cudaMemcpyToSymbolAsync(c_params, &ev->params, sizeof(glc_params), 0,
                                          cudaMemcpyHostToDevice, ev->stream));
    GLC_CHECK(ev, cudaMemcpyToSymbolAsync(c_tables, &ev->tables, sizeof(DeviceTables), 0,
                                          cudaMemcpyHostToDevice, ev->stream));
    return 0;
}

// ------------------------------------------------------------------------------- C ABI
extern "C" {

int glc_abi_version(void) { return GLC_ABI_VERSION; }

const char *glc_last_error(const glc_evolver *ev) { return ev ? ev->err.c_str() : "null evolver"; }

int glc_evolver_create(glc_evolver **out, int32_t device_ordinal) {
    if (!out) return -1;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return -2;  // no CPU fallback, by design
    if (device_ordinal < 0 || device_ordinal >= count) return -3;
    glc_evolver *ev = new glc_evolver();
    ev->device = device_ordinal;
    if (cudaSetDevice(device_ordinal) != cudaSuccess) {
        delete ev;
        return -4;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device_ordinal);
    ev->num_sms = prop.multiProcessorCount;
    if (const char *e = getenv("GLC_SLICE_BUDGET")) ev->slice_budget = atoi(e);
    if (const char *e = getenv("GLC_SORT_QUEUE")) ev->sort_queue = atoi(e);
    if (const char *e = getenv("GLC_SLICE_LOG")) ev->slice_log = atoi(e);
    if (const char *e = getenv("GLC_MAX_SLICES")) ev->max_slices = atoi(e);
    if (const char *e = getenv("GLC_MACHINE")) ev->use_machine = atoi(e);
    if (const char *e = getenv("GLC_MACHINE_MIN_NODES")) ev->machine_min_nodes = atoll(e);
    if (const char *e = getenv("GLC_DRAIN")) ev->drain_handover = atoi(e);
    if (const char *e = getenv("GLC_DRAIN_BELOW")) ev->drain_threshold = atoll(e);
    if (const char *e = getenv("GLC_DRAIN_DENSE_BUDGET")) ev->drain_dense_budget = atoi(e);
    if (const char *e = getenv("GLC_HYBRID_BUDGET")) ev->hybrid_budget = atoi(e);
    if (const char *e = getenv("GLC_DRAIN_EXPRESS")) ev->drain_express = atoi(e);
    if (const char *e = getenv("GLC_DRAIN_SPREAD")) ev->drain_spread = atoi(e);
    if (const char *e = getenv("GLC_DRAIN_BLOCK_SYNC")) ev->drain_block_sync = atoi(e);
    if (const char *e = getenv("GLC_DRAIN_LANES_MAX")) ev->drain_lanes_max = std::max(1, std::min(32, atoi(e)));
    if (const char *e = getenv("GLC_DRAIN_AGE_WEIGHT")) ev->drain_age_weight = (float)atof(e);
    if (const char *e = getenv("GLC_L2_PERSIST")) ev->l2_persist = atoi(e);
    if (const char *e = getenv("GLC_STREAM_SPARSE_BUDGET")) ev->stream_sparse_budget = atoi(e);
    if (const char *e = getenv("GLC_STREAM_DENSE_BUDGET")) ev->stream_dense_budget = atoi(e);
    if (const char *e = getenv("GLC_STREAM_MACHINE_ABOVE")) ev->stream_machine_above = atoll(e);
    if (const char *e = getenv("GLC_STREAM_MACHINE_BUDGET")) ev->stream_machine_budget = atoi(e);
    if (const char *e = getenv("GLC_STREAM_EXPRESS")) ev->stream_express = atoi(e);
    if (const char *e = getenv("GLC_STREAM_SPREAD")) ev->stream_spread = atoi(e);
    if (const char *e = getenv("GLC_STREAM_SORT")) ev->stream_sort = atoi(e);
    if (const char *e = getenv("GLC_STREAM_EXPRESS_BUDGET")) ev->stream_express_budget = atoi(e);
    if (const char *e = getenv("GLC_STREAM_PRIORITY_EXPRESS")) ev->stream_priority_express = atoi(e);
    cudaStreamCreateWithFlags(&ev->stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&ev->stream2, cudaStreamNonBlocking);
    cudaEventCreate(&ev->ev0);
    cudaEventCreate(&ev->ev1);
    cudaEventCreate(&ev->ev_mid);
    cudaMalloc(&ev->d_work, sizeof(int));
    cudaMalloc(&ev->d_counters, sizeof(unsigned long long) * 16);
    if (cudaGetLastError() != cudaSuccess) {
        delete ev;
        return -5;
    }
    *out = ev;
    return 0;
}

int glc_evolver_destroy(glc_evolver *ev) {
    if (!ev) return 0;
    cudaSetDevice(ev->device);
    cudaStreamSynchronize(ev->stream);
    for (auto &t : ev->host_tables) {
        cudaFree(t.d_x0);
        cudaFree(t.d_x1);
        cudaFree(t.d_v);
    }
    cudaFree(ev->d_props);
    cudaFree(ev->d_flags);
    cudaFree(ev->d_status);
    cudaFree(ev->d_interrupt);
    cudaFree(ev->d_time_end);
    cudaFree(ev->d_stage);
    cudaFree(ev->d_dydt);
    cudaFree(ev->d_snap_props);
    cudaFree(ev->d_snap_flags);
    cudaFree(ev->d_ws);
    cudaFree(ev->d_work);
    cudaFree(ev->d_counters);
    cudaFree(ev->d_held);
    cudaFree(ev->d_collected);
    cudaFree(ev->d_collect_list);
    cudaFree(ev->d_collect_meta);
    cudaFreeHost(ev->h_collect_meta);
    cudaFree(ev->d_pow_ac);
    cudaFree(ev->d_pow_kmt);
    cudaFree(ev->d_nfw_jx);
    cudaFree(ev->d_nfw_jv);
    cudaFree(ev->d_profile);
    cudaFree(ev->d_lanes);
    free_slots(ev);
    cudaFree(ev->d_order);
    cudaFree(ev->d_sort);
    cudaEventDestroy(ev->ev0);
    cudaEventDestroy(ev->ev1);
    cudaEventDestroy(ev->ev_mid);
    cudaStreamDestroy(ev->stream);
    cudaStreamDestroy(ev->stream2);
    cudaFree(ev->d_held_score);
    cudaFree(ev->d_stream_priority);
    delete ev;
    return 0;
}

int glc_evolver_set_params(glc_evolver *ev, const glc_params *params) {
    if (!ev || !params) return -1;
    if (params->abi_version != GLC_ABI_VERSION) {
        ev->err = "glc_params.abi_version mismatch";
        return -6;
    }
    if (params->model != GLC_MODEL_BOX && params->model != GLC_MODEL_STANDARD) {
        ev->err = "unknown model";
        return -7;
    }
    if (!(params->odeToleranceAbsolute > 0.0) && !(params->odeToleranceRelative > 0.0)) {
        // solver.F90:373-374
        ev->err = "at least one of absolute and relative tolerance must be greater than zero";
        return -8;
    }
    cudaSetDevice(ev->device);
    if (params->model == GLC_MODEL_STANDARD && (!ev->d_pow_ac || ev->pow_ac_exponent != params->adiabaticOmega)) {
        // fastExponentiator tables, tabulated once as the reference's constructors do
        const std::vector<double> ac = build_pow_table(1.0e-3, 1.0, params->adiabaticOmega, 1.0e4);
        const std::vector<double> kmt = build_pow_table(1.0, 1000.0, 0.33, 100.0);
        cudaFree(ev->d_pow_ac);
        cudaFree(ev->d_pow_kmt);
        ev->d_pow_ac = ev->d_pow_kmt = nullptr;
        GLC_CHECK(ev, cudaMalloc(&ev->d_pow_ac, sizeof(double) * ac.size()));
        GLC_CHECK(ev, cudaMalloc(&ev->d_pow_kmt, sizeof(double) * kmt.size()));
        GLC_CHECK(ev, cudaMemcpy(ev->d_pow_ac, ac.data(), sizeof(double) * ac.size(), cudaMemcpyHostToDevice));
        GLC_CHECK(ev, cudaMemcpy(ev->d_pow_kmt, kmt.data(), sizeof(double) * kmt.size(), cudaMemcpyHostToDevice));
        ev->tables.powAc = ev->d_pow_ac;
        ev->tables.powAcN = (int)ac.size();
        ev->tables.powKmt = ev->d_pow_kmt;
        ev->tables.powKmtN = (int)kmt.size();
        pow_table_spacing(1.0e-3, 1.0, (int)ac.size(), ev->tables.powAcDx, ev->tables.powAcInvDx);
        pow_table_spacing(1.0, 1000.0, (int)kmt.size(), ev->tables.powKmtDx, ev->tables.powKmtInvDx);
        ev->pow_ac_exponent = params->adiabaticOmega;
    }
    if (params->model == GLC_MODEL_STANDARD && !ev->d_nfw_jx) {
        std::vector<double> xs, js;
        build_nfw_j_table(xs, js);
        GLC_CHECK(ev, cudaMalloc(&ev->d_nfw_jx, sizeof(double) * xs.size()));
        GLC_CHECK(ev, cudaMalloc(&ev->d_nfw_jv, sizeof(double) * js.size()));
        GLC_CHECK(ev, cudaMemcpy(ev->d_nfw_jx, xs.data(), sizeof(double) * xs.size(), cudaMemcpyHostToDevice));
        GLC_CHECK(ev, cudaMemcpy(ev->d_nfw_jv, js.data(), sizeof(double) * js.size(), cudaMemcpyHostToDevice));
        ev->tables.nfwJx = ev->d_nfw_jx;
        ev->tables.nfwJv = ev->d_nfw_jv;
        ev->tables.nfwJN = (int)xs.size();
    }
    ev->tables.profile = nullptr;
    ev->tables.profBins = 0;
    if (params->profileOdeEvolver) {
        if (!(params->profilerTimeStepMinimum > 0.0) || !(params->profilerTimeStepMaximum > params->profilerTimeStepMinimum) ||
            params->profilerTimeStepPointsPerDecade < 1) {
            ev->err = "profileOdeEvolver needs 0 < profilerTimeStepMinimum < profilerTimeStepMaximum and profilerTimeStepPointsPerDecade >= 1";
            return -8;
        }
        if (!ev->d_profile) {
            GLC_CHECK(ev, cudaMalloc(&ev->d_profile, sizeof(unsigned long long) * kProfWords));
            GLC_CHECK(ev, cudaMemset(ev->d_profile, 0, sizeof(unsigned long long) * kProfWords));
            GLC_CHECK(ev, cudaMemset(ev->d_profile + kProfSmallest, 0x7f, sizeof(unsigned long long)));  // "huge": no step yet
        }
        ev->tables.profBins = build_profile_edges(*params, ev->tables.profEdges);
        ev->tables.profile = ev->d_profile;
    }
    ev->tables.lnThinDiskMin = params->accretionRateThinDiskMinimum > 0.0 ? dm_log(params->accretionRateThinDiskMinimum) : 0.0;
    ev->tables.lnThinDiskMax = params->accretionRateThinDiskMaximum > 0.0 ? dm_log(params->accretionRateThinDiskMaximum) : 0.0;
    ev->params = *params;
    ev->params_set = true;
    return 0;
}

int glc_evolver_set_table(glc_evolver *ev, int32_t id, int32_t n0, int32_t n1, const double *x0,
                          const double *x1, const double *values) {
    if (!ev) return -1;
    cudaSetDevice(ev->device);
    PreparedTable pt;
    if (prepare_table(id, n0, n1, x0, x1, values, pt) != 0) return -1;
    HostTable &t = ev->host_tables[id];
    cudaFree(t.d_x0);
    cudaFree(t.d_x1);
    cudaFree(t.d_v);
    t = HostTable();
    GLC_CHECK(ev, cudaMalloc(&t.d_x0, sizeof(double) * n0));
    GLC_CHECK(ev, cudaMemcpy(t.d_x0, pt.x0.data(), sizeof(double) * n0, cudaMemcpyHostToDevice));
    if (!pt.x1.empty()) {
        GLC_CHECK(ev, cudaMalloc(&t.d_x1, sizeof(double) * n1));
        GLC_CHECK(ev, cudaMemcpy(t.d_x1, pt.x1.data(), sizeof(double) * n1, cudaMemcpyHostToDevice));
    }
    GLC_CHECK(ev, cudaMalloc(&t.d_v, sizeof(double) * (size_t)n0 * n1));
    GLC_CHECK(ev, cudaMemcpy(t.d_v, pt.v.data(), sizeof(double) * (size_t)n0 * n1, cudaMemcpyHostToDevice));
    t.n0 = n0;
    t.n1 = n1;
    install_table(ev->tables, id, pt, DeviceTable2D{n0, n1, t.d_x0, t.d_x1, t.d_v});
    if (id == GLC_TABLE_HALO_MEAN_DENSITY) ev->halo_host.set(n0, x0, values);
    return 0;
}

int glc_arena_reserve(glc_evolver *ev, int64_t capacity) {
    if (!ev || capacity < 1) return -1;
    cudaSetDevice(ev->device);
    if (capacity <= ev->cap) return 0;
    cudaStreamSynchronize(ev->stream);
    cudaFree(ev->d_props);
    cudaFree(ev->d_flags);
    cudaFree(ev->d_status);
    cudaFree(ev->d_interrupt);
    cudaFree(ev->d_time_end);
    cudaFree(ev->d_stage);
    cudaFree(ev->d_dydt);
    ev->d_dydt = nullptr;
    ev->cap = 0;
    GLC_CHECK(ev, cudaMalloc(&ev->d_props, sizeof(double) * NPROP * capacity));
    GLC_CHECK(ev, cudaMalloc(&ev->d_stage, sizeof(double) * NPROP * capacity));
    GLC_CHECK(ev, cudaMalloc(&ev->d_flags, sizeof(int32_t) * capacity));
    GLC_CHECK(ev, cudaMalloc(&ev->d_status, sizeof(int32_t) * capacity));
    GLC_CHECK(ev, cudaMalloc(&ev->d_interrupt, sizeof(int32_t) * capacity));
    GLC_CHECK(ev, cudaMalloc(&ev->d_time_end, sizeof(double) * capacity));
    ev->cap = capacity;
    return 0;
}

int glc_arena_upload(glc_evolver *ev, int64_t n, const double *props, const int32_t *flags,
                     const double *time_end) {
    if (!ev || n < 0 || !props || !flags || !time_end) return -1;
    if (ev->stream_active) {
        ev->err = "arena upload during a streaming session (the session owns the arena): call glc_stream_end first";
        return GLC_ERR_BUSY;
    }
    if (n == 0) return 0;
    cudaSetDevice(ev->device);
    int rc = glc_arena_reserve(ev, n);
    if (rc) return rc;
    GLC_CHECK(ev, cudaMemcpyAsync(ev->d_stage, props, sizeof(double) * NPROP * n, cudaMemcpyHostToDevice, ev->stream));
    GLC_CHECK(ev, cudaMemcpyAsync(ev->d_flags, flags, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ev->stream));
    GLC_CHECK(ev, cudaMemcpyAsync(ev->d_time_end, time_end, sizeof(double) * n, cudaMemcpyHostToDevice, ev->stream));
    aos_to_soa_kernel<<<(int)((n + 63) / 64), 256, 0, ev->stream>>>(ev->d_stage, ev->d_props, (int)n, ev->cap);
    ev->launches++;
    GLC_CHECK(ev, cudaGetLastError());
    return 0;
}

int glc_arena_download(glc_evolver *ev, int64_t n, double *props, int32_t *flags, int32_t *status,
                       int32_t *interrupt) {
    if (!ev || n < 0 || n > ev->cap) return -1;
    if (n == 0) return 0;
    cudaSetDevice(ev->device);
    if (props) {
        soa_to_aos_kernel<<<(int)((n + 63) / 64), 256, 0, ev->stream>>>(ev->d_props, ev->d_stage, (int)n, ev->cap);
        ev->launches++;
        GLC_CHECK(ev, cudaGetLastError());
        GLC_CHECK(ev, cudaMemcpyAsync(props, ev->d_stage, sizeof(double) * NPROP * n, cudaMemcpyDeviceToHost, ev->stream));
    }
    if (flags) GLC_CHECK(ev, cudaMemcpyAsync(flags, ev->d_flags, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ev->stream));
    if (status) GLC_CHECK(ev, cudaMemcpyAsync(status, ev->d_status, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ev->stream));
    if (interrupt) GLC_CHECK(ev, cudaMemcpyAsync(interrupt, ev->d_interrupt, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ev->stream));
    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
    return 0;
}

int glc_evolve_arena(glc_evolver *ev, int64_t n, glc_counters *counters) {
    if (!ev || n < 0 || n > ev->cap) return -1;
    if (n > 0x7fffffff) {
        ev->err = "more than 2^31-1 nodes in one batch";
        return -1;
    }
    if (ev->stream_active) {
        ev->err = "batch call during a streaming session (the session owns the arena): call glc_stream_end first";
        return GLC_ERR_BUSY;
    }
    if (!ev->params_set) {
        ev->err = "glc_evolver_set_params has not been called";
        return -9;
    }
    if (n == 0) {
        if (counters) memset(counters, 0, sizeof(*counters));
        return 0;
    }
    cudaSetDevice(ev->device);
    int rc = upload_constants(ev);
    if (rc) return rc;
    unsigned long long hc[16] = {0};
    if (ev->params.model == GLC_MODEL_BOX)
        rc = launch_evolve<ModelBox>(ev, (int)n, hc);
    else if (ev->use_machine == 1 ||
             (ev->use_machine == 2 && n >= (ev->machine_min_nodes >= 0 ? ev->machine_min_nodes : ev->drain_threshold)))
        rc = launch_machine(ev, (int)n, hc);
    else
        rc = launch_evolve<ModelStandard>(ev, (int)n, hc);
    if (rc) return rc;
    GLC_CHECK(ev, cudaEventElapsedTime(&ev->last_ms, ev->ev0, ev->ev1));
    if (counters) {
        counters->steps_accepted = hc[0];
        counters->steps_rejected = hc[1];
        counters->rhs_evaluations = hc[2];
        counters->segments = hc[3];
        counters->trials_failed = hc[4];
        counters->nodes = hc[5];
    }
    return 0;
}

int glc_arena_snapshot(glc_evolver *ev, int64_t n) {
    if (!ev || n < 0 || n > ev->cap) return -1;
    cudaSetDevice(ev->device);
    if (ev->snap_cap != ev->cap) {
        cudaFree(ev->d_snap_props);
        cudaFree(ev->d_snap_flags);
        ev->d_snap_props = nullptr;
        ev->d_snap_flags = nullptr;
        GLC_CHECK(ev, cudaMalloc(&ev->d_snap_props, sizeof(double) * NPROP * ev->cap));
        GLC_CHECK(ev, cudaMalloc(&ev->d_snap_flags, sizeof(int32_t) * ev->cap));
        ev->snap_cap = ev->cap;
    }
    GLC_CHECK(ev, cudaMemcpyAsync(ev->d_snap_props, ev->d_props, sizeof(double) * NPROP * ev->cap, cudaMemcpyDeviceToDevice, ev->stream));
    GLC_CHECK(ev, cudaMemcpyAsync(ev->d_snap_flags, ev->d_flags, sizeof(int32_t) * ev->cap, cudaMemcpyDeviceToDevice, ev->stream));
    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
    return 0;
}

int glc_arena_restore(glc_evolver *ev, int64_t n) {
    if (!ev || n < 0 || n > ev->cap || ev->snap_cap != ev->cap) return -1;
    cudaSetDevice(ev->device);
    GLC_CHECK(ev, cudaMemcpyAsync(ev->d_props, ev->d_snap_props, sizeof(double) * NPROP * ev->cap, cudaMemcpyDeviceToDevice, ev->stream));
    GLC_CHECK(ev, cudaMemcpyAsync(ev->d_flags, ev->d_snap_flags, sizeof(int32_t) * ev->cap, cudaMemcpyDeviceToDevice, ev->stream));
    return 0;
}

int64_t glc_arena_capacity(const glc_evolver *ev) { return ev ? ev->cap : 0; }
int64_t glc_slice_count(const glc_evolver *ev) { return ev ? ev->slices : 0; }

int glc_evolver_set_option(glc_evolver *ev, int32_t option, int64_t value) {
    if (!ev) return -1;
    switch (option) {
        case GLC_OPT_SLICE_BUDGET: ev->slice_budget = (int32_t)std::max<int64_t>(0, std::min<int64_t>(value, 0x7fffffff)); return 0;
        case GLC_OPT_SORT_QUEUE: ev->sort_queue = value ? 1 : 0; return 0;
        case GLC_OPT_MICROTASK_MACHINE: ev->use_machine = value == 2 ? 2 : (value ? 1 : 0); return 0;
        case GLC_OPT_FOREST_SCHEDULE: ev->forest_schedule = value ? 1 : 0; return 0;
        default: ev->err = "unknown option"; return -10;
    }
}
int64_t glc_kernel_launch_count(const glc_evolver *ev) { return ev ? ev->launches : 0; }

double glc_measure_fp64_peak_tflops(glc_evolver *ev) {
    if (!ev) return 0.0;
    cudaSetDevice(ev->device);
    double *d_out = nullptr;
    const int grid = ev->num_sms * 16, block = 256, iters = 16384;
    if (cudaMalloc(&d_out, sizeof(double) * grid * block) != cudaSuccess) return 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 12; rep++) {  // the first launches also ramp the clocks
        cudaEventRecord(e0, ev->stream);
        fp64_peak_kernel<<<grid, block, 0, ev->stream>>>(d_out, iters);
        cudaEventRecord(e1, ev->stream);
        cudaStreamSynchronize(ev->stream);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 16.0 * (double)iters * (double)grid * block;  // 16 independent FMA chains
        if (rep >= 2 && ms > 0.f) best = std::max(best, flops / (ms * 1.0e-3) / 1.0e12);
    }
    ev->launches += 12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return best;
}

float glc_last_kernel_ms(const glc_evolver *ev) { return ev ? ev->last_ms : 0.f; }
int glc_last_phase_stats(const glc_evolver *ev, double *out6) {
    if (!ev || !out6) return -1;
    out6[6] = ev->phase_nodes[0];
    out6[7] = ev->phase_nodes[1];
    out6[0] = ev->phase_ms[0];
    out6[1] = ev->phase_ms[1];
    out6[2] = ev->phase_rhs[0];
    out6[3] = ev->phase_rhs[1];
    out6[4] = ev->phase_steps[0];
    out6[5] = ev->phase_steps[1];
    return 0;
}
void *glc_arena_device_props(glc_evolver *ev) { return ev ? (void *)ev->d_props : nullptr; }
void *glc_evolver_stream(glc_evolver *ev) { return ev ? (void *)ev->stream : nullptr; }

int glc_evolve_batch(glc_evolver *ev, int64_t n, double *props, int32_t *flags, const double *time_end,
                     int32_t *status, int32_t *interrupt, glc_counters *counters) {
    if (!ev || n < 0) return -1;
    if (n == 0) {
        if (counters) memset(counters, 0, sizeof(*counters));
        return 0;
    }
    if (!props || !flags || !time_end || !status || !interrupt) return -1;
    int rc = glc_arena_upload(ev, n, props, flags, time_end);
    if (rc) return rc;
    rc = glc_evolve_arena(ev, n, counters);
    if (rc == GLC_ERR_STALLED) {
        // nothing was lost: hand back what was evolved; the others carry GLC_STATUS_PENDING and their input records
        const std::string why = ev->err;
        glc_arena_download(ev, n, props, flags, status, interrupt);
        ev->err = why;
        return rc;
    }
    if (rc) return rc;
    return glc_arena_download(ev, n, props, flags, status, interrupt);
}

int glc_error_report_node(glc_evolver *ev, const double *record, int32_t flags, double time_step, glc_error_report *out) {
    if (!ev || !record || !out) return -1;
    if (!ev->params_set) return -9;
    if (ev->stream_active) {
        ev->err = "glc_error_report_node: a streaming session owns the arena";
        return GLC_ERR_BUSY;
    }
    cudaSetDevice(ev->device);
    std::vector<double> row(record, record + NPROP);
    const double te = 0.0;
    int rc = glc_arena_upload(ev, 1, row.data(), &flags, &te);
    if (rc) return rc;
    if (!ev->d_dydt) GLC_CHECK(ev, cudaMalloc(&ev->d_dydt, sizeof(double) * NY * ev->cap));
    rc = upload_constants(ev);
    if (rc) return rc;
    double *d_out = nullptr;
    GLC_CHECK(ev, cudaMalloc(&d_out, sizeof(double) * (7 * NY + 1)));
    KernelArgs A{};
    A.props = ev->d_props;
    A.flags = ev->d_flags;
    A.cap = ev->cap;
    A.n = 1;
    if (ev->params.model == GLC_MODEL_BOX)
        error_report_kernel<ModelBox><<<1, 32, 0, ev->stream>>>(A, time_step, d_out);
    else
        error_report_kernel<ModelStandard><<<1, 32, 0, ev->stream>>>(A, time_step, d_out);
    ev->launches++;
    double h[7 * NY + 1];
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_out, sizeof h, cudaMemcpyDeviceToHost, ev->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ev->stream);
    cudaFree(d_out);
    GLC_CHECK(ev, e);
    memset(out, 0, sizeof(*out));
    out->time = record[GLC_P_TIME];
    out->time_step = time_step;
    for (int i = 0; i < NY; i++) {
        out->y[i] = h[0 * NY + i];
        out->dydt[i] = h[1 * NY + i];
        out->scale[i] = h[2 * NY + i];
        out->tolerance[i] = h[3 * NY + i];
        out->error[i] = h[4 * NY + i];
        out->error_scaled[i] = h[5 * NY + i];
        out->active[i] = h[6 * NY + i] != 0.0 ? 1 : 0;
    }
    out->interrupt = (int32_t)h[7 * NY];
    return 0;
}

int glc_rhs_batch(glc_evolver *ev, int64_t n, double *props, const int32_t *flags, double *dydt,
                  int32_t *interrupt) {
    if (!ev || n <= 0 || !props || !flags || !dydt || !interrupt) return -1;
    if (!ev->params_set) return -9;
    cudaSetDevice(ev->device);
    std::vector<double> te((size_t)n, 0.0);
    int rc = glc_arena_upload(ev, n, props, flags, te.data());
    if (rc) return rc;
    if (!ev->d_dydt) GLC_CHECK(ev, cudaMalloc(&ev->d_dydt, sizeof(double) * NY * ev->cap));
    rc = upload_constants(ev);
    if (rc) return rc;
    KernelArgs A{};
    A.props = ev->d_props;
    A.flags = ev->d_flags;
    A.interrupt = ev->d_interrupt;
    A.cap = ev->cap;
    A.n = (int)n;
    if (ev->params.model == GLC_MODEL_BOX)
        rhs_kernel<ModelBox><<<(int)((n + 127) / 128), 128, 0, ev->stream>>>(A, ev->d_dydt);
    else
        rhs_kernel<ModelStandard><<<(int)((n + 127) / 128), 128, 0, ev->stream>>>(A, ev->d_dydt);
    GLC_CHECK(ev, cudaGetLastError());
    GLC_CHECK(ev, cudaMemcpyAsync(dydt, ev->d_dydt, sizeof(double) * NY * n, cudaMemcpyDeviceToHost, ev->stream));
    return glc_arena_download(ev, n, props, nullptr, nullptr, interrupt);
}

// ------------------------------------------------------------------ streaming interface
// The production host (the batching tree evolver of INTEGRATION.md) never waits for a batch to drain: it submits
// evolvable nodes as the tree walk produces them, lets the device run time slices, and collects finished nodes.
// A ticket is the node's index in the arena; the arena capacity bounds the tickets of one session.
int glc_stream_begin(glc_evolver *ev, int64_t capacity) {
    if (!ev || capacity < 1 || capacity > 0x7fffffff) return -1;
    if (!ev->params_set) return -9;
    if (ev->params.model != GLC_MODEL_STANDARD || !ev->use_machine) {
        ev->err = "streaming needs the micro-task machine (standard model)";
        return -11;
    }
    cudaSetDevice(ev->device);
    int rc = glc_arena_reserve(ev, capacity);
    if (rc) return rc;
    if (ev->collected_cap < ev->cap) {
        cudaFree(ev->d_collected);
        cudaFree(ev->d_collect_list);
        ev->d_collected = nullptr;
        ev->d_collect_list = nullptr;
        GLC_CHECK(ev, cudaMalloc(&ev->d_collected, (size_t)ev->cap));
        GLC_CHECK(ev, cudaMalloc(&ev->d_collect_list, sizeof(int64_t) * ((size_t)ev->cap + 1)));
        ev->collected_cap = ev->cap;
    }
    GLC_CHECK(ev, cudaMemsetAsync(ev->d_collected, 0, (size_t)ev->cap, ev->stream));
    GLC_CHECK(ev, cudaMemsetAsync(ev->d_status, 0x80, sizeof(int32_t) * (size_t)ev->cap, ev->stream));  // GLC_STATUS_PENDING
    rc = upload_constants(ev);
    if (rc) return rc;
    ev->stream_active = true;
    ev->stream_started = false;
    ev->stream_lane_mode = false;
    ev->stream_live = 0;
    ev->stream_n = 0;
    ev->stream_collected = 0;
    return 0;
}

int glc_stream_submit(glc_evolver *ev, int64_t n, const double *props, const int32_t *flags, const double *time_end,
                      int64_t *first_ticket) {
    if (!ev || !ev->stream_active || n < 0 || !props || !flags || !time_end) return -1;
    if (ev->stream_n + n > ev->cap) {
        ev->err = "stream capacity exhausted";
        return -12;
    }
    if (first_ticket) *first_ticket = ev->stream_n;
    if (n == 0) return 0;
    cudaSetDevice(ev->device);
    const int64_t off = ev->stream_n;
    GLC_CHECK(ev, cudaMemcpyAsync(ev->d_stage + off * NPROP, props, sizeof(double) * NPROP * n, cudaMemcpyHostToDevice, ev->stream));
    GLC_CHECK(ev, cudaMemcpyAsync(ev->d_flags + off, flags, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ev->stream));
    GLC_CHECK(ev, cudaMemcpyAsync(ev->d_time_end + off, time_end, sizeof(double) * n, cudaMemcpyHostToDevice, ev->stream));
    aos_to_soa_kernel<<<(int)((n + 63) / 64), 256, 0, ev->stream>>>(ev->d_stage + off * NPROP, ev->d_props + off, (int)n, ev->cap);
    ev->launches++;
    GLC_CHECK(ev, cudaGetLastError());
    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));  // the host buffers may be reused on return
    ev->stream_n += n;
    return 0;
}

int glc_stream_run(glc_evolver *ev, int32_t pops_per_warp, int64_t *n_finished_total, glc_counters *counters) {
    if (!ev || !ev->stream_active) return -1;
    cudaSetDevice(ev->device);
    unsigned long long hc[16] = {0};
    if (ev->stream_n > 0) {
        // pops_per_warp > 0: one machine slice of that budget; 0: an adaptive tick (machine slice or lane pass, stream_tick)
        int rc = pops_per_warp > 0 ? launch_machine(ev, (int)ev->stream_n, hc, 1, pops_per_warp) : stream_tick(ev, (int)ev->stream_n, hc);
        if (rc) return rc;
        if (pops_per_warp > 0) {
            ev->stream_lane_mode = false;
            ev->stream_live = (int64_t)hc[7];
        }
    }
    if (n_finished_total) *n_finished_total = (int64_t)hc[6];
    if (counters) {
        counters->steps_accepted = hc[0];
        counters->steps_rejected = hc[1];
        counters->rhs_evaluations = hc[2];
        counters->segments = hc[3];
        counters->trials_failed = hc[4];
        counters->nodes = hc[5];
    }
    return 0;
}

int glc_stream_finish(glc_evolver *ev, glc_counters *counters) {
    if (!ev || !ev->stream_active) return -1;
    cudaSetDevice(ev->device);
    unsigned long long hc[16] = {0};
    if (ev->stream_n > 0) {
        int rc = launch_machine(ev, (int)ev->stream_n, hc, 2, 0);
        if (rc) return rc;
    }
    if (counters) {
        counters->steps_accepted = hc[0];
        counters->steps_rejected = hc[1];
        counters->rhs_evaluations = hc[2];
        counters->segments = hc[3];
        counters->trials_failed = hc[4];
        counters->nodes = hc[5];
    }
    return 0;
}

__global__ void collect_list_kernel(const int32_t *__restrict__ status, unsigned char *collected, int64_t n, int64_t maxNodes,
                                    int64_t *list) {
    // list[0] = count, list[1..] = tickets
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (status[i] != GLC_STATUS_PENDING && !collected[i]) {
            const unsigned long long k = atomicAdd(reinterpret_cast<unsigned long long *>(list), 1ull);
            if ((int64_t)k < maxNodes) {
                list[1 + k] = i;
                collected[i] = 1;
            }
        }
}
__global__ void collect_gather_kernel(const double *__restrict__ soa, int64_t cap, const int64_t *__restrict__ list, int64_t m,
                                      double *__restrict__ rows, const int32_t *__restrict__ flags,
                                      const int32_t *__restrict__ status, const int32_t *__restrict__ interrupt,
                                      int32_t *__restrict__ meta) {
    // one warp per collected node: rows[k][NPROP] and meta[k][3] = {flags, status, interrupt}
    const int64_t k = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= m) return;
    const int64_t node = list[1 + k];
    for (int p = lane; p < NPROP; p += 32) rows[k * NPROP + p] = soa[(int64_t)p * cap + node];
    if (lane == 0) {
        meta[3 * k + 0] = flags[node];
        meta[3 * k + 1] = status[node];
        meta[3 * k + 2] = interrupt[node];
    }
}

int glc_stream_collect(glc_evolver *ev, int64_t max_nodes, int64_t *tickets, double *props, int32_t *flags, int32_t *status,
                       int32_t *interrupt, int64_t *n_out) {
    if (!ev || !ev->stream_active || max_nodes < 0 || !tickets || !props || !flags || !status || !interrupt || !n_out) return -1;
    *n_out = 0;
    if (max_nodes == 0 || ev->stream_n == 0) return 0;
    cudaSetDevice(ev->device);
    GLC_CHECK(ev, cudaMemsetAsync(ev->d_collect_list, 0, sizeof(int64_t), ev->stream));
    const int grid = (int)std::min<int64_t>((ev->stream_n + 255) / 256, (int64_t)ev->num_sms * 16);
    collect_list_kernel<<<grid, 256, 0, ev->stream>>>(ev->d_status, ev->d_collected, ev->stream_n, max_nodes, ev->d_collect_list);
    int64_t count = 0;
    GLC_CHECK(ev, cudaMemcpyAsync(&count, ev->d_collect_list, sizeof(int64_t), cudaMemcpyDeviceToHost, ev->stream));
    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
    const int64_t m = std::min(count, max_nodes);
    ev->launches++;
    if (m == 0) return 0;
    // rows are staged in d_stage (node-major, sized for the whole arena); meta in a buffer that grows geometrically
    if (m > ev->collect_meta_cap) {
        cudaFree(ev->d_collect_meta);
        cudaFreeHost(ev->h_collect_meta);
        ev->d_collect_meta = ev->h_collect_meta = nullptr;
        const int64_t want = std::max<int64_t>(2 * m, 1 << 16);
        GLC_CHECK(ev, cudaMalloc(&ev->d_collect_meta, sizeof(int32_t) * 3 * (size_t)want));
        GLC_CHECK(ev, cudaHostAlloc(&ev->h_collect_meta, sizeof(int32_t) * 3 * (size_t)want, cudaHostAllocDefault));
        ev->collect_meta_cap = want;
    }
    int32_t *d_meta = ev->d_collect_meta;
    collect_gather_kernel<<<(int)((m * 32 + 255) / 256), 256, 0, ev->stream>>>(ev->d_props, ev->cap, ev->d_collect_list, m, ev->d_stage,
                                                                              ev->d_flags, ev->d_status, ev->d_interrupt, d_meta);
    ev->launches++;
    GLC_CHECK(ev, cudaGetLastError());
    int32_t *meta = ev->h_collect_meta;
    GLC_CHECK(ev, cudaMemcpyAsync(props, ev->d_stage, sizeof(double) * NPROP * m, cudaMemcpyDeviceToHost, ev->stream));
    GLC_CHECK(ev, cudaMemcpyAsync(tickets, ev->d_collect_list + 1, sizeof(int64_t) * m, cudaMemcpyDeviceToHost, ev->stream));
    GLC_CHECK(ev, cudaMemcpyAsync(meta, d_meta, sizeof(int32_t) * 3 * m, cudaMemcpyDeviceToHost, ev->stream));
    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
    for (int64_t k = 0; k < m; k++) {
        flags[k] = meta[3 * k + 0];
        status[k] = meta[3 * k + 1];
        interrupt[k] = meta[3 * k + 2];
    }
    ev->stream_collected += m;
    *n_out = m;
    return 0;
}

int glc_stream_end(glc_evolver *ev) {
    if (!ev) return -1;
    ev->stream_active = false;
    ev->stream_started = false;
    ev->stream_n = 0;
    return 0;
}

int glc_histogram_accumulate(glc_evolver *ev, int64_t n, int32_t prop, double log10_min,
                             double log10_max, int32_t n_bins, double *device_hist) {
    if (!ev || n < 0 || n > ev->cap || prop < 0 || prop >= NPROP || n_bins < 1 || !device_hist) return -1;
    if (n == 0) return 0;
    cudaSetDevice(ev->device);
    histogram_kernel<<<(int)((n + 255) / 256), 256, 0, ev->stream>>>(
        ev->d_props + (int64_t)prop * ev->cap, (int)n, log10_min, log10_max, n_bins, device_hist);
    GLC_CHECK(ev, cudaGetLastError());
    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
    return 0;
}

int glc_profiler_reset(glc_evolver *ev) {
    if (!ev) return -1;
    if (!ev->d_profile) return 0;
    cudaSetDevice(ev->device);
    GLC_CHECK(ev, cudaMemsetAsync(ev->d_profile, 0, sizeof(unsigned long long) * kProfWords, ev->stream));
    GLC_CHECK(ev, cudaMemsetAsync(ev->d_profile + kProfSmallest, 0x7f, sizeof(unsigned long long), ev->stream));
    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
    return 0;
}

int glc_profiler_read(glc_evolver *ev, glc_profile *out) {
    if (!ev || !out) return -1;
    memset(out, 0, sizeof(*out));
    if (!ev->d_profile || !ev->params.profileOdeEvolver) {
        ev->err = "profileOdeEvolver is not set";
        return -9;
    }
    cudaSetDevice(ev->device);
    unsigned long long h[kProfWords];
    GLC_CHECK(ev, cudaMemcpyAsync(h, ev->d_profile, sizeof h, cudaMemcpyDeviceToHost, ev->stream));
    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
    out->n_bins = ev->tables.profBins;
    for (int i = 0; i < GLC_PROFILE_BINS; i++) {
        out->time_step[i] = ev->tables.profEdges[i];
        out->time_step_count[i] = h[0 * GLC_PROFILE_BINS + i];
        out->evaluation_count[i] = h[1 * GLC_PROFILE_BINS + i];
        out->time_step_count_interrupted[i] = h[2 * GLC_PROFILE_BINS + i];
        out->evaluation_count_interrupted[i] = h[3 * GLC_PROFILE_BINS + i];
    }
    for (int i = 0; i < GLC_NY; i++) out->property_hits[i] = h[kProfHits + i];
    out->property_hits_unknown = h[kProfUnknown];
    memcpy(&out->time_step_smallest, &h[kProfSmallest], sizeof(double));
    return 0;
}

int glc_params_default(glc_params *P, int32_t model);  // defined in glc_params.cpp

}  // extern "C"

// ---------------------------------------------------------------- forest interface (host/glc_forest.hpp)
int glc_forest_evolve(glc_evolver *ev, int64_t n_nodes, const int32_t *parent, const double *mass, const double *time,
                      const double *scale_radius, const double *angular_momentum, double *records, int32_t *flags,
                      int32_t *state, glc_forest_counters *forest_counters, glc_counters *counters) {
    if (!ev || n_nodes < 1 || !parent || !mass || !time || !scale_radius || !angular_momentum || !records || !flags || !state)
        return -1;
    if (!ev->params_set || ev->params.model != GLC_MODEL_STANDARD || !ev->params.resolveInterruptsOnDevice) {
        ev->err = "glc_forest_evolve needs the standard model with resolveInterruptsOnDevice = 1";
        return -9;
    }
    if (ev->halo_host.n0 < 2) {
        ev->err = "glc_forest_evolve needs GLC_TABLE_HALO_MEAN_DENSITY";
        return -9;
    }
    if (n_nodes > 0x7fffffff) {
        ev->err = "glc_forest_evolve: more than 2^31-1 nodes in one call";
        return GLC_ERR_BAD_FOREST;
    }
    {
        // the parent array must describe a forest: indices in [-1, n), no cycles (every walk towards a root ends)
        std::vector<int32_t> mark((size_t)n_nodes, 0);  // 0 = unseen, k > 0 = reached in walk k, -1 = known to end at a root
        for (int64_t i = 0; i < n_nodes; i++) {
            if (parent[i] < -1 || parent[i] >= n_nodes) {
                ev->err = "glc_forest_evolve: parent index out of range";
                return GLC_ERR_BAD_FOREST;
            }
        }
        for (int64_t i = 0; i < n_nodes; i++) {
            if (mark[i] != 0) continue;
            const int32_t walk = (int32_t)(i % 0x7ffffffe) + 1;
            int64_t q = i;
            while (q >= 0 && mark[q] == 0) {
                mark[q] = walk;
                q = parent[q];
            }
            if (q >= 0 && mark[q] == walk) {
                ev->err = "glc_forest_evolve: the parent array contains a cycle";
                return GLC_ERR_BAD_FOREST;
            }
            for (q = i; q >= 0 && mark[q] == walk; q = parent[q]) mark[q] = -1;
        }
    }
    glcf::Forest F;
    F.init(&ev->params, &ev->halo_host, n_nodes, parent, mass, time, scale_radius, angular_momentum, records, flags, state);
    glc_counters total{};
    const bool forest_log = getenv("GLC_FOREST_LOG") != nullptr;
    std::vector<double, PinnedAllocator<double>> buf, tpin;
    std::vector<int32_t, PinnedAllocator<int32_t>> bflags, status, interrupt;
    auto evolve = [&](const std::vector<int32_t> &list, const std::vector<double> &te_in) -> int {
        const int64_t m = (int64_t)list.size();
        if ((size_t)m * GLC_NPROP > buf.capacity()) {  // grow geometrically: page-locking is expensive
            const size_t cap = std::max<size_t>((size_t)m + (size_t)m / 4, 1024);
            buf.reserve(cap * GLC_NPROP);
            tpin.reserve(cap);
            bflags.reserve(cap);
            status.reserve(cap);
            interrupt.reserve(cap);
        }
        buf.resize((size_t)m * GLC_NPROP);
        tpin.assign(te_in.begin(), te_in.end());
        const auto &te = tpin;
        bflags.resize(m);
        status.resize(m);
        interrupt.resize(m);
        for (int64_t k = 0; k < m; k++) {
            memcpy(&buf[(size_t)k * GLC_NPROP], F.R(list[k]), sizeof(double) * GLC_NPROP);
            bflags[k] = flags[list[k]];
        }
        glc_counters c{};
        const double t_batch = now_s();
        static int n_batches = 0;
        if (const char *mb = getenv("GLC_FOREST_MAX_BATCHES"))  // debugging aid: stop after that many batched calls
            if (n_batches++ >= atoi(mb)) return GLC_ERR_BUSY;
        std::vector<double> input;
        const char *dumpPath = getenv("GLC_DUMP_PENDING");
        if (dumpPath) input.assign(buf.begin(), buf.end());
        int rc = glc_evolve_batch(ev, m, buf.data(), bflags.data(), te.data(), status.data(), interrupt.data(), &c);
        if (rc && dumpPath) {
            // debugging aid: the input records of the nodes that did not come back
            if (FILE *f = fopen(dumpPath, "wb")) {
                int64_t cnt = 0;
                for (int64_t k = 0; k < m; k++) cnt += status[k] != GLC_STATUS_SUCCESS;
                const int64_t head[3] = {cnt, GLC_NPROP, m};
                fwrite(head, sizeof(int64_t), 3, f);
                for (int64_t k = 0; k < m; k++)
                    if (status[k] != GLC_STATUS_SUCCESS) {
                        const int64_t kk = k;
                        fwrite(&kk, sizeof(int64_t), 1, f);
                        fwrite(&status[k], sizeof(int32_t), 1, f);
                        fwrite(&flags[list[k]], sizeof(int32_t), 1, f);
                        fwrite(&te[k], sizeof(double), 1, f);
                        fwrite(&input[(size_t)k * GLC_NPROP], sizeof(double), GLC_NPROP, f);
                    }
                fclose(f);
            }
        }
        if (rc) return rc;
        if (forest_log)
            fprintf(stderr, "[glc forest] batch of %lld nodes: %.1f ms (kernels %.1f ms), %llu RHS evaluations, %llu accepted steps\n",
                    (long long)m, 1e3 * (now_s() - t_batch), ev->last_ms, (unsigned long long)c.rhs_evaluations,
                    (unsigned long long)c.steps_accepted);
        for (int64_t k = 0; k < m; k++) {
            memcpy(F.R(list[k]), &buf[(size_t)k * GLC_NPROP], sizeof(double) * GLC_NPROP);
            if (status[k] != GLC_STATUS_SUCCESS || interrupt[k] != GLC_INT_NONE) {
                // The reference aborts the run here or, with tolerateFailures, drops the tree (tasks/evolve_forests/
                // _class.F90:887-897).  The walk goes on with the node moved to its end time so that the other trees finish;
                // the call then returns GLC_WARN_EVOLVE_FAILED and the count is in failed_evolves.
                F.fc.failed_evolves++;
                F.R(list[k])[GLC_P_TIME] = te[k];
            }
            flags[list[k]] = bflags[k];
        }
        total.steps_accepted += c.steps_accepted;
        total.steps_rejected += c.steps_rejected;
        total.rhs_evaluations += c.rhs_evaluations;
        total.segments += c.segments;
        total.trials_failed += c.trials_failed;
        total.nodes += c.nodes;
        return 0;
    };
    int rc;
    const char *envAsync = getenv("GLC_FOREST_ASYNC");
    const bool asynchronous = (envAsync ? atoi(envAsync) != 0 : ev->forest_schedule != 0) && ev->use_machine != 0 && !ev->stream_active;
    if (asynchronous) {
        // ---- asynchronous groups over the streaming machine (Forest::run_async): the device never drains between batches
        struct StreamEngine {
            glc_evolver *ev;
            glcf::Forest &F;
            glc_counters &total;
            bool log;
            int64_t capacity = 0, inflight = 0, submitted = 0, collected = 0;
            std::vector<int32_t> ticket_node;  // ticket of the current session -> node
            std::vector<int32_t> q_nodes;
            std::vector<double, PinnedAllocator<double>> q_props, q_tend, c_props;
            std::vector<int32_t, PinnedAllocator<int32_t>> q_flags, c_flags, c_status, c_interrupt;
            std::vector<int64_t, PinnedAllocator<int64_t>> c_tickets;
            std::vector<float> q_prio;
            glc_counters session{};
            double t0 = now_s(), t_run = 0.0, t_collect = 0.0, t_submit = 0.0;
            int64_t polls = 0;
            StreamEngine(glc_evolver *e, glcf::Forest &f, glc_counters &t, bool l) : ev(e), F(f), total(t), log(l) {}
            int begin() {
                // tickets are arena rows: room for several evolve calls per node; a full arena restarts the session (finish,
                // collect, begin) -- see flush()
                const int64_t want = std::min<int64_t>(std::max<int64_t>(8 * F.n + 65536, 1 << 20), 48ll << 20);
                capacity = want;
                ticket_node.clear();
                submitted = collected = 0;
                memset(&session, 0, sizeof(session));
                return glc_stream_begin(ev, capacity);
            }
            void submit(int32_t node, double tend) {
                q_nodes.push_back(node);
                q_tend.push_back(tend);
                inflight++;
            }
            int64_t in_flight() const { return inflight; }
            void add_session() {
                total.steps_accepted += session.steps_accepted;
                total.steps_rejected += session.steps_rejected;
                total.rhs_evaluations += session.rhs_evaluations;
                total.segments += session.segments;
                total.trials_failed += session.trials_failed;
                total.nodes += session.nodes;
            }
            int collect(std::vector<int32_t> &done, std::vector<int32_t> &st, std::vector<int32_t> &in) {
                const int64_t outstanding = submitted - collected;
                if (outstanding <= 0) return 0;
                if ((int64_t)c_tickets.size() < outstanding) {
                    const size_t cap = (size_t)std::max<int64_t>(2 * outstanding, 1 << 16);
                    c_tickets.resize(cap);
                    c_props.resize(cap * GLC_NPROP);
                    c_flags.resize(cap);
                    c_status.resize(cap);
                    c_interrupt.resize(cap);
                }
                int64_t nout = 0;
                int rc = glc_stream_collect(ev, outstanding, c_tickets.data(), c_props.data(), c_flags.data(), c_status.data(),
                                            c_interrupt.data(), &nout);
                if (rc) return rc;
                for (int64_t k = 0; k < nout; k++) {
                    const int32_t node = ticket_node[(size_t)c_tickets[k]];
                    memcpy(F.R(node), &c_props[(size_t)k * GLC_NPROP], sizeof(double) * GLC_NPROP);
                    F.flags[node] = c_flags[k];
                    done.push_back(node);
                    st.push_back(c_status[k]);
                    in.push_back(c_interrupt[k]);
                }
                collected += nout;
                inflight -= nout;
                return 0;
            }
            int flush() {
                const int64_t m = (int64_t)q_nodes.size();
                if (m == 0) return 0;
                const double t = now_s();
                if (submitted + m > capacity) {
                    // arena full: run everything in flight to completion, hand it back at the next poll, start a new session
                    restart = true;
                    return 0;
                }
                q_props.resize((size_t)m * GLC_NPROP);
                q_flags.resize(m);
                for (int64_t k = 0; k < m; k++) {
                    memcpy(&q_props[(size_t)k * GLC_NPROP], F.R(q_nodes[k]), sizeof(double) * GLC_NPROP);
                    q_flags[k] = F.flags[q_nodes[k]];
                    ticket_node.push_back(q_nodes[k]);
                }
                int64_t first = 0;
                int rc = glc_stream_submit(ev, m, q_props.data(), q_flags.data(), q_tend.data(), &first);
                if (rc) return rc;
                if (ev->stream_priority_express > 0) {
                    // priority of a ticket = the halo mass of its node: the main branches are what every tree waits for
                    if (ev->stream_priority_cap < ev->cap) {
                        cudaFree(ev->d_stream_priority);
                        ev->d_stream_priority = nullptr;
                        if (cudaMalloc(&ev->d_stream_priority, sizeof(float) * (size_t)ev->cap) != cudaSuccess) return -2;
                        ev->stream_priority_cap = ev->cap;
                    }
                    q_prio.resize(m);
                    for (int64_t k = 0; k < m; k++) q_prio[k] = (float)F.R(q_nodes[k])[GLC_P_BASIC_MASS];
                    if (cudaMemcpyAsync(ev->d_stream_priority + first, q_prio.data(), sizeof(float) * m, cudaMemcpyHostToDevice, ev->stream) != cudaSuccess) return -2;
                    cudaStreamSynchronize(ev->stream);
                }
                submitted += m;
                q_nodes.clear();
                q_tend.clear();
                t_submit += now_s() - t;
                return 0;
            }
            bool restart = false;
            int poll(std::vector<int32_t> &done, std::vector<int32_t> &st, std::vector<int32_t> &in) {
                polls++;
                if (restart) {
                    int rc = glc_stream_finish(ev, &session);
                    if (rc) return rc;
                    rc = collect(done, st, in);
                    if (rc) return rc;
                    add_session();
                    glc_stream_end(ev);
                    rc = begin();
                    if (rc) return rc;
                    restart = false;
                    return flush();  // the held-back submissions open the new session
                }
                for (int64_t idle = 0;; idle++) {
                    if (idle > 200000) {
                        ev->err = "glc_forest_evolve: 200 000 device time slices without a finished node";
                        return GLC_ERR_STALLED;
                    }
                    // one time slice; shorter slices when few nodes are in flight, so that finished nodes are seen early
                    int64_t nfin = 0;
                    double t = now_s();
                    int rc = glc_stream_run(ev, 0, &nfin, &session);  // adaptive tick: machine slice or lane pass
                    if (rc) return rc;
                    t_run += now_s() - t;
                    t = now_s();
                    if (nfin > collected) {
                        rc = collect(done, st, in);
                        if (rc) return rc;
                    }
                    t_collect += now_s() - t;
                    if (!done.empty()) return 0;
                }
            }
            int end() {
                add_session();
                if (log) {
                    fprintf(stderr, "[glc forest async] %lld polls, %.2f s in device slices, %.2f s collecting, %.2f s submitting, %.2f s total\n",
                            (long long)polls, t_run, t_collect, t_submit, now_s() - t0);
                    const int64_t ticks = ev->tick_machine + ev->tick_lane;
                    fprintf(stderr, "[glc forest async] ticks: %lld machine slices (%.2f s), %lld lane passes (%.2f s, mean %.1f nodes per warp, %.0f express warps), %lld hold slices; mean nodes queued or in flight per tick %.0f\n",
                            (long long)ev->tick_machine, ev->tick_machine_s, (long long)ev->tick_lane, ev->tick_lane_s,
                            ev->tick_lane ? ev->tick_lanes_sum / (double)ev->tick_lane : 0.0,
                            ev->tick_lane ? ev->tick_express_sum / (double)ev->tick_lane : 0.0, (long long)ev->tick_hold,
                            ticks ? ev->tick_live_sum / (double)ticks : 0.0);
                }
                return glc_stream_end(ev);
            }
        };
        StreamEngine E(ev, F, total, forest_log);
        rc = E.begin();
        if (rc == 0) rc = F.run_async(E);
        const int rc2 = E.end();
        if (rc == 0) rc = rc2;
    } else
        rc = F.run(evolve);
    if (forest_counters) *forest_counters = F.fc;
    if (counters) *counters = total;
    if (rc == 0 && !F.all_roots_finished()) {
        // no node can move although a tree has not reached its final time: the reference's deadlock report
        // (merger_trees/evolver/standard.F90:606-625)
        ev->err = "glc_forest_evolve: deadlock -- trees not at their final time although no node can be evolved";
        return GLC_ERR_DEADLOCK;
    }
    if (rc == 0 && F.fc.failed_evolves > 0) {
        ev->err = "glc_forest_evolve: node evolves came back with a status other than success (see failed_evolves)";
        return GLC_WARN_EVOLVE_FAILED;
    }
    return rc;
}
