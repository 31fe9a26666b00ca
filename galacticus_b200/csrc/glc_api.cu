// glc_api.cu -- C-ABI of libglcb200 (include/glc_b200.h): context, tables, arena, launches.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "glc_common.cuh"

#include "glc_evolve_kernel.cuh"
#include "glc_model_box.cuh"
#include "glc_model_standard.cuh"

using namespace glc;

namespace {

constexpr int kBlock = 128;

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
    float last_ms = 0.f;
    std::string err;
};

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
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= A.n) return;
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
    const int code = Model::rates(ctx, time, y, rate);
    const uint32_t mask = Model::active_mask(ctx.flags);
    for (int i = 0; i < NY; i++) dydt[(int64_t)node * NY + i] = (mask & (1u << i)) ? rate[i] : rate[i];
    A.interrupt[node] = code;
    AR(GLC_P_DISK_RADIUS) = ctx.diskRadius;
    AR(GLC_P_DISK_VELOCITY) = ctx.diskVelocity;
    AR(GLC_P_SPH_RADIUS) = ctx.sphRadius;
    AR(GLC_P_SPH_VELOCITY) = ctx.sphVelocity;
    AR(GLC_P_BASIC_MASS) = ctx.basicMass;
}

// FP64 FMA-chain microbenchmark (8 independent chains per thread): the measured FP64 roofline denominator
__global__ void fp64_peak_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1.0e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1.0e-7;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c);
        a1 = fma(a1, b, c);
        a2 = fma(a2, b, c);
        a3 = fma(a3, b, c);
        a4 = fma(a4, b, c);
        a5 = fma(a5, b, c);
        a6 = fma(a6, b, c);
        a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// debugging aid for the bit-exact parity work: selected intermediates of one RHS evaluation
__global__ void probe_kernel(KernelArgs A, double *out) {
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= A.n) return;
    auto AR = [&](int prop) -> double & { return A.props[(int64_t)prop * A.cap + node]; };
    NodeCtx c;
    double y[NY];
    for (int i = 0; i < NY; i++) y[i] = AR(i);
    c.flags = A.flags[node];
    c.massTarget = AR(GLC_P_MASS_TARGET);
    c.massRate = AR(GLC_P_MASS_RATE);
    c.timeTarget = AR(GLC_P_TIME_TARGET);
    c.scaleTarget = AR(GLC_P_DMSCALE_TARGET);
    c.scaleRate = AR(GLC_P_DMSCALE_RATE);
    c.spinTarget = AR(GLC_P_SPIN_TARGET);
    c.spinRate = AR(GLC_P_SPIN_RATE);
    c.timeLastIsolated = AR(GLC_P_TIME_LAST_ISOLATED);
    c.diskRadius = AR(GLC_P_DISK_RADIUS);
    c.diskVelocity = AR(GLC_P_DISK_VELOCITY);
    c.sphRadius = AR(GLC_P_SPH_RADIUS);
    c.sphVelocity = AR(GLC_P_SPH_VELOCITY);
    c.basicMass = AR(GLC_P_BASIC_MASS);
    c.massBaryonicSubhalos = AR(GLC_P_MASS_BARYONIC_SUBHALOS);
    c.numericsFailed = 0;
    const double time = AR(GLC_P_TIME);
    c.timeNode = time;
    typedef ModelStandard M;
    M::solve_analytics(c, time);
    Work w;
    int bad = 0;
    M::halo_scales(c, time, w);
    M::hh_profile(c, y, w);
    const double r0 = c.diskRadius > 0.0 ? c.diskRadius : 0.01 * w.rvir;
    const double nn = M::nfw_norm(c, w);
    double *o = out + (int64_t)node * 16;
    o[0] = w.rvir;
    o[1] = w.vvir;
    o[2] = w.tvir;
    o[3] = w.hhRho0;
    o[4] = M::nfw_mass(nn, c.dmScale, r0);
    o[5] = M::ac_orbital_mean(w, r0);
    o[6] = M::baryonic_vc2(c, y, w, r0);
    o[7] = M::dark_matter_mass_enclosed(c, y, w, nn, r0, bad);
    o[8] = M::disk_bessel_factor(0.37);
    o[9] = M::hh_mass_enclosed(w, r0);
    o[10] = fast_exponentiate(1.0e-3, 1.0, 0.7, 1.0e4, 0.0123);
    o[11] = M::nfw_radius_from_j(c, w, nn, 0.3 * w.rvir * w.vvir);
    o[12] = (c.flags & GLC_F_HAS_DISK) ? M::sfr_disk(c, y, bad) : 0.0;
    double ls;
    if ((c.flags & GLC_F_HAS_HOTHALO) && y[GLC_P_HH_MASS] > 0) {
        M::cooling_prepare(y, w, ls);
        o[13] = M::cooling_radius(y, w, bad);
    } else
        o[13] = 0.0;
    o[14] = sqrt(kGInternal * o[7] / r0 + o[6]);
    o[15] = dm_log(o[14] / r0);
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

// ---------------------------------------------------------------------------- helpers
static int ensure_workspace(glc_evolver *ev, int grid) {
    const int64_t need = (int64_t)grid * kBlock;
    if (need <= ev->nslots) return 0;
    if (ev->d_ws) cudaFree(ev->d_ws);
    ev->d_ws = nullptr;
    GLC_CHECK(ev, cudaMalloc(&ev->d_ws, sizeof(double) * WS_NVEC * NY * need));
    ev->nslots = need;
    return 0;
}

template <class Model>
static int launch_evolve(glc_evolver *ev, int n) {
    int blocksPerSm = 0;
    GLC_CHECK(ev, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocksPerSm, evolve_kernel<Model>,
                                                                 kBlock, 0));
    if (blocksPerSm < 1) blocksPerSm = 1;
    int grid = ev->num_sms * blocksPerSm;
    grid = std::min(grid, (n + kBlock - 1) / kBlock);
    if (grid < 1) grid = 1;
    int rc = ensure_workspace(ev, ev->num_sms * blocksPerSm);
    if (rc) return rc;
    KernelArgs A;
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
    GLC_CHECK(ev, cudaMemsetAsync(ev->d_work, 0, sizeof(int), ev->stream));
    GLC_CHECK(ev, cudaEventRecord(ev->ev0, ev->stream));
    evolve_kernel<Model><<<grid, kBlock, 0, ev->stream>>>(A);
    ev->launches++;
    GLC_CHECK(ev, cudaGetLastError());
    GLC_CHECK(ev, cudaEventRecord(ev->ev1, ev->stream));
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
    cudaStreamCreateWithFlags(&ev->stream, cudaStreamNonBlocking);
    cudaEventCreate(&ev->ev0);
    cudaEventCreate(&ev->ev1);
    cudaMalloc(&ev->d_work, sizeof(int));
    cudaMalloc(&ev->d_counters, sizeof(unsigned long long) * 8);
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
    cudaEventDestroy(ev->ev0);
    cudaEventDestroy(ev->ev1);
    cudaStreamDestroy(ev->stream);
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
    ev->params = *params;
    ev->params_set = true;
    return 0;
}

int glc_evolver_set_table(glc_evolver *ev, int32_t id, int32_t n0, int32_t n1, const double *x0,
                          const double *x1, const double *values) {
    if (!ev || id < 0 || id >= GLC_NTABLES || n0 < 2 || n1 < 1 || !x0 || !values) return -1;
    cudaSetDevice(ev->device);
    HostTable &t = ev->host_tables[id];
    cudaFree(t.d_x0);
    cudaFree(t.d_x1);
    cudaFree(t.d_v);
    t = HostTable();
    std::vector<double> hx0(x0, x0 + n0), hx1, hv(values, values + (size_t)n0 * n1);
    if (x1) hx1.assign(x1, x1 + n1);
    int is_log = 0, first_zero = 0;
    double first_nonzero = 0.0;
    const double zmin = hx0.front(), zmax = hx0.back();
    const double tmin = x1 ? hx1.front() : 0.0, tmax = x1 ? hx1.back() : 0.0;
    if (id == GLC_TABLE_COOLING_FUNCTION || id == GLC_TABLE_ELECTRON_FRACTION) {
        // cieFileReadFile, cooling/cooling_function/CIE_file.F90:627-659
        if (!x1 || n1 < 2) return -1;
        is_log = 1;
        for (double v : hv)
            if (!(v > 0.0)) is_log = 0;
        if (is_log) {
            first_zero = (hx0[0] == 0.0);
            if (first_zero) first_nonzero = hx0[1];
            for (auto &z : hx0) z = (z > 0.0) ? dm_log(z) : -999.0;
            for (auto &T : hx1) T = dm_log(T);
            for (auto &v : hv) v = dm_log(v);
        }
    }
    GLC_CHECK(ev, cudaMalloc(&t.d_x0, sizeof(double) * n0));
    GLC_CHECK(ev, cudaMemcpy(t.d_x0, hx0.data(), sizeof(double) * n0, cudaMemcpyHostToDevice));
    if (x1) {
        GLC_CHECK(ev, cudaMalloc(&t.d_x1, sizeof(double) * n1));
        GLC_CHECK(ev, cudaMemcpy(t.d_x1, hx1.data(), sizeof(double) * n1, cudaMemcpyHostToDevice));
    }
    GLC_CHECK(ev, cudaMalloc(&t.d_v, sizeof(double) * (size_t)n0 * n1));
    GLC_CHECK(ev, cudaMemcpy(t.d_v, hv.data(), sizeof(double) * (size_t)n0 * n1, cudaMemcpyHostToDevice));
    t.n0 = n0;
    t.n1 = n1;
    DeviceTable2D d{n0, n1, t.d_x0, t.d_x1, t.d_v};
    if (id == GLC_TABLE_COOLING_FUNCTION) {
        ev->tables.cooling = d;
        ev->tables.cooling_log = is_log;
        ev->tables.cooling_first_z_zero = first_zero;
        ev->tables.cooling_first_nonzero_z = first_nonzero;
        ev->tables.cooling_z_min = zmin;
        ev->tables.cooling_z_max = zmax;
        ev->tables.cooling_t_min = tmin;
        ev->tables.cooling_t_max = tmax;
    } else if (id == GLC_TABLE_ELECTRON_FRACTION) {
        ev->tables.electron = d;
        ev->tables.electron_log = is_log;
        ev->tables.electron_first_z_zero = first_zero;
        ev->tables.electron_first_nonzero_z = first_nonzero;
        ev->tables.electron_z_min = zmin;
        ev->tables.electron_z_max = zmax;
        ev->tables.electron_t_min = tmin;
        ev->tables.electron_t_max = tmax;
    } else if (id == GLC_TABLE_HALO_MEAN_DENSITY) {
        if (n1 != 2) return -1;
        // store ln t on the device; the grid must be log-uniform
        std::vector<double> lnt(n0);
        for (int i = 0; i < n0; i++) lnt[i] = dm_log(hx0[i]);
        GLC_CHECK(ev, cudaMemcpy(t.d_x0, lnt.data(), sizeof(double) * n0, cudaMemcpyHostToDevice));
        ev->tables.density = d;
        ev->tables.density_lnt0 = lnt[0];
        ev->tables.density_inv_dlnt = (double)(n0 - 1) / (lnt[n0 - 1] - lnt[0]);
    } else if (id == GLC_TABLE_DISK_ROTATION_CURVE) {
        if (n1 != 1) return -1;
        ev->tables.diskrc = d;
        ev->tables.diskrc_lnx0 = dm_log(hx0[0]);
        ev->tables.diskrc_inv_dlnx = (double)(n0 - 1) / (dm_log(hx0[n0 - 1]) - dm_log(hx0[0]));
    }
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
    GLC_CHECK(ev, cudaMemsetAsync(ev->d_counters, 0, sizeof(unsigned long long) * 8, ev->stream));
    if (ev->params.model == GLC_MODEL_BOX)
        rc = launch_evolve<ModelBox>(ev, (int)n);
    else
        rc = launch_evolve<ModelStandard>(ev, (int)n);
    if (rc) return rc;
    unsigned long long hc[8];
    GLC_CHECK(ev, cudaMemcpyAsync(hc, ev->d_counters, sizeof(hc), cudaMemcpyDeviceToHost, ev->stream));
    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
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
int64_t glc_kernel_launch_count(const glc_evolver *ev) { return ev ? ev->launches : 0; }

double glc_measure_fp64_peak_tflops(glc_evolver *ev) {
    if (!ev) return 0.0;
    cudaSetDevice(ev->device);
    double *d_out = nullptr;
    const int grid = ev->num_sms * 8, block = 256, iters = 4096;
    if (cudaMalloc(&d_out, sizeof(double) * grid * block) != cudaSuccess) return 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, ev->stream);
        fp64_peak_kernel<<<grid, block, 0, ev->stream>>>(d_out, iters);
        cudaEventRecord(e1, ev->stream);
        cudaStreamSynchronize(ev->stream);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8.0 * (double)iters * (double)grid * block;  // 8 independent FMA chains
        if (ms > 0.f) best = std::max(best, flops / (ms * 1.0e-3) / 1.0e12);
    }
    ev->launches += 5;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    return best;
}

float glc_last_kernel_ms(const glc_evolver *ev) { return ev ? ev->last_ms : 0.f; }
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
    if (rc) return rc;
    return glc_arena_download(ev, n, props, flags, status, interrupt);
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

int glc_params_default(glc_params *P, int32_t model);  // defined in glc_params.cpp

// not part of the public header: debugging aid (16 intermediates per node)
int glc_debug_probe(glc_evolver *ev, int64_t n, const double *props, const int32_t *flags, double *out) {
    std::vector<double> te((size_t)n, 0.0);
    int rc = glc_arena_upload(ev, n, props, flags, te.data());
    if (rc) return rc;
    rc = upload_constants(ev);
    if (rc) return rc;
    double *d_out = nullptr;
    GLC_CHECK(ev, cudaMalloc(&d_out, sizeof(double) * 16 * n));
    KernelArgs A{};
    A.props = ev->d_props;
    A.flags = ev->d_flags;
    A.cap = ev->cap;
    A.n = (int)n;
    probe_kernel<<<(int)((n + 63) / 64), 64, 0, ev->stream>>>(A, d_out);
    GLC_CHECK(ev, cudaGetLastError());
    GLC_CHECK(ev, cudaMemcpyAsync(out, d_out, sizeof(double) * 16 * n, cudaMemcpyDeviceToHost, ev->stream));
    GLC_CHECK(ev, cudaStreamSynchronize(ev->stream));
    cudaFree(d_out);
    return 0;
}

}  // extern "C"
