// glc_emu.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the CUDA kernels' per-lane logic (galacticus_b200/csrc/*.cuh: the lane state machine, the rate
// functions, the nested numerics) with g++ through the platform layer of glc_common.cuh, and drives it the way
// the evolve kernel does: warps of 32 lanes in lock-step, one heavy call per lane per iteration, a shared node
// queue in component-sorted order, optional time slices with lane states parked between them.  This lets the
// CPU-only test suite (-m "not gpu") check the kernel source bit-for-bit against the oracle, including the
// refill / time-slice / resume logic, without a GPU.  It is never loaded by the product path (the product
// fails loudly when CUDA is missing) and is not a fallback: no entry point of include/glc_b200.h is served
// from here.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <random>
#include <vector>

#include "../../galacticus_b200/csrc/glc_common.cuh"
#include "../../galacticus_b200/csrc/glc_tables_host.h"

#include "../../galacticus_b200/csrc/glc_evolve_kernel.cuh"
#include "../../galacticus_b200/csrc/glc_model_box.cuh"
#include "../../galacticus_b200/csrc/glc_model_standard.cuh"
#include "../../galacticus_b200/csrc/glc_machine.cuh"

using namespace glc;

#include <cstdio>
#include <cstdlib>
#include "../../galacticus_b200/csrc/host/glc_forest.hpp"

struct Emu {
    glcf::HaloTable halo_host;
    glc_params params;
    PreparedTable tables[GLC_NTABLES];
    DeviceTables dt;
    std::vector<double> powAc, powKmt, nfwJx, nfwJv;
    std::vector<unsigned long long> profile;
};

template <class Model>
static void run(Emu *e, int64_t n, double *props, int32_t *flags, const double *time_end, int32_t *status,
                int32_t *interrupt, glc_counters *counters, int nslots, int budget, int sort, int64_t *slices) {
    const int64_t cap = n;
    std::vector<double> soa((size_t)NPROP * cap);
    for (int64_t i = 0; i < n; i++)
        for (int p = 0; p < NPROP; p++) soa[(size_t)p * cap + i] = props[i * NPROP + p];
    std::vector<double> ws((size_t)WS_NVEC * NY * nslots);
    std::vector<LaneState> lanes(nslots);
    std::vector<int32_t> order;
    if (sort) {
        order.resize(n);
        for (int64_t i = 0; i < n; i++) order[i] = (int32_t)i;
        std::stable_sort(order.begin(), order.end(),
                         [&](int32_t a, int32_t b) { return queue_bucket(flags[a]) < queue_bucket(flags[b]); });
    }
    int work = 0;
    unsigned long long hc[8] = {0};
    KernelArgs A;
    A.props = soa.data();
    A.flags = flags;
    A.time_end = time_end;
    A.status = status;
    A.interrupt = interrupt;
    A.cap = cap;
    A.n = (int)n;
    A.ws = ws.data();
    A.nslots = nslots;
    A.work_counter = &work;
    A.counters = hc;
    A.order = sort ? order.data() : nullptr;
    A.lanes = lanes.data();
    A.resume = 0;
    A.budget = budget > 0 ? budget : 0x7fffffff;
    *slices = 0;
    for (;;) {
        for (int w0 = 0; w0 < nslots; w0 += 32) {
            const int w1 = std::min(w0 + 32, nslots);
            for (int s = w0; s < w1; s++) {
                LaneState &L = lanes[s];
                if (!A.resume) lane_reset(L);
                L.nAcc = L.nRej = L.nRhs = L.nSeg = L.nTrialFail = L.nNodes = L.nDone = 0;
                if (L.phase == PH_IDLE) L.phase = PH_FETCH;
            }
            for (int it = 0; it < A.budget; ++it) {
                bool any = false;
                for (int s = w0; s < w1; s++) {
                    LaneMem M{&A, A.ws + s, A.nslots};
                    any |= lane_iterate<Model>(lanes[s], M);
                }
                if (!any) break;
            }
            for (int s = w0; s < w1; s++) {
                const LaneState &L = lanes[s];
                hc[0] += L.nAcc;
                hc[1] += L.nRej;
                hc[2] += L.nRhs;
                hc[3] += L.nSeg;
                hc[4] += L.nTrialFail;
                hc[5] += L.nNodes;
                hc[6] += L.nDone;
            }
        }
        (*slices)++;
        if (hc[6] >= (unsigned long long)n) break;
        A.resume = 1;
    }
    for (int64_t i = 0; i < n; i++)
        for (int p = 0; p < NPROP; p++) props[i * NPROP + p] = soa[(size_t)p * cap + i];
    if (counters) {
        counters->steps_accepted = hc[0];
        counters->steps_rejected = hc[1];
        counters->rhs_evaluations = hc[2];
        counters->segments = hc[3];
        counters->trials_failed = hc[4];
        counters->nodes = hc[5];
    }
}

// The micro-task machine (glc_machine.cuh) driven on the host: every iteration each slot executes one unit; the
// order in which the slots of a block are visited is shuffled (seeded) to mimic the device's regrouping, which
// must not change any result.
static void run_machine(Emu *e, int64_t n, double *props, int32_t *flags, const double *time_end, int32_t *status,
                        int32_t *interrupt, glc_counters *counters, int nslots, int budget, int sort, int64_t *slices,
                        bool handover) {
    const int64_t cap = n;
    std::vector<double> soa((size_t)NPROP * cap);
    for (int64_t i = 0; i < n; i++)
        for (int p = 0; p < NPROP; p++) soa[(size_t)p * cap + i] = props[i * NPROP + p];
    std::vector<double> ws((size_t)WS_NVEC * NY * nslots);
    std::vector<LaneState> sL(nslots);
    std::vector<RhsState> sR(nslots);
    std::vector<RootState> sRoot(nslots);
    std::vector<double> sYt((size_t)nslots * NY);
    std::vector<QagState> sQ(nslots);
    std::vector<int> sUnit(nslots);
    // device memory comes from cudaMalloc uninitialised: poison the continuations and the workspace with
    // pseudo-random bytes so that any read-before-write in the machine shows up here as a mismatch or a hang
    {
        std::mt19937 g(987654321u);
        auto poison = [&](void *ptr, size_t bytes) {
            unsigned char *b = (unsigned char *)ptr;
            for (size_t i = 0; i < bytes; i++) b[i] = (unsigned char)(g() & 0xff);
        };
        poison(sL.data(), sL.size() * sizeof(LaneState));
        poison(sR.data(), sR.size() * sizeof(RhsState));
        poison(sRoot.data(), sRoot.size() * sizeof(RootState));
        poison(sYt.data(), sYt.size() * sizeof(double));
        poison(sQ.data(), sQ.size() * sizeof(QagState));
        poison(sUnit.data(), sUnit.size() * sizeof(int));
        poison(ws.data(), ws.size() * sizeof(double));
    }
    SlotArrays slots{sL.data(), sR.data(), sRoot.data(), sYt.data(), sQ.data(), sUnit.data()};
    std::vector<int32_t> order;
    if (sort) {
        order.resize(n);
        for (int64_t i = 0; i < n; i++) order[i] = (int32_t)i;
        std::stable_sort(order.begin(), order.end(),
                         [&](int32_t a, int32_t b) { return queue_bucket(flags[a]) < queue_bucket(flags[b]); });
    }
    int work = 0;
    unsigned long long hc[8] = {0};
    KernelArgs A;
    A.props = soa.data();
    A.flags = flags;
    A.time_end = time_end;
    A.status = status;
    A.interrupt = interrupt;
    A.cap = cap;
    A.n = (int)n;
    A.ws = ws.data();
    A.nslots = nslots;
    A.work_counter = &work;
    A.counters = hc;
    A.order = sort ? order.data() : nullptr;
    A.lanes = nullptr;
    A.resume = 0;
    A.budget = budget > 0 ? budget : 0x7fffffff;
    *slices = 0;
    std::mt19937 rng(12345);
    std::vector<int> perm(nslots);
    for (int s = 0; s < nslots; s++) perm[s] = s;
    for (;;) {
        for (int s = 0; s < nslots; s++) {
            const SlotRef S = slot_ref(slots, s);
            S.unit = machine_rearm(S, A.resume != 0);  // the kernel's own slice-start code
        }
        for (int it = 0; it < A.budget; ++it) {
            bool any = false;
            if (handover && work >= A.n && it >= 3) break;  // node queue dry: go to the drain hand-over below
            std::shuffle(perm.begin(), perm.end(), rng);
            for (int p = 0; p < nslots; p++) {
                const int s = perm[p];
                LaneMem M{&A, A.ws + (int64_t)s * (WS_NVEC * NY), 1};
                const SlotRef S = slot_ref(slots, s);
                any |= machine_step(S, M);
                // mimic the device's sticky mode on a pseudo-random subset of the steps
                if ((rng() & 3u) == 0u)
                    for (int k = 0; k < 64 && S.unit != U_IDLE && S.unit != U_RK; k++) machine_step(S, M);
            }
            if (!any) break;
        }
        for (int s = 0; s < nslots; s++) {
            LaneState &L = sL[s];
            hc[0] += L.nAcc;
            hc[1] += L.nRej;
            hc[2] += L.nRhs;
            hc[3] += L.nSeg;
            hc[4] += L.nTrialFail;
            hc[5] += L.nNodes;
            hc[6] += L.nDone;
            L.nAcc = L.nRej = L.nRhs = L.nSeg = L.nTrialFail = L.nNodes = L.nDone = 0;
        }
        (*slices)++;
        if (hc[6] >= (unsigned long long)n) break;
        {
            // stall guard: every slot idle although nodes are missing (would repeat empty slices for ever)
            bool occupied = false;
            for (int s = 0; s < nslots; s++) occupied |= (sUnit[s] != U_IDLE);
            if (!occupied && work >= A.n) {
                fprintf(stderr, "[emu] machine stalled: %llu of %lld nodes done, no slot occupied\n", hc[6], (long long)n);
                for (int64_t i = 0; i < n; i++) status[i] = GLC_STATUS_FAIL;
                break;
            }
        }
        A.resume = 1;
        if (handover && work >= A.n) {
            // ---- hold: bring every active slot to an RK boundary, as the device's hold slices do
            for (bool moved = true; moved;) {
                moved = false;
                for (int s = 0; s < nslots; s++) {
                    const SlotRef S = slot_ref(slots, s);
                    if (S.unit != U_IDLE && S.unit != U_RHS_BEGIN) {
                        LaneMem M{&A, A.ws + (int64_t)s * (WS_NVEC * NY), 1};
                        machine_step(S, M);
                        moved = true;
                    }
                }
            }
            for (int s = 0; s < nslots; s++) {  // counters gathered so far
                LaneState &L = sL[s];
                hc[0] += L.nAcc; hc[1] += L.nRej; hc[2] += L.nRhs; hc[3] += L.nSeg; hc[4] += L.nTrialFail; hc[5] += L.nNodes; hc[6] += L.nDone;
                L.nAcc = L.nRej = L.nRhs = L.nSeg = L.nTrialFail = L.nNodes = L.nDone = 0;
            }
            // ---- drain: the held slots are finished by whole evaluations (drain_iterate), 7 "lanes"
            std::vector<int32_t> held;
            for (int s = 0; s < nslots; s++)
                if (sUnit[s] == U_RHS_BEGIN) held.push_back(s);
            int cursor = 0;
            A.held = held.data();
            A.nheld = (int)held.size();
            A.held_counter = &cursor;
            A.slotL = sL.data();
            A.slotYt = sYt.data();
            A.slotUnit = sUnit.data();
            A.drainLanes = 0;
            const int nl = 7;
            std::vector<LaneState> dl(nl);
            std::vector<LaneMem> dm(nl, LaneMem{&A, A.ws, 1});
            std::vector<std::array<double, NY>> dyt(nl);
            std::vector<char> fresh(nl, 0);
            std::vector<int> hs(nl, -1);
            unsigned int tot[7] = {0, 0, 0, 0, 0, 0, 0};
            for (int l = 0; l < nl; l++) lane_reset(dl[l]);
            // passes of a bounded number of evaluations: unfinished nodes are parked (drain_park) and picked up
            // again by the next pass, as the device does
            for (int pass = 0; !held.empty(); pass++) {
                cursor = 0;
                A.held = held.data();
                A.nheld = (int)held.size();
                for (int l = 0; l < nl; l++) {
                    lane_reset(dl[l]);
                    hs[l] = -1;
                    fresh[l] = 0;
                }
                for (int it = 0; it < 5 + 3 * pass; it++) {
                    bool any = false;
                    for (int l = 0; l < nl; l++) {
                        bool f = fresh[l] != 0;
                        double(&y)[NY] = *reinterpret_cast<double(*)[NY]>(dyt[l].data());
                        any |= drain_iterate<ModelStandard>(dl[l], dm[l], A, y, f, hs[l], tot, true);
                        fresh[l] = f ? 1 : 0;
                    }
                    if (!any) break;
                }
                for (int l = 0; l < nl; l++) {
                    double(&y)[NY] = *reinterpret_cast<double(*)[NY]>(dyt[l].data());
                    drain_park<ModelStandard>(dl[l], dm[l], A, y, fresh[l] != 0, hs[l], tot);
                }
                held.clear();
                for (int s = 0; s < nslots; s++)
                    if (sUnit[s] == U_RHS_BEGIN) held.push_back(s);
            }
            for (int k = 0; k < 7; k++) hc[k] += tot[k];
            break;
        }
    }
    for (int64_t i = 0; i < n; i++)
        for (int p = 0; p < NPROP; p++) props[i * NPROP + p] = soa[(size_t)p * cap + i];
    if (counters) {
        counters->steps_accepted = hc[0];
        counters->steps_rejected = hc[1];
        counters->rhs_evaluations = hc[2];
        counters->segments = hc[3];
        counters->trials_failed = hc[4];
        counters->nodes = hc[5];
    }
}

// A streaming session as the product's adaptive ticks run it (glc_api.cu stream_tick), on the host: nodes are submitted in
// chunks; a tick is a machine slice ('M': a few unit executions per slot) or a lane pass ('L': every occupied slot is brought
// to an RK boundary, then drain lanes take the held slots AND free slots for queued nodes -- drainRefill -- run a bounded
// number of evaluations and park); the session is finished by the machine.  Exercises what the batch path never does: slots
// that go machine -> drain -> machine, slots released by the drain (stale lane state), refill from a growing queue.
static int run_stream_session(Emu *e, int64_t n, double *props, int32_t *flags, const double *time_end, int32_t *status,
                              int32_t *interrupt, glc_counters *counters, int nslots, int chunk, const char *pattern, int lane_budget,
                              int machine_budget) {
    const int64_t cap = n;
    std::vector<double> soa((size_t)NPROP * cap);
    for (int64_t i = 0; i < n; i++)
        for (int p = 0; p < NPROP; p++) soa[(size_t)p * cap + i] = props[i * NPROP + p];
    std::vector<double> ws((size_t)WS_NVEC * NY * nslots);
    std::vector<LaneState> sL(nslots);
    std::vector<RhsState> sR(nslots);
    std::vector<RootState> sRoot(nslots);
    std::vector<double> sYt((size_t)nslots * NY);
    std::vector<QagState> sQ(nslots);
    std::vector<int> sUnit(nslots);
    {
        std::mt19937 g(424242u);
        auto poison = [&](void *ptr, size_t bytes) {
            unsigned char *b = (unsigned char *)ptr;
            for (size_t i = 0; i < bytes; i++) b[i] = (unsigned char)(g() & 0xff);
        };
        poison(sL.data(), sL.size() * sizeof(LaneState));
        poison(sR.data(), sR.size() * sizeof(RhsState));
        poison(sRoot.data(), sRoot.size() * sizeof(RootState));
        poison(sYt.data(), sYt.size() * sizeof(double));
        poison(sQ.data(), sQ.size() * sizeof(QagState));
        poison(sUnit.data(), sUnit.size() * sizeof(int));
        poison(ws.data(), ws.size() * sizeof(double));
    }
    SlotArrays slots{sL.data(), sR.data(), sRoot.data(), sYt.data(), sQ.data(), sUnit.data()};
    for (int64_t i = 0; i < n; i++) status[i] = GLC_STATUS_PENDING;
    int work = 0;
    unsigned long long hc[8] = {0};
    KernelArgs A{};
    A.props = soa.data();
    A.flags = flags;
    A.time_end = time_end;
    A.status = status;
    A.interrupt = interrupt;
    A.cap = cap;
    A.n = 0;
    A.ws = ws.data();
    A.nslots = nslots;
    A.work_counter = &work;
    A.counters = hc;
    A.order = nullptr;
    A.resume = 0;
    A.slotL = sL.data();
    A.slotYt = sYt.data();
    A.slotUnit = sUnit.data();
    std::mt19937 rng(777);
    std::vector<int> perm(nslots);
    for (int s = 0; s < nslots; s++) perm[s] = s;
    auto tally_slots = [&]() {
        for (int s = 0; s < nslots; s++) {
            LaneState &L = sL[s];
            hc[0] += L.nAcc; hc[1] += L.nRej; hc[2] += L.nRhs; hc[3] += L.nSeg; hc[4] += L.nTrialFail; hc[5] += L.nNodes; hc[6] += L.nDone;
            L.nAcc = L.nRej = L.nRhs = L.nSeg = L.nTrialFail = L.nNodes = L.nDone = 0;
        }
    };
    // one machine time slice: `budget` unit executions per slot at most; hold = slots stop at the next RK boundary
    auto machine_slice = [&](int budget, bool hold) {
        for (int s = 0; s < nslots; s++) {
            const SlotRef S = slot_ref(slots, s);
            S.unit = machine_rearm(S, A.resume != 0);
        }
        for (int it = 0; it < budget; ++it) {
            bool any = false;
            std::shuffle(perm.begin(), perm.end(), rng);
            for (int p = 0; p < nslots; p++) {
                const int s = perm[p];
                const SlotRef S = slot_ref(slots, s);
                if (hold && S.unit == U_RHS_BEGIN) continue;
                LaneMem M{&A, A.ws + (int64_t)s * (WS_NVEC * NY), 1};
                any |= machine_step(S, M);
            }
            if (!any) break;
        }
        tally_slots();
        A.resume = 1;
    };
    bool laneMode = false;  // every occupied slot stands at an RK boundary (the previous tick was a lane pass)
    auto lane_pass = [&]() {
        // entering lane mode: hold slices bring every occupied slot to an RK boundary (idle slots fetch queued nodes on the
        // way, as on the device); in lane mode newly submitted nodes are fetched by the drain lanes into free slots
        for (int guard = 0; !laneMode && guard < 100000; guard++) {
            machine_slice(1, true);
            bool mid = false;
            for (int s = 0; s < nslots; s++) mid |= (sUnit[s] >= 0 && sUnit[s] != U_IDLE && sUnit[s] != U_RHS_BEGIN);
            if (!mid) break;
        }
        laneMode = true;
        const int queued = std::max(0, A.n - std::min(work, A.n));
        std::vector<int32_t> held;
        int fresh = 0;
        for (int s = 0; s < nslots; s++) {
            if (sUnit[s] == U_RHS_BEGIN)
                held.push_back(s);
            else if ((sUnit[s] == U_IDLE || sUnit[s] < 0) && fresh < queued) {
                held.push_back(s | kHeldFresh);
                fresh++;
            }
        }
        if (held.empty()) return;
        std::shuffle(held.begin(), held.end(), rng);
        int cursor = 0;
        A.held = held.data();
        A.nheld = (int)held.size();
        A.held_counter = &cursor;
        A.drainRefill = 1;
        A.drainLanes = 0;
        const int nl = 9;
        std::vector<LaneState> dl(nl);
        std::vector<LaneMem> dm(nl, LaneMem{&A, A.ws, 1});
        std::vector<std::array<double, NY>> dyt(nl);
        std::vector<char> fr(nl, 0);
        std::vector<int> hs(nl, -1);
        unsigned int tot[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int l = 0; l < nl; l++) lane_reset(dl[l]);
        for (int it = 0; it < lane_budget; it++) {
            bool any = false;
            for (int l = 0; l < nl; l++) {
                bool f = fr[l] != 0;
                double(&y)[NY] = *reinterpret_cast<double(*)[NY]>(dyt[l].data());
                any |= drain_iterate<ModelStandard>(dl[l], dm[l], A, y, f, hs[l], tot, true);
                fr[l] = f ? 1 : 0;
            }
            if (!any) break;
        }
        for (int l = 0; l < nl; l++) {
            double(&y)[NY] = *reinterpret_cast<double(*)[NY]>(dyt[l].data());
            drain_park<ModelStandard>(dl[l], dm[l], A, y, fr[l] != 0, hs[l], tot);
        }
        for (int k = 0; k < 7; k++) hc[k] += tot[k];
        A.drainRefill = 0;
    };
    for (int64_t first = 0; first < n; first += chunk) {
        A.n = (int)std::min<int64_t>(n, first + chunk);
        for (const char *t = pattern; *t; t++) {
            if (*t == 'M') {
                machine_slice(machine_budget, false);
                laneMode = false;
            } else
                lane_pass();
        }
    }
    // finish on the machine alone
    for (int guard = 0; guard < 1000000 && hc[6] < (unsigned long long)n; guard++) machine_slice(64, false);
    machine_slice(64, false);  // one more slice: a ghost (a stale lane state taken for a live one) would write its node back again
    for (int64_t i = 0; i < n; i++)
        for (int p = 0; p < NPROP; p++) props[i * NPROP + p] = soa[(size_t)p * cap + i];
    if (counters) {
        counters->steps_accepted = hc[0];
        counters->steps_rejected = hc[1];
        counters->rhs_evaluations = hc[2];
        counters->segments = hc[3];
        counters->trials_failed = hc[4];
        counters->nodes = hc[5];
    }
    if (hc[6] != (unsigned long long)n) return hc[6] > (unsigned long long)n ? 2 : 1;  // nodes written back twice / lost
    // every node is done: a slot that is still occupied is evolving a GHOST (a stale lane state taken for a live one)
    for (int s = 0; s < nslots; s++)
        if (sUnit[s] >= 0 && sUnit[s] != U_IDLE) return 3;
    return 0;
}

extern "C" {

#ifdef GLC_EMU_COUNTERS
// profiling aid: evaluation counts of the nested solvers (GLC_COUNT sites)
void emu_get_counts(long long *out) {
    for (int k = 0; k < 8; k++) out[k] = g_emu_count[k];
}
void emu_reset_counts(void) {
    for (int k = 0; k < 8; k++) g_emu_count[k] = 0;
}
#endif

void *emu_create(void) { return new Emu(); }
void emu_destroy(void *h) { delete (Emu *)h; }
void emu_set_params(void *h, const glc_params *p) {
    Emu *e = (Emu *)h;
    e->params = *p;
    if (p->model == GLC_MODEL_STANDARD) {
        e->powAc = build_pow_table(1.0e-3, 1.0, p->adiabaticOmega, 1.0e4);
        e->powKmt = build_pow_table(1.0, 1000.0, 0.33, 100.0);
        e->dt.powAc = e->powAc.data();
        e->dt.powAcN = (int)e->powAc.size();
        e->dt.powKmt = e->powKmt.data();
        e->dt.powKmtN = (int)e->powKmt.size();
        pow_table_spacing(1.0e-3, 1.0, e->dt.powAcN, e->dt.powAcDx, e->dt.powAcInvDx);
        pow_table_spacing(1.0, 1000.0, e->dt.powKmtN, e->dt.powKmtDx, e->dt.powKmtInvDx);
        e->dt.lnThinDiskMin = p->accretionRateThinDiskMinimum > 0.0 ? dm_log(p->accretionRateThinDiskMinimum) : 0.0;
        e->dt.lnThinDiskMax = p->accretionRateThinDiskMaximum > 0.0 ? dm_log(p->accretionRateThinDiskMaximum) : 0.0;
        e->dt.profile = nullptr;
        e->dt.profBins = 0;
        if (p->profileOdeEvolver) {
            e->profile.assign(kProfWords, 0ull);
            e->profile[kProfSmallest] = 0x7f7f7f7f7f7f7f7full;
            e->dt.profBins = build_profile_edges(*p, e->dt.profEdges);
            e->dt.profile = e->profile.data();
        }
        build_nfw_j_table(e->nfwJx, e->nfwJv);
        e->dt.nfwJx = e->nfwJx.data();
        e->dt.nfwJv = e->nfwJv.data();
        e->dt.nfwJN = (int)e->nfwJx.size();
    }
}
int emu_set_table(void *h, int id, int n0, int n1, const double *x0, const double *x1, const double *v) {
    Emu *e = (Emu *)h;
    if (id == GLC_TABLE_HALO_MEAN_DENSITY) e->halo_host.set(n0, x0, v);
    if (prepare_table(id, n0, n1, x0, x1, v, e->tables[id]) != 0) return -1;
    PreparedTable &t = e->tables[id];
    install_table(e->dt, id, t, DeviceTable2D{n0, n1, t.x0.data(), t.x1.empty() ? nullptr : t.x1.data(), t.v.data()});
    return 0;
}

int emu_evolve_batch(void *h, int64_t n, double *props, int32_t *flags, const double *time_end, int32_t *status,
                     int32_t *interrupt, glc_counters *counters, int nslots, int budget, int sort, int machine,
                     int64_t *slices) {
    Emu *e = (Emu *)h;
    c_params = e->params;
    c_tables = e->dt;
    if (n <= 0) {
        if (counters) memset(counters, 0, sizeof(*counters));
        return 0;
    }
    if (e->params.model == GLC_MODEL_BOX)
        run<ModelBox>(e, n, props, flags, time_end, status, interrupt, counters, nslots, budget, sort, slices);
    else if (machine)
        run_machine(e, n, props, flags, time_end, status, interrupt, counters, nslots, budget, sort, slices, machine == 2);
    else
        run<ModelStandard>(e, n, props, flags, time_end, status, interrupt, counters, nslots, budget, sort, slices);
    return 0;
}

int emu_stream_session(void *h, int64_t n, double *props, int32_t *flags, const double *time_end, int32_t *status,
                       int32_t *interrupt, glc_counters *counters, int nslots, int chunk, const char *pattern, int lane_budget,
                       int machine_budget) {
    Emu *e = (Emu *)h;
    c_params = e->params;
    c_tables = e->dt;
    return run_stream_session(e, n, props, flags, time_end, status, interrupt, counters, nslots, chunk, pattern, lane_budget,
                              machine_budget);
}

// the host scheduler of the product (host/glc_forest.hpp) driven with the host-executed kernel source as evolve call-back
int emu_forest_evolve(void *h, int64_t n_nodes, const int32_t *parent, const double *mass, const double *time,
                      const double *scale_radius, const double *angular_momentum, double *records, int32_t *flags,
                      int32_t *state, glc_forest_counters *fc, glc_counters *counters, int nslots, int budget, int sort,
                      int machine) {
    Emu *e = (Emu *)h;
    glcf::Forest F;
    F.init(&e->params, &e->halo_host, n_nodes, parent, mass, time, scale_radius, angular_momentum, records, flags, state);
    glc_counters total{};
    std::vector<double> buf;
    std::vector<int32_t> bflags, status, interrupt;
    auto evolve = [&](const std::vector<int32_t> &list, const std::vector<double> &te) -> int {
        const int64_t m = (int64_t)list.size();
        buf.resize((size_t)m * GLC_NPROP);
        bflags.resize(m);
        status.resize(m);
        interrupt.resize(m);
        for (int64_t k = 0; k < m; k++) {
            memcpy(&buf[(size_t)k * GLC_NPROP], F.R(list[k]), sizeof(double) * GLC_NPROP);
            bflags[k] = flags[list[k]];
        }
        glc_counters c{};
        int64_t slices = 0;
        static int n_batches = 0;
        if (const char *mb = getenv("EMU_FOREST_MAX_BATCHES")) {
            if (n_batches >= atoi(mb)) return 1;
            n_batches++;
        }
        int rc = emu_evolve_batch(h, m, buf.data(), bflags.data(), te.data(), status.data(), interrupt.data(), &c, nslots, budget,
                                  sort, machine, &slices);
        if (rc) return rc;
        if (getenv("EMU_FOREST_MAX_BATCHES"))
            fprintf(stderr, "[emu forest] batch of %lld nodes: %llu RHS, %lld slices, nodes counted %llu\n", (long long)m,
                    (unsigned long long)c.rhs_evaluations, (long long)slices, (unsigned long long)c.nodes);
        for (int64_t k = 0; k < m; k++) {
            memcpy(F.R(list[k]), &buf[(size_t)k * GLC_NPROP], sizeof(double) * GLC_NPROP);
            if (status[k] != GLC_STATUS_SUCCESS || interrupt[k] != GLC_INT_NONE) {
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
    const int rc = F.run(evolve);
    if (fc) *fc = F.fc;
    if (counters) *counters = total;
    return rc;
}
int emu_profiler_read(void *h, glc_profile *out) {
    Emu *e = (Emu *)h;
    memset(out, 0, sizeof(*out));
    if (e->profile.empty()) return -9;
    const unsigned long long *P = e->profile.data();
    out->n_bins = e->dt.profBins;
    for (int i = 0; i < GLC_PROFILE_BINS; i++) {
        out->time_step[i] = e->dt.profEdges[i];
        out->time_step_count[i] = P[0 * GLC_PROFILE_BINS + i];
        out->evaluation_count[i] = P[1 * GLC_PROFILE_BINS + i];
        out->time_step_count_interrupted[i] = P[2 * GLC_PROFILE_BINS + i];
        out->evaluation_count_interrupted[i] = P[3 * GLC_PROFILE_BINS + i];
    }
    for (int i = 0; i < GLC_NY; i++) out->property_hits[i] = P[kProfHits + i];
    out->property_hits_unknown = P[kProfUnknown];
    memcpy(&out->time_step_smallest, &P[kProfSmallest], sizeof(double));
    return 0;
}

// Debugging aid: replays slots dumped by a GLC_LEDGER build of the product (GLC_LEDGER_DUMP=file) on the host, one unit at a
// time, printing the unit sequence -- to see on the CPU what a slot that never finishes on the device is doing.
int emu_replay_slots(void *h, const char *path, int max_steps, int verbose) {
    Emu *e = (Emu *)h;
    c_params = e->params;
    c_tables = e->dt;
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    int sizes[8];
    if (fread(sizes, sizeof(int), 8, f) != 8) return -2;
    if (sizes[1] != (int)sizeof(LaneState) || sizes[2] != (int)sizeof(RhsState) || sizes[3] != (int)sizeof(RootState) ||
        sizes[4] != (int)sizeof(QagState) || sizes[5] != NY || sizes[6] != WS_NVEC * NY || sizes[7] != NPROP) {
        fprintf(stderr, "[emu replay] struct sizes differ: file %d %d %d %d, here %zu %zu %zu %zu\n", sizes[1], sizes[2], sizes[3], sizes[4],
                sizeof(LaneState), sizeof(RhsState), sizeof(RootState), sizeof(QagState));
        return -3;
    }
    for (int k = 0; k < sizes[0]; k++) {
        int node, slot, unit, flags;
        double tEnd, yt[NY], ws[WS_NVEC * NY], rec[NPROP];
        LaneState L;
        RhsState R;
        RootState root;
        QagState Q;
        size_t ok = fread(&node, sizeof(int), 1, f) + fread(&slot, sizeof(int), 1, f) + fread(&unit, sizeof(int), 1, f) +
                    fread(&flags, sizeof(int), 1, f) + fread(&tEnd, sizeof(double), 1, f) + fread(&L, sizeof L, 1, f) +
                    fread(&R, sizeof R, 1, f) + fread(&root, sizeof root, 1, f) + fread(&Q, sizeof Q, 1, f) + fread(yt, sizeof yt, 1, f) +
                    fread(ws, sizeof ws, 1, f) + fread(rec, sizeof rec, 1, f);
        if (ok != 12) break;
        int32_t aflags = flags, status = 0, interrupt = 0;
        int work = 1;
        unsigned long long hc[16] = {0};
        KernelArgs A{};
        A.props = rec;  // cap = 1: the record is the arena
        A.flags = &aflags;
        A.time_end = &tEnd;
        A.status = &status;
        A.interrupt = &interrupt;
        A.cap = 1;
        A.n = 1;
        A.ws = ws;
        A.nslots = 1;
        A.work_counter = &work;
        A.counters = hc;
        L.node = 0;
        SlotArrays slots{&L, &R, &root, yt, &Q, &unit};
        const SlotRef S = slot_ref(slots, 0);
        LaneMem M{&A, ws, 1};
        fprintf(stderr, "[emu replay] node %d slot %d unit %d phase %d heavy %d stage %d: ", node, slot, unit, L.phase, L.heavy, L.stage);
        int step = 0;
        unsigned int rhs0 = L.nRhs;
        for (; step < max_steps && S.unit != U_IDLE; step++) {
            if (verbose) fprintf(stderr, "%d ", S.unit);
            machine_step(S, M);
            if (L.nDone) break;
        }
        fprintf(stderr, "-> %d steps, unit %d, nDone %u, rhs +%u, B.state %d it %d busy %d x %.9g | R.count %d comp %d fit %.3g\n", step, S.unit,
                L.nDone, L.nRhs - rhs0, root.B.state, root.B.iteration, root.B.busy, root.B.x, R.count, R.comp, R.fit);
    }
    fclose(f);
    return 0;
}

// scheduler-only run (no evolution: every node just arrives at its end time): host-side cost and round structure
int emu_forest_dryrun(void *h, int64_t n_nodes, const int32_t *parent, const double *mass, const double *time,
                      const double *scale_radius, const double *angular_momentum, double *records, int32_t *flags,
                      int32_t *state, glc_forest_counters *fc, int64_t *batch_sizes, int max_batches) {
    Emu *e = (Emu *)h;
    glcf::Forest F;
    F.init(&e->params, &e->halo_host, n_nodes, parent, mass, time, scale_radius, angular_momentum, records, flags, state);
    int nb = 0;
    auto evolve = [&](const std::vector<int32_t> &list, const std::vector<double> &te) -> int {
        if (nb < max_batches) batch_sizes[nb] = (int64_t)list.size();
        nb++;
        for (size_t k = 0; k < list.size(); k++) F.R(list[k])[GLC_P_TIME] = te[k];
        return 0;
    };
    const int rc = F.run(evolve);
    if (fc) *fc = F.fc;
    return rc ? rc : nb;
}
}  // extern "C"

// The asynchronous scheduler (Forest::run_async) over the host-executed kernel source.  The engine evolves what it is given at
// once but, to exercise the schedule independence the product relies on, reports finished nodes late and out of order: each
// finished node is held back for a pseudo-random number of polls (`straggle` > 0) -- results must not depend on it.
struct EmuAsyncEngine {
    void *h;
    glcf::Forest &F;
    int nslots, budget, sort, machine, straggle;
    std::vector<int32_t> q_nodes;
    std::vector<double> q_tend;
    struct Held { int32_t node, status, interrupt; int polls; };
    std::vector<Held> held;
    glc_counters total{};
    std::mt19937 rng{20240607u};
    int64_t inflight = 0;
    EmuAsyncEngine(void *h_, glcf::Forest &F_, int ns, int b, int so, int m, int st) : h(h_), F(F_), nslots(ns), budget(b), sort(so), machine(m), straggle(st) {}
    void submit(int32_t node, double tend) { q_nodes.push_back(node); q_tend.push_back(tend); inflight++; }
    int64_t in_flight() const { return inflight; }
    int flush() {
        const int64_t m = (int64_t)q_nodes.size();
        if (m == 0) return 0;
        std::vector<double> buf((size_t)m * GLC_NPROP);
        std::vector<int32_t> bflags(m), status(m), interrupt(m);
        for (int64_t k = 0; k < m; k++) {
            memcpy(&buf[(size_t)k * GLC_NPROP], F.R(q_nodes[k]), sizeof(double) * GLC_NPROP);
            bflags[k] = F.flags[q_nodes[k]];
        }
        glc_counters c{};
        int64_t slices = 0;
        int rc = emu_evolve_batch(h, m, buf.data(), bflags.data(), q_tend.data(), status.data(), interrupt.data(), &c, nslots, budget, sort, machine, &slices);
        if (rc) return rc;
        for (int64_t k = 0; k < m; k++) {
            memcpy(F.R(q_nodes[k]), &buf[(size_t)k * GLC_NPROP], sizeof(double) * GLC_NPROP);
            F.flags[q_nodes[k]] = bflags[k];
            held.push_back(Held{q_nodes[k], status[k], interrupt[k], straggle > 0 ? (int)(rng() % (unsigned)(straggle + 1)) : 0});
        }
        total.steps_accepted += c.steps_accepted; total.steps_rejected += c.steps_rejected; total.rhs_evaluations += c.rhs_evaluations;
        total.segments += c.segments; total.trials_failed += c.trials_failed; total.nodes += c.nodes;
        q_nodes.clear();
        q_tend.clear();
        return 0;
    }
    int poll(std::vector<int32_t> &done, std::vector<int32_t> &st, std::vector<int32_t> &in) {
        for (;;) {
            size_t keep = 0;
            for (size_t k = 0; k < held.size(); k++) {
                if (held[k].polls <= 0) {
                    done.push_back(held[k].node); st.push_back(held[k].status); in.push_back(held[k].interrupt);
                    inflight--;
                } else {
                    held[k].polls--;
                    held[keep++] = held[k];
                }
            }
            held.resize(keep);
            if (!done.empty() || held.empty()) return 0;
        }
    }
};

extern "C" int emu_forest_evolve_async(void *h, int64_t n_nodes, const int32_t *parent, const double *mass, const double *time,
                                       const double *scale_radius, const double *angular_momentum, double *records, int32_t *flags,
                                       int32_t *state, glc_forest_counters *fc, glc_counters *counters, int nslots, int budget, int sort,
                                       int machine, int straggle) {
    Emu *e = (Emu *)h;
    glcf::Forest F;
    F.init(&e->params, &e->halo_host, n_nodes, parent, mass, time, scale_radius, angular_momentum, records, flags, state);
    EmuAsyncEngine E(h, F, nslots, budget, sort, machine, straggle);
    const int rc = F.run_async(E);
    if (fc) *fc = F.fc;
    if (counters) *counters = E.total;
    return rc;
}

// Scheduling analysis (scripts/forest_critical_path.py): the asynchronous scheduler over a VIRTUAL clock with unlimited lanes.
// Every submitted node starts at once and lasts (its rate-function evaluations) x 1 time unit; poll() returns the node that
// finishes first.  The makespan is the critical path of the forest in evaluations -- the time the device would need with
// infinitely many lone lanes, in units of the lone-lane time per evaluation -- and (total evaluations) / makespan is the mean
// number of nodes the schedule can keep in flight.
struct EmuClockEngine {
    void *h;
    glcf::Forest &F;
    struct Ev { double finish; int32_t node, status, interrupt; };
    std::vector<Ev> heap;  // min-heap on finish
    double now = 0.0, total = 0.0, overhead;
    int64_t inflight = 0, evolves = 0;
    EmuClockEngine(void *h_, glcf::Forest &F_, double ov) : h(h_), F(F_), overhead(ov) {}
    static bool later(const Ev &a, const Ev &b) { return a.finish > b.finish; }
    void submit(int32_t node, double tend) {
        double buf[GLC_NPROP];
        memcpy(buf, F.R(node), sizeof buf);
        int32_t fl = F.flags[node], st = 0, in = 0;
        glc_counters c{};
        int64_t slices = 0;
        emu_evolve_batch(h, 1, buf, &fl, &tend, &st, &in, &c, 1, 0, 0, 0, &slices);
        memcpy(F.R(node), buf, sizeof buf);
        F.flags[node] = fl;
        heap.push_back(Ev{now + overhead + (double)c.rhs_evaluations, node, st, in});
        std::push_heap(heap.begin(), heap.end(), later);
        total += (double)c.rhs_evaluations;
        inflight++;
        evolves++;
    }
    int64_t in_flight() const { return inflight; }
    int flush() { return 0; }
    int poll(std::vector<int32_t> &done, std::vector<int32_t> &st, std::vector<int32_t> &in) {
        if (heap.empty()) return 0;
        const double t = heap.front().finish;
        while (!heap.empty() && heap.front().finish <= t) {
            std::pop_heap(heap.begin(), heap.end(), later);
            const Ev e = heap.back();
            heap.pop_back();
            done.push_back(e.node); st.push_back(e.status); in.push_back(e.interrupt);
            inflight--;
        }
        now = t;
        return 0;
    }
};

extern "C" int emu_forest_critical_path(void *h, int64_t n_nodes, const int32_t *parent, const double *mass, const double *time,
                                        const double *scale_radius, const double *angular_momentum, double *records, int32_t *flags,
                                        int32_t *state, double overhead, double *out3) {
    Emu *e = (Emu *)h;
    glcf::Forest F;
    F.init(&e->params, &e->halo_host, n_nodes, parent, mass, time, scale_radius, angular_momentum, records, flags, state);
    EmuClockEngine E(h, F, overhead);
    const int rc = F.run_async(E);
    out3[0] = E.now;            // makespan = critical path, in evaluations
    out3[1] = E.total;          // all evaluations
    out3[2] = (double)E.evolves;
    return rc;
}
