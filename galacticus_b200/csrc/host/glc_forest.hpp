// glc_forest.hpp -- host side of the path: the batching tree evolver ("mergerTreeEvolverB200").
//
// replaces: the tree walk of mergerTreeEvolverStandard::evolve (source/merger_trees/evolver/standard.F90:291-635) over a
// SET of forests, which hands one node at a time to mergerTreeNodeEvolver%evolve (:452).  Here every round gathers all
// nodes that the reference's rules allow to move (standardNodeIsEvolvable :723-760, standardTimeEvolveTo :762-1035)
// and evolves them in one call of the batched node evolver; promotions and node mergers (node_evolver/standard.F90:
// 1241-1356) and the node-operator hooks that go with them are bookkeeping on the host.  Citations are relative to
// /root/reference/source.
//
// The evolve call-back is a template parameter: the product (glc_forest.cpp) passes glc_evolve_batch (CUDA); the
// CPU-only test harness (tests/emu, test infrastructure) passes the host-driven kernel source.
//
// Rules restated (and what is left out):
//   * a node moves only when it has no children left (:741) and, if it is not a satellite, at most to its parent's time
//     (:905-914); every call is also capped by mergerTreeEvolveTimestepSimple (evolve/timesteps/simple.F90:
//     min(timeStepRelative/H(t), timeStepAbsolute)) and by the final time of its tree;
//   * a satellite may not pass its host while the host still has a child (:927-941), otherwise it may lead the host by
//     timeStepHost = min(timestepHostRelative/H(t_host), timestepHostAbsolute) (:942-968);
//   * a host may not pass any of its satellites (:1003-1030);
//   * a non-primary progenitor that reaches its parent's time becomes a satellite of the parent at once (standardMerge,
//     mergerTreeNodeMergerSingleLevelHierarchy: its own satellites are re-hosted too); the primary progenitor is promoted
//     when it has reached the parent's time and all its siblings have merged (:1241-1327);
//   * hooks: nodeOperatorDMOInterpolate (dark_matter_only_mass/interpolate.F90:84-291), the scale-radius and
//     angular-momentum interpolators (same scheme), nodeOperatorCGMAccretion nodeInitialize / nodePromote / nodesMerge
//     (circumgalactic_medium/accretion.F90:144-426) with accretionHaloSimple.
//   Satellites move before their hosts in every round (the reference's walk visits a node's satellites first), sums over
//   satellites run in ascending node index.  Not restated: the "primary may not lead its siblings" cap (:984-1000; it only
//   changes where the primary waits), satellite merging times / galaxy mergers (satellites live to the end of
//   the tree), tree and node events.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../../include/glc_b200.h"
#include "../glc_detmath.h"

namespace glcf {

constexpr double kPi = 3.14159265358979323846;
constexpr double kGInternal = 6.673e-11 * 1.98892e30 / (1.0e3 * 1.0e3) / (1.0e6 * 3.08567758135e16);
constexpr double kMpcPerKmPerSToGyr = (1.0e6 * 3.08567758135e16) / 1.0e3 / (1.0e9 * 3.15581497635456e7);

enum NodeState : int32_t {
    NS_PENDING = GLC_FOREST_NODE_PENDING,     // still has children (a future halo of the tree)
    NS_ACTIVE = GLC_FOREST_NODE_ISOLATED,     // childless, not a satellite: being evolved towards its parent
    NS_SATELLITE = GLC_FOREST_NODE_SATELLITE, // hosted by another node
    NS_DONE = GLC_FOREST_NODE_PROMOTED        // promoted into its parent (destroyed)
};

struct HaloTable {  // GLC_TABLE_HALO_MEAN_DENSITY as uploaded (times, {rho_mean, dln rho/dt})
    int n0 = 0;
    std::vector<double> lnt, v;
    double lnt0 = 0.0, inv_dlnt = 0.0;
    void set(int n, const double *t, const double *values) {
        n0 = n;
        lnt.resize(n);
        for (int i = 0; i < n; i++) lnt[i] = dm_log(t[i]);
        v.assign(values, values + 2 * (size_t)n);
        lnt0 = lnt[0];
        inv_dlnt = (double)(n - 1) / (lnt[n - 1] - lnt[0]);
    }
    // virialDensityContrastDefinition: dark_matter_halos/scales/virial_density_contrast.F90:195-417 (same look-up as
    // the device's halo_scales)
    double virial_velocity(double mass, double time) const {
        const double lt = dm_log(time);
        const double x = (lt - lnt0) * inv_dlnt;
        int i = (int)x;
        if (lt < lnt0) i = 0;
        i = std::max(std::min(i, n0 - 2), 0);
        const double h = x - (double)i;
        const double rho = v[2 * i] * (1.0 - h) + v[2 * (i + 1)] * h;
        const double rvir = dm_cbrt(3.0 * mass / 4.0 / kPi / rho);
        return sqrt(kGInternal * mass / rvir);
    }
};

struct Forest {
    const glc_params *P = nullptr;
    const HaloTable *halo = nullptr;
    int64_t n = 0;
    const int32_t *parent = nullptr;
    const double *mass = nullptr, *time = nullptr, *scale = nullptr, *angmom = nullptr;
    double *rec = nullptr;     // [n][GLC_NPROP]
    int32_t *flags = nullptr;  // [n]
    int32_t *state = nullptr;  // [n] NodeState
    std::vector<int32_t> first_child, sibling, host, children_left, root_of;
    std::vector<std::vector<int32_t>> sats;  // ascending node index
    std::vector<double> time_end;            // per node: final time of its tree
    // work lists (so that a round costs O(nodes that can move), not O(all nodes)): childless isolated nodes, nodes that host
    // satellites; entries that no longer qualify are dropped when the list is next walked
    std::vector<int32_t> active_list, host_list, sib_rank;
    std::vector<uint8_t> in_host_list;
    // what cgmAccretionNodesMerge needs of a merged progenitor (see merge()): unaccreted mass / metals and halo mass at merger
    std::vector<double> merged_unaccreted, merged_unaccreted_abund, merged_mass;
    std::vector<uint8_t> merged_hot;
    mutable double memo_t[2][2] = {{-1.0, -1.0}, {-1.0, -1.0}}, memo_v[2][2] = {{0.0, 0.0}, {0.0, 0.0}};  // timestep memos
    glc_forest_counters fc{};

    double *R(int64_t i) const { return rec + i * GLC_NPROP; }

    // cosmologyFunctionsMatterLambda in closed form (flat): H(t) = H0 sqrt(OL) coth(1.5 sqrt(OL) H0 t)
    double expansion_timescale(double t) const {
        const double OL = 1.0 - P->OmegaMatter;
        const double H0 = P->HubbleConstant / kMpcPerKmPerSToGyr;
        const double e = dm_exp(2.0 * (1.5 * sqrt(OL) * H0 * t));
        return 1.0 / (H0 * sqrt(OL) * ((e + 1.0) / (e - 1.0)));
    }
    // mergerTreeEvolveTimestepSimple (evolve/timesteps/simple.F90: min(timeStepRelative / H(t), timeStepAbsolute)) and the host
    // timestep of mergerTreeEvolverStandard (evolver/standard.F90:942-968, same form with timestepHostRelative/Absolute; a
    // non-positive relative value switches the expansion timescale off, :951-954); each memoised like timeHostPrevious (:945-958)
    double timestep_of(double t, double relative, double absolute, int which) const {
        double *mt = memo_t[which], *mv = memo_v[which];
        if (t == mt[0]) return mv[0];
        if (t == mt[1]) return mv[1];
        mt[1] = mt[0];
        mv[1] = mv[0];
        mt[0] = t;
        mv[0] = relative > 0.0 ? std::min(relative * expansion_timescale(t), absolute) : absolute;
        return mv[0];
    }
    double timestep(double t) const { return timestep_of(t, P->timestepSimpleRelative, P->timestepSimpleAbsolute, 0); }
    double timestep_host(double t) const { return timestep_of(t, P->timestepHostRelative, P->timestepHostAbsolute, 1); }
    double failed_fraction(double m, double t) const {  // accretion/halo/simple.F90: simpleFailedFraction
        return (t > P->timeReionization && halo->virial_velocity(m, t) < P->velocitySuppressionReionization) ? 1.0 : 0.0;
    }
    bool is_primary(int32_t i) const { return parent[i] >= 0 && first_child[parent[i]] == i; }
    double node_time(int32_t i) const { return state[i] == NS_PENDING ? time[i] : R(i)[GLC_P_TIME]; }

    // interpolation targets of node i towards `p` (its parent in the tree), nodeOperator{DMO,darkMatterProfileScale,
    // haloAngularMomentum}Interpolate::nodeInitialize
    void set_targets(int32_t i, double *r) const {
        const int32_t p = parent[i];
        r[GLC_P_MASS_TARGET] = mass[i];
        r[GLC_P_MASS_RATE] = 0.0;
        r[GLC_P_TIME_TARGET] = time[i];
        r[GLC_P_DMSCALE_TARGET] = scale[i];
        r[GLC_P_DMSCALE_RATE] = 0.0;
        r[GLC_P_SPIN_TARGET] = angmom[i];
        r[GLC_P_SPIN_RATE] = 0.0;
        if (p < 0) return;
        double unresolved = mass[p];
        for (int32_t c = first_child[p]; c >= 0; c = sibling[c]) unresolved = unresolved - mass[c];
        const double dt = time[p] - time[i];
        r[GLC_P_TIME_TARGET] = time[p];
        if (unresolved > 0.0) {
            if (is_primary(i)) {
                if (dt > 0.0) r[GLC_P_MASS_RATE] = unresolved / dt;
                r[GLC_P_MASS_TARGET] = mass[i] + unresolved;
            }
        } else {
            const double total = mass[p] - unresolved;
            if (dt > 0.0) r[GLC_P_MASS_RATE] = (unresolved / dt) * (mass[i] / total);
            r[GLC_P_MASS_TARGET] = mass[i] + unresolved * mass[i] / total;
        }
        if (is_primary(i) && dt > 0.0) {
            r[GLC_P_DMSCALE_TARGET] = scale[p];
            r[GLC_P_DMSCALE_RATE] = (scale[p] - scale[i]) / dt;
            r[GLC_P_SPIN_TARGET] = angmom[p];
            r[GLC_P_SPIN_RATE] = (angmom[p] - angmom[i]) / dt;
        }
    }

    void init(const glc_params *params, const HaloTable *h, int64_t n_nodes, const int32_t *par, const double *m, const double *t,
              const double *s, const double *j, double *records, int32_t *fl, int32_t *st) {
        P = params; halo = h; n = n_nodes; parent = par; mass = m; time = t; scale = s; angmom = j;
        rec = records; flags = fl; state = st;
        first_child.assign(n, -1); sibling.assign(n, -1); host.assign(n, -1); children_left.assign(n, 0);
        root_of.assign(n, -1); sats.assign(n, {}); time_end.assign(n, 0.0);
        active_list.clear(); host_list.clear(); sib_rank.assign(n, 0); in_host_list.assign(n, 0);
        merged_unaccreted.assign(n, 0.0); merged_unaccreted_abund.assign(n, 0.0); merged_mass.assign(n, 0.0); merged_hot.assign(n, 0);
        memset(&fc, 0, sizeof(fc));
        // children ordered by descending mass (ties: ascending index); the first one is the primary progenitor
        std::vector<int32_t> order(n);
        for (int64_t i = 0; i < n; i++) order[i] = (int32_t)i;
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return mass[a] > mass[b]; });
        for (auto it = order.rbegin(); it != order.rend(); ++it) {  // pushed at the front in reverse => descending lists
            const int32_t i = *it;
            const int32_t p = parent[i];
            if (p < 0) continue;
            sibling[i] = first_child[p];
            first_child[p] = i;
            children_left[p]++;
        }
        for (int64_t p = 0; p < n; p++) {
            int32_t k = 0;
            for (int32_t c = first_child[p]; c >= 0; c = sibling[c]) sib_rank[c] = k++;
        }
        for (int64_t i = 0; i < n; i++) {
            int32_t r = (int32_t)i;
            while (parent[r] >= 0 && root_of[r] < 0) r = parent[r];
            const int32_t root = root_of[r] >= 0 ? root_of[r] : r;
            for (int32_t q = (int32_t)i; q != r && root_of[q] < 0;) {  // path compression: every node is walked once
                const int32_t nx = parent[q];
                root_of[q] = root;
                q = nx;
            }
            root_of[r] = root;
            root_of[i] = root;
            time_end[i] = time[root];
        }
        memset(rec, 0, sizeof(double) * (size_t)n * GLC_NPROP);
        const double fb = P->OmegaBaryon / P->OmegaMatter;
        for (int64_t i = 0; i < n; i++) {
            flags[i] = 0;
            state[i] = children_left[i] > 0 ? NS_PENDING : NS_ACTIVE;
            if (state[i] != NS_ACTIVE) continue;
            active_list.push_back((int32_t)i);
            fc.trees += parent[i] < 0 ? 1 : 0;
            double *r = R(i);
            r[GLC_P_TIME] = time[i];
            r[GLC_P_TIME_STEP] = -1.0;
            r[GLC_P_BASIC_MASS] = mass[i];
            r[GLC_P_DMSCALE] = scale[i];
            r[GLC_P_SPIN] = angmom[i];
            r[GLC_P_SAT_BOUND_MASS] = mass[i];
            set_targets((int32_t)i, r);
            // nodeOperatorCGMAccretion::nodeInitialize (accretion.F90:144-200) for branch tips
            const double failed = failed_fraction(mass[i], time[i]);
            const double m_hot = fb * mass[i] * (1.0 - failed), m_failed = fb * mass[i] * failed;
            if (m_hot > 0.0 || m_failed > 0.0) {
                flags[i] |= GLC_F_HAS_HOTHALO;
                r[GLC_P_HH_MASS] = m_hot;
                r[GLC_P_HH_UNACCRETED_MASS] = m_failed;
                r[GLC_P_HH_ANGMOM] = angmom[i] * m_hot / mass[i];
            }
        }
        for (int64_t i = 0; i < n; i++)
            if (parent[i] < 0 && children_left[i] > 0) fc.trees++;
        fc.nodes = (uint64_t)n;
    }

    // standardTimeEvolveTo for a satellite (:916-982)
    double satellite_limit(int32_t s) const {
        const int32_t h = host[s];
        const double tn = R(s)[GLC_P_TIME];
        double to = std::min(time_end[s], tn + timestep(tn));
        if (to == tn) return tn;
        const double th = parent[h] >= 0 ? node_time(h) : std::max(node_time(h), tn);
        double limit;
        if (children_left[h] > 0)
            limit = std::max(th, tn);
        else
            limit = std::max(th + timestep_host(th), tn);
        return std::min(to, limit);
    }
    // ... for a childless, isolated node (:905-914, :1003-1030)
    double isolated_limit(int32_t a) const {
        const double tn = R(a)[GLC_P_TIME];
        if (parent[a] < 0) return tn;  // standardNodeIsEvolvable: no parent
        double to = std::min(time_end[a], tn + timestep(tn));
        to = std::min(to, time[parent[a]]);
        for (int32_t s : sats[a]) {
            const double ts = R(s)[GLC_P_TIME];
            if (ts < to) to = std::max(ts, tn);
        }
        return to;
    }
    double baryons(int32_t i) const {
        const double *r = R(i);
        double m = 0.0;
        if (flags[i] & GLC_F_HAS_HOTHALO) m += r[GLC_P_HH_MASS] + r[GLC_P_HH_OUTFLOWED_MASS];
        if (flags[i] & GLC_F_HAS_DISK) m += r[GLC_P_DISK_MASS_GAS] + r[GLC_P_DISK_MASS_STELLAR];
        if (flags[i] & GLC_F_HAS_SPHEROID) m += r[GLC_P_SPH_MASS_GAS] + r[GLC_P_SPH_MASS_STELLAR];
        if (flags[i] & GLC_F_HAS_BH) m += r[GLC_P_BH_MASS];
        return m;
    }
    void add_satellite(int32_t h, int32_t s) {
        host[s] = h;
        if (!in_host_list[h]) {
            in_host_list[h] = 1;
            host_list.push_back(h);
        }
        sats[h].insert(std::lower_bound(sats[h].begin(), sats[h].end(), s), s);
    }

    // standardMerge (:1329-1356): node i (non-primary, at its parent's time) becomes a satellite of the parent
    void merge(int32_t i) {
        const int32_t p = parent[i];
        double *r = R(i);
        // nodeOperatorCGMAccretion::nodesMerge (accretion.F90:265-426): unaccreted gas goes to the parent, part of the
        // parent's unaccreted reservoir is re-accreted into its hot phase.  The parent's side of this hook is order
        // dependent (each merger sees what the previous ones left); the reference applies it in arrival order, which is a
        // property of its walk.  Here the node's side happens now and the parent's side is applied at the parent's own
        // promotion, in progenitor order (apply_merged_progenitors): the same result for every schedule, and identical to
        // arrival order whenever a node has at most one non-primary progenitor.
        if (flags[i] & GLC_F_HAS_HOTHALO) {
            merged_hot[i] = 1;
            merged_unaccreted[i] = r[GLC_P_HH_UNACCRETED_MASS];
            merged_unaccreted_abund[i] = r[GLC_P_HH_UNACCRETED_ABUND];
            merged_mass[i] = r[GLC_P_BASIC_MASS];
            r[GLC_P_HH_UNACCRETED_MASS] = 0.0;
            r[GLC_P_HH_UNACCRETED_ABUND] = 0.0;
        }
        // dmoInterpolateNodesMerge (:277-291) and the scale / angular-momentum interpolators: growth stops
        r[GLC_P_MASS_RATE] = 0.0;
        r[GLC_P_MASS_TARGET] = r[GLC_P_BASIC_MASS];
        r[GLC_P_DMSCALE_RATE] = 0.0;
        r[GLC_P_DMSCALE_TARGET] = r[GLC_P_DMSCALE];
        r[GLC_P_SPIN_RATE] = 0.0;
        r[GLC_P_SPIN_TARGET] = r[GLC_P_SPIN];
        r[GLC_P_TIME_LAST_ISOLATED] = r[GLC_P_TIME];
        r[GLC_P_SAT_BOUND_MASS] = r[GLC_P_BASIC_MASS];
        r[GLC_P_MASS_BARYONIC_SUBHALOS] = 0.0;
        // mergerTreeNodeMergerSingleLevelHierarchy: the node and its own satellites all become satellites of the parent
        flags[i] |= GLC_F_IS_SATELLITE;
        state[i] = NS_SATELLITE;
        add_satellite(p, i);
        for (int32_t s : sats[i]) add_satellite(p, s);
        sats[i].clear();
        children_left[p]--;
        fc.node_mergers++;
    }
    // the parent's side of cgmAccretionNodesMerge (:281-362) for every merged progenitor of p, in progenitor order; until its
    // own promotion the parent's record only carries this pending hot-halo content
    void apply_merged_progenitors(int32_t p) {
        double *rp = R(p);
        const double fb = P->OmegaBaryon / P->OmegaMatter;
        for (int32_t c = first_child[p]; c >= 0; c = sibling[c]) {
            if (!merged_hot[c]) continue;
            flags[p] |= GLC_F_HAS_HOTHALO;
            rp[GLC_P_HH_UNACCRETED_MASS] = rp[GLC_P_HH_UNACCRETED_MASS] + merged_unaccreted[c];
            rp[GLC_P_HH_UNACCRETED_ABUND] = rp[GLC_P_HH_UNACCRETED_ABUND] + merged_unaccreted_abund[c];
            const double failed = failed_fraction(mass[p], time[p]);
            const double acc_hot = fb * mass[p] * (1.0 - failed), acc = acc_hot, unacc = fb * mass[p] * failed;
            if (acc_hot > 0.0) {
                const double fraction = acc_hot / (acc + unacc);
                const double re = rp[GLC_P_HH_UNACCRETED_MASS] * fraction * merged_mass[c] / mass[p];
                rp[GLC_P_HH_UNACCRETED_MASS] = rp[GLC_P_HH_UNACCRETED_MASS] - re;
                rp[GLC_P_HH_MASS] = rp[GLC_P_HH_MASS] + re;
                // accreted metals: zero (IGM metallicity zero) => nothing moves between the abundance reservoirs
                rp[GLC_P_HH_ANGMOM] = rp[GLC_P_HH_ANGMOM] + re * angmom[p] / mass[p];
            }
        }
    }
    // standardPromote (:1241-1327): the primary progenitor i takes over its parent
    void promote(int32_t i) {
        const int32_t p = parent[i];
        double *r = R(i), *rp = R(p);
        apply_merged_progenitors(p);
        // nodeOperatorCGMAccretion::nodePromote (accretion.F90:202-263): add the parent's pending hot halo
        if (flags[p] & GLC_F_HAS_HOTHALO) {
            if (!(flags[i] & GLC_F_HAS_HOTHALO)) flags[i] |= GLC_F_HAS_HOTHALO;
            if (r[GLC_P_HH_MASS] <= 0.0) r[GLC_P_HH_MASS] = r[GLC_P_HH_ANGMOM] = r[GLC_P_HH_ABUND] = 0.0;
            r[GLC_P_HH_UNACCRETED_MASS] = r[GLC_P_HH_UNACCRETED_MASS] + rp[GLC_P_HH_UNACCRETED_MASS];
            r[GLC_P_HH_MASS] = r[GLC_P_HH_MASS] + rp[GLC_P_HH_MASS];
            r[GLC_P_HH_ANGMOM] = r[GLC_P_HH_ANGMOM] + rp[GLC_P_HH_ANGMOM];
            r[GLC_P_HH_UNACCRETED_ABUND] = r[GLC_P_HH_UNACCRETED_ABUND] + rp[GLC_P_HH_UNACCRETED_ABUND];
            r[GLC_P_HH_ABUND] = r[GLC_P_HH_ABUND] + rp[GLC_P_HH_ABUND];
        }
        // moveComponentsTo(parent); dmoInterpolateNodePromote (:241-275): mass, target and rate of the parent
        memcpy(rp, r, sizeof(double) * GLC_NPROP);
        flags[p] = flags[i];
        rp[GLC_P_BASIC_MASS] = mass[p];
        rp[GLC_P_DMSCALE] = scale[p];
        rp[GLC_P_SPIN] = angmom[p];
        rp[GLC_P_SAT_BOUND_MASS] = mass[p];
        set_targets(p, rp);
        // satellites follow (standardPromote :1281-1303)
        for (int32_t s : sats[i]) add_satellite(p, s);
        sats[i].clear();
        state[i] = NS_DONE;
        state[p] = NS_ACTIVE;
        active_list.push_back(p);
        children_left[p]--;
        fc.promotions++;
    }

    // every root at the final time of its tree (else the walk stopped with nodes that cannot move: a deadlock)
    bool all_roots_finished() const {
        for (int64_t i = 0; i < n; i++)
            if (parent[i] < 0 && !(state[i] == NS_ACTIVE && R(i)[GLC_P_TIME] == time[i])) return false;
        return true;
    }

    // Rounds over the forest.  evolve(idx, time_end) evolves the listed nodes' records in place and returns 0.
    template <class Evolve>
    int run(Evolve &&evolve) {
        std::vector<int32_t> list, arrived;
        std::vector<double> tend;
        for (;;) {
            bool progressed = false;
            // phase A: satellites, host by host (a host that still has a progenitor pins its satellites, :927-941)
            list.clear(); tend.clear();
            {
                size_t keep = 0;
                for (size_t k = 0; k < host_list.size(); k++) {
                    const int32_t h = host_list[k];
                    if (sats[h].empty()) { in_host_list[h] = 0; continue; }
                    host_list[keep++] = h;
                    if (children_left[h] > 0) continue;
                    for (int32_t s : sats[h]) {
                        const double to = satellite_limit(s);
                        if (to > R(s)[GLC_P_TIME]) { list.push_back(s); tend.push_back(to); }
                    }
                }
                host_list.resize(keep);
            }
            if (!list.empty()) {
                if (int rc = evolve(list, tend)) return rc;
                fc.evolve_calls += list.size();
                progressed = true;
            }
            // phase B: isolated childless nodes
            list.clear(); tend.clear();
            {
                size_t keep = 0;
                for (size_t k = 0; k < active_list.size(); k++) {
                    const int32_t a = active_list[k];
                    if (state[a] != NS_ACTIVE) continue;
                    active_list[keep++] = a;
                    const double to = isolated_limit(a);
                    if (to > R(a)[GLC_P_TIME]) {
                        double sub = 0.0;
                        for (int32_t s : sats[a]) sub += baryons(s);
                        R(a)[GLC_P_MASS_BARYONIC_SUBHALOS] = sub;
                        list.push_back(a); tend.push_back(to);
                    }
                }
                active_list.resize(keep);
            }
            if (!list.empty()) {
                if (int rc = evolve(list, tend)) return rc;
                fc.evolve_calls += list.size();
                progressed = true;
            }
            // arrivals: node mergers first -- siblings of one parent in progenitor order, as the reference's walk meets
            // them --, then promotions
            arrived.clear();
            for (int32_t i : active_list)
                if (parent[i] >= 0 && R(i)[GLC_P_TIME] == time[parent[i]]) arrived.push_back(i);
            std::sort(arrived.begin(), arrived.end(), [&](int32_t a, int32_t b) {
                return parent[a] != parent[b] ? parent[a] < parent[b] : sib_rank[a] < sib_rank[b];
            });
            for (int32_t i : arrived)
                if (!is_primary(i)) {
                    merge(i);
                    progressed = true;
                }
            for (int32_t i : arrived)
                if (is_primary(i) && children_left[parent[i]] == 1) {
                    promote(i);
                    progressed = true;
                }
            fc.rounds++;
            if (!progressed) break;
        }
        return 0;
    }
    // ------------------------------------------------------------------------------------------ asynchronous schedule
    // The rounds of run() end with a barrier: every batched call lasts as long as the chain of its slowest node (one node
    // advances at ~4 000 rate-function evaluations per second on a GPU lane), whichever tree that node belongs to.  The rules
    // above couple a node only to its own GROUP -- an isolated, childless host and the satellites it carries -- between
    // arrival events, and arrivals are schedule independent by construction (a parent's time is fixed by the tree; the
    // parent's side of the merger hooks is applied at its promotion in progenitor order).  run_async() therefore lets every
    // group cycle on its own through the SAME two phases as a round of run() -- (A) its satellites that may move, all
    // together; (B) the host itself; then the arrival test -- and only waits for the nodes of that group.  Every node sees the
    // sequence of (state, end time) it sees under run() and under the reference's walk, so the results are bit-identical;
    // what changes is that a slow node delays its own group, not the forest.
    //
    // Engine: submit(node, t_end) queues a node whose record is in R(node); flush() hands the queue to the device; poll(done,
    // status, interrupt) makes the device work and returns nodes that have finished (records and flags written back in
    // place); in_flight() = submitted and not yet returned.
    template <class Engine>
    int run_async(Engine &E) {
        enum : uint8_t { NEED_A = 0, IN_A = 1, NEED_B = 2, IN_B = 3, AFTER_B = 4 };
        std::vector<uint8_t> phase((size_t)n, (uint8_t)NEED_A), queued((size_t)n, 0);
        std::vector<int32_t> pending((size_t)n, 0), group_of((size_t)n, -1), ready, done, dstatus, dinterrupt;
        std::vector<double> tsub((size_t)n, 0.0);
        auto wake = [&](int32_t h) {
            if (!queued[h]) {
                queued[h] = 1;
                ready.push_back(h);
            }
        };
        for (int32_t a : active_list) wake(a);
        // a primary progenitor that has arrived takes over its parent once all its siblings have merged
        auto try_promote = [&](int32_t p) {
            const int32_t c = first_child[p];
            if (c >= 0 && children_left[p] == 1 && state[c] == NS_ACTIVE && pending[c] == 0 && R(c)[GLC_P_TIME] == time[p]) {
                promote(c);
                phase[p] = NEED_A;
                wake(p);
            }
        };
        for (;;) {
            for (size_t k = 0; k < ready.size(); k++) {  // (grows while it is walked: promotions wake the parent)
                const int32_t h = ready[k];
                queued[h] = 0;
                if (state[h] != NS_ACTIVE || pending[h] > 0) continue;
                for (int tries = 0; tries < 3; tries++) {
                    if (phase[h] == AFTER_B) {
                        // arrival at the parent: node merger at once, promotion when the siblings are gone
                        const int32_t p = parent[h];
                        if (p >= 0 && R(h)[GLC_P_TIME] == time[p]) {
                            if (!is_primary(h)) {
                                merge(h);
                                try_promote(p);
                            } else
                                try_promote(p);
                            break;  // merged, promoted, or waiting for its siblings
                        }
                        phase[h] = NEED_A;
                    }
                    if (phase[h] == NEED_A) {
                        int32_t cnt = 0;
                        for (int32_t s : sats[h]) {
                            const double to = satellite_limit(s);
                            if (to > R(s)[GLC_P_TIME]) {
                                E.submit(s, to);
                                tsub[s] = to;
                                group_of[s] = h;
                                cnt++;
                            }
                        }
                        if (cnt) {
                            pending[h] = cnt;
                            phase[h] = IN_A;
                            fc.evolve_calls += (uint64_t)cnt;
                            break;
                        }
                        phase[h] = NEED_B;
                    }
                    if (phase[h] == NEED_B) {
                        const double to = isolated_limit(h);
                        if (to > R(h)[GLC_P_TIME]) {
                            double sub = 0.0;
                            for (int32_t s : sats[h]) sub += baryons(s);
                            R(h)[GLC_P_MASS_BARYONIC_SUBHALOS] = sub;
                            E.submit(h, to);
                            tsub[h] = to;
                            group_of[h] = h;
                            pending[h] = 1;
                            phase[h] = IN_B;
                            fc.evolve_calls++;
                            break;
                        }
                        phase[h] = AFTER_B;  // cannot move: perhaps it stands at its parent's time already
                        if (tries >= 1) {
                            // neither its satellites nor the host can move: the arrival test, then sleep until an event (a
                            // promotion into this node) wakes the group
                            const int32_t p = parent[h];
                            if (p >= 0 && R(h)[GLC_P_TIME] == time[p]) continue;
                            phase[h] = NEED_A;
                            break;
                        }
                    }
                }
            }
            ready.clear();
            if (E.in_flight() == 0) break;
            if (int rc = E.flush()) return rc;
            done.clear();
            dstatus.clear();
            dinterrupt.clear();
            if (int rc = E.poll(done, dstatus, dinterrupt)) return rc;
            fc.rounds++;
            for (size_t k = 0; k < done.size(); k++) {
                const int32_t x = done[k], g = group_of[x];
                if (dstatus[k] != GLC_STATUS_SUCCESS || dinterrupt[k] != GLC_INT_NONE) {
                    fc.failed_evolves++;
                    R(x)[GLC_P_TIME] = tsub[x];
                }
                if (--pending[g] == 0) {
                    phase[g] = (phase[g] == IN_A) ? (uint8_t)NEED_B : (uint8_t)AFTER_B;
                    wake(g);
                }
            }
        }
        return 0;
    }
};

}  // namespace glcf
