"""Tree level: the product's batching tree scheduler (csrc/host/glc_forest.hpp: bulk-synchronous rounds over
many trees) against the CPU checker's restatement of the reference's walk (oracle/orc_tree.c: one tree at a time, depth
first, one node per call).  Integer bookkeeping (promotions, node mergers, evolve calls, node states) must be identical,
the surviving node records bit-identical.  On the CPU the scheduler drives the host-executed kernel source (tests/emu);
the `-m gpu` test repeats it through glc_forest_evolve."""
import numpy as np
import pytest

from galacticus_b200 import abi, synthetic
from tests import cases

P = abi.P
PROMOTED = abi.GLC_FOREST_NODE_PROMOTED
SATELLITE = abi.GLC_FOREST_NODE_SATELLITE
ISOLATED = abi.GLC_FOREST_NODE_ISOLATED


def _forest(p, n_trees=6, seed=5, mass=(3.0e11, 2.0e12), resolution=2.0e10):
    return synthetic.binary_split_forest(p, n_trees, mass[0], resolution, seed=seed, mass_root_max=mass[1])


def _check_bookkeeping(f, state, fc):
    parent = f["parent"]
    n = parent.shape[0]
    roots = np.where(parent < 0)[0]
    has_child = np.zeros(n, dtype=bool)
    has_child[parent[parent >= 0]] = True
    # every non-root node either was promoted into its parent or became a satellite: integer-exact
    assert fc["trees"] == roots.size and fc["nodes"] == n
    assert fc["promotions"] + fc["node_mergers"] == n - roots.size
    assert fc["promotions"] == (state == PROMOTED).sum() and fc["node_mergers"] == (state == SATELLITE).sum()
    assert (state[roots] == ISOLATED).all() and (state != abi.GLC_FOREST_NODE_PENDING).all()
    # one primary progenitor per node that has progenitors
    assert fc["promotions"] == has_child.sum()


@pytest.mark.parametrize("machine,nslots,budget", [(2, 64, 0), (True, 48, 9), (False, 32, 0)])
def test_scheduler_matches_reference_walk(oracle_lib, machine, nslots, budget):
    from tests import emu

    p = cases.standard_params(with_black_holes=True)
    f = _forest(p)
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    ro, fo, so, fco, co = o.forest_evolve(f, n_threads=4)
    e = emu.EmuEvolver(nslots, budget, True, machine)
    synthetic.install(e, p)
    re, fe, se, fce, ce = e.forest_evolve(f)
    np.testing.assert_array_equal(se, so)
    np.testing.assert_array_equal(fe, fo)
    for k in ("trees", "nodes", "evolve_calls", "promotions", "node_mergers", "failed_evolves"):
        assert fce[k] == fco[k], k
    assert fco["failed_evolves"] == 0
    assert ce == co  # segments, accepted / rejected steps, RHS evaluations
    alive = so != PROMOTED
    assert np.array_equal(re[alive], ro[alive]), "surviving node records not bit-identical"
    _check_bookkeeping(f, so, fco)
    # every surviving node has reached the final time of its tree
    roots = np.where(f["parent"] < 0)[0]
    t_end = f["time"][roots][f["tree"]]
    np.testing.assert_array_equal(ro[alive, P["TIME"]], t_end[alive])
    assert ((fo[so == SATELLITE] & abi.GLC_F_IS_SATELLITE) != 0).all() and ((fo[roots] & abi.GLC_F_IS_SATELLITE) == 0).all()


def test_scheduler_matches_walk_on_trees_with_many_progenitors(oracle_lib):
    """Nodes with three and more progenitors (N-body-like trees): several siblings can become satellites of one parent in
    the same round; their hooks touch the parent's pending hot halo, so they must run in the walk's (progenitor) order."""
    from tests import emu

    p = cases.standard_params(with_black_holes=True)
    f = _forest(p, n_trees=8, seed=29, resolution=1.5e10)
    rng = np.random.default_rng(3)
    parent = f["parent"].copy()
    grand = np.where(parent >= 0, parent[np.maximum(parent, 0)], -1)
    move = (grand >= 0) & (rng.random(parent.size) < 0.3)
    parent[move] = grand[move]  # re-hang on the grandparent: still earlier than its new parent
    f["parent"] = parent.astype(np.int32)
    assert np.bincount(parent[parent >= 0]).max() >= 4
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    ro, fo, so, fco, co = o.forest_evolve(f, n_threads=4)
    e = emu.EmuEvolver(64, 0, True, 2)
    synthetic.install(e, p)
    re, fe, se, fce, ce = e.forest_evolve(f)
    np.testing.assert_array_equal(se, so)
    np.testing.assert_array_equal(fe, fo)
    assert ce == co and {k: fce[k] for k in fce if k != "rounds"} == {k: fco[k] for k in fco if k != "rounds"}
    alive = so != PROMOTED
    assert np.array_equal(re[alive], ro[alive])
    assert fco["promotions"] + fco["node_mergers"] == parent.size - (parent < 0).sum()


def test_tree_level_invariants(oracle_lib):
    p = cases.standard_params(with_black_holes=True)
    f = _forest(p, n_trees=10, seed=11)
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    rec, flags, state, fc, c = o.forest_evolve(f, n_threads=4)
    _check_bookkeeping(f, state, fc)
    roots = np.where(f["parent"] < 0)[0]
    # dmoInterpolate: the root carries the tree's root mass; promoted galaxies keep growing (hot gas present in every root)
    np.testing.assert_array_equal(rec[roots, P["BASIC_MASS"]], f["mass"][roots])
    assert (rec[roots, P["HH_MASS"]] > 0).all()
    # baryons never exceed the universal fraction of the halo by more than the ODE tolerance
    fb = p.OmegaBaryon / p.OmegaMatter
    alive = state != PROMOTED
    tree_baryons = np.zeros(roots.size)
    cols = [P[k] for k in ("HH_MASS", "HH_OUTFLOWED_MASS", "HH_UNACCRETED_MASS", "HH_STRIPPED_MASS", "DISK_MASS_GAS",
                           "DISK_MASS_STELLAR", "SPH_MASS_GAS", "SPH_MASS_STELLAR", "BH_MASS")]
    np.add.at(tree_baryons, f["tree"][alive], rec[alive][:, cols].sum(axis=1))
    assert (tree_baryons < 1.05 * fb * f["mass"][roots]).all() and (tree_baryons > 0.3 * fb * f["mass"][roots]).all()
    # independent trees: evolving a subset of the forest gives the same records (no cross-tree coupling)
    keep = f["tree"] < 3
    idx = np.where(keep)[0]
    remap = -np.ones(f["parent"].shape[0], dtype=np.int64)
    remap[idx] = np.arange(idx.size)
    sub = {k: v[idx] for k, v in f.items()}
    sub["parent"] = np.where(sub["parent"] >= 0, remap[sub["parent"]], -1).astype(np.int32)
    rec2, flags2, state2, fc2, c2 = o.forest_evolve(sub, n_threads=2)
    np.testing.assert_array_equal(state2, state[idx])
    a2 = state2 != PROMOTED
    assert np.array_equal(rec2[a2], rec[idx][a2])


@pytest.mark.gpu
def test_forest_evolve_cuda_matches_reference_walk(oracle_lib):
    from galacticus_b200.evolver import Evolver

    p = cases.standard_params(with_black_holes=True)
    f = _forest(p, n_trees=40, seed=23, resolution=1.0e10)
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    ro, fo, so, fco, co = o.forest_evolve(f, n_threads=16)
    ev = Evolver(0)
    synthetic.install(ev, p)
    rg, fg, sg, fcg, cg = ev.forest_evolve(f)
    np.testing.assert_array_equal(sg, so)
    np.testing.assert_array_equal(fg, fo)
    for k in ("trees", "nodes", "evolve_calls", "promotions", "node_mergers", "failed_evolves"):
        assert fcg[k] == fco[k], k
    assert cg == co
    alive = so != PROMOTED
    assert np.array_equal(rg[alive], ro[alive]), "surviving node records not bit-identical"
    _check_bookkeeping(f, sg, fcg)
    ev.close()


@pytest.mark.parametrize("straggle", [0, 3, 11])
def test_asynchronous_schedule_matches_reference_walk(oracle_lib, straggle):
    """Forest::run_async: every group (a host and its satellites) cycles on its own and finished nodes come back late and out
    of order (`straggle` polls); integer bookkeeping and surviving records must still equal the reference walk's."""
    from tests import emu

    p = cases.standard_params(with_black_holes=True)
    f = _forest(p, n_trees=7, seed=13)
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    ro, fo, so, fco, co = o.forest_evolve(f, n_threads=4)
    e = emu.EmuEvolver(64, 0, True, 2)
    synthetic.install(e, p)
    re, fe, se, fce, ce = e.forest_evolve(f, asynchronous=True, straggle=straggle)
    np.testing.assert_array_equal(se, so)
    np.testing.assert_array_equal(fe, fo)
    for k in ("trees", "nodes", "evolve_calls", "promotions", "node_mergers", "failed_evolves"):
        assert fce[k] == fco[k], k
    assert ce == co
    alive = so != PROMOTED
    assert np.array_equal(re[alive], ro[alive]), "surviving node records not bit-identical"
    _check_bookkeeping(f, se, fce)


def test_asynchronous_schedule_many_progenitors(oracle_lib):
    from tests import emu

    p = cases.standard_params(with_black_holes=True)
    f = _forest(p, n_trees=8, seed=29, resolution=1.5e10)
    rng = np.random.default_rng(3)
    parent = f["parent"].copy()
    grand = np.where(parent >= 0, parent[np.maximum(parent, 0)], -1)
    move = (grand >= 0) & (rng.random(parent.size) < 0.3)
    parent[move] = grand[move]
    f["parent"] = parent.astype(np.int32)
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    ro, fo, so, fco, co = o.forest_evolve(f, n_threads=4)
    e = emu.EmuEvolver(64, 0, True, 2)
    synthetic.install(e, p)
    re, fe, se, fce, ce = e.forest_evolve(f, asynchronous=True, straggle=5)
    np.testing.assert_array_equal(se, so)
    assert ce == co and {k: fce[k] for k in fce if k != "rounds"} == {k: fco[k] for k in fco if k != "rounds"}
    alive = so != PROMOTED
    assert np.array_equal(re[alive], ro[alive])


@pytest.mark.gpu
def test_forest_schedules_agree_on_gpu(oracle_lib):
    """glc_forest_evolve under both schedules (GLC_OPT_FOREST_SCHEDULE: asynchronous groups over the streaming machine, the
    default, and bulk-synchronous rounds) against the checker's walk: identical bookkeeping, bit-identical records."""
    from galacticus_b200.evolver import Evolver

    p = cases.standard_params(with_black_holes=True)
    f = _forest(p, n_trees=60, seed=31, resolution=8.0e9)
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    ro, fo, so, fco, co = o.forest_evolve(f, n_threads=16)
    alive = so != PROMOTED
    for schedule in (1, 0):
        ev = Evolver(0)
        synthetic.install(ev, p)
        ev.set_option(abi.GLC_OPT_FOREST_SCHEDULE, schedule)
        rg, fg, sg, fcg, cg = ev.forest_evolve(f)
        ev.close()
        np.testing.assert_array_equal(sg, so)
        np.testing.assert_array_equal(fg, fo)
        for k in ("trees", "nodes", "evolve_calls", "promotions", "node_mergers", "failed_evolves"):
            assert fcg[k] == fco[k], (schedule, k)
        assert cg == co, schedule
        assert np.array_equal(rg[alive], ro[alive]), f"schedule {schedule}: surviving node records not bit-identical"
