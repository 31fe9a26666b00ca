"""The micro-task machine at target scale.

Round 1's unit queues could be lapped: ring cells carried a 5-bit lap tag and producers did not wait for the consumer of
the position one lap back.  In batches with very fast turnover (hundreds of thousands of nodes that finish within a few
units, as in every batch of a large forest) a consumer that was slow between reserving and reading its cell lost its
slot (the node never finished) and, 32 laps later, took another slot's entry (the slot then ran on two lanes at once and
its lane state was corrupted): profiles/r02a_ledger_4000_trees_round1_queues.txt.  test_forest_4000_trees is the reproducer
(the round-1 library never finishes its third machine batch); the two batch tests put the machine in the same regime -- they
happen to pass on the round-1 protocol too, lapping needs the timing of that forest -- and demand the checker's results bit for
bit for EVERY node."""
import numpy as np
import pytest

from galacticus_b200 import abi, synthetic
from tests import cases

pytestmark = pytest.mark.gpu
P = abi.P


def fast_turnover_batch(p, n, seed, trivial_fraction=0.9):
    """Mostly nodes that finish within a few units (hot halo only, a few per cent of a Gyr) around a minority of ordinary ones."""
    props, flags, t_end = cases.standard_bh_nodes(p, n, seed=seed)
    rng = np.random.default_rng(seed + 77)
    trivial = rng.random(n) < trivial_fraction
    flags[trivial] &= ~(abi.GLC_F_HAS_DISK | abi.GLC_F_HAS_SPHEROID | abi.GLC_F_HAS_BH)
    for k, col in P.items():
        if k.startswith(("DISK_", "SPH_", "BH_")) and col < abi.NPROP:
            props[trivial, col] = 0.0
    t_end[trivial] = props[trivial, P["TIME"]] + rng.uniform(0.002, 0.03, int(trivial.sum()))
    return props, flags, t_end


def _run(oracle_lib, n, seed, budget=0, env=None, trivial_fraction=0.9):
    import os

    from galacticus_b200.evolver import Evolver

    p = cases.standard_params(with_black_holes=True)
    saved = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        ev = Evolver(0)  # (the execution knobs are read when the evolver is created)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    synthetic.install(ev, p)
    ev.set_option(abi.GLC_OPT_MICROTASK_MACHINE, 1)
    if budget:
        ev.set_option(abi.GLC_OPT_SLICE_BUDGET, budget)
    o = oracle_lib.Oracle(fast=False)
    synthetic.install(o, p)
    props, flags, t_end = fast_turnover_batch(p, n, seed, trivial_fraction)
    pg, fg = props.copy(), flags.copy()
    po, fo = props.copy(), flags.copy()
    sg, ig, cg = ev.evolve_batch(pg, fg, t_end)
    so, io, co = o.evolve_batch(po, fo, t_end, n_threads=os.cpu_count() or 1)
    ev.close()
    assert cg["nodes"] == n, "every node is fetched exactly once"
    np.testing.assert_array_equal(sg, so)
    np.testing.assert_array_equal(ig, io)
    np.testing.assert_array_equal(fg, fo)
    bad = np.argwhere(pg != po)
    assert bad.size == 0, f"{len(bad)} record entries differ from the checker, first {bad[:5].tolist()}"
    for k in ("steps_accepted", "steps_rejected", "rhs_evaluations", "segments"):
        assert cg[k] == co[k], k


def test_fast_turnover_batch_600k(oracle_lib):
    """600 000 nodes, 90 per cent of which finish within a few units: the regime in which round 1 lost nodes."""
    _run(oracle_lib, 600_000, seed=4000)


def test_hand_over_with_express_launch(oracle_lib):
    """The hand-over variant that is no longer the default (glc_evolver::drain_express): the held nodes with the most predicted
    steps one per warp on a second stream beside a dense launch on the other block per SM, then dense passes over what is left."""
    _run(oracle_lib, 300_000, seed=4003, trivial_fraction=0.5, env={"GLC_DRAIN_EXPRESS": "1"})


def test_fast_turnover_batch_user_slices(oracle_lib):
    """The same regime in user time slices (no drain hand-over: the machine alone runs every node to its end)."""
    _run(oracle_lib, 330_000, seed=4001, budget=2048)


def test_forest_4000_trees(oracle_lib):
    """The case that exposed the defect: 4000 Milky-Way-mass trees (6.9 million nodes) through glc_forest_evolve.  In round 1
    the third machine batch of this forest (551 585 nodes) lost 132 nodes and never finished.  Every tree must reach its
    final time with no failed evolve, and -- trees being independent -- the records of the first trees must equal the
    checker's depth-first walk over those trees alone, bit for bit."""
    import os

    from galacticus_b200.evolver import Evolver

    p = cases.standard_params(with_black_holes=True)
    forest = synthetic.binary_split_forest(p, 4000, 1.52e12, 1.0e9, seed=219)
    ev = Evolver(0)
    synthetic.install(ev, p)
    rec, flags, state, fc, c = ev.forest_evolve(forest)
    ev.close()
    assert fc["failed_evolves"] == 0 and fc["trees"] == 4000
    roots = forest["parent"] < 0
    assert (state[roots] == abi.GLC_FOREST_NODE_ISOLATED).all()
    assert (rec[roots, P["TIME"]] == forest["time"][roots]).all()
    n_sub = 120
    sub = synthetic.forest_subset(forest, n_sub)
    o = oracle_lib.Oracle(fast=False)
    synthetic.install(o, p)
    ro, fo, so, fco, co = o.forest_evolve(sub, n_threads=os.cpu_count() or 1)
    keep = forest["tree"] < n_sub
    np.testing.assert_array_equal(state[keep], so)
    np.testing.assert_array_equal(flags[keep], fo)
    alive = so != abi.GLC_FOREST_NODE_PROMOTED
    bad = np.argwhere(rec[keep][alive] != ro[alive])
    assert bad.size == 0, f"{len(bad)} record entries of the first {n_sub} trees differ from the checker"
