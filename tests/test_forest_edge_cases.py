"""Edge-shaped forests and the error surface of glc_forest_evolve: trees that are a single node, chains with one
progenitor per node, a forest mixing them with ordinary trees; malformed parent arrays (GLC_ERR_BAD_FOREST) and trees that
cannot reach their final time (GLC_ERR_DEADLOCK: the reference's deadlock report, merger_trees/evolver/standard.F90:606-625)."""
import numpy as np
import pytest

from galacticus_b200 import abi, synthetic
from tests import cases

P = abi.P
PROMOTED = abi.GLC_FOREST_NODE_PROMOTED


def _edge_forest(p):
    """One ordinary binary-split tree, then a single-node tree, then a chain of five nodes (one progenitor each), then a node
    with three progenitors of equal mass (ties in the progenitor order)."""
    f = synthetic.binary_split_forest(p, 1, 6.0e11, 4.0e10, seed=3)
    n0 = f["parent"].shape[0]
    t_end = float(f["time"][np.where(f["parent"] < 0)[0][0]])
    parent = list(f["parent"])
    mass, time = list(f["mass"]), list(f["time"])
    scale, angmom = list(f["scale_radius"]), list(f["angular_momentum"])
    rs0, j0 = float(f["scale_radius"][0]), float(f["angular_momentum"][0])

    def add(par, m, t):
        parent.append(par)
        mass.append(m)
        time.append(t)
        scale.append(rs0 * (m / f["mass"][0]) ** (1.0 / 3.0))
        angmom.append(j0 * (m / f["mass"][0]) ** (5.0 / 3.0))
        return len(parent) - 1

    add(-1, 3.0e11, t_end)                       # a tree that is only its root
    c = add(-1, 5.0e11, t_end)                   # a chain: every node has exactly one progenitor
    for k in range(1, 5):
        c = add(c, 5.0e11 * (1.0 - 0.1 * k), t_end - 2.0 * k)
    r = add(-1, 9.0e11, t_end)                   # three progenitors of equal mass
    for _ in range(3):
        add(r, 2.5e11, t_end - 3.0)
    return {"parent": np.array(parent, dtype=np.int32), "mass": np.array(mass), "time": np.array(time),
            "scale_radius": np.array(scale), "angular_momentum": np.array(angmom)}, n0


def _same(a, b):
    ra, fa, sa, fca, ca = a
    rb, fb, sb, fcb, cb = b
    np.testing.assert_array_equal(sb, sa)
    np.testing.assert_array_equal(fb, fa)
    for k in ("trees", "nodes", "evolve_calls", "promotions", "node_mergers", "failed_evolves"):
        assert fcb[k] == fca[k], k
    assert cb == ca
    alive = sa != PROMOTED
    assert np.array_equal(rb[alive], ra[alive])


def test_edge_shaped_forest_on_the_host_driven_kernels(oracle_lib):
    from tests import emu

    p = cases.standard_params(with_black_holes=True)
    f, n0 = _edge_forest(p)
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    walk = o.forest_evolve(f, n_threads=3)
    assert walk[3]["trees"] == 4 and walk[3]["failed_evolves"] == 0
    # the lone root needs no evolve call and stays what it was initialised to; the chain promotes four times
    assert walk[2][n0] == abi.GLC_FOREST_NODE_ISOLATED and walk[0][n0, P["TIME"]] == f["time"][n0]
    assert (walk[2][n0 + 2:n0 + 6] == PROMOTED).all()
    # three equal-mass progenitors: one is promoted (the first in index order), two become satellites
    trio = walk[2][-3:]
    assert (trio == PROMOTED).sum() == 1 and trio[0] == PROMOTED and (trio == abi.GLC_FOREST_NODE_SATELLITE).sum() == 2
    e = emu.EmuEvolver(nslots=24, machine=2)
    synthetic.install(e, p)
    _same(walk, e.forest_evolve(f))
    _same(walk, e.forest_evolve(f, asynchronous=True, straggle=5))


@pytest.mark.gpu
def test_edge_shaped_forest_on_the_device(oracle_lib):
    from galacticus_b200.evolver import Evolver

    p = cases.standard_params(with_black_holes=True)
    f, _ = _edge_forest(p)
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    walk = o.forest_evolve(f, n_threads=3)
    ev = Evolver(0)
    synthetic.install(ev, p)
    for schedule in (1, 0):
        ev.set_option(abi.GLC_OPT_FOREST_SCHEDULE, schedule)
        _same(walk, ev.forest_evolve(f))
    ev.close()


@pytest.mark.gpu
def test_malformed_forests_are_refused(oracle_lib):
    from galacticus_b200.evolver import Evolver, GlcError

    p = cases.standard_params(with_black_holes=True)
    f, _ = _edge_forest(p)
    ev = Evolver(0)
    synthetic.install(ev, p)
    n = f["parent"].shape[0]
    for bad_parent in (n, -2):
        g = {k: v.copy() for k, v in f.items()}
        g["parent"][3] = bad_parent
        with pytest.raises(GlcError, match=r"\(%d\)" % abi.GLC_ERR_BAD_FOREST):
            ev.forest_evolve(g)
    g = {k: v.copy() for k, v in f.items()}
    roots = np.where(g["parent"] < 0)[0]
    g["parent"][roots[0]] = int(np.where(g["parent"] == roots[0])[0][0])  # the root's parent is its own child: a cycle
    with pytest.raises(GlcError, match=r"\(%d\)" % abi.GLC_ERR_BAD_FOREST):
        ev.forest_evolve(g)
    # a progenitor that lives AFTER its parent can never arrive: the tree cannot reach its final time
    for schedule in (1, 0):
        ev.set_option(abi.GLC_OPT_FOREST_SCHEDULE, schedule)
        g = {k: v.copy() for k, v in f.items()}
        leaf = n - 1
        g["time"][leaf] = g["time"][g["parent"][leaf]] + 0.5
        with pytest.raises(GlcError, match=r"\(%d\)" % abi.GLC_ERR_DEADLOCK):
            ev.forest_evolve(g)
    # the evolver is still usable afterwards
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    _same(o.forest_evolve(f, n_threads=3), ev.forest_evolve(f))
    ev.close()
