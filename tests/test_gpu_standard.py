"""CUDA path vs the CPU oracle on the quickTest operator set (through the C-ABI)."""
import numpy as np
import pytest

from galacticus_b200 import abi, synthetic
from tests import cases

pytestmark = pytest.mark.gpu
P = abi.P
RTOL = 1.0e-6  # north_star tolerance on per-galaxy properties


def make(orc, with_black_holes=False, machine=1, **kw):
    from galacticus_b200.evolver import Evolver

    p = cases.standard_params(with_black_holes=with_black_holes)
    for k, v in kw.items():
        setattr(p, k, v)
    ev = Evolver(0)
    synthetic.install(ev, p)
    ev.set_option(abi.GLC_OPT_MICROTASK_MACHINE, machine)
    o = orc.Oracle()
    synthetic.install(o, p)
    return ev, o, p


def compare(pg, po, fg, fo, sg, so, ig, io, what):
    """north_star asks for 1e-6 relative; both sides are built from the same IEEE-only elementary
    functions with contraction off, so we can (and do) demand bit-for-bit identity, which is the only
    robust criterion given the chaotic accept/reject decisions of the nested loose-tolerance solvers."""
    np.testing.assert_array_equal(sg, so, err_msg=f"{what}: status")
    np.testing.assert_array_equal(ig, io, err_msg=f"{what}: interrupt")
    np.testing.assert_array_equal(fg, fo, err_msg=f"{what}: flags")
    np.testing.assert_array_equal(pg[:, P["TIME"]], po[:, P["TIME"]], err_msg=f"{what}: time")
    scale = cases.y_scale(po)[:, None]
    cases.assert_close(pg[:, :abi.NY], po[:, :abi.NY], RTOL, scale=scale * 1e3, what=f"{what}: y (1e-6)")
    bad = np.argwhere(pg != po)
    assert bad.size == 0, f"{what}: {len(bad)} entries not bit-identical, first {bad[:5].tolist()}"


def test_rhs_parity(oracle_lib):
    """One evaluation of standardODEs per node: every operator, no time stepping."""
    ev, o, p = make(oracle_lib)
    props, flags, _ = synthetic.standard_nodes(p, 2000, seed=5)
    dg, ig, pg = ev.rhs_batch(props, flags)
    do = np.zeros_like(dg)
    io = np.zeros_like(ig)
    po = props.copy()
    for i in range(props.shape[0]):
        do[i], io[i], po[i] = o.rhs(props[i], flags[i])
    np.testing.assert_array_equal(ig, io)
    scale = np.abs(do).max(axis=1, keepdims=True)
    cases.assert_close(dg, do, RTOL, scale=scale * 1e-3, what="dydt")
    assert np.array_equal(dg, do), f"dydt not bit-identical in {(dg != do).sum()} entries"
    for k in ("DISK_RADIUS", "DISK_VELOCITY", "SPH_RADIUS", "SPH_VELOCITY"):
        assert np.array_equal(pg[:, P[k]], po[:, P[k]]), k


def test_rhs_parity_black_holes(oracle_lib):
    """One evaluation with the full operator list (Bondi-Hoyle-Lyttleton accretion, switched thin-disk /
    ADAF efficiencies and spin-up, Ciotti 2009 winds, jet-power heating, seed interrupt)."""
    ev, o, p = make(oracle_lib, with_black_holes=True)
    props, flags, _ = cases.standard_bh_nodes(p, 3000, seed=6, fresh_fraction=0.0)
    dg, ig, pg = ev.rhs_batch(props, flags)
    do = np.zeros_like(dg)
    io = np.zeros_like(ig)
    for i in range(props.shape[0]):
        do[i], io[i], _ = o.rhs(props[i], flags[i])
    np.testing.assert_array_equal(ig, io)
    assert (io == abi.GLC_INT_BH_CREATE).any() and (do[:, P["BH_MASS"]] != 0).any()
    assert np.array_equal(dg, do), f"dydt not bit-identical in {(dg != do).sum()} entries"


@pytest.mark.parametrize("n", [37, 3000, 20000])
def test_evolve_parity_black_holes(oracle_lib, n):
    ev, o, p = make(oracle_lib, with_black_holes=True)
    props, flags, t_end = cases.standard_bh_nodes(p, n, seed=200 + n)
    pg, fg = props.copy(), flags.copy()
    po, fo = props.copy(), flags.copy()
    sg, ig, cg = ev.evolve_batch(pg, fg, t_end)
    so, io, co = o.evolve_batch(po, fo, t_end, n_threads=8)
    compare(pg, po, fg, fo, sg, so, ig, io, f"black holes n={n}")
    assert cg == co
    assert ((fg & abi.GLC_F_HAS_BH) != 0).all()


def test_black_hole_seed_interrupt_returned_to_host(oracle_lib):
    """resolveInterruptsOnDevice=0: blackHoleCreate comes back to the host, which seeds mass and spin
    (black_holes/seed.F90:184-206) and re-submits; same result as the on-device resolution."""
    ev, o, p = make(oracle_lib, with_black_holes=True, resolveInterruptsOnDevice=0)
    ev2, _, _ = make(oracle_lib, with_black_holes=True)
    props, flags, t_end = cases.standard_bh_nodes(p, 1200, seed=19)
    p_ref, f_ref = props.copy(), flags.copy()
    ev2.evolve_batch(p_ref, f_ref, t_end)
    ph, fh = props.copy(), flags.copy()
    pending = np.arange(props.shape[0])
    rounds, saw = 0, False
    while pending.size and rounds < 16:
        sub_p, sub_f = ph[pending].copy(), fh[pending].copy()
        s, i, _ = ev.evolve_batch(sub_p, sub_f, t_end[pending])
        assert (s == 0).all()
        saw |= bool((i == abi.GLC_INT_BH_CREATE).any())
        for code, bit in ((abi.GLC_INT_HOTHALO_CREATE, abi.GLC_F_HAS_HOTHALO), (abi.GLC_INT_DISK_CREATE, abi.GLC_F_HAS_DISK),
                          (abi.GLC_INT_SPHEROID_CREATE, abi.GLC_F_HAS_SPHEROID), (abi.GLC_INT_BH_CREATE, abi.GLC_F_HAS_BH)):
            sub_f[i == code] |= bit
        sub_p[i == abi.GLC_INT_BH_CREATE, P["BH_MASS"]] = p.bhSeedMass
        sub_p[i == abi.GLC_INT_BH_CREATE, P["BH_SPIN"]] = p.bhSeedSpin
        ph[pending], fh[pending] = sub_p, sub_f
        pending = pending[(i != 0) & (sub_p[:, P["TIME"]] < t_end[pending])]
        rounds += 1
    assert saw and pending.size == 0
    np.testing.assert_array_equal(fh, f_ref)
    assert np.array_equal(ph[:, :abi.NY], p_ref[:, :abi.NY])


@pytest.mark.parametrize("n", [1, 37, 3000, 20000])
def test_evolve_parity(oracle_lib, n):
    ev, o, p = make(oracle_lib)
    props, flags, t_end = synthetic.standard_nodes(p, n, seed=100 + n)
    pg, fg = props.copy(), flags.copy()
    po, fo = props.copy(), flags.copy()
    sg, ig, cg = ev.evolve_batch(pg, fg, t_end)
    so, io, co = o.evolve_batch(po, fo, t_end, n_threads=8)
    compare(pg, po, fg, fo, sg, so, ig, io, f"n={n}")
    # integer bookkeeping: same number of segments (component creations) and same step counts
    assert cg["segments"] == co["segments"]
    assert cg["steps_accepted"] == co["steps_accepted"]
    assert cg["steps_rejected"] == co["steps_rejected"]
    assert cg["rhs_evaluations"] == co["rhs_evaluations"]


@pytest.mark.parametrize("n", [500, 150000])
def test_evolve_parity_kernel_chosen_by_batch_size(oracle_lib, n):
    """Default execution option: small batches run on the warp-synchronous kernel, large ones on the micro-task machine
    with drain hand-over; either way the records equal the oracle's bit for bit."""
    ev, o, p = make(oracle_lib, with_black_holes=True, machine=2)
    props, flags, t_end = cases.standard_bh_nodes(p, n, seed=77 + n)
    pg, fg = props.copy(), flags.copy()
    po, fo = props.copy(), flags.copy()
    sg, ig, cg = ev.evolve_batch(pg, fg, t_end)
    so, io, co = o.evolve_batch(po, fo, t_end, n_threads=16)
    compare(pg, po, fg, fo, sg, so, ig, io, f"auto n={n}")
    assert cg == co


def test_interrupts_returned_to_host(oracle_lib):
    """resolveInterruptsOnDevice=0: component-creation interrupts come back to the host loop
    (evolver/standard.F90:425-476); the host applies them and re-submits; the result equals the on-device path."""
    ev, o, p = make(oracle_lib, resolveInterruptsOnDevice=0)
    ev2, _, _ = make(oracle_lib)
    props, flags, t_end = synthetic.standard_nodes(p, 1500, seed=9, fresh_fraction=0.6)
    # strip some components so that creation interrupts fire
    flags[::3] &= ~(abi.GLC_F_HAS_DISK | abi.GLC_F_HAS_SPHEROID)
    props[::3, P["DISK_MASS_STELLAR"]:P["DISK_ANGMOM"] + 1] = 0.0
    props[::3, P["SPH_MASS_STELLAR"]:P["SPH_ANGMOM"] + 1] = 0.0
    p_ref, f_ref = props.copy(), flags.copy()
    ev2.evolve_batch(p_ref, f_ref, t_end)
    ph, fh = props.copy(), flags.copy()
    po, fo = props.copy(), flags.copy()
    pending = np.arange(props.shape[0])
    n_round = 0
    saw_interrupt = False
    while pending.size and n_round < 16:
        sub_p, sub_f = ph[pending].copy(), fh[pending].copy()
        s, i, _ = ev.evolve_batch(sub_p, sub_f, t_end[pending])
        sub_po, sub_fo = po[pending].copy(), fo[pending].copy()
        so, io, _ = o.evolve_batch(sub_po, sub_fo, t_end[pending], n_threads=8)
        np.testing.assert_array_equal(i, io)
        assert (s == 0).all()
        saw_interrupt |= bool((i != 0).any())
        for code, bit in ((abi.GLC_INT_HOTHALO_CREATE, abi.GLC_F_HAS_HOTHALO), (abi.GLC_INT_DISK_CREATE, abi.GLC_F_HAS_DISK),
                          (abi.GLC_INT_SPHEROID_CREATE, abi.GLC_F_HAS_SPHEROID)):
            sub_f[i == code] |= bit
            sub_fo[io == code] |= bit
        ph[pending], fh[pending] = sub_p, sub_f
        po[pending], fo[pending] = sub_po, sub_fo
        pending = pending[(i != 0) & (sub_p[:, P["TIME"]] < t_end[pending])]
        n_round += 1
    assert saw_interrupt
    np.testing.assert_array_equal(fh, f_ref)
    assert np.array_equal(ph[:, :abi.NY], p_ref[:, :abi.NY]), "host-loop vs on-device interrupt resolution"
    assert np.array_equal(ph[:, :abi.NY], po[:, :abi.NY]), "vs oracle"


def test_baryon_budget_full_size():
    """Invariant at a size the oracle is not run at: baryons change only by accretion (and by the
    reference's own negative-mass clamps), cf. testSuite/test-mass-conservation-standard.py:49-85."""
    from galacticus_b200.evolver import Evolver

    p = cases.standard_params()
    ev = Evolver(0)
    synthetic.install(ev, p)
    n = 200_000
    props, flags, t_end = synthetic.standard_nodes(p, n, seed=77)
    p0 = props.copy()
    s, i, c = ev.evolve_batch(props, flags, t_end)
    assert (s == 0).all() and (i == 0).all()
    assert not np.isnan(props).any()

    def bary(q):
        return sum(q[:, P[k]] for k in ("HH_MASS", "HH_OUTFLOWED_MASS", "HH_UNACCRETED_MASS", "HH_STRIPPED_MASS",
                                        "DISK_MASS_GAS", "DISK_MASS_STELLAR", "SPH_MASS_GAS", "SPH_MASS_STELLAR"))

    fb = p.OmegaBaryon / p.OmegaMatter
    sat = (flags & abi.GLC_F_IS_SATELLITE) != 0
    acc = np.where(sat, 0.0, fb * (props[:, P["BASIC_MASS"]] - p0[:, P["BASIC_MASS"]]))
    tot = bary(p0) + np.abs(acc)
    ok = tot > 0
    d = np.abs(bary(props) - bary(p0) - acc)[ok] / tot[ok]
    assert np.median(d) < 1e-12
    assert (d < 1e-3).mean() > 0.9  # only nodes that hit the negative-mass clamps deviate
    for k in ("HH_MASS", "DISK_MASS_GAS", "DISK_MASS_STELLAR", "SPH_MASS_GAS", "SPH_MASS_STELLAR"):
        assert (props[:, P[k]] >= 0).all(), k


def test_execution_options_do_not_change_results(oracle_lib):
    """The three execution paths -- hybrid run-to-completion (machine slices + hold + drain kernel), the machine in
    user time slices, and the warp-synchronous lane kernel -- and the queue order are scheduling choices only."""
    from galacticus_b200.evolver import Evolver

    p = cases.standard_params()
    props, flags, t_end = synthetic.standard_nodes(p, 6000, seed=31337)
    ref = None
    for budget, sort, machine in [(0, 1, 1), (777, 1, 1), (0, 0, 1), (0, 1, 0), (333, 0, 0)]:
        ev = Evolver(0)
        synthetic.install(ev, p)
        ev.set_option(abi.GLC_OPT_SLICE_BUDGET, budget)
        ev.set_option(abi.GLC_OPT_SORT_QUEUE, sort)
        ev.set_option(abi.GLC_OPT_MICROTASK_MACHINE, machine)
        pg, fg = props.copy(), flags.copy()
        s, i, c = ev.evolve_batch(pg, fg, t_end)
        if budget:
            assert ev.slice_count() > 1
        out = (pg, fg, s, i, c)
        if ref is None:
            ref = out
        else:
            assert np.array_equal(out[0], ref[0]) and np.array_equal(out[1], ref[1])
            assert np.array_equal(out[2], ref[2]) and np.array_equal(out[3], ref[3]) and out[4] == ref[4]
        ev.close()


def test_ragged_and_degenerate_nodes(oracle_lib):
    """Edge cases of the reference's own tests: zero-length evolutions (timeStart == timeEnd), nodes without any
    component, freshly created components, an empty batch."""
    ev, o, p = make(oracle_lib)
    props, flags, t_end = synthetic.standard_nodes(p, 1200, seed=4)
    t_end[::5] = props[::5, P["TIME"]]  # zero-length intervals
    flags[::7] = 0  # no components at all (the state is ignored, components are created by interrupts)
    props[1::9, P["DISK_RADIUS"]] = 0.0  # cold structure solves
    props[1::9, P["DISK_VELOCITY"]] = 0.0
    pg, fg = props.copy(), flags.copy()
    po, fo = props.copy(), flags.copy()
    sg, ig, cg = ev.evolve_batch(pg, fg, t_end)
    so, io, co = o.evolve_batch(po, fo, t_end, n_threads=8)
    compare(pg, po, fg, fo, sg, so, ig, io, "ragged")
    assert cg == co
    z = np.zeros((0, abi.NPROP))
    s0, i0, c0 = ev.evolve_batch(z, np.zeros(0, dtype=np.int32), np.zeros(0))
    assert s0.size == 0 and c0["nodes"] == 0
