"""Kernel source vs the oracle ON THE CPU: the per-lane logic of the CUDA kernels (lane state machine, rate
functions, nested numerics -- galacticus_b200/csrc/*.cuh) is compiled by g++ through the platform layer and
driven like the evolve kernel drives it (warps in lock-step, shared queue, time slices).  This is test
infrastructure (tests/emu); the `-m gpu` tests repeat the same comparisons through the C-ABI on the device."""
import numpy as np
import pytest

from galacticus_b200 import abi, synthetic
from tests import cases, emu

P = abi.P


def _both(orc, p, nslots, budget, sort, tables=True, machine=True):
    e = emu.EmuEvolver(nslots, budget, sort, machine)
    o = orc.Oracle()
    if tables:
        synthetic.install(e, p)
        synthetic.install(o, p)
    else:
        e.set_params(p)
        o.set_params(p)
    return e, o


@pytest.mark.parametrize("nslots,budget,sort,machine", [
    (64, 0, True, False),   # lane state machine + warp-synchronous rate function (evolve_kernel)
    (96, 7, True, True),    # micro-task machine, time slices, shuffled slot order, sticky stepping
    (33, 50, False, True),
    (40, 0, True, 2),       # machine, then hold at RK boundaries and hand over to drain_iterate
    (300, 9, False, 2),
])
def test_standard_lane_logic_is_bit_identical_to_oracle(oracle_lib, nslots, budget, sort, machine):
    p = cases.standard_params()
    e, o = _both(oracle_lib, p, nslots, budget, sort, machine=machine)
    props, flags, t_end = synthetic.standard_nodes(p, 1500, seed=103)
    pe, fe = props.copy(), flags.copy()
    po, fo = props.copy(), flags.copy()
    se, ie, ce = e.evolve_batch(pe, fe, t_end)
    so, io, co = o.evolve_batch(po, fo, t_end, n_threads=8)
    np.testing.assert_array_equal(se, so)
    np.testing.assert_array_equal(ie, io)
    np.testing.assert_array_equal(fe, fo)
    assert ce == co  # integer bookkeeping: segments, accepted/rejected steps, RHS evaluations
    assert np.array_equal(pe, po), "records not bit-identical"
    if budget and machine != 2:
        assert e.slices > 1  # the time-slice / park / resume path was exercised


@pytest.mark.parametrize("nslots,budget,sort,machine", [
    (64, 0, True, False),
    (96, 7, True, True),
    (40, 0, True, 2),
    (130, 13, False, 2),
])
def test_black_hole_chain_lane_logic_is_bit_identical_to_oracle(oracle_lib, nslots, budget, sort, machine):
    """Full quickTest operator list: blackHolesSeed (creation interrupt), blackHolesAccretion (Bondi-Hoyle-Lyttleton rates
    from spheroid gas and hot halo, switched thin-disk/ADAF efficiencies, spin-up), blackHolesWinds (Ciotti 2009) and the
    jet-power heating term of CGMCoolingHeating -- SURVEY 8a a19."""
    p = cases.standard_params(with_black_holes=True)
    e, o = _both(oracle_lib, p, nslots, budget, sort, machine=machine)
    props, flags, t_end = cases.standard_bh_nodes(p, 1500, seed=303)
    pe, fe = props.copy(), flags.copy()
    po, fo = props.copy(), flags.copy()
    se, ie, ce = e.evolve_batch(pe, fe, t_end)
    so, io, co = o.evolve_batch(po, fo, t_end, n_threads=8)
    np.testing.assert_array_equal(se, so)
    np.testing.assert_array_equal(ie, io)
    np.testing.assert_array_equal(fe, fo)
    assert ce == co
    assert np.array_equal(pe, po), "records not bit-identical"
    assert (so == 0).all() and ((fo & abi.GLC_F_HAS_BH) != 0).all()
    grown = po[:, P["BH_MASS"]] > 1.01 * np.maximum(props[:, P["BH_MASS"]], 100.0)
    assert grown.mean() > 0.05  # accretion did something


def test_black_hole_seed_interrupt_returned_to_host(oracle_lib):
    p = cases.standard_params(with_black_holes=True)
    p.resolveInterruptsOnDevice = 0
    e, o = _both(oracle_lib, p, 64, 11, True)
    props, flags, t_end = cases.standard_bh_nodes(p, 500, seed=12)
    pe, fe = props.copy(), flags.copy()
    po, fo = props.copy(), flags.copy()
    se, ie, ce = e.evolve_batch(pe, fe, t_end)
    so, io, co = o.evolve_batch(po, fo, t_end, n_threads=8)
    assert (io == abi.GLC_INT_BH_CREATE).any()
    np.testing.assert_array_equal(ie, io)
    assert ce == co and np.array_equal(pe, po)


def test_standard_interrupts_returned_to_host(oracle_lib):
    p = cases.standard_params()
    p.resolveInterruptsOnDevice = 0
    e, o = _both(oracle_lib, p, 64, 11, True)
    props, flags, t_end = synthetic.standard_nodes(p, 600, seed=9, fresh_fraction=0.6)
    flags[::3] &= ~(abi.GLC_F_HAS_DISK | abi.GLC_F_HAS_SPHEROID)
    props[::3, P["DISK_MASS_STELLAR"]:P["DISK_ANGMOM"] + 1] = 0.0
    props[::3, P["SPH_MASS_STELLAR"]:P["SPH_ANGMOM"] + 1] = 0.0
    pe, fe = props.copy(), flags.copy()
    po, fo = props.copy(), flags.copy()
    se, ie, ce = e.evolve_batch(pe, fe, t_end)
    so, io, co = o.evolve_batch(po, fo, t_end, n_threads=8)
    assert (ie != 0).any()
    np.testing.assert_array_equal(ie, io)
    assert ce == co and np.array_equal(pe, po)


@pytest.mark.parametrize("leaky", [False, True])
def test_box_lane_logic_is_bit_identical_to_oracle(oracle_lib, leaky):
    from galacticus_b200.evolver import params_default

    p = params_default(abi.GLC_MODEL_BOX)
    if leaky:
        p.box_timescaleStarFormation = 0.5
        p.box_fractionOutflow = 1.0
    e, o = _both(oracle_lib, p, 64, 5, True, tables=False)
    props, flags, t_end = cases.box_nodes(3000, seed=220, leaky=leaky)
    pe, fe = props.copy(), flags.copy()
    po, fo = props.copy(), flags.copy()
    se, ie, ce = e.evolve_batch(pe, fe, t_end)
    so, io, co = o.evolve_batch(po, fo, t_end, n_threads=8)
    np.testing.assert_array_equal(se, so)
    assert ce == co and np.array_equal(pe, po) and np.array_equal(fe, fo)


@pytest.mark.parametrize("pattern,chunk,lane_budget", [("LL", 16, 120), ("MLL", 16, 120), ("LML", 16, 80), ("LLLL", 40, 30), ("MMM", 40, 6)])
def test_stream_session_matches_oracle(oracle_lib, pattern, chunk, lane_budget):
    """The adaptive ticks of a streaming session (glc_api.cu stream_tick) on the host: slots go machine -> drain -> machine, the
    drain lanes refill from a queue that grows between ticks and take free slots for newly submitted nodes, slots released by
    the drain are re-armed by the machine's own slice-start code (machine_rearm).  Every node must be written back exactly once
    and bit-identical to the checker.  (The stale-slot defect of round 2 -- a released slot re-evolved its old node when the
    machine took over -- fails this test with "written back twice".)"""
    from oracle import orc
    from tests import emu

    p = cases.standard_params(orc, with_black_holes=True)
    props, flags, t_end = cases.standard_bh_nodes(p, 240, seed=31)
    o = orc.Oracle()
    synthetic.install(o, p)
    po, fo = props.copy(), flags.copy()
    so, io, co = o.evolve_batch(po, fo, t_end)
    e = emu.EmuEvolver(nslots=24, machine=True)
    synthetic.install(e, p)
    pe, fe = props.copy(), flags.copy()
    rc, se, ie, ce = e.stream_session(pe, fe, t_end, chunk=chunk, pattern=pattern, lane_budget=lane_budget)
    assert rc == 0, "nodes lost (1), written back twice (2) or a stale slot taken for a live one (3): %d" % rc
    np.testing.assert_array_equal(se, so)
    np.testing.assert_array_equal(ie, io)
    np.testing.assert_array_equal(fe, fo)
    assert np.array_equal(pe, po), "streamed records differ from the checker"
    assert ce == co
