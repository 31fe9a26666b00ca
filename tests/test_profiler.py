"""profileOdeEvolver: the step error analyzer / simple profiler of the reference
(node_evolver/standard.F90:1187-1239, merger_trees/evolve/profiler/simple.F90) restated in the checker and in the kernel
source; the integer accumulators must agree exactly."""
import numpy as np
import pytest

from galacticus_b200 import abi, synthetic
from tests import cases

P = abi.P


def _case(orc):
    p = cases.standard_params(orc, with_black_holes=True)
    p.profileOdeEvolver = 1
    props, flags, t_end = cases.standard_bh_nodes(p, 400, seed=77)
    return p, props, flags, t_end


def _check_sane(pr, counters):
    assert pr["n_bins"] == 22  # int(log10(1e1 / 1e-6) * 3) + 1
    assert abs(pr["time_step"][0] - 1.0e-6) < 1e-18 and abs(pr["time_step"][-1] / 10.0 - 1.0) < 1e-12
    assert pr["time_step_count"].sum() == counters["steps_accepted"]
    assert pr["evaluation_count"].sum() >= pr["time_step_count"].sum()
    assert pr["property_hits"].sum() + pr["property_hits_unknown"] == counters["steps_accepted"]
    assert 0.0 < pr["time_step_smallest"] < 1.0


def test_profiler_oracle_vs_kernel_source(oracle_lib):
    from tests import emu

    p, props, flags, t_end = _case(oracle_lib)
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    o.profiler_reset()
    _, _, co = o.evolve_batch(props.copy(), flags.copy(), t_end)
    po = o.profiler_read()
    _check_sane(po, co)
    for machine in (False, True, 2):
        e = emu.EmuEvolver(nslots=96, machine=machine)
        synthetic.install(e, p)
        _, _, ce = e.evolve_batch(props.copy(), flags.copy(), t_end)
        pe = e.profiler_read()
        assert ce == co
        for k in ("time_step", "time_step_count", "evaluation_count", "time_step_count_interrupted", "evaluation_count_interrupted",
                  "property_hits"):
            np.testing.assert_array_equal(pe[k], po[k], err_msg=f"{k} (machine={machine})")
        assert pe["property_hits_unknown"] == po["property_hits_unknown"]
        assert pe["time_step_smallest"] == po["time_step_smallest"]
    # the hot halo and the disk limit most steps of this workload; the limiting property is always an active one
    hits = po["property_hits"]
    assert hits[P["SAT_BOUND_MASS"]] >= 0 and hits.sum() > 0


@pytest.mark.gpu
def test_profiler_gpu(oracle_lib):
    from galacticus_b200.evolver import Evolver

    p, props, flags, t_end = _case(oracle_lib)
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    o.profiler_reset()
    _, _, co = o.evolve_batch(props.copy(), flags.copy(), t_end)
    po = o.profiler_read()
    for machine in (0, 1):
        ev = Evolver(0)
        synthetic.install(ev, p)
        ev.set_option(abi.GLC_OPT_MICROTASK_MACHINE, machine)
        _, _, cg = ev.evolve_batch(props.copy(), flags.copy(), t_end)
        pg = ev.profiler_read()
        ev.close()
        assert cg == co
        for k in ("time_step", "time_step_count", "evaluation_count", "time_step_count_interrupted", "evaluation_count_interrupted",
                  "property_hits"):
            np.testing.assert_array_equal(pg[k], po[k], err_msg=f"{k} (machine={machine})")
        assert pg["time_step_smallest"] == po["time_step_smallest"]


@pytest.mark.gpu
def test_wall_clock_guard_returns_xcpu(oracle_lib):
    """systemClockMaximum (node_evolver/standard.F90:694-705): nodes unfinished when the wall-clock budget of a call expires
    come back with errorStatusXCPU (1025), their records untouched; the others are evolved as usual."""
    from galacticus_b200.evolver import Evolver

    p = cases.standard_params(with_black_holes=True)
    p.wallClockMaximumSeconds = 1.0e-5
    props, flags, t_end = cases.standard_bh_nodes(p, 60000, seed=5)
    for machine in (0, 1):
        ev = Evolver(0)
        synthetic.install(ev, p)
        ev.set_option(abi.GLC_OPT_MICROTASK_MACHINE, machine)
        pg, fg = props.copy(), flags.copy()
        sg, ig, _ = ev.evolve_batch(pg, fg, t_end)
        ev.close()
        assert set(np.unique(sg)) <= {abi.GLC_STATUS_SUCCESS, abi.GLC_STATUS_XCPU}
        late = sg == abi.GLC_STATUS_XCPU
        assert late.any(), "the budget of 10 microseconds must expire before 60 000 nodes are done"
        assert abi.GLC_STATUS_XCPU == 1025
        assert (ig[late] == abi.GLC_INT_NONE).all()
