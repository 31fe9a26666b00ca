"""standardErrorHandler / standardODEStepTolerances (node_evolver/standard.F90:1063-1158): the "ODE system
parameters" table of a node -- y, dy/dt, yScale, yTolerance, yError, |yError| / yTolerance per property -- restated in the
checker (orc_error_report) and on the device (glc_error_report_node)."""
import numpy as np
import pytest

from galacticus_b200 import abi, synthetic
from tests import cases

P = abi.P


def _nodes(orc):
    p = cases.standard_params(orc, with_black_holes=True)
    props, flags, t_end = cases.standard_bh_nodes(p, 64, seed=5)
    return p, props, flags, t_end


def test_checker_table_is_consistent(oracle_lib):
    from oracle import orc

    p, props, flags, t_end = _nodes(orc)
    o = orc.Oracle()
    synthetic.install(o, p)
    seen_disk = False
    for i in range(props.shape[0]):
        h = 0.01
        rep = o.error_report(props[i], flags[i], h)
        act = rep["active"] != 0
        # the ODE system is the node's component set (treeNodeSerializeValuesToArray over the components it has)
        assert act[P["SAT_BOUND_MASS"]]
        assert act[P["DISK_MASS_GAS"]] == bool(flags[i] & abi.GLC_F_HAS_DISK)
        assert act[P["BH_MASS"]] == bool(flags[i] & abi.GLC_F_HAS_BH)
        seen_disk |= bool(flags[i] & abi.GLC_F_HAS_DISK)
        # standardODEStepTolerances: rel |y| + abs scale, with positive scales for every active property
        assert np.all(rep["scale"][act] > 0.0)
        np.testing.assert_allclose(rep["tolerance"][act], p.odeToleranceRelative * np.abs(rep["y"][act]) + p.odeToleranceAbsolute * rep["scale"][act], rtol=1e-15)
        np.testing.assert_allclose(rep["error_scaled"][act], np.abs(rep["error"][act]) / rep["tolerance"][act], rtol=1e-15)
        assert np.all(rep["y"][~act] == 0.0) and np.all(rep["error"][~act] == 0.0)
        # dy/dt is the rate function at the node's time (after the pre-evolve hooks, which may create components: compare
        # only where the hooks leave the record alone)
        dydt, code, _ = o.rhs(props[i], flags[i])
        if code == abi.GLC_INT_NONE and rep["interrupt"] == abi.GLC_INT_NONE and (flags[i] & abi.GLC_F_HH_INITIALIZED):
            np.testing.assert_array_equal(rep["dydt"][act], dydt[act])
        # the embedded error estimate is second order small against the step itself: h*|dydt| bounds it loosely
        assert np.all(np.abs(rep["error"][act]) <= 10.0 * h * np.abs(rep["dydt"][act]).max() + 1e-300)
    assert seen_disk


def test_error_estimate_order_on_the_box_model(oracle_lib):
    """closedBox / leakyBox rates are smooth (no nested solvers), so Cash-Karp's embedded estimate must fall like h^5: halving
    the step shrinks it ~32-fold.  (The standard model's rate function carries the 1e-2 noise of its root finders, which the
    estimate picks up linearly in h -- the reason its controller rejects a third of all steps.)"""
    from oracle import orc

    props, flags, t_end = cases.box_nodes(32, seed=3, leaky=True, ragged=False)
    o = orc.Oracle()
    o.set_params(orc.params_default(abi.GLC_MODEL_BOX))
    checked = 0
    for i in range(props.shape[0]):
        a = o.error_report(props[i], flags[i], 0.2)
        b = o.error_report(props[i], flags[i], 0.1)
        act = a["active"] != 0
        big = act & (np.abs(a["error"]) > 1e-3 * np.abs(a["error"]).max()) & (np.abs(a["error"]) > 0.0)
        if not big.any():
            continue
        ratio = np.abs(a["error"][big]) / np.abs(b["error"][big])
        assert np.all(ratio > 16.0) and np.all(ratio < 64.0), ratio
        checked += 1
    assert checked >= 16


@pytest.mark.gpu
def test_cuda_table_equals_checker(oracle_lib):
    from galacticus_b200.evolver import Evolver
    from oracle import orc

    p, props, flags, t_end = _nodes(orc)
    o = orc.Oracle()
    synthetic.install(o, p)
    ev = Evolver(0)
    synthetic.install(ev, p)
    for i in range(props.shape[0]):
        h = float(props[i, P["TIME_STEP"]]) if props[i, P["TIME_STEP"]] > 0.0 else 0.01
        ro = o.error_report(props[i], flags[i], h)
        rg = ev.error_report_node(props[i], flags[i], h)
        assert rg["interrupt"] == ro["interrupt"]
        np.testing.assert_array_equal(rg["active"], ro["active"])
        for key in ("y", "dydt", "scale", "tolerance"):
            np.testing.assert_array_equal(rg[key], ro[key], err_msg=f"node {i} {key}")
        if ro["interrupt"] == abi.GLC_INT_NONE:
            for key in ("error", "error_scaled"):
                np.testing.assert_array_equal(rg[key], ro[key], err_msg=f"node {i} {key}")
    ev.close()
