"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py): the CPU checker, the kernel
source driven on the host (tests/emu) and -- on the GPU box -- the CUDA path through the C-ABI must all reproduce
them bit for bit."""
import os

import numpy as np
import pytest

from galacticus_b200 import abi, synthetic
from tests import cases

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _case(name):
    g = np.load(os.path.join(HERE, name + ".npz"))
    if name.startswith("standard"):
        p = cases.standard_params(with_black_holes="_bh_" in name)
        tables = True
    else:
        from galacticus_b200.evolver import params_default

        p = params_default(abi.GLC_MODEL_BOX)
        p.box_timescaleStarFormation = 0.5
        p.box_fractionOutflow = 1.0
        tables = False
    return g, p, tables


def _check(g, props, flags, status, interrupt, counters):
    np.testing.assert_array_equal(status, g["status"])
    np.testing.assert_array_equal(interrupt, g["interrupt"])
    np.testing.assert_array_equal(flags, g["flags_out"])
    assert np.array_equal(props, g["props_out"]), "records differ from the golden vector"
    want = dict(zip([str(k) for k in g["counter_names"]], [int(v) for v in g["counters"]]))
    assert counters == want


@pytest.mark.parametrize("name", ["standard_96", "standard_bh_96", "box_leaky_96"])
def test_oracle_reproduces_golden(oracle_lib, name):
    g, p, tables = _case(name)
    o = oracle_lib.Oracle()
    synthetic.install(o, p) if tables else o.set_params(p)
    props, flags = g["props_in"].copy(), g["flags_in"].copy()
    s, i, c = o.evolve_batch(props, flags, g["t_end"], n_threads=4)
    _check(g, props, flags, s, i, c)


@pytest.mark.parametrize("name", ["standard_96", "standard_bh_96", "box_leaky_96"])
def test_kernel_source_on_host_reproduces_golden(name):
    from tests import emu

    g, p, tables = _case(name)
    e = emu.EmuEvolver(nslots=40, budget=11, sort=True, machine=2 if tables else True)
    synthetic.install(e, p) if tables else e.set_params(p)
    props, flags = g["props_in"].copy(), g["flags_in"].copy()
    s, i, c = e.evolve_batch(props, flags, g["t_end"])
    _check(g, props, flags, s, i, c)


@pytest.mark.gpu
@pytest.mark.parametrize("machine", [1, 2])  # micro-task machine + drain hand-over / kernel chosen by batch size
@pytest.mark.parametrize("name", ["standard_96", "standard_bh_96", "box_leaky_96"])
def test_cuda_reproduces_golden(name, machine):
    from galacticus_b200.evolver import Evolver

    g, p, tables = _case(name)
    ev = Evolver(0)
    synthetic.install(ev, p) if tables else ev.set_params(p)
    ev.set_option(abi.GLC_OPT_MICROTASK_MACHINE, machine)
    props, flags = g["props_in"].copy(), g["flags_in"].copy()
    s, i, c = ev.evolve_batch(props, flags, g["t_end"])
    _check(g, props, flags, s, i, c)
    ev.close()


# ---------------------------------------------------------------- tree level
def _forest_case():
    g = np.load(os.path.join(HERE, "forest_5.npz"))
    f = {k[3:]: g[k] for k in g.files if k.startswith("in_")}
    return g, f, cases.standard_params(with_black_holes=True)


def _check_forest(g, rec, flags, state, fc, c):
    np.testing.assert_array_equal(state, g["state"])
    np.testing.assert_array_equal(flags, g["flags"])
    assert {str(k): int(v) for k, v in zip(g["forest_counter_names"], g["forest_counters"])} == {k: v for k, v in fc.items() if k != "rounds"}
    assert {str(k): int(v) for k, v in zip(g["counter_names"], g["counters"])} == c
    alive = g["state"] != abi.GLC_FOREST_NODE_PROMOTED
    assert np.array_equal(rec[alive], g["records"][alive]), "surviving node records differ from the golden forest"


def test_oracle_walk_reproduces_golden_forest(oracle_lib):
    g, f, p = _forest_case()
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    _check_forest(g, *o.forest_evolve(f, n_threads=3))


def test_scheduler_on_host_reproduces_golden_forest():
    from tests import emu

    g, f, p = _forest_case()
    e = emu.EmuEvolver(nslots=48, budget=13, sort=True, machine=2)
    synthetic.install(e, p)
    _check_forest(g, *e.forest_evolve(f))


@pytest.mark.gpu
def test_cuda_forest_reproduces_golden_forest():
    from galacticus_b200.evolver import Evolver

    g, f, p = _forest_case()
    ev = Evolver(0)
    synthetic.install(ev, p)
    _check_forest(g, *ev.forest_evolve(f))
    ev.close()

