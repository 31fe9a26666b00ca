"""Table-independent invariants of the reference's own test suite, run on the CPU checker for the standard
(quickTest) operator set: baryon conservation (testSuite/test-mass-conservation-standard.py:49-85: the baryons of a
halo change only by accretion), sign invariants (test-enforceNonNegativity.py:70-83) and, at the level of one RHS
evaluation, exact closure of the mass rates."""
import numpy as np

from galacticus_b200 import abi, synthetic
from tests import cases

P = abi.P
MASSES = ("HH_MASS", "HH_OUTFLOWED_MASS", "HH_UNACCRETED_MASS", "HH_STRIPPED_MASS", "DISK_MASS_GAS", "DISK_MASS_STELLAR",
          "SPH_MASS_GAS", "SPH_MASS_STELLAR")


def _oracle(orc):
    p = cases.standard_params()
    o = orc.Oracle()
    synthetic.install(o, p)
    return o, p


def test_mass_rates_close_at_every_evaluation(oracle_lib):
    o, p = _oracle(oracle_lib)
    props, flags, _ = synthetic.standard_nodes(p, 1500, seed=11)
    fb = p.OmegaBaryon / p.OmegaMatter
    worst = 0.0
    checked = 0
    for i in range(props.shape[0]):
        dydt, code, _ = o.rhs(props[i], flags[i])
        if code != 0 or not (flags[i] & abi.GLC_F_HAS_HOTHALO):
            continue  # a creation interrupt zeroes the rates; without a hot halo accretion waits for the interrupt
        total = sum(dydt[P[k]] for k in MASSES)
        sat = bool(flags[i] & abi.GLC_F_IS_SATELLITE)
        expect = 0.0 if sat else fb * props[i, P["MASS_RATE"]]
        scale = sum(abs(dydt[P[k]]) for k in MASSES) + abs(expect)
        if scale > 0:
            worst = max(worst, abs(total - expect) / scale)
            checked += 1
    assert checked > 1000
    assert worst < 1.0e-12, worst


def test_baryon_budget_and_signs_after_evolution(oracle_lib):
    o, p = _oracle(oracle_lib)
    n = 4000
    props, flags, t_end = synthetic.standard_nodes(p, n, seed=12)
    p0 = props.copy()
    s, i, c = o.evolve_batch(props, flags, t_end, n_threads=8)
    assert (s == 0).all() and (i == 0).all() and not np.isnan(props).any()
    bary = lambda q: sum(q[:, P[k]] for k in MASSES)
    fb = p.OmegaBaryon / p.OmegaMatter
    sat = (flags & abi.GLC_F_IS_SATELLITE) != 0
    acc = np.where(sat, 0.0, fb * (props[:, P["BASIC_MASS"]] - p0[:, P["BASIC_MASS"]]))
    tot = bary(p0) + np.abs(acc)
    ok = tot > 0
    d = np.abs(bary(props) - bary(p0) - acc)[ok] / tot[ok]
    assert np.median(d) < 1.0e-12
    assert (d < 1.0e-3).mean() > 0.9  # only nodes that hit the reference's negative-mass clamps deviate
    for k in ("HH_MASS", "DISK_MASS_GAS", "DISK_MASS_STELLAR", "SPH_MASS_GAS", "SPH_MASS_STELLAR"):
        assert (props[:, P[k]] >= 0).all(), k
    np.testing.assert_array_equal(props[:, P["TIME"]], t_end)
