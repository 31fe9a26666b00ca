"""General beta for the hot-halo beta profile: the reference evaluates the density normalisation, the
enclosed mass and the radial moments through Hypergeometric_2F1 when beta differs from 2/3
(source/mass_distributions/spherical/beta_profile.F90:258,425-436,600-612).  Here they go through
I_m(x) = int_0^x t^m (1+t^2)^(-3 beta/2) dt = x^(m+1)/(m+1) 2F1((m+1)/2, 3 beta/2; (m+3)/2; -x^2)
(galacticus_b200/csrc/glc_specfun.h, shared by the CUDA code and the checker)."""
import ctypes as C

import numpy as np
import pytest

from galacticus_b200 import abi, synthetic
from tests import cases

P = abi.P


def beta_moment(orc, m, x, beta):
    L = orc.lib()
    L.orc_dm_beta_moment.restype = None
    L.orc_dm_beta_moment.argtypes = [C.c_int, C.c_long, np.ctypeslib.ndpointer(np.float64), C.c_double,
                                     np.ctypeslib.ndpointer(np.float64)]
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.zeros_like(x)
    L.orc_dm_beta_moment(m, x.size, x, beta, out)
    return out


def test_against_the_hypergeometric_function(oracle_lib):
    scipy_special = pytest.importorskip("scipy.special")
    x = np.concatenate([10.0 ** np.linspace(-5, -3.01, 9), 10.0 ** np.linspace(-2.99, 1.5, 60), [1.0, 3.0 + 1.0 / 3.0, 10.0]])
    for beta in (0.35, 0.5, 2.0 / 3.0, 0.8, 1.0, 1.4):
        for m in (2, 3):
            ours = beta_moment(oracle_lib, m, x, beta)
            ref = x ** (m + 1) / (m + 1) * scipy_special.hyp2f1((m + 1) / 2.0, 1.5 * beta, (m + 3) / 2.0, -x * x)
            np.testing.assert_allclose(ours, ref, rtol=2.0e-12, err_msg=f"beta={beta} m={m}")


def test_two_thirds_limit_is_the_closed_form(oracle_lib):
    """radialMomentTwoThirds (beta_profile.F90:667-728): I_2 = x - atan x, I_3 = (x^2 - ln(1 + x^2)) / 2."""
    x = 10.0 ** np.linspace(-2, 1.2, 40)
    np.testing.assert_allclose(beta_moment(oracle_lib, 2, x, 2.0 / 3.0), x - np.arctan(x), rtol=1.0e-11)
    np.testing.assert_allclose(beta_moment(oracle_lib, 3, x, 2.0 / 3.0), 0.5 * (x * x - np.log1p(x * x)), rtol=1.0e-11)


def _probe(orc, beta, mass, rcore, router, radius):
    L = orc.lib()
    L.orc_mass_distribution_probe.restype = None
    L.orc_mass_distribution_probe.argtypes = [C.POINTER(abi.glc_params), C.c_double, C.c_double, C.c_double, C.c_double,
                                              np.ctypeslib.ndpointer(np.float64)]
    p = orc.params_default(abi.GLC_MODEL_STANDARD)
    p.hotHaloBeta = beta
    out = np.zeros(8)
    L.orc_mass_distribution_probe(C.byref(p), mass, rcore, router, radius, out)
    return out


@pytest.mark.parametrize("beta", [0.4, 0.55, 0.9])
def test_normalisation_identity(oracle_lib, beta):
    """M(< r_outer) = M (the assertion of beta_profile.F90:263-281, relTol 1e-6), and the density integrates to the mass."""
    mass, rcore, router = 3.0e10, 0.03, 0.1
    out = _probe(oracle_lib, beta, mass, rcore, router, router)
    assert abs(out[0] / mass - 1.0) < 1.0e-12
    half = _probe(oracle_lib, beta, mass, rcore, router, 0.5 * router)
    r = np.linspace(0.0, 0.5 * router, 20001)
    rho = out[4] / (1.0 + (r / rcore) ** 2) ** (1.5 * beta)
    numeric = np.trapezoid(4.0 * np.pi * r * r * rho, r)
    assert abs(half[0] / numeric - 1.0) < 1.0e-6
    # beta within 1e-3 of 2/3 takes the closed forms (betaIsTwoThirds), just outside the general ones: continuous across
    a = _probe(oracle_lib, 2.0 / 3.0 * (1.0 + 0.9e-3), mass, rcore, router, 0.5 * router)[0]
    b = _probe(oracle_lib, 2.0 / 3.0 * (1.0 + 1.1e-3), mass, rcore, router, 0.5 * router)[0]
    assert abs(a / b - 1.0) < 2.0e-3


def _general_beta_case(orc, beta):
    p = cases.standard_params(orc, with_black_holes=True)
    p.hotHaloBeta = beta
    synthetic.finalize_params(p)
    props, flags, t_end = cases.standard_bh_nodes(p, 72, seed=17)
    return p, props, flags, t_end


@pytest.mark.parametrize("beta", [0.5, 0.85])
def test_kernel_source_matches_checker_for_general_beta(oracle_lib, beta):
    from tests import emu

    p, props, flags, t_end = _general_beta_case(oracle_lib, beta)
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    po, fo = props.copy(), flags.copy()
    so, io, co = o.evolve_batch(po, fo, t_end)
    # a different beta is a different model: the results must differ from the default's ...
    p23, _, _, _ = _general_beta_case(oracle_lib, 2.0 / 3.0)
    o23 = oracle_lib.Oracle()
    synthetic.install(o23, p23)
    q = props.copy()
    o23.evolve_batch(q, flags.copy(), t_end)
    assert not np.array_equal(q, po)
    # ... and the kernel source must reproduce them bit for bit, on the machine with the drain hand-over
    e = emu.EmuEvolver(nslots=40, machine=2)
    synthetic.install(e, p)
    pe, fe = props.copy(), flags.copy()
    se, ie, ce = e.evolve_batch(pe, fe, t_end)
    np.testing.assert_array_equal(se, so)
    np.testing.assert_array_equal(fe, fo)
    assert np.array_equal(pe, po)
    assert ce == co


@pytest.mark.gpu
def test_cuda_matches_checker_for_general_beta(oracle_lib):
    from galacticus_b200.evolver import Evolver

    p, props, flags, t_end = _general_beta_case(oracle_lib, 0.5)
    props = np.tile(props, (8, 1))
    flags = np.tile(flags, 8)
    t_end = np.tile(t_end, 8)
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    po, fo = props.copy(), flags.copy()
    so, io, co = o.evolve_batch(po, fo, t_end, n_threads=4)
    ev = Evolver(0)
    synthetic.install(ev, p)
    for machine in (0, 1):
        ev.set_option(abi.GLC_OPT_MICROTASK_MACHINE, machine)
        pg, fg = props.copy(), flags.copy()
        sg, ig, cg = ev.evolve_batch(pg, fg, t_end)
        np.testing.assert_array_equal(sg, so)
        np.testing.assert_array_equal(fg, fo)
        assert np.array_equal(pg, po), f"machine={machine}"
        assert cg == co
    ev.close()
