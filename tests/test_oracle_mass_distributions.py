"""Known answers the reference's own unit tests hold for the mass distributions on the path
(source/tests/mass_distributions.F90), asserted on the checker's restatements at the reference's tolerances."""
import ctypes as C

import numpy as np
import pytest

from galacticus_b200 import abi

PI = np.pi


def probe(orc, mass, rcore, router, radius):
    L = orc.lib()
    L.orc_mass_distribution_probe.restype = None
    L.orc_mass_distribution_probe.argtypes = [C.POINTER(abi.glc_params), C.c_double, C.c_double, C.c_double, C.c_double,
                                              np.ctypeslib.ndpointer(np.float64)]
    p = orc.params_default(abi.GLC_MODEL_STANDARD)
    out = np.zeros(8)
    L.orc_mass_distribution_probe(C.byref(p), mass, rcore, router, radius, out)
    return out


def test_beta_profile_dimensionless(oracle_lib):
    # mass_distributions.F90:240-300: beta = 2/3, dimensionless (rho_0 = 1, r_c = 1, no truncation)
    router = 1.0e3
    mass_total = 4.0 * PI * (router - np.arctan(router))  # the mass that makes rho_0 = 1 for this outer radius
    out = probe(oracle_lib, mass_total, 1.0, router, 1.0)
    assert abs(out[4] - 1.0) < 1.0e-12  # density normalisation
    assert abs(out[0] - (4.0 - PI) * PI) < 1.0e-6  # "Mass within scale radius" :252-257
    assert abs(out[2] - 0.2146018366025517) < 1.0e-6  # "Radial moment, m=2, from 0 to 1" :271-276
    assert abs(out[3] - 0.1534264097200273) < 1.0e-6  # "Radial moment, m=3, from 0 to 1" :277-282
    # "Density gradient (logarithmic)" = -1 at the scale radius :289-294: rho ~ (1 + x^2)^-1
    eps = 1.0e-6
    lo, hi = probe(oracle_lib, mass_total, 1.0, router, 1.0 - eps)[1], probe(oracle_lib, mass_total, 1.0, router, 1.0 + eps)[1]
    assert abs((np.log(hi) - np.log(lo)) / (np.log(1.0 + eps) - np.log(1.0 - eps)) + 1.0) < 1.0e-6


def test_beta_profile_mass_within_outer_radius(oracle_lib):
    # mass_distributions.F90:333-343: the quickTest configuration itself (core radius 0.3 of the outer radius)
    out = probe(oracle_lib, 182582297.19568533, 3.8492316686261747e-002, 0.12829569846196026, 0.12829569846196026)
    assert abs(out[0] / 182582297.19568533 - 1.0) < 1.0e-6


def test_hernquist(oracle_lib):
    # mass_distributions.F90:176-189: a quarter of the mass lies within the scale radius, half within 1 + sqrt(2)
    assert abs(probe(oracle_lib, 1.0, 1.0, 1.0, 1.0)[5] - 0.25) < 1.0e-6
    assert abs(probe(oracle_lib, 1.0, 1.0, 1.0, 1.0 + np.sqrt(2.0))[5] - 0.5) < 1.0e-6


def test_nfw_scale_free_mass(oracle_lib):
    # NFW.F90:550-571: m(1) = ln 2 - 1/2 (nfwNormalizationFactorUnitRadius), series branch continuous with the exact form
    assert abs(probe(oracle_lib, 1.0, 1.0, 1.0, 1.0)[6] - (np.log(2.0) - 0.5)) < 1.0e-15
    a, b = probe(oracle_lib, 1.0, 1.0, 1.0, 0.999999e-6)[6], probe(oracle_lib, 1.0, 1.0, 1.0, 1.000001e-6)[6]
    assert abs(a / b - (0.999999 / 1.000001) ** 2) < 1.0e-3  # the closed form loses ~4 digits to cancellation there


def test_nfw_known_answers_of_the_reference(oracle_lib):
    """source/tests/dark_matter_profiles.F90:60-73,301-362: NFW halo of concentration 8 at z = 0; enclosed mass fractions at
    r / r_s = 1/8 ... 8 (relTol 1e-6) and the radius recovered from the specific angular momentum of a circular orbit (relTol
    2e-4: the inverse tabulation of NFW.F90:589-639 that gives the structure solver its first guess)."""
    from galacticus_b200 import synthetic

    orc = oracle_lib
    p = orc.params_default(abi.GLC_MODEL_STANDARD)
    o = orc.Oracle()
    synthetic.install(o, p)
    L = orc.lib()
    L.orc_nfw_probe.restype = None
    L.orc_nfw_probe.argtypes = [C.POINTER(abi.glc_params), C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                np.ctypeslib.ndpointer(np.float64)]
    radius = [0.125, 0.25, 0.5, 1.0, 2.0, 4.0, 8.0]
    mass_expected = [5.099550982355504e-3, 1.768930674181593e-2, 5.513246746363203e-2, 1.476281525188409e-1,
                     3.301489257042704e-1, 6.186775455118112e-1, 1.000000000000000e+0]
    out = np.zeros(2)
    for x, m in zip(radius, mass_expected):
        L.orc_nfw_probe(C.byref(o.params), o.T, 1.0e12, 13.8, 8.0, x, out)
        assert abs(out[0] / m - 1.0) < 1.0e-6, (x, out[0], m)
        assert abs(out[1] / x - 1.0) < 2.0e-4, (x, out[1])
