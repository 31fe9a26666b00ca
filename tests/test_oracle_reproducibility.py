"""Pin the structure solver / adiabatic contraction of the checker to the reference's own golden model
testSuite/test-reproducibility.py:70-117 ("adiabaticContraction": parameters
testSuite/parameters/reproducibility/adiabaticContraction.xml, tree adiabaticContractionTree.xml): an isothermal dark
matter halo of 1e12 Msun with a 1e10 Msun Hernquist spheroid of angular momentum 1e10, Gnedin et al. (2004) contraction
with A = omega = 1, structure solver tolerance 1e-4, no hot halo, no disk.  The model needs no external dataset: the
only tabulated input is the virial density contrast of the spherical collapse model, which
galacticus_b200.synthetic.spherical_collapse_virial_density_contrast restates from the reference's solver.  All four
assertions of the reference test are made at the reference's tolerances."""
import ctypes as C

import numpy as np
import pytest

from galacticus_b200 import abi, synthetic

P = abi.P
G = 4.3011827419096073e-9  # testSuite/test-reproducibility.py:15


def adiabatic_contraction_case(orc):
    p = orc.params_default(abi.GLC_MODEL_STANDARD)
    p.OmegaMatter, p.OmegaBaryon, p.HubbleConstant = 0.3, 0.05, 70.0  # adiabaticContraction.xml cosmologyParameters
    p.darkMatterProfileDMO = abi.GLC_DMO_ISOTHERMAL
    p.adiabaticA, p.adiabaticOmega = 1.0, 1.0
    p.structureSolutionTolerance = 1.0e-4
    # the operator list of the parameter file: stellar feedback (disks, spheroids), bar instability; no star formation
    p.operatorMask = (abi.GLC_OP_STELLAR_FEEDBACK_DISKS | abi.GLC_OP_STELLAR_FEEDBACK_SPHEROIDS | abi.GLC_OP_BAR_INSTABILITY)
    synthetic.finalize_params(p)
    tables = synthetic.standard_tables(p)
    # virialDensityContrast = sphericalCollapseClsnlssMttrCsmlgclCnstnt
    tables[abi.GLC_TABLE_HALO_MEAN_DENSITY] = synthetic.spherical_collapse_mean_density_table(p, t_min=5.0, t_max=20.0)
    cosmo = synthetic.Cosmology(p)
    props = np.zeros((1, abi.NPROP))
    flags = np.array([abi.GLC_F_HAS_SPHEROID], dtype=np.int32)
    r = props[0]
    r[P["TIME"]] = 13.46  # adiabaticContractionTree.xml node 1
    r[P["TIME_STEP"]] = -1.0
    r[P["BASIC_MASS"]] = r[P["MASS_TARGET"]] = 1.0e12
    r[P["TIME_TARGET"]] = 13.48
    r[P["DMSCALE"]] = r[P["DMSCALE_TARGET"]] = 0.03
    r[P["SAT_BOUND_MASS"]] = 1.0e12
    r[P["SPH_MASS_STELLAR"]] = 1.0e10
    r[P["SPH_ANGMOM"]] = 1.0e10
    t_out = np.array([float(cosmo.time_of_redshift(1.0e-4))])  # output 1 of outputRedshifts "0.00 0.0001"
    return p, tables, props, flags, t_out


@pytest.fixture(scope="module")
def solved(oracle_lib):
    p, tables, props, flags, t_out = adiabatic_contraction_case(oracle_lib)
    o = oracle_lib.Oracle()
    o.set_params(p)
    for tid, (x0, x1, v) in tables.items():
        o.set_table(tid, x0, x1, v)
    status, interrupt, _ = o.evolve_batch(props, flags, t_out)
    assert status[0] == 0 and interrupt[0] == 0
    row = props[0]
    out = np.zeros(5)
    L = o.L
    L.orc_rotation_curve_probe.restype = None
    L.orc_rotation_curve_probe.argtypes = [C.POINTER(abi.glc_params), C.c_void_p, np.ctypeslib.ndpointer(np.float64), C.c_int, C.c_double,
                                           np.ctypeslib.ndpointer(np.float64)]
    L.orc_rotation_curve_probe(C.byref(p), o.T, row.copy(), int(flags[0]), float(row[P["SPH_RADIUS"]]), out)
    return row, out


def test_spheroid_radius_golden(solved):
    row, _ = solved
    assert abs(row[P["SPH_RADIUS"]] / 0.00360702918954165 - 1.0) < 2.0e-4  # test-reproducibility.py:74-80


def test_spheroid_angular_momentum(solved):
    row, _ = solved
    v = row[P["SPH_RADIUS"]] * row[P["SPH_VELOCITY"]] * row[P["SPH_MASS_STELLAR"]] / row[P["SPH_ANGMOM"]]
    assert abs(v / 0.5 - 1.0) < 2.0e-4  # :81-87


def test_rotation_curve_at_spheroid_radius(solved):
    row, rc = solved
    assert abs(rc[0] / row[P["SPH_VELOCITY"]] - 1.0) < 2.0e-4  # :88-94


def test_initial_specific_angular_momentum(solved):
    # :95-116: the adiabatic invariant r_i M_i(r_i) = r M(r) of Gnedin et al. (2004) for the isothermal halo
    row, rc = solved
    r, v, m = row[P["SPH_RADIUS"]], row[P["SPH_VELOCITY"]], row[P["BASIC_MASS"]]
    rvir, vvir = rc[3], rc[4]
    value = np.sqrt(0.84333333) * vvir / v * (rvir * (rc[1] ** 2 * r / G / 0.83333333) / m) / r
    assert abs(value - 1.0) < 3.0e-3
