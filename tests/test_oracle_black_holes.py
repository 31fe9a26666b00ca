"""Known answers and invariants pinning the black-hole chain of the CPU checker:
black_holes/fundamentals.F90, accretion/Bondi_Hoyle_Lyttleton.F90, thermodynamics/ideal_gases.F90,
black_holes/accretion_rates/standard.F90, accretion_disks/{switched,Shakura_Sunyaev,ADAF}.F90, black_holes/winds/Ciotti2009.F90,
black_holes/CGM_heating/jet_power.F90, nodes/operators/physics/black_holes/{seed,accretion,winds}.F90."""
import ctypes as C

import numpy as np
import pytest

from galacticus_b200 import abi, synthetic
from tests import cases

P = abi.P


def _probe(orc, mass, spin, temperature):
    L = orc.lib()
    out = np.zeros(8)
    L.orc_bh_probe.argtypes = [C.c_double, C.c_double, C.c_double, np.ctypeslib.ndpointer(np.float64)]
    L.orc_bh_probe.restype = None
    L.orc_bh_probe(mass, spin, temperature, out)
    return out


def test_kerr_isco_known_answers(oracle_lib):
    """Bardeen, Press & Teukolsky (1972): Schwarzschild r=6, E=sqrt(8/9), L=2 sqrt(3); extreme Kerr r=1, E=1/sqrt(3),
    L=2/sqrt(3) (the reference's own series coefficients, fundamentals.F90:213-214,252-253)."""
    r, e, l = _probe(oracle_lib, 1.0e8, 0.0, 1.0e4)[:3]
    assert r == pytest.approx(6.0, rel=1e-14)
    assert e == pytest.approx(np.sqrt(8.0 / 9.0), rel=1e-14)
    assert l == pytest.approx(2.0 * np.sqrt(3.0), rel=1e-14)
    r, e, l = _probe(oracle_lib, 1.0e8, 1.0, 1.0e4)[:3]
    assert r == pytest.approx(1.0, abs=1e-12)
    assert e == pytest.approx(1.0 / np.sqrt(3.0), rel=1e-9)
    assert l == pytest.approx(2.0 / np.sqrt(3.0), rel=1e-9)
    # monotone: the ISCO moves inwards and the binding energy grows with prograde spin
    js = np.linspace(0.0, 0.9999, 200)
    v = np.array([_probe(oracle_lib, 1.0e8, j, 1.0e4)[:3] for j in js])
    assert (np.diff(v[:, 0]) < 0).all() and (np.diff(v[:, 1]) < 0).all() and (np.diff(v[:, 2]) < 0).all()
    # continuity across the switch to the near-extremal series at j = 0.99999
    a = _probe(oracle_lib, 1.0e8, 0.99999 - 1e-9, 1.0e4)[:3]
    b = _probe(oracle_lib, 1.0e8, 0.99999 + 1e-9, 1.0e4)[:3]
    np.testing.assert_allclose(a[1:], b[1:], rtol=1e-3)  # the reference's two-term series is good to 7e-4 there


def test_eddington_and_bondi_known_answers(oracle_lib):
    """Eddington rate 4 pi G M m_H / (sigma_T c) (fundamentals.F90:123-138) and the Bondi-Hoyle-Lyttleton radius and
    rate (Bondi_Hoyle_Lyttleton.F90:34-78) against an independent evaluation in SI units."""
    G, c, mH, sT, Msun, Gyr, Mpc, kB, amu = (6.673e-11, 2.99792458e8, 1.0078250322 * 1.660538782e-27, 6.65245893699e-29,
                                             1.98892e30, 1.0e9 * 3.15581497635456e7, 3.08567758135e22, 1.3806504e-23, 1.660538782e-27)
    mass, T = 3.0e7, 2.0e6
    out = _probe(oracle_lib, mass, 0.3, T)
    edd_si = 4.0 * np.pi * G * (mass * Msun) * mH / (sT * c)  # kg/s
    assert out[3] == pytest.approx(edd_si * Gyr / Msun, rel=1e-12)
    # the Salpeter e-folding time at 10 % efficiency is ~45 Myr
    assert 0.1 * mass / out[3] == pytest.approx(0.045, rel=0.02)
    mu = 1.0 / (2.0 * 0.7514 / 1.0078250322 + 3.0 * 0.2486 / 4.0026032545)
    cs = np.sqrt(5.0 * kB * T / 3.0 / mu / amu)  # m/s
    assert out[6] == pytest.approx(cs / 1.0e3, rel=1e-13)
    assert out[4] == pytest.approx(G * mass * Msun / cs**2 / Mpc, rel=1e-12)
    rate_si = 4.0 * np.pi * (G * mass * Msun) ** 2 * (Msun / Mpc**3) / cs**3  # kg/s at 1 Msun/Mpc^3
    assert out[5] == pytest.approx(rate_si * Gyr / Msun, rel=1e-12)
    assert out[7] == pytest.approx(cs / np.sqrt(G * Msun / Mpc**3) / Mpc, rel=1e-12)


def _rhs_all(orc, p, props, flags):
    o = orc.Oracle()
    synthetic.install(o, p)
    d = np.zeros((props.shape[0], abi.NY))
    code = np.zeros(props.shape[0], dtype=int)
    for i in range(props.shape[0]):
        d[i], code[i], _ = o.rhs(props[i], flags[i])
    return d, code


def test_accretion_moves_mass_from_reservoirs_to_the_black_hole(oracle_lib):
    """blackHolesAccretion (accretion.F90:113-173): the black hole gains (1 - eps_rad - eps_jet) of what the spheroid gas
    and the hot halo lose; spin rates vanish with the accretion rate; rates are finite for every regime."""
    p_off = cases.standard_params(with_black_holes=False)
    p_off.operatorMask &= ~abi.GLC_OP_CGM_COOLING_HEATING  # no jet heating in the reference run of the comparison
    p_on = cases.standard_params(with_black_holes=True)
    p_on.operatorMask &= ~(abi.GLC_OP_CGM_COOLING_HEATING | abi.GLC_OP_BLACK_HOLES_WINDS | abi.GLC_OP_BLACK_HOLES_SEED)
    props, flags, _ = cases.standard_bh_nodes(p_on, 1500, seed=77, fresh_fraction=0.0)
    d0, c0 = _rhs_all(oracle_lib, p_off, props, flags)
    d1, c1 = _rhs_all(oracle_lib, p_on, props, flags)
    assert np.isfinite(d1).all()
    has = (flags & abi.GLC_F_HAS_BH) != 0
    growth = d1[:, P["BH_MASS"]]
    assert (growth[~has] == 0).all() and (growth[has] != 0).mean() > 0.5
    removed = (d0 - d1)[:, [P["SPH_MASS_GAS"], P["HH_MASS"]]].sum(axis=1)  # what the sinks take
    ok = has & (growth != 0) & (c0 == 0) & (c1 == 0)
    # where the accretion rate is not lost in the rounding of the (much larger) other rates of the reservoirs
    big = ok & (np.abs(growth) > 1.0e-6 * (np.abs(d0[:, P["SPH_MASS_GAS"]]) + np.abs(d0[:, P["HH_MASS"]])))
    assert big.sum() > 100
    eff = 1.0 - growth[big] / removed[big]  # radiative + jet efficiency
    assert (removed[big] > 0).all()
    assert (eff > 0.0).all() and (eff < 2.5).all()  # jet efficiency is capped at 2 (efficiencyJetMaximum)
    # nothing but the black hole, the spheroid gas (mass, metals, angular momentum) and the hot gas changes
    touched = [P[k] for k in ("BH_MASS", "BH_SPIN", "SPH_MASS_GAS", "SPH_ABUND_GAS", "SPH_ANGMOM", "HH_MASS", "HH_ABUND", "HH_ANGMOM")]
    others = [k for k in range(abi.NY) if k not in touched]
    assert np.array_equal(d0[np.ix_(ok, others)], d1[np.ix_(ok, others)])
    # thin-disk spin-up has the sign of L_isco - 2 j E_isco > 0 below the equilibrium spin
    slow = ok & (props[:, P["BH_SPIN"]] < 0.5)
    assert (d1[slow, P["BH_SPIN"]] > 0).all()


def test_seed_operator_requests_creation(oracle_lib):
    """blackHolesSeed (seed.F90:153-182): nodes without a black hole interrupt with blackHoleCreate; the host procedure
    seeds mass 100 Msun and spin 0 (quickTest.xml:266-271), after which evolution proceeds."""
    p = cases.standard_params(with_black_holes=True)
    props, flags, t_end = cases.standard_bh_nodes(p, 400, seed=5)
    none = (flags & abi.GLC_F_HAS_BH) == 0
    _, code = _rhs_all(oracle_lib, p, props[none][:50], flags[none][:50])
    assert (code != abi.GLC_INT_NONE).all()  # some node may ask for another component first, the seed wins if asked last
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    po, fo = props.copy(), flags.copy()
    s, i, c = o.evolve_batch(po, fo, t_end, n_threads=4)
    assert (s == 0).all() and (i == 0).all()
    assert ((fo & abi.GLC_F_HAS_BH) != 0).all()
    assert (po[none, P["BH_MASS"]] >= 100.0 * (1 - 1e-12)).all()
    assert (po[:, P["BH_SPIN"]] >= 0).all() and (po[:, P["BH_SPIN"]] <= 0.9999).all()  # post-step clamp :422-462


def test_jet_heating_offsets_cooling(oracle_lib):
    """circumgalacticMediumHeatingAGNFeedback over blackHoleCGMHeatingJetPower: with a massive black hole the net
    cooling rate onto the disk drops (or turns into an outflow), never rises."""
    p = cases.standard_params(with_black_holes=True)
    p.operatorMask &= ~(abi.GLC_OP_BLACK_HOLES_ACCRETION | abi.GLC_OP_BLACK_HOLES_WINDS | abi.GLC_OP_BLACK_HOLES_SEED)
    props, flags, _ = cases.standard_bh_nodes(p, 1200, seed=31, fresh_fraction=0.0)
    d1, c1 = _rhs_all(oracle_lib, p, props, flags)
    f0 = flags & ~abi.GLC_F_HAS_BH
    d0, c0 = _rhs_all(oracle_lib, p, props, f0)
    ok = (c0 == 0) & (c1 == 0) & ((flags & abi.GLC_F_HAS_DISK) != 0)
    assert (d1[ok, P["DISK_MASS_GAS"]] <= d0[ok, P["DISK_MASS_GAS"]]).all()
    assert (d1[ok, P["DISK_MASS_GAS"]] < d0[ok, P["DISK_MASS_GAS"]]).any()


def test_adaf_jet_power_known_answers():
    """source/tests/accretion_disks.F90:37-66: jet power efficiency of an ADAF (pureADAF energy, exponential field enhancement,
    fitted viscosity, adiabatic index 1.444, efficiencyJetMaximum 2) at six spins, relTol 1e-3 -- the reference's own known
    answers for the construction-time tabulation the device interpolates (GLC_TABLE_ADAF), restated in galacticus_b200/adaf.py."""
    from galacticus_b200 import adaf

    disk = adaf.ADAF(energy="pureADAF", field="exponential", viscosity="fit", adiabatic_index=1.444, efficiency_jet_maximum=2.0)
    spin = [0.0, 0.2, 0.4, 0.6, 0.8, 0.95]
    expected = [2.993e-3, 3.916e-3, 6.571e-3, 1.564e-2, 5.246e-2, 4.119e-1]
    for j, e in zip(spin, expected):
        assert abs(disk.jet_efficiency(j) / e - 1.0) < 1.0e-3, (j, disk.jet_efficiency(j), e)
    # Kerr-metric helpers against the reference's black_hole_fundamentals unit test (:41-50)
    assert abs(adaf.isco_radius(0.0) - 6.0) < 1e-6 and abs(adaf.isco_radius(1.0) - 1.0) < 1e-6
    assert abs(adaf.horizon_radius(0.0) - 2.0) < 1e-6 and abs(adaf.horizon_radius(1.0) - 1.0) < 1e-6
    # the spin-up function changes sign at the equilibrium spin of Benson & Babul (2009), j ~ 0.92
    assert disk.spin_up(0.90) > 0.0 > disk.spin_up(0.93)


def test_path_runs_on_the_restated_adaf_tabulation(oracle_lib):
    """The restated tabulation uploaded as GLC_TABLE_ADAF: the kernel source reproduces the checker bit for bit on it too
    (black holes of at least 1e5 Msun: the spin equation of seed-mass holes is stiff with this table, see synthetic.adaf_table)."""
    from galacticus_b200 import adaf
    from tests import cases, emu

    p = cases.standard_params(oracle_lib, with_black_holes=True)
    props, flags, t_end = cases.standard_bh_nodes(p, 160, seed=23)
    keep = ((flags & abi.GLC_F_HAS_BH) != 0) & (props[:, P["BH_MASS"]] >= 1.0e5) & (props[:, P["BH_SPIN"]] < 0.99)
    props, flags, t_end = np.ascontiguousarray(props[keep]), np.ascontiguousarray(flags[keep]), np.ascontiguousarray(t_end[keep])
    assert props.shape[0] >= 20
    x, _, v = adaf.adaf_tabulations(count=2000)
    o = oracle_lib.Oracle()
    synthetic.install(o, p)
    o.set_table(abi.GLC_TABLE_ADAF, x, None, v)
    po, fo = props.copy(), flags.copy()
    so, io, co = o.evolve_batch(po, fo, t_end)
    assert (so == 0).all()
    e = emu.EmuEvolver(nslots=32, machine=True)
    synthetic.install(e, p)
    e.set_table(abi.GLC_TABLE_ADAF, x, None, v)
    pe, fe = props.copy(), flags.copy()
    se, ie, ce = e.evolve_batch(pe, fe, t_end)
    np.testing.assert_array_equal(se, so)
    assert np.array_equal(pe, po) and ce == co
    # and it is a different model from the stand-in: black-hole spins end elsewhere
    o2 = oracle_lib.Oracle()
    synthetic.install(o2, p)
    q = props.copy()
    o2.evolve_batch(q, flags.copy(), t_end)
    assert not np.array_equal(q[:, P["BH_SPIN"]], po[:, P["BH_SPIN"]])
