"""Seeded synthetic node records shared by the oracle-vs-CUDA parity tests."""
import numpy as np

from galacticus_b200 import abi

P = abi.P


def box_nodes(n, seed=219, leaky=True, ragged=True):
    """Records for the closedBox/leakyBox operator set: disk gas + optional hot halo."""
    rng = np.random.default_rng(seed)
    props = np.zeros((n, abi.NPROP))
    flags = np.zeros(n, dtype=np.int32)
    t0 = rng.uniform(1.0, 12.0, n)
    dt = rng.uniform(0.05, 1.5, n)
    if ragged:
        dt[::7] = 0.0  # zero-length evolutions (timeStart == timeEnd)
    mgas = 10.0 ** rng.uniform(6, 11.5, n)
    mstar = np.where(rng.random(n) < 0.5, 0.0, 10.0 ** rng.uniform(5, 11, n))
    z = rng.uniform(0, 0.03, n)
    props[:, P["TIME"]] = t0
    props[:, P["TIME_STEP"]] = np.where(rng.random(n) < 0.5, -1.0, rng.uniform(1e-3, 0.5, n))
    props[:, P["DISK_MASS_GAS"]] = mgas
    props[:, P["DISK_ABUND_GAS"]] = z * mgas
    props[:, P["DISK_MASS_STELLAR"]] = mstar
    props[:, P["DISK_ABUND_STELLAR"]] = z * mstar * 0.5
    props[:, P["MASS_TARGET"]] = 10.0 ** rng.uniform(10, 13, n)
    props[:, P["MASS_RATE"]] = np.where(rng.random(n) < 0.5, 0.0, props[:, P["MASS_TARGET"]] * 0.05)
    props[:, P["TIME_TARGET"]] = t0 + dt
    props[:, P["BASIC_MASS"]] = props[:, P["MASS_TARGET"]] - props[:, P["MASS_RATE"]] * dt
    flags[:] = abi.GLC_F_HAS_DISK
    if leaky:
        hh = rng.random(n) < 0.8
        flags[hh] |= abi.GLC_F_HAS_HOTHALO
    if ragged:
        flags[::11] = 0  # nodes with no evolvable component at all
    return props, flags, t0 + dt


def reproducibility_box(leaky):
    """The single-node trees of testSuite/parameters/reproducibility/{closedBox,leakyBox}Tree.xml."""
    props = np.zeros((1, abi.NPROP))
    flags = np.array([abi.GLC_F_HAS_DISK | (abi.GLC_F_HAS_HOTHALO if leaky else 0)], dtype=np.int32)
    props[0, P["TIME"]] = 12.47
    props[0, P["TIME_STEP"]] = -1.0
    props[0, P["DISK_MASS_GAS"]] = 1.0e11
    props[0, P["MASS_TARGET"]] = 1.0e12
    props[0, P["TIME_TARGET"]] = 13.47
    props[0, P["BASIC_MASS"]] = 1.0e12
    props[0, P["SPIN_TARGET"]] = 1.635e12
    return props, flags, np.array([13.47])


def assert_close(a, b, rtol, scale=None, what=""):
    """|a-b| <= rtol*max(|a|,|b|) + rtol*scale  (scale = the ODE absolute-tolerance scale)."""
    a = np.asarray(a)
    b = np.asarray(b)
    s = 0.0 if scale is None else np.asarray(scale)
    err = np.abs(a - b)
    tol = rtol * np.maximum(np.abs(a), np.abs(b)) + rtol * s
    bad = err > tol
    if np.any(bad):
        idx = np.argwhere(bad)[:10]
        msg = "\n".join(f"  {tuple(i)}: {a[tuple(i)]!r} vs {b[tuple(i)]!r}" for i in idx)
        raise AssertionError(f"{what}: {bad.sum()} mismatches (rtol={rtol})\n{msg}")


def smoke_case(model_name):
    """(props, flags, t_end, params, tables) for __graft_entry__.smoke()."""
    from galacticus_b200.evolver import params_default

    if model_name == "box":
        p = params_default(abi.GLC_MODEL_BOX)
        p.box_timescaleStarFormation = 0.5
        p.box_fractionOutflow = 1.0
        props, flags, t_end = box_nodes(512, seed=11, leaky=True)
        return props, flags, t_end, p, {}
    if model_name == "standard":
        from galacticus_b200 import synthetic

        p = standard_params()
        props, flags, t_end = synthetic.standard_nodes(p, 512, seed=17)
        return props, flags, t_end, p, synthetic.standard_tables(p)
    raise NotImplementedError(model_name)


def standard_params(orc_or_none=None, with_black_holes=False):
    """quickTest.xml operator set; with_black_holes=False masks blackHolesSeed/Accretion/Winds (the reduced set the
    first golden vectors were made with), True is the full <nodeOperator value="multi"> list of quickTest.xml."""
    from galacticus_b200 import synthetic

    if orc_or_none is not None:
        p = orc_or_none.params_default(abi.GLC_MODEL_STANDARD)
    else:
        from galacticus_b200.evolver import params_default

        p = params_default(abi.GLC_MODEL_STANDARD)
    if not with_black_holes:
        p.operatorMask = abi.GLC_OP_ALL & ~(abi.GLC_OP_BLACK_HOLES_SEED | abi.GLC_OP_BLACK_HOLES_ACCRETION
                                            | abi.GLC_OP_BLACK_HOLES_WINDS)
    return synthetic.finalize_params(p)


BH_FRACTION = 0.7  # share of the synthetic nodes that start with a black hole; the others get the seed by interrupt


def standard_bh_nodes(p, n, seed, **kw):
    """Node records for the full operator set: black-hole masses from the seed to 1e9.5 Msun, spins over [0, 0.9999]."""
    from galacticus_b200 import synthetic

    return synthetic.standard_nodes(p, n, seed=seed, black_hole_fraction=BH_FRACTION, **kw)


def y_scale(props):
    """Per-node magnitude used as the absolute floor of the 1e-6 comparison (the ODE's own scales)."""
    m = np.abs(props[:, [P["HH_MASS"], P["DISK_MASS_GAS"], P["DISK_MASS_STELLAR"], P["SPH_MASS_GAS"],
                         P["SPH_MASS_STELLAR"], P["HH_OUTFLOWED_MASS"]]]).sum(axis=1)
    return np.maximum(m, 1.0)
