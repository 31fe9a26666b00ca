"""Synthetic tabulated inputs and merger trees of the same SHAPE as the reference's datasets.

The external datasets the reference reads (Cloudy CIE cooling/chemical-state tables,
``atomic_CIE_Cloudy.F90:63``) are not available offline; BASELINE.json's north_star prescribes
synthetic tables of the same shape, fed identically to the CPU checker and to the CUDA path.  Everything
here is host-side input generation: it never evaluates the hot path.
"""
from __future__ import annotations

import numpy as np

from . import abi

P = abi.P

# numerical constants as in source/numerical/constants (GSL 2.6 MKSA values)
G_INTERNAL = 6.673e-11 * 1.98892e30 / 1.0e6 / (1.0e6 * 3.08567758135e16)
MPC_PER_KMS_TO_GYR = (1.0e6 * 3.08567758135e16) / 1.0e3 / (1.0e9 * 3.15581497635456e7)


# ------------------------------------------------------------------ cosmology (matterLambda)
class Cosmology:
    """Flat matter+Lambda cosmology (cosmologyFunctionsMatterLambda) in closed form."""

    def __init__(self, params: abi.glc_params):
        self.Om = params.OmegaMatter
        self.OL = 1.0 - self.Om
        self.Ob = params.OmegaBaryon
        self.H0 = params.HubbleConstant / MPC_PER_KMS_TO_GYR  # 1/Gyr
        self.rho_crit0 = 3.0 * params.HubbleConstant**2 / (8.0 * np.pi * G_INTERNAL)  # Msun/Mpc^3

    def expansion_factor(self, t):
        return (self.Om / self.OL) ** (1.0 / 3.0) * np.sinh(1.5 * np.sqrt(self.OL) * self.H0 * t) ** (2.0 / 3.0)

    def time_of_redshift(self, z):
        a = 1.0 / (1.0 + z)
        return np.arcsinh(np.sqrt(self.OL / self.Om) * a**1.5) / (1.5 * np.sqrt(self.OL) * self.H0)

    def hubble(self, a):
        return self.H0 * np.sqrt(self.Om / a**3 + self.OL)

    def virial_mean_density(self, t):
        """Stand-in for virialDensityContrastSphericalCollapse...: Bryan & Norman (1998) fit."""
        a = self.expansion_factor(t)
        e2 = self.Om / a**3 + self.OL
        om_a = self.Om / a**3 / e2
        x = om_a - 1.0
        delta_mean = (18.0 * np.pi**2 + 82.0 * x - 39.0 * x**2) / om_a
        return delta_mean * self.Om * self.rho_crit0 / a**3


def halo_mean_density_table(params, t_min=0.01, t_max=30.0, per_decade=100):
    """GLC_TABLE_HALO_MEAN_DENSITY: log-uniform in time, 100 points per decade
    (dark_matter_halos/scales/virial_density_contrast.F90:386-414)."""
    cosmo = Cosmology(params)
    n = int(np.ceil(np.log10(t_max / t_min) * per_decade)) + 1
    t = t_min * 10.0 ** (np.arange(n) / per_decade)
    rho = cosmo.virial_mean_density(t)
    eps = 1.0e-5
    dln = (np.log(cosmo.virial_mean_density(t * (1 + eps))) - np.log(cosmo.virial_mean_density(t * (1 - eps)))) / (2 * eps * t)
    return t, None, np.stack([rho, dln], axis=1)


# ------------------------------------------------------------------ CIE tables
def cie_grids():
    temperatures = 10.0 ** np.arange(2.0, 9.0001, 0.1)
    metallicities = np.array([0.0, 1.0e-3, 1.0e-2, 1.0e-1, 0.3, 1.0, 3.0, 10.0, 30.0])  # Solar units
    return metallicities, temperatures


def cooling_function_table():
    """GLC_TABLE_COOLING_FUNCTION with the layout of CIE_file.F90:30-134 (Lambda/n_H^2, erg cm^3/s)."""
    Z, T = cie_grids()
    lgT = np.log10(T)
    cutoff = 1.0 / (1.0 + (1.2e4 / T) ** 8) + 1.0e-12
    prim = (2.5e-27 * np.sqrt(T) + 7.0e-23 * np.exp(-(((lgT - 4.3) / 0.25) ** 2))
            + 2.0e-23 * np.exp(-(((lgT - 5.0) / 0.3) ** 2))) * cutoff
    metal = 1.2e-22 * np.exp(-(((lgT - 5.4) / 0.6) ** 2)) * cutoff
    table = prim[None, :] + Z[:, None] * metal[None, :]
    return Z, T, table


def electron_fraction_table():
    """GLC_TABLE_ELECTRON_FRACTION: n_e / n_H on the same grid (chemical/state/CIE_file.F90)."""
    Z, T = cie_grids()
    ion = 1.0 / (1.0 + (1.5e4 / T) ** 6)
    table = (1.17 * ion[None, :] + 1.0e-4) + 0.01 * Z[:, None] * ion[None, :]
    return Z, T, table


def disk_rotation_curve_table(per_decade=100, x_min=1.0e-6, x_max=1.0e2):
    """GLC_TABLE_DISK_ROTATION_CURVE: x^2 [I0 K0 - I1 K1](x) on the reference's lattice
    (exponential_disk.F90:112-113,686: rotationCurveHalfRadiusMinimumDefault=1e-6, 100 points/decade)."""
    from scipy import special

    n = int(round(np.log10(x_max / x_min) * per_decade)) + 1
    x = x_min * 10.0 ** (np.arange(n) / per_decade)
    f = x**2 * (special.i0e(x) * special.k0e(x) - special.i1e(x) * special.k1e(x))
    return x, None, f


def adaf_table(count=10000, x_min=1.0e-6, x_max=1.0):
    """GLC_TABLE_ADAF: the two tabulations the accretionDisksADAF constructor builds (accretion_disks/ADAF.F90:394-447)
    on its lattice (table1DLogarithmicLinear in the inverse spin 1-j on [1e-6, 1], countTable = 10000 points).  This is a
    smooth STAND-IN of the same shape for the synthetic workloads: a jet efficiency that rises steeply towards j = 1 and is
    capped at efficiencyJetMaximum = 2, and a spin-up function that changes sign at the equilibrium spin j ~ 0.93.  The
    reference's own tabulation (the Benson & Babul 2009 structure functions) is restated in galacticus_b200/adaf.py
    (adaf_tabulations, pinned by the reference's unit test) and can be uploaded instead; with it the spin equation of
    seed-mass black holes in massive haloes is stiff (spin-up ratio -53 at j = 0.999, -1.2e5 at 0.9999) and 3 of 20 000 bench
    nodes end in errorStatusUnderflow in the checker and on the device alike, so the benchmark workloads keep this one."""
    x = np.exp(np.linspace(np.log(x_min), np.log(x_max), count))
    x[0], x[-1] = x_min, x_max
    j = 1.0 - x
    speed_light_kms = 2.99792458e8 / 1.0e3
    efficiency_jet = np.minimum(2.0, 0.002 + 0.1 * j**2 + 0.02 * j**2 / x**0.7)
    power_jet = efficiency_jet * speed_light_kms**2
    spin_up = 2.0 * (1.0 - j / 0.93)
    return x, None, np.stack([power_jet, spin_up], axis=1)


def standard_tables(params):
    return {
        abi.GLC_TABLE_COOLING_FUNCTION: cooling_function_table(),
        abi.GLC_TABLE_ELECTRON_FRACTION: electron_fraction_table(),
        abi.GLC_TABLE_HALO_MEAN_DENSITY: halo_mean_density_table(params),
        abi.GLC_TABLE_DISK_ROTATION_CURVE: disk_rotation_curve_table(),
        abi.GLC_TABLE_ADAF: adaf_table(),
    }


def finalize_params(params):
    """Fill host-computed constants: timeReionization from redshiftReionization=10.5
    (parameters/quickTest.xml:101-104)."""
    if params.model == abi.GLC_MODEL_STANDARD:
        params.timeReionization = float(Cosmology(params).time_of_redshift(10.5))
    return params


def install(target, params):
    """Give parameters and tables to any object exposing set_params/set_table (e.g. an Evolver)."""
    finalize_params(params)
    target.set_params(params)
    if params.model == abi.GLC_MODEL_STANDARD:
        for tid, (x0, x1, v) in standard_tables(params).items():
            target.set_table(tid, x0, x1, v)


# ------------------------------------------------------------------ halo helpers (inputs only)
def virial_radius(params, mass, t):
    rho = Cosmology(params).virial_mean_density(t)
    return np.cbrt(3.0 * mass / (4.0 * np.pi * rho))


def standard_nodes(params, n, seed=219, fresh_fraction=0.3, satellite_fraction=0.2, black_hole_fraction=0.0):
    """Seeded node records spanning the quickTest mass range (1e10..1e13 Msun) at 1..13 Gyr.  A fraction
    black_hole_fraction of the nodes carries a black hole (drawn from a separate stream so that the other properties do
    not depend on it); the others get their seed by the blackHolesSeed interrupt when that operator is enabled."""
    rng = np.random.default_rng(seed)
    cosmo = Cosmology(params)
    props = np.zeros((n, abi.NPROP))
    flags = np.zeros(n, dtype=np.int32)
    t0 = rng.uniform(0.8, 12.5, n)
    dt = rng.uniform(0.05, 0.8, n)
    mass = 10.0 ** rng.uniform(10.0, 13.0, n)
    growth = rng.uniform(0.0, 0.4, n) * mass / 1.0  # Msun/Gyr
    sat = rng.random(n) < satellite_fraction
    growth[sat] = 0.0
    rvir = virial_radius(params, mass, t0)
    vvir = np.sqrt(G_INTERNAL * mass / rvir)
    conc = rng.uniform(4.0, 15.0, n)
    lam = 0.04326 * np.exp(rng.normal(0.0, 0.5, n))
    jhalo = np.sqrt(2.0) * lam * mass * rvir * vvir
    fb = params.OmegaBaryon / params.OmegaMatter

    props[:, P["TIME"]] = t0
    props[:, P["TIME_STEP"]] = np.where(rng.random(n) < 0.5, -1.0, rng.uniform(1e-3, 0.3, n))
    props[:, P["TIME_TARGET"]] = t0 + dt * rng.uniform(1.0, 1.5, n)
    props[:, P["MASS_RATE"]] = growth
    props[:, P["MASS_TARGET"]] = mass + growth * (props[:, P["TIME_TARGET"]] - t0)
    props[:, P["BASIC_MASS"]] = mass
    props[:, P["DMSCALE"]] = rvir / conc
    props[:, P["DMSCALE_RATE"]] = np.where(sat, 0.0, rng.uniform(-0.1, 0.1, n) * rvir / conc)
    props[:, P["DMSCALE_TARGET"]] = props[:, P["DMSCALE"]] + props[:, P["DMSCALE_RATE"]] * (props[:, P["TIME_TARGET"]] - t0)
    props[:, P["SPIN"]] = jhalo
    # the halo gains angular momentum with its mass at (roughly) fixed spin parameter, J ~ M^(5/3): the specific
    # angular momentum of freshly accreted gas is then that of the halo (haloAngularMomentumInterpolate tracks
    # the same relation between neighbouring tree nodes); an uncorrelated rate would hand newly created hot
    # haloes almost no angular momentum and make parsec-sized, pathologically stiff disks
    props[:, P["SPIN_RATE"]] = np.where(sat, 0.0, (5.0 / 3.0) * jhalo * growth / mass * rng.uniform(0.7, 1.3, n))
    props[:, P["SPIN_TARGET"]] = jhalo + props[:, P["SPIN_RATE"]] * (props[:, P["TIME_TARGET"]] - t0)
    props[:, P["TIME_LAST_ISOLATED"]] = np.where(sat, t0 * rng.uniform(0.6, 1.0, n), 0.0)
    props[:, P["SAT_BOUND_MASS"]] = mass
    props[:, P["MASS_BARYONIC_SUBHALOS"]] = np.where(rng.random(n) < 0.3, fb * mass * rng.uniform(0, 0.2, n), 0.0)

    # hot halo
    has_hh = rng.random(n) < 0.95
    mhh = fb * mass * rng.uniform(0.05, 1.0, n)
    zhh = rng.uniform(0.0, 0.02, n) * (rng.random(n) < 0.8)
    props[:, P["HH_MASS"]] = np.where(has_hh, mhh, 0.0)
    props[:, P["HH_ABUND"]] = np.where(has_hh, mhh * zhh, 0.0)
    props[:, P["HH_ANGMOM"]] = np.where(has_hh, jhalo * mhh / mass, 0.0)
    mout = np.where(rng.random(n) < 0.5, mhh * rng.uniform(0, 0.5, n), 0.0)
    props[:, P["HH_OUTFLOWED_MASS"]] = np.where(has_hh, mout, 0.0)
    props[:, P["HH_OUTFLOWED_ABUND"]] = np.where(has_hh, mout * zhh, 0.0)
    props[:, P["HH_OUTFLOWED_ANGMOM"]] = np.where(has_hh, jhalo * mout / mass, 0.0)
    props[:, P["HH_UNACCRETED_MASS"]] = np.where(has_hh & (rng.random(n) < 0.3), fb * mass * 0.1, 0.0)
    init = has_hh & (rng.random(n) > fresh_fraction)
    props[:, P["HH_OUTER_RADIUS"]] = np.where(init, rvir * rng.uniform(0.3, 1.0, n), 0.0)
    flags[has_hh] |= abi.GLC_F_HAS_HOTHALO
    flags[init] |= abi.GLC_F_HH_INITIALIZED
    flags[sat] |= abi.GLC_F_IS_SATELLITE

    # disk
    has_d = has_hh & (rng.random(n) < 0.7)
    md = fb * mass * 10.0 ** rng.uniform(-3.0, -0.5, n)
    fgas = rng.uniform(0.05, 1.0, n)
    zd = rng.uniform(1.0e-4, 0.03, n)
    rd = rvir * lam / np.sqrt(2.0) * rng.uniform(0.5, 1.5, n)
    props[:, P["DISK_MASS_GAS"]] = np.where(has_d, md * fgas, 0.0)
    props[:, P["DISK_MASS_STELLAR"]] = np.where(has_d, md * (1 - fgas), 0.0)
    props[:, P["DISK_ABUND_GAS"]] = np.where(has_d, md * fgas * zd, 0.0)
    props[:, P["DISK_ABUND_STELLAR"]] = np.where(has_d, md * (1 - fgas) * zd * 0.7, 0.0)
    props[:, P["DISK_ANGMOM"]] = np.where(has_d, 2.0 * md * rd * vvir * rng.uniform(0.8, 1.3, n), 0.0)
    warm = has_d & (rng.random(n) > fresh_fraction)
    props[:, P["DISK_RADIUS"]] = np.where(warm, rd, 0.0)
    props[:, P["DISK_VELOCITY"]] = np.where(warm, vvir * rng.uniform(0.8, 1.5, n), 0.0)
    flags[has_d] |= abi.GLC_F_HAS_DISK

    # spheroid
    has_s = has_d & (rng.random(n) < 0.4)
    ms = md * 10.0 ** rng.uniform(-2.0, 0.5, n)
    fgs = rng.uniform(0.0, 0.5, n)
    rs = rd * rng.uniform(0.1, 0.6, n)
    props[:, P["SPH_MASS_GAS"]] = np.where(has_s, ms * fgs, 0.0)
    props[:, P["SPH_MASS_STELLAR"]] = np.where(has_s, ms * (1 - fgs), 0.0)
    props[:, P["SPH_ABUND_GAS"]] = np.where(has_s, ms * fgs * zd, 0.0)
    props[:, P["SPH_ABUND_STELLAR"]] = np.where(has_s, ms * (1 - fgs) * zd, 0.0)
    props[:, P["SPH_ANGMOM"]] = np.where(has_s, ms * rs * vvir * 2.0 * rng.uniform(0.8, 1.3, n), 0.0)
    warm_s = has_s & (rng.random(n) > fresh_fraction)
    props[:, P["SPH_RADIUS"]] = np.where(warm_s, rs, 0.0)
    props[:, P["SPH_VELOCITY"]] = np.where(warm_s, vvir * rng.uniform(0.8, 1.8, n), 0.0)
    flags[has_s] |= abi.GLC_F_HAS_SPHEROID

    # black hole: masses from the seed (100 Msun) to 1e9.5 Msun, capped at 1 % of the halo's baryons so that all
    # accretion regimes (ADAF / thin disk / Eddington-limited) occur; spins over [0, 0.9999]
    if black_hole_fraction > 0.0:
        rb = np.random.default_rng(seed + 100003)
        has_b = rb.random(n) < black_hole_fraction
        mb = np.minimum(10.0 ** rb.uniform(2.0, 9.5, n), 0.01 * fb * mass)
        jb = rb.uniform(0.0, 0.998, n)
        u = rb.random(n)
        jb = np.where(u < 0.1, 0.0, np.where(u > 0.95, 0.9999, jb))
        props[:, P["BH_MASS"]] = np.where(has_b, mb, 0.0)
        props[:, P["BH_SPIN"]] = np.where(has_b, jb, 0.0)
        flags[has_b] |= abi.GLC_F_HAS_BH
    return props, flags, t0 + dt


# ------------------------------------------------------------------ synthetic forests (inputs of glc_forest_evolve)
def binary_split_forest(params, n_trees, mass_root, mass_resolution, seed=219, time_min=0.6, step=(0.04, 0.10),
                        mass_root_max=None):
    """Seeded binary-split merger trees as flat arrays (parent, mass, time, scale_radius, angular_momentum), roots at the
    present day.  Stand-in for mergerTreeBuilderCole2000 (tree building stays on the host and is outside the hot path,
    SURVEY 8d/8f): going back in time every halo takes a step dt = t U(step), loses a smoothly accreted fraction and, with
    a probability that grows with M / m_res, splits off a secondary progenitor with mass ratio q drawn from
    dn/dq ~ q^-1.5 above the resolution; node counts scale like M / m_res.  Built level by level (vectorised).
    Roots have mass mass_root (or log-uniform in [mass_root, mass_root_max])."""
    rng = np.random.default_rng(seed)
    cosmo = Cosmology(params)
    t0 = float(cosmo.time_of_redshift(0.0))
    if np.ndim(mass_root) == 1:  # one root mass per tree (e.g. drawn from a halo mass function)
        m_root = np.asarray(mass_root, dtype=float).copy()
        assert m_root.shape[0] == n_trees
    elif mass_root_max is None:
        m_root = np.full(n_trees, float(mass_root))
    else:
        m_root = 10.0 ** rng.uniform(np.log10(mass_root), np.log10(mass_root_max), n_trees)
    parent = [np.full(n_trees, -1, dtype=np.int64)]
    mass = [m_root]
    time = [np.full(n_trees, t0)]
    tree = [np.arange(n_trees)]
    front_idx = np.arange(n_trees)
    front_m, front_t, front_tree = m_root.copy(), np.full(n_trees, t0), np.arange(n_trees)
    n_total = n_trees
    while front_idx.size:
        k = front_idx.size
        dt = front_t * rng.uniform(step[0], step[1], k)
        t_child = front_t - dt
        alive = (t_child > time_min) & (front_m > mass_resolution)
        smooth = rng.uniform(0.0, 0.04, k)
        m_avail = front_m * (1.0 - smooth)
        q_min = mass_resolution / np.maximum(m_avail, mass_resolution)
        p_split = np.clip(0.25 + 0.12 * np.log10(np.maximum(front_m / mass_resolution, 1.0)), 0.0, 0.85)
        split = alive & (rng.random(k) < p_split) & (q_min < 0.5)
        u = rng.random(k)
        qm = np.minimum(q_min, 0.5)
        q = (qm ** -0.5 - u * (qm ** -0.5 - 0.5 ** -0.5)) ** -2.0  # dn/dq ~ q^-1.5 on [q_min, 0.5]
        m1 = np.where(split, m_avail * (1.0 - q), m_avail)
        m2 = m_avail * q
        keep1 = alive & (m1 > mass_resolution)
        keep2 = split & (m2 > mass_resolution) & keep1
        n1, n2 = int(keep1.sum()), int(keep2.sum())
        idx1 = n_total + np.arange(n1)
        idx2 = n_total + n1 + np.arange(n2)
        n_total += n1 + n2
        parent += [front_idx[keep1], front_idx[keep2]]
        mass += [m1[keep1], m2[keep2]]
        time += [t_child[keep1], t_child[keep2]]
        tree += [front_tree[keep1], front_tree[keep2]]
        front_idx = np.concatenate([idx1, idx2])
        front_m = np.concatenate([m1[keep1], m2[keep2]])
        front_t = np.concatenate([t_child[keep1], t_child[keep2]])
        front_tree = np.concatenate([front_tree[keep1], front_tree[keep2]])
    parent = np.concatenate(parent).astype(np.int32)
    mass = np.concatenate(mass)
    time = np.concatenate(time)
    tree = np.concatenate(tree).astype(np.int32)
    lam = (0.04326 * np.exp(rng.normal(0.0, 0.5, n_trees)))[tree]
    rvir = virial_radius(params, mass, time)
    vvir = np.sqrt(G_INTERNAL * mass / rvir)
    conc = np.clip(9.0 * (mass / 1.0e12) ** -0.1 * (time / t0) ** 0.7, 3.0, 20.0)
    return {"parent": parent, "mass": mass, "time": time, "scale_radius": rvir / conc,
            "angular_momentum": np.sqrt(2.0) * lam * mass * rvir * vvir, "tree": tree}



def forest_subset(forest, n_trees):
    """The first n_trees trees of a forest (trees are independent: a sample of the SAME trees for the CPU baseline)."""
    keep = forest["tree"] < n_trees
    new_index = np.cumsum(keep) - 1
    parent = forest["parent"][keep]
    parent = np.where(parent >= 0, new_index[np.maximum(parent, 0)], -1).astype(np.int32)
    out = {k: v[keep] for k, v in forest.items() if k != "parent"}
    out["parent"] = parent
    return out


def mass_function_roots(n_trees, mass_min=1.0e10, mass_max=1.0e14, slope=-0.9, mass_star=1.0e14, seed=219):
    """Root masses of a Monte Carlo volume (BASELINE.json configs[3]): dn/dlnM ~ M^slope exp(-M/M*) on
    [mass_min, mass_max], a Tinker et al. (2008)-like shape, by inverse-transform sampling on a fine lattice in ln M."""
    rng = np.random.default_rng(seed)
    lnm = np.linspace(np.log(mass_min), np.log(mass_max), 4097)
    m = np.exp(lnm)
    pdf = m**slope * np.exp(-m / mass_star)
    cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(lnm))])
    cdf /= cdf[-1]
    return np.exp(np.interp(rng.random(n_trees), cdf, lnm))


def mass_function_forest(params, n_trees, mass_resolution=5.0e9, seed=219, **kw):
    """configs[3]: halo-mass-function-sampled Monte Carlo trees at the quickTest mass resolution (mergerTreeMassResolution
    fixed, default 5e9 Msun: merger_trees/construct/build/mass_resolution/fixed.F90:61-64)."""
    roots = mass_function_roots(n_trees, seed=seed, **kw)
    return binary_split_forest(params, n_trees, roots, mass_resolution, seed=seed + 1)


# ------------------------------------------------------------------ spherical collapse (host-side table input)
def spherical_collapse_virial_density_contrast(params, t):
    """Virial density contrast (relative to the mean matter density at time t [Gyr]) of the spherical collapse model
    for collisionless matter plus a cosmological constant: what virialDensityContrastSphericalCollapseClsnlssMttrCsmlgclCnstnt
    tabulates (structure_formation/spherical_collapse/solver/collisionlessMatter_cosmologicalConstant.F90:348-497,
    520-640).  For each epoch: find the perturbation amplitude epsilon whose collapse time -- twice the time to
    turnaround, the integral of sqrt(r / (Om + eps r + OL r^3)) in units of the epochal Hubble time with the analytic
    correction near turnaround -- equals t; the turnaround radius solves Om/r + eps + OL r^2 = 0; the ratio of virial
    to turnaround radius solves the cubic energy equation 2 eta x^3 - (2 + eta) x + 1 = 0 with eta = 2 (OL/Om) r_ta^3
    (Lahav et al. 1991); Delta = 1 / (x r_ta)^3.  Host-side input generation for GLC_TABLE_HALO_MEAN_DENSITY: the
    reference builds the same table once at start-up, outside the hot path."""
    from scipy import integrate, optimize

    cosmo = Cosmology(params)
    t = np.atleast_1d(np.asarray(t, dtype=float))
    out = np.empty_like(t)
    for i, ti in enumerate(t):
        a = float(cosmo.expansion_factor(ti))
        e2 = cosmo.Om / a**3 + cosmo.OL
        om, ol = cosmo.Om / a**3 / e2, cosmo.OL / e2
        hubble = float(cosmo.hubble(a))  # 1/Gyr

        def radius_max(eps):
            lo, hi = -om / eps, (om / ol / 2.0) ** (1.0 / 3.0)
            f = lambda r: om / r + eps + ol * r * r
            if f(hi) > 0.0:
                return hi
            if f(lo) < 0.0:
                return lo
            return optimize.brentq(f, lo, hi, xtol=1.0e-300, rtol=1.0e-12)

        def time_collapse(eps):
            rta = radius_max(eps)
            ru = (1.0 - 1.0e-4) * rta

            def integrand(r):
                s = om + eps * r + ol * r**3
                return np.sqrt(r / s) if s > 0.0 else 0.0

            tt = integrate.quad(integrand, 0.0, ru, epsrel=1.0e-9, limit=400)[0] / hubble
            tt -= 2.0 * np.sqrt(om / ru + eps + ol * ru**2) / (2.0 * ol * ru - om / ru**2) / hubble
            return 2.0 * tt

        eps_max = -((27.0 / 4.0 * ol * om**2) ** (1.0 / 3.0))
        lo, hi = -10.0, eps_max * (1.0 + 1.0e-9)
        while time_collapse(lo) > ti:  # more negative = collapses earlier
            lo *= 2.0
        eps = optimize.brentq(lambda e: time_collapse(e) - ti, lo, hi, xtol=1.0e-300, rtol=1.0e-10)
        rta = radius_max(eps)
        eta = 2.0 * (ol / om) * rta**3
        roots = np.roots([2.0 * eta, 0.0, -(2.0 + eta), 1.0])
        x = min(float(r.real) for r in roots if abs(r.imag) < 1.0e-9 and 0.0 < r.real < 1.0)
        out[i] = 1.0 / (x * rta) ** 3
    return out


def spherical_collapse_mean_density_table(params, t_min=0.01, t_max=30.0, per_decade=100):
    """GLC_TABLE_HALO_MEAN_DENSITY with the spherical-collapse contrast instead of the Bryan & Norman fit (the layout of
    halo_mean_density_table): mean matter density x Delta_vir(t), and its logarithmic time derivative."""
    cosmo = Cosmology(params)
    n = int(np.ceil(np.log10(t_max / t_min) * per_decade)) + 1
    t = t_min * 10.0 ** (np.arange(n) / per_decade)

    def rho(tt):
        a = cosmo.expansion_factor(tt)
        return spherical_collapse_virial_density_contrast(params, tt) * cosmo.Om * cosmo.rho_crit0 / a**3

    r = rho(t)
    eps = 1.0e-4
    dln = (np.log(rho(t * (1 + eps))) - np.log(rho(t * (1 - eps)))) / (2 * eps * t)
    return t, None, np.stack([r, dln], axis=1)
