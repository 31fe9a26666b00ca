/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * "Standard" model: CPU restatement of the RHS of parameters/quickTest.xml as evaluated by
 * standardDerivativesCompute (source/merger_trees/node_evolver/standard.F90:1019-1061) through
 * nodeOperatorMulti (source/nodes/operators/multi.F90:313-332), with the standard components'
 * scale-set, pre-evolve, post-step and post-evolve hooks.  Each function cites what it restates
 * (paths relative to /root/reference/source).
 *
 * PARITY: unpinned against a reference run (no Fortran toolchain, no Cloudy/ADAF datasets here);
 * pinned only through analytic limits and invariants in tests/ (mass conservation a la
 * testSuite/test-mass-conservation-standard.py, beta-profile normalisation identity
 * mass_distributions/spherical/beta_profile.F90:265, closed-form cooling radius).
 *
 * Documented deviations from quickTest.xml (see DESIGN.md "out of scope / next"):
 *  - hotHaloRamPressureStripping: the "virialRadius" class instead of font2008
 *    (hot_halo/ramPressureStripping/virialRadius: stripping radius == virial radius);
 *  - black-hole operators are gated by glc_params.operatorMask (not yet restated);
 *  - the first-guess radius of a just-created component solves j^2 = G M_NFW(<r) r directly
 *    instead of inverting the reference's tabulated relation (NFW.F90:589-625);
 *  - the equilibrium solver's oscillation history is reset at every solve
 *    (equilibrium.F90:441-476 keeps it in thread-level saved storage);
 *  - beta-profile supports beta = 2/3 only (the quickTest default; beta_profile.F90:245-259).
 */
#include <float.h>
#include <math.h>

#include "../galacticus_b200/csrc/glc_detmath.h"
#include "../galacticus_b200/csrc/glc_specfun.h"
#include <string.h>

#include "orc_constants.h"
#include "orc_node.h"
#include "orc_numerics.h"

#define F_HH_INIT GLC_F_HH_INITIALIZED

typedef struct std_work {
    orc_evolve_ctx *c;
    const glc_params *P;
    const orc_tables *T;
    double *p;
    int flags;
    double time;
    /* halo scales (virialDensityContrastDefinition), memoised per Calculations_Reset */
    int halo_done;
    double rho_mean, rvir, vvir, tdyn, tvir, dlnrho_dt;
    /* hot-halo beta profile */
    int hh_done, hh_valid;
    double hh_router, hh_rcore, hh_rho0, hh_mass;
    /* cooling */
    int rcool_done, rcool_rate_done;
    double rcool, rcool_rate, cool_z, cool_tavail;
    /* disk SFR memo (krumholz2009Unchanged) */
    int sfr_done;
    double sfr_disk;
    /* Krumholz factors */
    double k_xh, k_zsolar, k_chi, k_sigma_norm, k_s_norm, k_sigma_trunc, k_mgas, k_rdisk;
    int plausible, solvable;
    /* adiabatic contraction factors for the current solve */
    double ac_rfinal, ac_bterm, ac_fi, ac_fd;
} std_work;

/* ------------------------------------------------------------------------ helpers */
static double mass_to_fraction(double ab, double mass) {
    /* Abundances_Mass_To_Mass_Fraction, objects/abundances.F90:811-828 */
    if (ab > mass) return 1.0;
    if (ab <= 0.0) return 0.0;
    return ab / mass;
}
static double hydrogen_mass_fraction(double z) {
    /* objects/abundances.F90:830-850 */
    double x = z / ORC_METALLICITY_SOLAR * (ORC_HYDROGEN_BY_MASS_SOLAR - ORC_HYDROGEN_BY_MASS_PRIMORDIAL) +
               ORC_HYDROGEN_BY_MASS_PRIMORDIAL;
    x = fmax(x, 0.7);
    return fmin(x, ORC_HYDROGEN_BY_MASS_PRIMORDIAL);
}
static double helium_mass_fraction(double z) {
    return fmin(z / ORC_METALLICITY_SOLAR * (ORC_HELIUM_BY_MASS_SOLAR - ORC_HELIUM_BY_MASS_PRIMORDIAL) +
                    ORC_HELIUM_BY_MASS_PRIMORDIAL,
                ORC_HELIUM_BY_MASS_PRIMORDIAL);
}
static double hydrogen_number_fraction(double z) {
    double nh = hydrogen_mass_fraction(z) / ORC_ATOMIC_MASS_HYDROGEN;
    double nhe = helium_mass_fraction(z) / ORC_ATOMIC_MASS_HELIUM;
    return nh / (nh + nhe);
}

static int has(const std_work *w, int f) { return (w->flags & f) != 0; }

/* ------------------------------------------------------------ halo scales */
static void halo_scales(std_work *w) {
    /* dark_matter_halos/scales/virial_density_contrast.F90:195-417; the mean density table is
       interpolated linearly in ln t (:414, table1DLogarithmicLinear) */
    const orc_table2d *t = &w->T->t[GLC_TABLE_HALO_MEAN_DENSITY];
    double time, lnt, x, h;
    int i;
    if (w->halo_done) return;
    w->halo_done = 1;
    time = w->p[GLC_P_TIME_LAST_ISOLATED];
    if (!has(w, GLC_F_IS_SATELLITE) || time <= 0.0) time = w->time;
    lnt = dm_log(time);
    {
        const double lnt0 = dm_log(t->x0[0]), lnt1 = dm_log(t->x0[t->n0 - 1]);
        const double inv = (double)(t->n0 - 1) / (lnt1 - lnt0);
        x = (lnt - lnt0) * inv;
        i = (int)x;
        if (lnt < lnt0) i = 0;
        if (i > t->n0 - 2) i = t->n0 - 2;
        if (i < 0) i = 0;
        h = x - (double)i;
    }
    w->rho_mean = t->v[2 * i] * (1.0 - h) + t->v[2 * (i + 1)] * h;
    w->dlnrho_dt = t->v[2 * i + 1] * (1.0 - h) + t->v[2 * (i + 1) + 1] * h;
    {
        double mass = w->p[GLC_P_BASIC_MASS];
        w->rvir = dm_cbrt(3.0 * mass / 4.0 / ORC_PI / w->rho_mean);
        w->vvir = sqrt(ORC_G_INTERNAL * mass / w->rvir);
        w->tdyn = w->rvir / w->vvir * ORC_MPC_PER_KMS_TO_GYR;
        w->tvir = 0.5 * ORC_ATOMIC_MASS_UNIT * ORC_MEAN_ATOMIC_MASS_PRIMORDIAL *
                  ((ORC_KILO * w->vvir) * (ORC_KILO * w->vvir)) / ORC_BOLTZMANN;
    }
}

static double rvir_growth_rate(std_work *w, double dlnrho_dt) {
    /* virialDensityContrastDefinitionVirialRadiusGrowthRate :338-354; density growth 0 for satellites */
    double g = has(w, GLC_F_IS_SATELLITE) ? 0.0 : dlnrho_dt;
    return (1.0 / 3.0) * w->rvir * (w->p[GLC_P_MASS_RATE] / w->p[GLC_P_BASIC_MASS] - g);
}

/* ------------------------------------------------------------ hot halo beta profile */
static int beta_is_two_thirds(const std_work *w);
static double hh_outer_radius(std_work *w) {
    /* Node_Component_Hot_Halo_Standard_Outer_Radius, hot_halo/standard/_class.F90:430-450 */
    halo_scales(w);
    return fmax(fmin(w->p[GLC_P_HH_OUTER_RADIUS], w->rvir), w->P->hotHaloScaleRadiusRelative * w->rvir);
}

static void hh_profile(std_work *w) {
    /* hotHaloMassDistributionBetaProfile::get (hot_halo/mass_distribution/beta_profile.F90:140-215)
       + massDistributionBetaProfile initialize (mass_distributions/spherical/beta_profile.F90:190-301) */
    double r;
    if (w->hh_done) return;
    w->hh_done = 1;
    halo_scales(w);
    w->hh_router = has(w, GLC_F_HAS_HOTHALO) ? hh_outer_radius(w) : 0.0;
    w->hh_mass = has(w, GLC_F_HAS_HOTHALO) ? w->p[GLC_P_HH_MASS] : 0.0;
    w->hh_rcore = w->P->coreRadiusOverVirialRadius * w->rvir;
    w->hh_valid = !(w->hh_router <= 0.0 || w->hh_mass <= 0.0);
    if (!w->hh_valid) return;
    r = w->hh_router / w->hh_rcore;
    if (beta_is_two_thirds(w)) {
        double nf = (r < 1.0e-6) ? 3.0 / (r * r * r) + 9.0 / 5.0 / r - 36.0 * r / 175.0 : 1.0 / (r - dm_atan(r));
        w->hh_rho0 = w->hh_mass / 4.0 / ORC_PI / (w->hh_rcore * w->hh_rcore * w->hh_rcore) * nf;
    } else {
        /* :258: rho_0 = 3 M / (4 pi r_outer^3) / 2F1(3/2, 3 beta/2; 5/2; -r^2), with r^3/3 2F1(...) = I_2(r) (glc_specfun.h) */
        w->hh_rho0 = w->hh_mass / 4.0 / ORC_PI / (w->hh_rcore * w->hh_rcore * w->hh_rcore) / dm_beta_moment(2, r, w->P->hotHaloBeta);
    }
}
static int beta_is_two_thirds(const std_work *w) {
    /* betaIsTwoThirds = Values_Agree(beta, 2/3, relTol = 1e-3), beta_profile.F90:214 (numerical/comparison.F90: |a - b| <=
       relTol * (|a| + |b|) / 2) */
    const double b = w->P->hotHaloBeta, t = 2.0 / 3.0;
    return fabs(b - t) <= 1.0e-3 * 0.5 * (fabs(b) + fabs(t));
}
static double hh_density(std_work *w, double radius) {
    /* betaProfileDensity :303-320 (truncateAtOuterRadius) */
    double x;
    hh_profile(w);
    if (!w->hh_valid) return 0.0;
    if (radius > w->hh_router) return 0.0;
    x = radius / w->hh_rcore;
    return w->hh_rho0 / dm_pow(1.0 + x * x, 1.5 * w->P->hotHaloBeta);
}
static double hh_mass_enclosed(std_work *w, double radius) {
    /* betaProfileMassEnclosedBySphere :377-437, beta = 2/3 */
    double x;
    hh_profile(w);
    if (!w->hh_valid) return 0.0;
    if (radius > w->hh_router) radius = w->hh_router;
    x = radius / w->hh_rcore;
    if (!beta_is_two_thirds(w)) /* :425-436: 4 pi rho_0 r^3 / 3 * 2F1(3/2, 3 beta/2; 5/2; -x^2) = 4 pi rho_0 r_c^3 I_2(x) */
        return 4.0 * ORC_PI * w->hh_rho0 * dm_beta_moment(2, x, w->P->hotHaloBeta) * (w->hh_rcore * w->hh_rcore * w->hh_rcore);
    if (x < 1.0e-6)
        return 4.0 * ORC_PI * w->hh_rho0 * (w->hh_rcore * w->hh_rcore * w->hh_rcore) * (x * x * x) *
               (1.0 / 3.0 + x * x * (-1.0 / 5.0 + x * x * (1.0 / 7.0)));
    return 4.0 * ORC_PI * w->hh_rho0 * (x - dm_atan(x)) * (w->hh_rcore * w->hh_rcore * w->hh_rcore);
}
static double hh_radial_moment23(int m, double x) {
    /* radialMomentTwoThirds :667-728 */
    if (x <= 0.0) return 0.0;
    if (m == 2) return (x < 1.0e-6) ? x * x * x * (1.0 / 3.0 - x * x / 5.0) : x - dm_atan(x);
    return (x < 1.0e-6) ? x * x * x * x * (1.0 / 4.0 - x * x / 6.0) : 0.5 * (x * x - dm_log(1.0 + x * x));
}

/* ------------------------------------------------------------ cooling function (CIE tables) */
static int locate(const double *x, int n, double v) {
    /* interpolator%locate (gsl_interp_bsearch): index i (1-based) with x(i) <= v < x(i+1), clamped */
    int lo = 0, hi = n - 1;
    while (hi > lo + 1) {
        int mid = (hi + lo) / 2;
        if (x[mid] > v)
            hi = mid;
        else
            lo = mid;
    }
    return lo + 1;
}

typedef struct {
    int iT, iZ;
    double hT, hZ;
} cie_factors;

static void cie_interp_factors(const double *lnZ, const double *lnT, int nZ, int nT, int is_log,
                               int first_zero, double first_nonzero, double temperature,
                               double metallicity, cie_factors *f) {
    /* cieFileInterpolatingFactors, cooling/cooling_function/CIE_file.F90:665-715 */
    double tu = temperature, zu;
    int i;
    if (is_log) tu = dm_log(tu);
    i = locate(lnT, nT, tu);
    if (i > nT - 1) i = nT - 1;
    if (i < 1) i = 1;
    f->iT = i;
    f->hT = (tu - lnT[i - 1]) / (lnT[i] - lnT[i - 1]);
    zu = fmax(metallicity, 0.0);
    if (first_zero && zu < first_nonzero) {
        f->iZ = 1;
        f->hZ = zu / first_nonzero;
    } else {
        if (is_log) zu = dm_log(zu);
        i = locate(lnZ, nZ, zu);
        if (i > nZ - 1) i = nZ - 1;
        if (i < 1) i = 1;
        f->iZ = i;
        f->hZ = (zu - lnZ[i - 1]) / (lnZ[i] - lnZ[i - 1]);
    }
}
static double cie_interpolate(const double *v, int nT, int is_log, const cie_factors *f) {
    /* cieFileInterpolate :717-735; v[iZ][iT] */
    const double *a = v + (size_t)(f->iZ - 1) * nT + (f->iT - 1);
    const double *b = a + nT;
    double r = a[0] * (1.0 - f->hT) * (1.0 - f->hZ) + b[0] * (1.0 - f->hT) * f->hZ +
               a[1] * f->hT * (1.0 - f->hZ) + b[1] * f->hT * f->hZ;
    return is_log ? dm_exp(r) : r;
}

/* Lambda(T,Z)/n_H^2: cieFileCoolingFunction :238-317 with all extrapolation types "fix" */
static double cooling_function_over_nh2(const std_work *w, double temperature, double z_fraction,
                                        double *log_slope_t) {
    const orc_tables *T = w->T;
    const orc_table2d *t = &T->t[GLC_TABLE_COOLING_FUNCTION];
    double tu = temperature, zu = z_fraction / ORC_METALLICITY_SOLAR, lam;
    cie_factors f;
    int outside_t = 0;
    if (tu < t->x1[0]) {
        tu = t->x1[0];
        outside_t = 1;
    }
    if (tu > t->x1[t->n1 - 1]) {
        tu = t->x1[t->n1 - 1];
        outside_t = 1;
    }
    if (zu < t->x0[0]) zu = t->x0[0];
    if (zu > t->x0[t->n0 - 1]) zu = t->x0[t->n0 - 1];
    cie_interp_factors(T->cooling_lnZ, T->cooling_lnT, t->n0, t->n1, T->cooling_log, T->cooling_first_z_zero,
                       T->cooling_first_nonzero_z, tu, zu, &f);
    lam = cie_interpolate(T->cooling_lnL, t->n1, T->cooling_log, &f);
    if (log_slope_t) {
        /* cieFileCoolingFunctionTemperatureLogSlope :410-512 */
        if (outside_t)
            *log_slope_t = 0.0;
        else {
            const double *a = T->cooling_lnL + (size_t)(f.iZ - 1) * t->n1 + (f.iT - 1);
            const double *b = a + t->n1;
            double s = ((a[1] - a[0]) * (1.0 - f.hZ) + (b[1] - b[0]) * f.hZ) /
                       (T->cooling_lnT[f.iT] - T->cooling_lnT[f.iT - 1]);
            if (!T->cooling_log) s = s * temperature / lam;
            *log_slope_t = s;
        }
    }
    return lam;
}
static double electron_fraction(const std_work *w, double temperature, double z_fraction) {
    /* cieFileElectronDensity (chemical/state/CIE_file.F90:258-320) / n_H, extrapolation "fix" */
    const orc_tables *T = w->T;
    const orc_table2d *t = &T->t[GLC_TABLE_ELECTRON_FRACTION];
    double tu = temperature, zu = z_fraction / ORC_METALLICITY_SOLAR;
    cie_factors f;
    if (tu < t->x1[0]) tu = t->x1[0];
    if (tu > t->x1[t->n1 - 1]) tu = t->x1[t->n1 - 1];
    if (zu < t->x0[0]) zu = t->x0[0];
    if (zu > t->x0[t->n0 - 1]) zu = t->x0[t->n0 - 1];
    cie_interp_factors(T->electron_lnZ, T->electron_lnT, t->n0, t->n1, T->electron_log,
                       T->electron_first_z_zero, T->electron_first_nonzero_z, tu, zu, &f);
    return cie_interpolate(T->electron_lnV, t->n1, T->electron_log, &f);
}

static double cooling_time(const std_work *w, double temperature, double density, double z_fraction) {
    /* coolingTimeSimple::time, cooling/cooling_time/simple.F90:128-179 */
    const double time_large = 1.0e10;
    double nh = density * hydrogen_mass_fraction(z_fraction) * ORC_MASS_SOLAR / ORC_MASS_HYDROGEN_ATOM /
                (ORC_HECTO * ORC_HECTO * ORC_HECTO) / (ORC_MEGAPARSEC * ORC_MEGAPARSEC * ORC_MEGAPARSEC);
    double nall = nh / hydrogen_number_fraction(z_fraction) + electron_fraction(w, temperature, z_fraction) * nh;
    double cf = cooling_function_over_nh2(w, temperature, z_fraction, NULL) * nh * nh;
    if (cf > 0.0) {
        double e = w->P->coolingDegreesOfFreedom / 2.0 * ORC_BOLTZMANN * temperature * nall / ORC_ERGS;
        return e / cf / ORC_GIGAYEAR;
    }
    return time_large;
}

static double cooling_radius_root(double radius, void *vw) {
    /* coolingRadiusRoot, cooling/cooling_radius/simple.F90:389-427 */
    std_work *w = (std_work *)vw;
    double density = hh_density(w, radius);
    return cooling_time(w, w->tvir, density, w->cool_z) - w->cool_tavail;
}

static double cooling_radius(std_work *w) {
    /* coolingRadiusSimple::radius :313-387; finder tolerance :133,174-178 */
    orc_root_finder rf;
    double router, root_outer, root_zero;
    int st;
    if (w->rcool_done) return w->rcool;
    w->rcool_done = 1;
    halo_scales(w);
    w->cool_tavail = w->tdyn; /* whiteFrenk1991TimeAvailable with ageFactor = 0, time_available/White-Frenk.F90:144-146 */
    w->cool_z = mass_to_fraction(w->p[GLC_P_HH_ABUND], w->p[GLC_P_HH_MASS]);
    router = hh_outer_radius(w);
    root_outer = cooling_radius_root(router, w);
    if (root_outer < 0.0) return w->rcool = router;
    root_zero = cooling_radius_root(0.0, w);
    if (root_zero > 0.0) return w->rcool = 0.0;
    orc_root_init(&rf, cooling_radius_root, w, 0.0, 1.0e-6);
    w->rcool = orc_root_find(&rf, 0.0, router, 1, root_zero, root_outer, &st);
    if (st != 0) w->c->nonfinite = 1;
    return w->rcool;
}

static double cooling_radius_growth_rate(std_work *w) {
    /* coolingRadiusSimple::radiusGrowthRate :229-311 with virial (isothermal) temperature profile:
       temperatureLogSlope = 0; coolingTime density slope = 1 - 2, temperature slope = -dlnLambda/dlnT */
    double router, rc, x, density_log_slope, slope, ls_t;
    if (w->rcool_rate_done) return w->rcool_rate;
    w->rcool_rate_done = 1;
    router = hh_outer_radius(w);
    rc = cooling_radius(w);
    if (rc >= router) return w->rcool_rate = 0.0;
    hh_profile(w);
    x = rc / w->hh_rcore;
    density_log_slope = -3.0 * w->P->hotHaloBeta * x * x / (x * x + 1.0); /* betaProfileDensityGradientRadial :346-351 */
    (void)cooling_function_over_nh2(w, w->tvir, w->cool_z, &ls_t);
    if (rc > 0.0) {
        slope = density_log_slope * (1.0 - 2.0) + 0.0 * (-ls_t);
        if (slope != 0.0)
            w->rcool_rate = rc / w->cool_tavail * 1.0 / slope;
        else
            w->rcool_rate = 0.0;
    } else
        w->rcool_rate = 0.0;
    return w->rcool_rate;
}

static double cooling_rate(std_work *w) {
    /* coolingRateWhiteFrenk1991::rate, cooling/cooling_rate/White-Frenk.F90:131-185 */
    double router, rinfall;
    halo_scales(w);
    if (w->vvir > w->P->coolingVelocityCutOff) return 0.0;
    router = hh_outer_radius(w);
    rinfall = cooling_radius(w); /* coolingInfallRadiusCoolingRadius */
    if (rinfall >= router) return w->p[GLC_P_HH_MASS] / w->tdyn;
    return 4.0 * ORC_PI * rinfall * rinfall * hh_density(w, rinfall) * cooling_radius_growth_rate(w);
}

static double cooling_specific_angular_momentum(std_work *w, double radius) {
    /* coolingSpecificAngularMomentumConstantRotation (hotGas, hotGas),
       cooling/specific_angular_momentum/constant_rotation.F90:198-286 */
    double jmean, x, norm;
    if (!(radius > 0.0)) return 0.0;
    jmean = w->p[GLC_P_HH_ANGMOM] / w->p[GLC_P_HH_MASS];
    hh_profile(w);
    x = w->hh_router / w->hh_rcore;
    /* densityRadialMoment(m) = I_m(x) rho0 rc^(1+m): ratio m=2 / m=3 */
    {
        const double rc = w->hh_rcore;
        /* general beta: the same moments through 2F1 (:600-612) = I_m(x) of glc_specfun.h */
        const double m2 = beta_is_two_thirds(w) ? hh_radial_moment23(2, x) : dm_beta_moment(2, x, w->P->hotHaloBeta);
        const double m3 = beta_is_two_thirds(w) ? hh_radial_moment23(3, x) : dm_beta_moment(3, x, w->P->hotHaloBeta);
        norm = (m2 * w->hh_rho0 * (rc * rc * rc)) / (m3 * w->hh_rho0 * (rc * rc * rc * rc));
    }
    return norm * jmean * radius;
}

/* ------------------------------------------------------------ galactic structure */
static double disk_mass(const std_work *w) {
    return fmax(0.0, w->p[GLC_P_DISK_MASS_STELLAR]) + fmax(0.0, w->p[GLC_P_DISK_MASS_GAS]);
}
static double sph_mass(const std_work *w) {
    return fmax(0.0, w->p[GLC_P_SPH_MASS_STELLAR]) + fmax(0.0, w->p[GLC_P_SPH_MASS_GAS]);
}

static double disk_bessel_factor(const std_work *w, double half_radius) {
    /* exponentialDiskBesselFactorRotationCurve, mass_distributions/cylindrical/exponential_disk.F90:675-733:
       table1DLogarithmicLinear lookup of x^2 [I0 K0 - I1 K1] */
    const orc_table2d *t = &w->T->t[GLC_TABLE_DISK_ROTATION_CURVE];
    const double ln2 = 0.69314718055994530942, euler = 0.57721566490153286061;
    double lx, lx0, lx1, inv, x, h;
    int i;
    if (half_radius <= 0.0) return 0.0;
    if (half_radius < 1.0e-3) return (ln2 - euler - 0.5 - dm_log(half_radius)) * half_radius * half_radius;
    lx = dm_log(half_radius);
    lx0 = dm_log(t->x0[0]);
    lx1 = dm_log(t->x0[t->n0 - 1]);
    inv = (double)(t->n0 - 1) / (lx1 - lx0);
    x = (lx - lx0) * inv;
    i = (int)x;
    if (i > t->n0 - 2) i = t->n0 - 2;
    if (i < 0) i = 0;
    h = x - (double)i;
    return t->v[i] * (1.0 - h) + t->v[i + 1] * h;
}

/* V^2 of all baryons (massType=massTypeBaryonic: gas + stars of disk, spheroid, hot halo) */
static double baryonic_vc2(std_work *w, double radius) {
    double v2 = 0.0;
    if (!(radius > 0.0)) return 0.0;
    if (has(w, GLC_F_HAS_DISK) && w->p[GLC_P_DISK_RADIUS] > 0.0) {
        /* exponentialDiskRotationCurve :518-556 through the scaler (dimensionless profile) */
        double rd = w->p[GLC_P_DISK_RADIUS], m = disk_mass(w), r = radius / rd;
        if (r > 30.0)
            v2 += ORC_G_INTERNAL * m / radius;
        else
            v2 += ORC_G_INTERNAL * 2.0 * (m / rd) * disk_bessel_factor(w, 0.5 * r);
    }
    if (has(w, GLC_F_HAS_SPHEROID) && w->p[GLC_P_SPH_RADIUS] > 0.0) {
        /* Hernquist, mass_distributions/spherical/Hernquist.F90: M(<r) = M r^2/(r+a)^2 */
        double a = w->p[GLC_P_SPH_RADIUS], m = sph_mass(w);
        v2 += ORC_G_INTERNAL * m * radius / ((radius + a) * (radius + a));
    }
    if (has(w, GLC_F_HAS_HOTHALO)) {
        double m = hh_mass_enclosed(w, radius);
        v2 += ORC_G_INTERNAL * m / radius;
    }
    return v2;
}

static double nfw_mass_scale_free(double x) {
    /* massEnclosedScaleFree, mass_distributions/spherical/NFW.F90:550-571 (without the 4 pi) */
    if (x == 1.0) return dm_log(2.0) - 0.5;
    if (x >= 1.0e-6) return dm_log(1.0 + x) - x / (1.0 + x);
    return x * x * (0.5 + x * (-2.0 / 3.0 + x * (0.75 + x * (-0.8))));
}
static int dmo_isothermal(const std_work *w) { return w->P->darkMatterProfileDMO == GLC_DMO_ISOTHERMAL; }
static double nfw_mass_enclosed(std_work *w, double radius) {
    /* darkMatterProfileDMO%get(node)%massEnclosedBySphere.
       NFW: nfwMassEnclosedBySphere :444-464 with normalisation :254-255: M(<r) = [M_vir / m(c)] m(r/r_s) (the
       4 pi rho_0 r_s^3 of :254-255,458-460 folded into one factor);
       isothermal (mass_distributions/spherical/isothermal.F90:150-205,287-302; mass = M_vir, lengthReference = r_vir):
       M(<r) = 4 pi rho_n L^2 r */
    double rs = w->p[GLC_P_DMSCALE]; /* current scale radius */
    double conc, norm;
    halo_scales(w);
    if (dmo_isothermal(w)) {
        const double L = w->rvir;
        const double rho_n = w->p[GLC_P_BASIC_MASS] / 4.0 / ORC_PI / (L * L * L);
        return (4.0 * ORC_PI * rho_n * (L * L)) * radius;
    }
    conc = w->rvir / rs;
    norm = w->p[GLC_P_BASIC_MASS] / (dm_log(1.0 + conc) - conc / (1.0 + conc));
    return norm * nfw_mass_scale_free(radius / rs);
}
static double dmo_rotation_curve(std_work *w, double radius) {
    /* massDistribution_%rotationCurve: isothermal.F90:196-204,431-442 (velocityRotation); NFW: sqrt(G M(<r)/r) */
    if (dmo_isothermal(w)) {
        const double L = (halo_scales(w), w->rvir);
        const double rho_n = w->p[GLC_P_BASIC_MASS] / 4.0 / ORC_PI / (L * L * L);
        return sqrt(ORC_G_INTERNAL * (4.0 * ORC_PI * rho_n * (L * L)));
    }
    return (radius > 0.0) ? sqrt(ORC_G_INTERNAL * nfw_mass_enclosed(w, radius) / radius) : 0.0;
}

static double ac_orbital_mean(std_work *w, double radius) {
    /* sphericalAdiabaticGnedin2004RadiusOrbitalMean :664-687 (radiusFractionalPivot = 1) */
    return w->P->adiabaticA * w->rvir *
           orc_fast_exponentiate(1.0e-3, 1.0, w->P->adiabaticOmega, 1.0e4, radius / w->rvir);
}
static double ac_solver(double radius_initial, void *vw) {
    /* sphericalAdiabaticGnedin2004Solver :707-727 */
    std_work *w = (std_work *)vw;
    double m = nfw_mass_enclosed(w, ac_orbital_mean(w, radius_initial));
    return m * (w->ac_fi * radius_initial - w->ac_fd * w->ac_rfinal) - w->ac_bterm;
}
static double baryonic_mass_self(const std_work *w) {
    /* node%massBaryonic(): disk + spheroid + hot halo (mass + outflowed) bound functions */
    double m = 0.0;
    if (has(w, GLC_F_HAS_DISK)) m += disk_mass(w);
    if (has(w, GLC_F_HAS_SPHEROID)) m += sph_mass(w);
    if (has(w, GLC_F_HAS_HOTHALO))
        m += fmax(0.0, w->p[GLC_P_HH_MASS]) + fmax(0.0, w->p[GLC_P_HH_OUTFLOWED_MASS]);
    return m;
}
static double dark_matter_mass_enclosed(std_work *w, double radius) {
    /* massDistribution(componentTypeDarkHalo,massTypeDark)%massEnclosedBySphere:
       adiabaticGnedin2004 decorator (mass_distributions/spherical/adiabatic_Gnedin2004.F90:410-530,
       dark_matter_profiles/adiabatic_Gnedin2004.F90:302-364) over NFW */
    const double f_dm = 1.0 - w->P->OmegaBaryon / w->P->OmegaMatter;
    double r_init;
    halo_scales(w);
    if (!w->P->adiabaticContraction) return nfw_mass_enclosed(w, radius);
    if (radius <= 0.0) return 0.0;
    if (radius >= w->rvir)
        r_init = radius;
    else {
        double m_self = fmax(baryonic_mass_self(w), 0.0);
        double m_tot = fmax(baryonic_mass_self(w) + w->p[GLC_P_MASS_BARYONIC_SUBHALOS], 0.0);
        double rmean, menc, rup;
        w->ac_fd = fmin(f_dm + (m_tot - m_self) / w->p[GLC_P_BASIC_MASS], 1.0);
        w->ac_fi = fmin(f_dm + m_tot / w->p[GLC_P_BASIC_MASS], 1.0);
        w->ac_rfinal = radius;
        rmean = ac_orbital_mean(w, radius);
        w->ac_bterm = baryonic_vc2(w, rmean) * rmean * radius / ORC_G_INTERNAL; /* computeFactors :633-662 */
        if (ac_solver(w->rvir, w) < 0.0)
            r_init = w->rvir;
        else {
            orc_root_finder rf;
            int st;
            menc = nfw_mass_enclosed(w, rmean);
            if (menc > 0.0) {
                rup = (w->ac_bterm / menc + w->ac_fd * radius) / w->ac_fi;
                if (rup < radius) rup = radius;
            } else
                rup = radius;
            orc_root_init(&rf, ac_solver, w, 0.0, 1.0e-2);
            rf.expand_type = ORC_EXPAND_MULTIPLICATIVE;
            rf.expand_upward = 1.1;
            rf.expand_downward = 0.9;
            rf.sign_expect_upward = ORC_SIGN_POSITIVE;
            rf.sign_expect_downward = ORC_SIGN_NEGATIVE;
            r_init = orc_root_find(&rf, radius, rup, 0, 0, 0, &st);
            if (st != 0) w->c->nonfinite = 1;
        }
    }
    return f_dm * nfw_mass_enclosed(w, r_init);
}

/* inverse tabulation of the scale-free NFW specific angular momentum on the octave lattice x_k = 2^(k/30)
   (NFW.F90:94,589-639; numerical/tabulations_inverse.F90:143-245; numerical/ranges.F90 Lattice_Value), tabulated once over
   2^-40 .. 2^40: the lattice is absolute, so these are the points the reference computes, whatever extent it has grown to */
#define NFWJ_PER_OCTAVE 30
#define NFWJ_OCTAVES 40
#define NFWJ_N (2 * NFWJ_OCTAVES * NFWJ_PER_OCTAVE + 1)
static double nfwj_x[NFWJ_N], nfwj_v[NFWJ_N];
static int nfwj_built = 0;
static void nfwj_build(void) {
    int i;
#pragma omp critical(orc_nfwj_build)
    if (!nfwj_built) {
        for (i = 0; i < NFWJ_N; i++) {
            const int k = i - NFWJ_OCTAVES * NFWJ_PER_OCTAVE;
            const double x = (k % NFWJ_PER_OCTAVE == 0) ? dm_scale2(1.0, k / NFWJ_PER_OCTAVE)
                                                        : dm_exp(((double)k / (double)NFWJ_PER_OCTAVE) * 0.69314718055994530942);
            nfwj_x[i] = x;
            nfwj_v[i] = sqrt(4.0 * ORC_PI * nfw_mass_scale_free(x) * x);
        }
        nfwj_built = 1;
    }
}
static double nfw_radius_from_j(std_work *w, double j) {
    /* massDistribution_%radiusFromSpecificAngularMomentum: nfwRadiusFromSpecificAngularMomentum (NFW.F90:589-625),
       isothermalRadiusFromSpecificAngularMomentum (isothermal.F90:350-370) */
    double rs, conc, rho_n, jsf, x;
    int lo, hi;
    if (!(j > 0.0)) return 0.0;
    halo_scales(w);
    if (dmo_isothermal(w)) {
        const double L = w->rvir;
        const double rn = w->p[GLC_P_BASIC_MASS] / 4.0 / ORC_PI / (L * L * L);
        return j / sqrt(4.0 * ORC_PI * rn * (L * L)) / sqrt(ORC_G_INTERNAL);
    }
    if (!nfwj_built) nfwj_build();
    rs = w->p[GLC_P_DMSCALE];
    conc = w->rvir / rs;
    rho_n = w->p[GLC_P_BASIC_MASS] / 4.0 / ORC_PI / (rs * rs * rs) / (dm_log(1.0 + conc) - conc / (1.0 + conc));
    jsf = j / sqrt(ORC_G_INTERNAL * rho_n) / (rs * rs);
    lo = 0;
    hi = NFWJ_N - 1;
    while (hi > lo + 1) {
        const int mid = (hi + lo) >> 1;
        if (nfwj_v[mid] > jsf)
            hi = mid;
        else
            lo = mid;
    }
    /* gsl_interp_linear: y_lo + (x - x_lo) / dx * dy */
    x = nfwj_x[lo] + (jsf - nfwj_v[lo]) / (nfwj_v[lo + 1] - nfwj_v[lo]) * (nfwj_x[lo + 1] - nfwj_x[lo]);
    return x * rs;
}

static void plausibility(std_work *w) {
    /* Node_Component_Basic_Standard_Plausibility (basic/standard/_class.F90:105-123),
       disk (disk/standard/_class.F90:999-1049) and spheroid (spheroid/standard/_class.F90:1249-1296) */
    w->plausible = 1;
    w->solvable = 1;
    if (w->p[GLC_P_BASIC_MASS] <= 0.0 || w->time <= 0.0) {
        w->plausible = 0;
        w->solvable = 0;
        return;
    }
    halo_scales(w);
    if (has(w, GLC_F_HAS_DISK)) {
        double m = w->p[GLC_P_DISK_MASS_STELLAR] + w->p[GLC_P_DISK_MASS_GAS], j = w->p[GLC_P_DISK_ANGMOM];
        if (m >= 0.0 && j > 0.0) {
            double s = m * w->rvir * w->vvir;
            if (j > 1.0e1 * s || j < 1.0e-6 * s) w->plausible = 0;
        }
    }
    if (w->plausible && has(w, GLC_F_HAS_SPHEROID)) {
        double m = w->p[GLC_P_SPH_MASS_STELLAR] + w->p[GLC_P_SPH_MASS_GAS], j = w->p[GLC_P_SPH_ANGMOM];
        if (m >= 0.0 && j > 0.0) {
            double s = m * w->rvir * w->vvir;
            if (j > 1.0e1 * s || j < 1.0e-6 * s) w->plausible = 0;
        }
    }
}

static double component_j(const std_work *w, int comp) {
    /* disk: Node_Component_Disk_Standard_Radius_Solver (disk/standard/_class.F90:1112-1177),
       ratioAngularMomentumSolverRadius = 1/(I2/I1) = 1/2 for the exponential disk (:345-356);
       spheroid: spheroid/standard/_class.F90:1359-1407 */
    double j, m;
    if (comp == 0) {
        j = w->p[GLC_P_DISK_ANGMOM];
        m = w->p[GLC_P_DISK_MASS_GAS] + w->p[GLC_P_DISK_MASS_STELLAR];
        if (!(j >= 0.0)) return 0.0;
        return ((m > 0.0) ? j / m : 0.0) * 0.5;
    }
    j = w->p[GLC_P_SPH_ANGMOM];
    m = w->p[GLC_P_SPH_MASS_GAS] + w->p[GLC_P_SPH_MASS_STELLAR];
    if (!(j >= 0.0)) return 0.0;
    return w->P->spheroidRatioAngularMomentumScaleRadius * ((m > 0.0) ? j / m : 0.0);
}

static void structure_solve(std_work *w) {
    /* galacticStructureSolverEquilibrium::solve, galactic_structure/radius_solver/equilibrium.F90:243-506 */
    const int iteration_maximum = 100;
    /* radiusHistory is a saved, thread-private array in the reference (:329) and is NOT reset between solves (the reset at
       :455 sits in the branch that only runs for countIterations >= 2).  That persistence is unobservable: an entry is read
       only when countIterations > 10 (:456-459), and by then iterations 2..10 of THIS solve have overwritten both entries
       of every active component (:477-478) -- so starting each solve from -1 gives the reference's results exactly */
    double history[2][2] = {{-1.0, -1.0}, {-1.0, -1.0}};
    double fit;
    int count = 0, comp;
    plausibility(w);
    if (!w->plausible) return;
    fit = 2.0 * w->P->structureSolutionTolerance;
    while (count <= 1 || (fit > w->P->structureSolutionTolerance && count < iteration_maximum)) {
        int active = 0;
        count++;
        if (count > 1) fit = 0.0;
        for (comp = 0; comp < 2; comp++) {
            const int pr = comp == 0 ? GLC_P_DISK_RADIUS : GLC_P_SPH_RADIUS;
            const int pv = comp == 0 ? GLC_P_DISK_VELOCITY : GLC_P_SPH_VELOCITY;
            double j, radius, velocity;
            if (!has(w, comp == 0 ? GLC_F_HAS_DISK : GLC_F_HAS_SPHEROID)) continue;
            j = component_j(w, comp);
            active++;
            if (count == 1) {
                radius = w->p[pr];
                if (radius <= 0.0) {
                    /* :358-376: guess from the dark-matter-only profile */
                    double jmax = dmo_rotation_curve(w, 1.0e10) * 1.0e10;
                    if (jmax < j)
                        radius = w->rvir;
                    else
                        radius = nfw_radius_from_j(w, j);
                    velocity = dmo_rotation_curve(w, radius);
                } else
                    velocity = w->p[pv];
                if (w->P->structureVelocityMaximumFactor > 0.0) /* :376,380 */
                    velocity = fmin(velocity, w->P->structureVelocityMaximumFactor * w->vvir);
            } else {
                double mdm, vdm2, vb2, radius_new;
                if (j <= 0.0) continue;
                radius = w->p[pr];
                mdm = dark_matter_mass_enclosed(w, radius);
                vdm2 = ORC_G_INTERNAL * mdm / radius;
                vb2 = w->P->includeBaryonGravity ? baryonic_vc2(w, radius) : 0.0;
                velocity = sqrt(vdm2 + vb2);
                if (w->P->structureVelocityMaximumFactor > 0.0) /* :429-433 */
                    velocity = fmin(velocity, w->P->structureVelocityMaximumFactor * w->vvir);
                radius_new = (radius > 0.0) ? sqrt(j / velocity * radius) : j / velocity;
                if (count > 10 && history[comp][0] >= 0.0 && history[comp][1] >= 0.0 &&
                    (history[comp][1] - history[comp][0]) * (history[comp][0] - radius) < 0.0) {
                    switch (count % 4) {
                    case 0: radius = sqrt(radius * history[comp][0]); break;
                    case 1: radius = 0.5 * (radius + history[comp][0]); break;
                    case 2: radius = sqrt(history[comp][0] * history[comp][1]); break;
                    default: radius = 0.5 * (history[comp][0] + history[comp][1]); break;
                    }
                    history[comp][0] = history[comp][1] = -1.0;
                }
                history[comp][1] = history[comp][0];
                history[comp][0] = radius;
                if (radius > 0.0 && radius_new > 0.0) fit += fabs(dm_log(radius_new / radius));
                radius = radius_new;
                if (!(radius > 0.0)) w->c->nonfinite = 1;
            }
            w->p[pr] = fmax(radius, 0.0); /* Radius_Solve_Set :1065-1078 */
            w->p[pv] = velocity;
        }
        if (active == 0) {
            fit = 0.0;
            break;
        }
        fit /= (double)active;
    }
}

/* ------------------------------------------------------------ star formation in disks */
static double kmt_fh2_fast(double s, void *u) {
    /* krumholz2009MolecularFractionFast :462-476 */
    (void)u;
    return (s < 2.0) ? 1.0 - 0.75 * s / (1.0 + 0.25 * s) : 0.0;
}
static double disk_sigma_gas(const std_work *w, double radius) {
    /* exponentialDiskSurfaceDensity :484-499 scaled: M/(2 pi Rd^2) dm_exp(-R/Rd) */
    double rd = w->k_rdisk;
    return fmax(0.0, w->k_mgas) / (2.0 * ORC_PI * rd * rd) * dm_exp(-radius / rd);
}
static void kmt_factors(std_work *w) {
    /* krumholz2009ComputeFactors :312-358 */
    double z;
    w->k_mgas = w->p[GLC_P_DISK_MASS_GAS];
    w->k_rdisk = w->p[GLC_P_DISK_RADIUS];
    z = mass_to_fraction(w->p[GLC_P_DISK_ABUND_GAS], w->k_mgas);
    w->k_xh = hydrogen_mass_fraction(z);
    w->k_zsolar = z / ORC_METALLICITY_SOLAR;
    w->k_sigma_norm = 0.0;
    w->k_sigma_trunc = 0.0;
    if (w->k_zsolar > 0.0) {
        w->k_chi = 0.77 * (1.0 + 3.1 * dm_pow(w->k_zsolar, 0.365));
        w->k_sigma_norm = w->k_xh * w->P->clumpingFactorMolecularComplex / (ORC_MEGA * ORC_MEGA);
        w->k_s_norm = dm_log(1.0 + 0.6 * w->k_chi + 0.01 * w->k_chi * w->k_chi) / (0.04 * w->k_zsolar);
        if (w->k_sigma_norm > 0.0)
            w->k_sigma_trunc = w->k_s_norm / w->k_sigma_norm / w->P->krumholzSTruncation;
        else
            w->k_sigma_trunc = DBL_MAX;
    }
}
static int kmt_degenerate(const std_work *w) {
    return w->k_mgas <= 0.0 || w->k_rdisk <= 0.0 || w->k_zsolar <= 0.0 || w->k_sigma_norm <= 0.0;
}
static double kmt_rate(std_work *w, double radius) {
    /* krumholz2009Rate :360-414 */
    double sg, sgd, s, fh2, factor;
    if (kmt_degenerate(w)) return 0.0;
    sg = disk_sigma_gas(w, radius);
    sgd = w->k_xh * sg / 85.0e12;
    if (sg <= 1.0e-100) return 0.0;
    s = w->k_s_norm / (w->k_sigma_norm * sg);
    if (s > 10.0)
        fh2 = kmt_fh2_fast(s, NULL);
    else
        fh2 = orc_linear_table_eval(kmt_fh2_fast, NULL, 0.0, 10.0, 1000, s, 1);
    if (sgd <= 0.0)
        factor = 0.0;
    else if (sgd < 1.0)
        factor = orc_fast_exponentiate(1.0, 1000.0, 0.33, 100.0, 1.0 / sgd);
    else
        factor = orc_fast_exponentiate(1.0, 1000.0, 0.33, 100.0, sgd);
    return w->P->frequencyStarFormation * sg * factor * fh2;
}
static double kmt_integrand(double radius, void *vw) {
    return radius * kmt_rate((std_work *)vw, radius); /* intgrtdSurfaceDensityIntegrand :192-202 */
}
static double kmt_molecular_root(double radius, void *vw) {
    std_work *w = (std_work *)vw;
    return disk_sigma_gas(w, radius) - w->k_sigma_trunc;
}
static double kmt_critical_root(double radius, void *vw) {
    std_work *w = (std_work *)vw;
    return w->k_xh * disk_sigma_gas(w, radius) / 85.0e12 - 1.0;
}
static void kmt_finder(orc_root_finder *rf, orc_fn1 f, std_work *w) {
    /* finderCritical / finderMolecules :245-264 */
    orc_root_init(rf, f, w, 0.0, 1.0e-4);
    rf->expand_type = ORC_EXPAND_MULTIPLICATIVE;
    rf->expand_upward = 2.0;
    rf->expand_downward = 0.5;
    rf->sign_expect_upward = ORC_SIGN_NEGATIVE;
    rf->sign_expect_downward = ORC_SIGN_POSITIVE;
}

static double sfr_disk(std_work *w) {
    /* starFormationRateDisksIntgrtdSurfaceDensity::rate :131-190 with krumholz2009Intervals :478-587 */
    double r_in, r_out, r_max, sgd_in, sgd, sg, total = 0.0, res, err;
    double iv[2][2];
    int n_iv = 0, i, st, ni;
    if (w->sfr_done) return w->sfr_disk;
    w->sfr_done = 1;
    w->sfr_disk = 0.0;
    if (w->p[GLC_P_DISK_MASS_GAS] <= 0.0 || w->p[GLC_P_DISK_RADIUS] <= 0.0) return 0.0;
    kmt_factors(w);
    if (kmt_degenerate(w)) return 0.0;
    r_in = 0.0;
    r_out = 10.0 * w->k_rdisk;
    sg = disk_sigma_gas(w, r_in);
    sgd_in = w->k_xh * sg / 85.0e12;
    if (sg <= w->k_sigma_trunc) return 0.0;
    sg = disk_sigma_gas(w, r_out);
    sgd = w->k_xh * sg / 85.0e12;
    if (sg <= w->k_sigma_trunc) {
        orc_root_finder rf;
        kmt_finder(&rf, kmt_molecular_root, w);
        r_max = orc_root_find(&rf, r_in, r_out, 0, 0, 0, &st);
        if (st != 0) w->c->nonfinite = 1;
        sgd = w->k_xh * disk_sigma_gas(w, r_max) / 85.0e12;
    } else
        r_max = r_out;
    if (sgd_in <= 1.0 || sgd >= 1.0) {
        iv[0][0] = r_in;
        iv[0][1] = r_max;
        n_iv = 1;
    } else {
        orc_root_finder rf;
        double r_crit;
        kmt_finder(&rf, kmt_critical_root, w);
        r_crit = orc_root_find(&rf, r_in, r_max, 0, 0, 0, &st);
        if (st != 0) w->c->nonfinite = 1;
        iv[0][0] = r_in;
        iv[0][1] = r_crit;
        iv[1][0] = r_crit;
        iv[1][1] = r_max;
        n_iv = 2;
    }
    for (i = 0; i < n_iv; i++) {
        st = orc_qag15(kmt_integrand, w, iv[i][0], iv[i][1], 1.0e-12, w->P->sfrIntegrationTolerance, 1000, &res,
                       &err, &ni);
        total += res;
    }
    w->sfr_disk = 2.0 * ORC_PI * total;
    return w->sfr_disk;
}

static double sfr_spheroid(const std_work *w) {
    /* starFormationRateSpheroidsTimescale (rates/spheroids/timescale.F90:109-130) with
       starFormationTimescaleDynamicalTime (timescales/dynamical_time.F90:121-189) */
    double v = w->p[GLC_P_SPH_VELOCITY], r = w->p[GLC_P_SPH_RADIUS], tau;
    if (v <= 0.0 || w->P->sfSpheroidEfficiency == 0.0) return 0.0;
    tau = fmax(ORC_MPC_PER_KMS_TO_GYR * r / v * dm_pow(v / 200.0, w->P->sfSpheroidExponentVelocity) /
                   w->P->sfSpheroidEfficiency,
               w->P->sfSpheroidTimescaleMinimum);
    return (tau > 0.0) ? w->p[GLC_P_SPH_MASS_GAS] / tau : 0.0;
}

/* ------------------------------------------------------------ black holes */
typedef struct bh_state {
    int on;          /* a black hole of positive mass exists */
    double mass, spin;
    double e_isco, l_isco; /* ISCO specific energy / angular momentum, gravitational units, prograde */
    double eddington;
    double acc_sph, acc_hot, acc; /* blackHoleAccretionRateStandard::rateAccretion */
} bh_state;

static double ideal_gas_sound_speed(double temperature) {
    /* Ideal_Gas_Sound_Speed, thermodynamics/ideal_gases.F90:46-69 (primordial mean atomic mass) */
    return sqrt(5.0 * ORC_BOLTZMANN * temperature / 3.0 / ORC_MEAN_ATOMIC_MASS_PRIMORDIAL / ORC_ATOMIC_MASS_UNIT) / ORC_KILO;
}
static double bhl_radius(double mass, double temperature) {
    /* Bondi_Hoyle_Lyttleton_Accretion_Radius, accretion/Bondi_Hoyle_Lyttleton.F90:61-78 */
    double cs;
    if (!(temperature > 0.0)) return DBL_MAX;
    cs = ideal_gas_sound_speed(temperature);
    return ORC_G_INTERNAL * mass / (cs * cs);
}
static double bhl_rate(double mass, double density, double velocity, double temperature, int with_radius, double radius) {
    /* Bondi_Hoyle_Lyttleton_Accretion_Rate, accretion/Bondi_Hoyle_Lyttleton.F90:34-59 */
    const double cs = ideal_gas_sound_speed(temperature);
    const double gm = ORC_G_INTERNAL * mass;
    if (with_radius)
        return (ORC_KILO * ORC_GIGAYEAR / ORC_MEGAPARSEC) * 4.0 * ORC_PI * (radius * radius) * density *
               sqrt(cs * cs + velocity * velocity);
    return (ORC_KILO * ORC_GIGAYEAR / ORC_MEGAPARSEC) * 4.0 * ORC_PI * (gm * gm) * density /
           dm_pow(cs * cs + velocity * velocity, 1.5);
}
static double bh_isco_radius(double j) {
    /* Black_Hole_ISCO_Radius_Spin (prograde), black_holes/fundamentals.F90:78-121; A1, A2 :553-573 */
    const double third = 1.0 / 3.0;
    double a1 = 1.0 + dm_pow(1.0 - j * j, third) * (dm_pow(1.0 + j, third) + dm_pow(1.0 - j, third));
    double a2 = sqrt(3.0 * (j * j) + a1 * a1);
    return 3.0 + a2 - sqrt((3.0 - a1) * (3.0 + a1 + 2.0 * a2));
}
static double bh_isco_energy(double j, double r) {
    /* Black_Hole_ISCO_Specific_Energy_Spin, black_holes/fundamentals.F90:207-233 */
    if (j >= 0.99999) return 0.5773502693 + 0.9164864242 * dm_pow(1.0 - j, 1.0 / 3.0);
    return (r * r - 2.0 * r + j * sqrt(r)) / r / sqrt(r * r - 3.0 * r + 2.0 * j * sqrt(r));
}
static double bh_isco_angular_momentum(double j, double r) {
    /* Black_Hole_ISCO_Specific_Angular_Momentum (gravitational units), black_holes/fundamentals.F90:235-282 */
    if (j > 0.99999) return 1.154700538 + 1.832972849 * dm_pow(1.0 - j, 1.0 / 3.0);
    return sqrt(r) * (r * r - 2.0 * j * sqrt(r) + j * j) / r / sqrt(r * r - 3.0 * r + 2.0 * j * sqrt(r));
}
static double adaf_table(const std_work *w, double spin, int column) {
    /* table1DLogarithmicLinear::interpolate with extrapolationTypeFix (objects/tables/_module.F90:1360-1405,
       1518-1533, 2505-2555) of the ADAF tabulations in 1-j (accretion_disks/ADAF.F90:394-447,481-523).
       1-j <= 0 (the reference would take log of a non-positive number) is treated as below the table. */
    const orc_table2d *t = &w->T->t[GLC_TABLE_ADAF];
    const double xinv = 1.0 - spin;
    const double lx0 = dm_log(t->x0[0]), lxn = dm_log(t->x0[t->n0 - 1]);
    const double inv = (double)(t->n0 - 1) / (lxn - lx0);
    double xe, h;
    int i;
    xe = (xinv > 0.0) ? dm_log(xinv) : lx0;
    if (xe < lx0) xe = lx0;
    if (xe > lxn) xe = lxn;
    if (xe >= lxn)
        i = t->n0 - 2;
    else {
        i = (int)((xe - lx0) * inv);
        if (i > t->n0 - 2) i = t->n0 - 2;
        if (i < 0) i = 0;
    }
    h = (xe - dm_log(t->x0[i])) * inv;
    return t->v[2 * i + column] * (1.0 - h) + t->v[2 * (i + 1) + column] * h;
}
static double disk_fraction_adaf(const std_work *w, const bh_state *b, double mdot) {
    /* switchedFractionADAF, accretion_disks/switched.F90:259-297 */
    const glc_params *P = w->P;
    double f = 0.0, lm, arg;
    if (!(b->eddington > 0.0 && mdot > 0.0)) return 0.0;
    lm = dm_log(mdot / b->eddington);
    if (P->accretionRateThinDiskMinimum > 0.0) {
        arg = fmin(+(lm - dm_log(P->accretionRateThinDiskMinimum)) / P->accretionRateTransitionWidth, 60.0);
        f = f + 1.0 / (1.0 + dm_exp(arg));
    }
    if (P->accretionRateThinDiskMaximum < DBL_MAX) {
        arg = fmin(-(lm - dm_log(P->accretionRateThinDiskMaximum)) / P->accretionRateTransitionWidth, 60.0);
        f = f + 1.0 / (1.0 + dm_exp(arg));
    }
    return f;
}
static double disk_efficiency_radiative(const std_work *w, const bh_state *b, double mdot) {
    /* switchedEfficiencyRadiative :199-226 over shakuraSunyaevEfficiencyRadiative (Shakura_Sunyaev.F90:69-93)
       and adafEfficiencyRadiative (ADAF.F90:453-479) with switchedEfficiencyRadiativeScalingADAF :299-331 */
    const glc_params *P = w->P;
    const double f = disk_fraction_adaf(w, b, mdot);
    const double eff_thin = 1.0 - b->e_isco;
    double eff_adaf = P->adafEfficiencyRadiationTypeThinDisk ? eff_thin : P->adafEfficiencyRadiation;
    double eff;
    if (P->scaleADAFRadiativeEfficiency) {
        double scaling = 1.0;
        if (b->eddington > 0.0 && mdot > 0.0) {
            const double md = mdot / b->eddington;
            if (P->accretionRateThinDiskMinimum > 0.0 && md < P->accretionRateThinDiskMinimum)
                scaling = md / P->accretionRateThinDiskMinimum;
        }
        eff_adaf = eff_adaf * scaling;
    }
    eff = 0.0;
    eff = eff + f * eff_adaf;
    eff = eff + (1.0 - f) * eff_thin;
    return eff;
}
static double disk_power_jet(const std_work *w, const bh_state *b, double mdot) {
    /* switchedPowerJet :228-242; shakuraSunyaevPowerJet (Shakura_Sunyaev.F90:95-155, Meier 2001);
       adafPowerJet (ADAF.F90:481-499) */
    /* (10**42.7, 10**41.7) * ergs * gigaYear / massSolar / kilo**2: compile-time constants of the reference */
    const double norm_kerr = 5.011872336272756e+42 * ORC_ERGS * ORC_GIGAYEAR / ORC_MASS_SOLAR / (ORC_KILO * ORC_KILO);
    const double norm_schw = 5.011872336272755e+41 * ORC_ERGS * ORC_GIGAYEAR / ORC_MASS_SOLAR / (ORC_KILO * ORC_KILO);
    const double f = disk_fraction_adaf(w, b, mdot);
    double thin = 0.0, adaf;
    if (mdot > 0.0) {
        const double md = mdot / b->eddington, mb = b->mass / 1.0e9;
        if (mb > 0.0 && md > 0.0) {
            if (b->spin > 0.8)
                thin = norm_kerr * dm_pow(mb, 0.9) * dm_pow(md, 1.2) / 1.0 *
                       (1.0 + 1.1 * b->spin + 0.29 * (b->spin * b->spin));
            else
                thin = norm_schw * dm_pow(mb, 0.9) * dm_pow(md, 1.2) / 1.0 * dm_exp(3.785 * b->spin);
        }
    }
    adaf = mdot * adaf_table(w, b->spin, 0);
    return (1.0 - f) * thin + f * adaf;
}
static double disk_rate_spin_up(const std_work *w, const bh_state *b, double mdot) {
    /* switchedRateSpinUp :244-257; shakuraSunyaevRateSpinUp (Shakura_Sunyaev.F90:157-179); adafRateSpinUp (ADAF.F90:501-523) */
    const double f = disk_fraction_adaf(w, b, mdot);
    double thin = 0.0, adaf;
    if (mdot != 0.0) thin = (b->l_isco - 2.0 * b->spin * b->e_isco) * mdot / b->mass;
    adaf = adaf_table(w, b->spin, 1) * mdot / b->mass;
    return (1.0 - f) * thin + f * adaf;
}
static double sph_gas_density(const std_work *w, double radius) {
    /* node%massDistribution(spheroid, gaseous)%density: Hernquist profile inside a spherical scaler
       (spheroid/standard/bound_functions.Inc: Node_Component_Spheroid_Standard_Mass_Distribution;
       mass_distributions/spherical/{scaler,Hernquist}.F90) */
    double a, m, x;
    if (!has(w, GLC_F_HAS_SPHEROID)) return 0.0;
    a = w->p[GLC_P_SPH_RADIUS];
    m = fmax(0.0, w->p[GLC_P_SPH_MASS_GAS]);
    if (a <= 0.0 || !(m > 0.0)) return 0.0;
    x = radius * (1.0 / a);
    return 0.5 / ORC_PI / x / ((1.0 + x) * (1.0 + x) * (1.0 + x)) * m / (a * a * a);
}
static void bh_accretion(std_work *w, bh_state *b) {
    /* blackHoleAccretionRateStandard::rateAccretion, black_holes/accretion_rates/standard.F90:221-440
       (no nuclear star cluster component, blackHoleBinarySeparationGrowthRate "zero" => velocityRelative = 0,
       radialPosition = 0 for the central black hole, cold mode not tracked) */
    const glc_params *P = w->P;
    const double density_gas_minimum = 1.0;
    const double velocity = 0.0 * ORC_MPC_PER_KMS_TO_GYR;
    double r_acc, rho, riso, jk;
    memset(b, 0, sizeof(*b));
    if (!(P->operatorMask & (GLC_OP_BLACK_HOLES_ACCRETION | GLC_OP_BLACK_HOLES_WINDS | GLC_OP_CGM_COOLING_HEATING))) return;
    if (!has(w, GLC_F_HAS_BH)) return;
    b->mass = w->p[GLC_P_BH_MASS];
    b->spin = w->p[GLC_P_BH_SPIN];
    if (!(b->mass > 0.0)) return;
    b->on = 1;
    /* a trial RK stage can carry the spin out of the range in which the Kerr expressions of the reference are finite
       (j > 1: cube roots of negative numbers; j < ~-0.63: negative radicand in the ISCO energy of the prograde formula).
       The ISCO quantities are evaluated at the spin clamped to [-0.5, 1]; states reached by accepted steps are in
       [0, 0.9999] (post-step clamp), so this only replaces NaNs of rejected trial stages. */
    jk = fmax(fmin(b->spin, 1.0), -0.5);
    riso = bh_isco_radius(jk);
    b->e_isco = bh_isco_energy(jk, riso);
    b->l_isco = bh_isco_angular_momentum(jk, riso);
    /* Black_Hole_Eddington_Accretion_Rate, fundamentals.F90:123-138 */
    b->eddington = 4.0 * ORC_PI * ORC_GRAVITATIONAL_CONSTANT * b->mass * ORC_MASS_HYDROGEN_ATOM * ORC_GIGAYEAR /
                   ORC_THOMSON_CROSS_SECTION / ORC_SPEED_LIGHT;
    /* spheroid */
    r_acc = fmax(bhl_radius(b->mass, P->bondiHoyleAccretionTemperatureSpheroid), 0.0);
    rho = sph_gas_density(w, r_acc);
    if (rho > density_gas_minimum) {
        double lj = ideal_gas_sound_speed(P->bondiHoyleAccretionTemperatureSpheroid) / sqrt(ORC_G_INTERNAL) / sqrt(rho);
        double eff;
        lj = fmin(lj, w->p[GLC_P_SPH_RADIUS]);
        if (lj > r_acc) rho = sph_gas_density(w, lj);
        b->acc_sph = fmax(P->bondiHoyleAccretionEnhancementSpheroid *
                              bhl_rate(b->mass, rho, velocity, P->bondiHoyleAccretionTemperatureSpheroid, 0, 0.0),
                          0.0);
        eff = disk_efficiency_radiative(w, b, b->acc_sph);
        if (eff > 0.0) b->acc_sph = fmin(b->acc_sph, b->eddington / eff);
    }
    /* hot halo */
    halo_scales(w);
    hh_profile(w);
    if (w->hh_valid) {
        const double t_hot = w->tvir; /* hotHaloTemperatureProfile virial */
        double f_hot = 1.0;
        r_acc = bhl_radius(b->mass, t_hot);
        r_acc = fmin(r_acc, hh_outer_radius(w));
        if (P->bondiHoyleAccretionHotModeOnly) {
            const double xf = cooling_radius(w) / w->rvir;
            if (xf < 0.9)
                f_hot = 1.0;
            else if (xf > 1.0)
                f_hot = 0.0;
            else {
                const double x = (xf - 0.9) / (1.0 - 0.9);
                f_hot = x * x * (2.0 * x - 3.0) + 1.0;
            }
        }
        rho = f_hot * hh_density(w, r_acc);
        if (rho > density_gas_minimum) {
            double rate_max, eff;
            b->acc_hot = fmax(P->bondiHoyleAccretionEnhancementHotHalo * bhl_rate(b->mass, rho, velocity, t_hot, 1, r_acc), 0.0);
            rate_max = fmax(w->p[GLC_P_HH_MASS] / (hh_outer_radius(w) / ideal_gas_sound_speed(t_hot) * ORC_MPC_PER_KMS_TO_GYR), 0.0);
            b->acc_hot = fmin(b->acc_hot, rate_max);
            eff = disk_efficiency_radiative(w, b, b->acc_hot);
            if (eff > 0.0) b->acc_hot = fmin(b->acc_hot, b->eddington / eff);
        }
    }
    b->acc = b->acc_sph + b->acc_hot;
}
static double bh_wind_power(std_work *w, const bh_state *b) {
    /* blackHoleWindCiotti2009::power, black_holes/winds/Ciotti2009.F90:153-246 */
    const glc_params *P = w->P;
    const double velocity_wind = 1.0e4, temperature_ism = 1.0e4;
    double eff = P->bhEfficiencyWind, coupled = 0.0, mgas, rs;
    if (P->bhEfficiencyWindScalesWithEfficiencyRadiative) eff = eff * disk_efficiency_radiative(w, b, b->acc);
    if (b->acc <= 0.0 || eff <= 0.0) return 0.0;
    mgas = has(w, GLC_F_HAS_SPHEROID) ? w->p[GLC_P_SPH_MASS_GAS] : 0.0;
    if (mgas > 0.0) {
        rs = w->p[GLC_P_SPH_RADIUS];
        if (rs > 0.0) {
            const double c2 = ORC_SPEED_LIGHT * ORC_SPEED_LIGHT;
            double p_wind = eff * b->acc * ORC_MASS_SOLAR / ORC_GIGAYEAR * c2 / 4.0 / ORC_PI / velocity_wind / ORC_KILO /
                            (rs * rs) / (ORC_MEGAPARSEC * ORC_MEGAPARSEC);
            double p_ism = 3.0 / 4.0 / ORC_PI * mgas * ORC_MASS_SOLAR / (rs * rs * rs) /
                           (ORC_MEGAPARSEC * ORC_MEGAPARSEC * ORC_MEGAPARSEC) / ORC_MASS_HYDROGEN_ATOM * 3.0 / 2.0 *
                           ORC_BOLTZMANN * temperature_ism;
            double x = p_ism / p_wind - 0.50;
            if (x <= 0.0)
                coupled = 0.0;
            else if (x >= 1.0)
                coupled = 1.0;
            else
                coupled = 3.0 * (x * x) - 2.0 * (x * x * x);
        }
    }
    eff = eff * coupled;
    return eff * b->acc * (ORC_SPEED_LIGHT * ORC_SPEED_LIGHT) / (ORC_KILO * ORC_KILO);
}

/* ------------------------------------------------------------ rate plumbing */
typedef struct {
    double *rate;
    int interrupt; /* last functionInterrupt requested */
    std_work *w;
} rate_ctx;

static void rate_add(rate_ctx *r, int prop, int component_flag, int create_if_needed, int code, double v) {
    /* generated <prop>Rate functions (Properties/Evolve.py:202-493): accumulate if the component
       exists; otherwise request creation when createIfNeeded and the rate is non-zero */
    if (r->w->flags & component_flag) {
        r->rate[prop] += v;
    } else if (create_if_needed && v != 0.0) {
        r->interrupt = code;
    }
}
#define HH(prop, v) rate_add(rc, prop, GLC_F_HAS_HOTHALO, 0, 0, v)
#define HHC(prop, v) rate_add(rc, prop, GLC_F_HAS_HOTHALO, 1, GLC_INT_HOTHALO_CREATE, v)
#define DK(prop, v) rate_add(rc, prop, GLC_F_HAS_DISK, 0, 0, v)
#define DKC(prop, v) rate_add(rc, prop, GLC_F_HAS_DISK, 1, GLC_INT_DISK_CREATE, v)
#define SP(prop, v) rate_add(rc, prop, GLC_F_HAS_SPHEROID, 0, 0, v)
#define SPC(prop, v) rate_add(rc, prop, GLC_F_HAS_SPHEROID, 1, GLC_INT_SPHEROID_CREATE, v)

static double fraction_outflow_stripped(std_work *w) {
    /* hotHaloOutflowStrippingStandard::fractionStripped, hot_halo/outflow_stripping/standard.F90:133-173 */
    double mo, mv;
    if (!has(w, GLC_F_IS_SATELLITE)) return 0.0;
    halo_scales(w);
    mo = hh_mass_enclosed(w, hh_outer_radius(w));
    mv = hh_mass_enclosed(w, w->rvir);
    return (mv > 0.0) ? w->P->outflowStrippingEfficiency * (1.0 - mo / mv) : w->P->outflowStrippingEfficiency;
}
static void hh_outflowing(rate_ctx *rc, double mass, double angmom, double abund) {
    /* deferred-rate functions Node_Component_Hot_Halo_Standard_Outflowing_{Mass,Ang_Mom,Abundances}_Rate,
       hot_halo/standard/_class.F90:598-722 */
    std_work *w = rc->w;
    double fs;
    if (!has(w, GLC_F_HAS_HOTHALO)) return;
    fs = fraction_outflow_stripped(w);
    HH(GLC_P_HH_STRIPPED_MASS, mass * fs);
    HH(GLC_P_HH_OUTFLOWED_MASS, mass * (1.0 - fs));
    HH(GLC_P_HH_OUTFLOWED_ANGMOM, angmom * (1.0 - fs) / (1.0 - w->P->fractionLossAngularMomentum));
    HH(GLC_P_HH_OUTFLOWED_ABUND, abund * (1.0 - fs));
}

static void feedback(rate_ctx *rc, int is_disk, double psi) {
    /* nodeOperatorStellarFeedback{Disks,Spheroids} (stellar_feedback/{disks,spheroids}.F90:116-206) with
       rateLimit(powerLaw): outflows/rate_limit.F90:111-170, outflows/power_law/_class.F90:118-164 */
    std_work *w = rc->w;
    const glc_params *P = w->P;
    const int pm = is_disk ? GLC_P_DISK_MASS_GAS : GLC_P_SPH_MASS_GAS;
    const int pa = is_disk ? GLC_P_DISK_ABUND_GAS : GLC_P_SPH_ABUND_GAS;
    const int pj = is_disk ? GLC_P_DISK_ANGMOM : GLC_P_SPH_ANGMOM;
    const int ps = is_disk ? GLC_P_DISK_MASS_STELLAR : GLC_P_SPH_MASS_STELLAR;
    const double radius = w->p[is_disk ? GLC_P_DISK_RADIUS : GLC_P_SPH_RADIUS];
    const double velocity = w->p[is_disk ? GLC_P_DISK_VELOCITY : GLC_P_SPH_VELOCITY];
    const double vchar = is_disk ? P->fbDiskVelocityCharacteristic : P->fbSpheroidVelocityCharacteristic;
    const double expo = is_disk ? P->fbDiskExponent : P->fbSpheroidExponent;
    double mass_gas = w->p[pm], z = mass_to_fraction(w->p[pa], mass_gas);
    double energy = ORC_FEEDBACK_ENERGY_INPUT_AT_INFINITY_CANONICAL * psi;
    double outflow, tdyn, outflow_max, mass_comp, j_out, z_out;
    outflow = (velocity <= 0.0) ? 0.0
                                : dm_pow(vchar / velocity, expo) * energy / ORC_FEEDBACK_ENERGY_INPUT_AT_INFINITY_CANONICAL;
    tdyn = (velocity <= 0.0 || radius <= 0.0) ? 1.0 : ORC_MPC_PER_KMS_TO_GYR * radius / velocity;
    outflow_max = fmax(mass_gas / tdyn / P->fbTimescaleOutflowFractionalMinimum, 0.0);
    if (outflow > outflow_max) outflow = outflow * outflow_max / outflow;
    if (!(outflow > 0.0)) return;
    mass_comp = mass_gas + w->p[ps];
    j_out = (mass_comp > 0.0) ? w->p[pj] * (outflow / mass_comp) : 0.0;
    z_out = (mass_gas > 0.0) ? z * outflow : 0.0;
    hh_outflowing(rc, outflow, j_out, z_out);
    if (is_disk) {
        DK(pm, -outflow);
        DK(pj, -j_out);
        DK(pa, -z_out);
    } else {
        SP(pm, -outflow);
        SP(pj, -j_out);
        SP(pa, -z_out);
    }
}

static void star_formation(rate_ctx *rc, int is_disk, double psi) {
    /* nodeOperatorStarFormation{Disks,Spheroids} + instantaneousRates (instantaneous.F90:173-182) */
    std_work *w = rc->w;
    const int pm = is_disk ? GLC_P_DISK_MASS_GAS : GLC_P_SPH_MASS_GAS;
    const int pa = is_disk ? GLC_P_DISK_ABUND_GAS : GLC_P_SPH_ABUND_GAS;
    const int ps = is_disk ? GLC_P_DISK_MASS_STELLAR : GLC_P_SPH_MASS_STELLAR;
    const int pz = is_disk ? GLC_P_DISK_ABUND_STELLAR : GLC_P_SPH_ABUND_STELLAR;
    double z = mass_to_fraction(w->p[pa], w->p[pm]);
    double rate_stellar = (1.0 - w->P->recycledFraction) * psi;
    double rate_z_stellar = z * rate_stellar;
    double rate_z_fuel = -rate_z_stellar + w->P->metalYield * psi;
    if (is_disk) {
        DK(ps, rate_stellar);
        DK(pm, -rate_stellar);
        DK(pz, rate_z_stellar);
        DK(pa, rate_z_fuel);
    } else {
        SP(ps, rate_stellar);
        SP(pm, -rate_stellar);
        SP(pz, rate_z_stellar);
        SP(pa, rate_z_fuel);
    }
}

/* ------------------------------------------------------------ model hooks */
static void work_init(std_work *w, orc_evolve_ctx *c, double time) {
    memset(w, 0, sizeof(*w));
    w->c = c;
    w->P = c->P;
    w->T = c->T;
    w->p = c->p;
    w->flags = c->flags;
    w->time = time;
}

int orc_std_active_list(const orc_evolve_ctx *c, int *active) {
    /* serialization order of TreeNodes/ODESolver.py:95-138; analytic properties removed (:262-275) */
    int n = 0, i;
    if (c->flags & GLC_F_HAS_BH) {
        active[n++] = GLC_P_BH_MASS;
        active[n++] = GLC_P_BH_SPIN;
    }
    if (c->flags & GLC_F_HAS_DISK)
        for (i = GLC_P_DISK_MASS_STELLAR; i <= GLC_P_DISK_ANGMOM; i++) active[n++] = i;
    if (c->flags & GLC_F_HAS_HOTHALO) {
        for (i = GLC_P_HH_MASS; i <= GLC_P_HH_OUTER_RADIUS; i++) active[n++] = i;
        active[n++] = GLC_P_HH_STRIPPED_MASS;
        active[n++] = GLC_P_HH_STRIPPED_ABUND;
    }
    active[n++] = GLC_P_SAT_BOUND_MASS;
    if (c->flags & GLC_F_HAS_SPHEROID)
        for (i = GLC_P_SPH_MASS_STELLAR; i <= GLC_P_SPH_ANGMOM; i++) active[n++] = i;
    return n;
}

void orc_std_solve_analytics(orc_evolve_ctx *c, double time) {
    /* {dmo,dmpScale,haloAngMom}Interpolate...SolveAnalytics: linear interpolation in time
       (dark_matter_only_mass/interpolate.F90:217-239, dark_matter_profile_scale/interpolate.F90:153-178,
       halo_angular_momentum_interpolate.F90:147-192); rate = 0 for non-primary progenitors/satellites */
    double *p = c->p;
    if (p[GLC_P_MASS_RATE] != 0.0)
        p[GLC_P_BASIC_MASS] = p[GLC_P_MASS_TARGET] + p[GLC_P_MASS_RATE] * (time - p[GLC_P_TIME_TARGET]);
    p[GLC_P_DMSCALE] = p[GLC_P_DMSCALE_TARGET] + (time - p[GLC_P_TIME_TARGET]) * p[GLC_P_DMSCALE_RATE];
    p[GLC_P_SPIN] = p[GLC_P_SPIN_TARGET] + (time - p[GLC_P_TIME_TARGET]) * p[GLC_P_SPIN_RATE];
}

void orc_std_pre_evolve(orc_evolve_ctx *c) {
    /* preEvolveTask: Node_Component_Hot_Halo_Standard_Pre_Evolve -> Initializor
       (hot_halo/standard/_class.F90:725-747,871-891): outerRadius := virial radius, once */
    std_work w;
    if ((c->flags & GLC_F_HAS_HOTHALO) && !(c->flags & F_HH_INIT)) {
        orc_std_solve_analytics(c, c->p[GLC_P_TIME]);
        work_init(&w, c, c->p[GLC_P_TIME]);
        halo_scales(&w);
        c->p[GLC_P_HH_OUTER_RADIUS] = w.rvir;
        c->flags |= F_HH_INIT;
    }
}

void orc_std_scales(orc_evolve_ctx *c, double *s) {
    std_work w;
    const double *p = c->p;
    const glc_params *P = c->P;
    work_init(&w, c, p[GLC_P_TIME]);
    halo_scales(&w);
    /* satellite: satellite/standard.F90:209-232 (massScaleFractional 1e-6) */
    s[GLC_P_SAT_BOUND_MASS] = 1.0e-6 * p[GLC_P_BASIC_MASS];
    if (c->flags & (GLC_F_HAS_DISK | GLC_F_HAS_SPHEROID)) {
        /* disk/standard/_class.F90:735-824, spheroid/standard/_class.F90:804-897 (identical expressions) */
        const int hd = (c->flags & GLC_F_HAS_DISK) != 0, hs = (c->flags & GLC_F_HAS_SPHEROID) != 0;
        double jd = hd ? p[GLC_P_DISK_ANGMOM] : 0.0, js = hs ? p[GLC_P_SPH_ANGMOM] : 0.0;
        double mgd = hd ? p[GLC_P_DISK_MASS_GAS] : 0.0, msd = hd ? p[GLC_P_DISK_MASS_STELLAR] : 0.0;
        double mgs = hs ? p[GLC_P_SPH_MASS_GAS] : 0.0, mss = hs ? p[GLC_P_SPH_MASS_STELLAR] : 0.0;
        double zgd = hd ? p[GLC_P_DISK_ABUND_GAS] : 0.0, zsd = hd ? p[GLC_P_DISK_ABUND_STELLAR] : 0.0;
        double zgs = hs ? p[GLC_P_SPH_ABUND_GAS] : 0.0, zss = hs ? p[GLC_P_SPH_ABUND_STELLAR] : 0.0;
        double sj = fmax(fabs(jd) + fabs(js), 0.1);
        double sm = fmax(fabs(mgd) + fabs(mgs) + fabs(msd) + fabs(mss), 1.0);
        double sz = fmax(fabs(zgd) + fabs(zsd) + fabs(zgs) + fabs(zss), fmax(sm * 1.0e-4, 1.0));
        if (hd) {
            s[GLC_P_DISK_ANGMOM] = sj;
            s[GLC_P_DISK_MASS_GAS] = s[GLC_P_DISK_MASS_STELLAR] = sm;
            s[GLC_P_DISK_ABUND_GAS] = s[GLC_P_DISK_ABUND_STELLAR] = sz;
        }
        if (hs) {
            s[GLC_P_SPH_ANGMOM] = sj;
            s[GLC_P_SPH_MASS_GAS] = s[GLC_P_SPH_MASS_STELLAR] = sm;
            s[GLC_P_SPH_ABUND_GAS] = s[GLC_P_SPH_ABUND_STELLAR] = sz;
        }
    }
    if (c->flags & GLC_F_HAS_HOTHALO) {
        /* hot_halo/standard/_class.F90:804-850 */
        double sm = p[GLC_P_BASIC_MASS] * P->hotHaloScaleMassRelative;
        double sj = p[GLC_P_BASIC_MASS] * w.rvir * w.vvir * P->hotHaloScaleMassRelative;
        s[GLC_P_HH_MASS] = s[GLC_P_HH_OUTFLOWED_MASS] = s[GLC_P_HH_UNACCRETED_MASS] = sm;
        s[GLC_P_HH_ABUND] = s[GLC_P_HH_UNACCRETED_ABUND] = s[GLC_P_HH_OUTFLOWED_ABUND] = sm;
        s[GLC_P_HH_ANGMOM] = s[GLC_P_HH_OUTFLOWED_ANGMOM] = sj;
        s[GLC_P_HH_OUTER_RADIUS] = w.rvir * P->hotHaloScaleRadiusRelative;
        /* stripped* scales are only set for satellites (:843-847); generated default scale is 1 */
        if (c->flags & GLC_F_IS_SATELLITE)
            s[GLC_P_HH_STRIPPED_MASS] = s[GLC_P_HH_STRIPPED_ABUND] = sm;
        else
            s[GLC_P_HH_STRIPPED_MASS] = s[GLC_P_HH_STRIPPED_ABUND] = 1.0;
    }
    if (c->flags & GLC_F_HAS_BH) {
        /* Node_Component_Black_Hole_Standard_Scale_Set, black_hole/standard/_class.F90:193-257 (no nuclear star
           cluster: massStellar = spheroid stellar mass if positive, else 0) */
        double mstar = (c->flags & GLC_F_HAS_SPHEROID) ? p[GLC_P_SPH_MASS_STELLAR] : 0.0;
        if (!(mstar > 0.0)) mstar = 0.0;
        s[GLC_P_BH_MASS] = fmax(fmax(1.0, 1.0e-4 * mstar), p[GLC_P_BH_MASS]);
        s[GLC_P_BH_SPIN] = 1.0;
    }
}

int orc_std_rates(orc_evolve_ctx *c, double time, double *rate) {
    std_work w;
    rate_ctx rcs, *rc = &rcs;
    const glc_params *P = c->P;
    const double *p = c->p;
    double dlnrho_dt;
    bh_state bh;
    work_init(&w, c, time);
    rcs.rate = rate;
    rcs.interrupt = GLC_INT_NONE;
    rcs.w = &w;
    halo_scales(&w);
    dlnrho_dt = w.dlnrho_dt;
    /* <eventHook preDerivative>: galactic structure solve (standard.F90:1045) */
    structure_solve(&w);
    if (!w.solvable) return GLC_INT_NONE;

    /* satelliteMassLoss: satellite/mass_loss/_class.F90:230-257, darkMatterHaloMassLossRate default "zero" */
    if (P->operatorMask & GLC_OP_SATELLITE_MASS_LOSS)
        rate[GLC_P_SAT_BOUND_MASS] += (c->flags & GLC_F_IS_SATELLITE) ? 0.0 : p[GLC_P_MASS_RATE];

    /* starFormationDisks: star_formation/disks.F90:200-284 */
    if ((P->operatorMask & GLC_OP_STAR_FORMATION_DISKS) && has(&w, GLC_F_HAS_DISK)) {
        if (!(p[GLC_P_DISK_ANGMOM] < 0.0 || p[GLC_P_DISK_RADIUS] < 0.0 || p[GLC_P_DISK_MASS_GAS] < 0.0))
            star_formation(rc, 1, sfr_disk(&w));
    }
    /* starFormationSpheroids: star_formation/spheroids.F90:152-240 */
    if ((P->operatorMask & GLC_OP_STAR_FORMATION_SPHEROIDS) && has(&w, GLC_F_HAS_SPHEROID)) {
        if (!(p[GLC_P_SPH_ANGMOM] < 1.0e-20 || p[GLC_P_SPH_RADIUS] < 1.0e-12 || p[GLC_P_SPH_MASS_GAS] < 1.0e-6))
            star_formation(rc, 0, sfr_spheroid(&w));
    }
    if ((P->operatorMask & GLC_OP_STELLAR_FEEDBACK_DISKS) && has(&w, GLC_F_HAS_DISK)) {
        if (!(p[GLC_P_DISK_ANGMOM] < 0.0 || p[GLC_P_DISK_RADIUS] < 0.0 || p[GLC_P_DISK_MASS_GAS] < 0.0))
            feedback(rc, 1, sfr_disk(&w));
    }
    if ((P->operatorMask & GLC_OP_STELLAR_FEEDBACK_SPHEROIDS) && has(&w, GLC_F_HAS_SPHEROID)) {
        if (!(p[GLC_P_SPH_ANGMOM] < 1.0e-20 || p[GLC_P_SPH_RADIUS] < 1.0e-12 || p[GLC_P_SPH_MASS_GAS] < 1.0e-6))
            feedback(rc, 0, sfr_spheroid(&w));
    }

    /* barInstability: bar_instability.F90:145-249 with efstathiou1982 (Efstathiou1982.F90:153-258) */
    if ((P->operatorMask & GLC_OP_BAR_INSTABILITY) && has(&w, GLC_F_HAS_DISK) &&
        !(p[GLC_P_DISK_ANGMOM] < 0.0 || p[GLC_P_DISK_RADIUS] < 0.0 || p[GLC_P_DISK_MASS_GAS] < 0.0)) {
        double timescale = -1.0;
        if (w.plausible) {
            const double stability_isolated = 0.6221297315, boost = 1.1800237580;
            double md = p[GLC_P_DISK_MASS_GAS] + p[GLC_P_DISK_MASS_STELLAR];
            if (p[GLC_P_DISK_ANGMOM] > 0.0 && p[GLC_P_DISK_VELOCITY] > 0.0 && p[GLC_P_DISK_RADIUS] > 0.0) {
                double fgas = p[GLC_P_DISK_MASS_GAS] / md;
                double thr = P->barStabilityThresholdStellar * (1.0 - fgas) + P->barStabilityThresholdGaseous * fgas;
                double est = DBL_MAX;
                if (md >= 0.0) {
                    double vself = sqrt(ORC_G_INTERNAL * md / p[GLC_P_DISK_RADIUS]);
                    if (vself > 0.0) est = fmax(stability_isolated, boost * p[GLC_P_DISK_VELOCITY] / vself);
                }
                if (est < thr) {
                    double tdyn = ORC_MPC_PER_KMS_TO_GYR * p[GLC_P_DISK_RADIUS] /
                                  fmin(p[GLC_P_DISK_VELOCITY], ORC_SPEED_LIGHT / ORC_KILO);
                    double a = thr - stability_isolated, b = thr - est, tdim;
                    tdim = (a > 1.0e10 * b) ? 1.0e10 : (a / b) * (a / b);
                    timescale = fmax(tdyn, 1.0e-9) * tdim;
                }
            }
        }
        if (!(timescale < 0.0)) {
            double tr;
            tr = fmax(0.0, p[GLC_P_DISK_MASS_GAS]) / timescale;
            DK(GLC_P_DISK_MASS_GAS, -tr);
            SPC(GLC_P_SPH_MASS_GAS, tr);
            tr = fmax(0.0, p[GLC_P_DISK_MASS_STELLAR]) / timescale;
            DK(GLC_P_DISK_MASS_STELLAR, -tr);
            SPC(GLC_P_SPH_MASS_STELLAR, tr);
            tr = fmax(0.0, p[GLC_P_DISK_ANGMOM]) / timescale;
            DK(GLC_P_DISK_ANGMOM, -(1.0 - 1.0) * tr);
            SPC(GLC_P_SPH_ANGMOM, 1.0 * tr);
            tr = fmax(0.0, p[GLC_P_DISK_ABUND_GAS]) / timescale;
            DK(GLC_P_DISK_ABUND_GAS, -tr);
            SPC(GLC_P_SPH_ABUND_GAS, tr);
            tr = fmax(0.0, p[GLC_P_DISK_ABUND_STELLAR]) / timescale;
            DK(GLC_P_DISK_ABUND_STELLAR, -tr);
            SPC(GLC_P_SPH_ABUND_STELLAR, tr);
            /* :240-246: external driving torque term is zero for efstathiou1982 */
        }
    }

    /* blackHolesSeed: black_holes/seed.F90:153-182 (blackHoleSeeds fixed): create the seed by interrupt */
    if ((P->operatorMask & GLC_OP_BLACK_HOLES_SEED) && !has(&w, GLC_F_HAS_BH) && P->bhSeedMass > 0.0)
        rcs.interrupt = GLC_INT_BH_CREATE;
    bh_accretion(&w, &bh);
    /* blackHolesAccretion: black_holes/accretion.F90:113-173 */
    if ((P->operatorMask & GLC_OP_BLACK_HOLES_ACCRETION) && has(&w, GLC_F_HAS_BH) && bh.on && bh.acc > 0.0) {
        const double eff_rad = disk_efficiency_radiative(&w, &bh, bh.acc);
        const double eff_jet = disk_power_jet(&w, &bh, bh.acc) / bh.acc / (ORC_SPEED_LIGHT * ORC_SPEED_LIGHT) / (ORC_KILO * ORC_KILO);
        const double reduced = bh.acc * (1.0 - eff_rad - eff_jet);
        const double spin_up = disk_rate_spin_up(&w, &bh, bh.acc);
        rate[GLC_P_BH_MASS] += reduced;
        /* Node_Component_Spheroid_Standard_Mass_Gas_Sink_Rate, spheroid/standard/_class.F90:638-671 */
        if (has(&w, GLC_F_HAS_SPHEROID) && -bh.acc_sph != 0.0) {
            const double mg = p[GLC_P_SPH_MASS_GAS], ms = p[GLC_P_SPH_MASS_STELLAR], r = -bh.acc_sph;
            if (mg > 0.0 && mg + ms > 0.0) {
                SP(GLC_P_SPH_MASS_GAS, r);
                SP(GLC_P_SPH_ANGMOM, (r / (mg + ms)) * p[GLC_P_SPH_ANGMOM]);
                SP(GLC_P_SPH_ABUND_GAS, (r / mg) * p[GLC_P_SPH_ABUND_GAS]);
            }
        }
        /* Node_Component_Hot_Halo_Standard_Mass_Sink -> Hot_Gas_All_Rate, hot_halo/standard/_class.F90:749-800 */
        if (has(&w, GLC_F_HAS_HOTHALO) && -bh.acc_hot != 0.0) {
            const double mg = p[GLC_P_HH_MASS], r = -bh.acc_hot;
            if (mg > 0.0) {
                HH(GLC_P_HH_MASS, r);
                HH(GLC_P_HH_ANGMOM, p[GLC_P_HH_ANGMOM] * (r / mg));
                HH(GLC_P_HH_ABUND, p[GLC_P_HH_ABUND] * (r / mg));
            }
        }
        rate[GLC_P_BH_SPIN] += spin_up;
    }
    /* blackHolesWinds: black_holes/winds.F90:103-134 -> Node_Component_Spheroid_Standard_Energy_Gas_Input_Rate,
       spheroid/standard/_class.F90:673-725 */
    if ((P->operatorMask & GLC_OP_BLACK_HOLES_WINDS) && has(&w, GLC_F_HAS_BH) && bh.on) {
        const double power = bh_wind_power(&w, &bh);
        if (power != 0.0 && has(&w, GLC_F_HAS_SPHEROID)) {
            const double mg = p[GLC_P_SPH_MASS_GAS], ms = p[GLC_P_SPH_MASS_STELLAR], vs = p[GLC_P_SPH_VELOCITY];
            if (mg > 0.0 && mg + ms > 0.0 && vs > 0.0) {
                const double out = P->spheroidEfficiencyEnergeticOutflow * power / (vs * vs);
                const double jout = (out / (mg + ms)) * p[GLC_P_SPH_ANGMOM];
                const double zout = (out / mg) * p[GLC_P_SPH_ABUND_GAS];
                SP(GLC_P_SPH_MASS_GAS, -out);
                SP(GLC_P_SPH_ANGMOM, -jout);
                SP(GLC_P_SPH_ABUND_GAS, -zout);
                hh_outflowing(rc, out, jout, zout);
            }
        }
    }

    /* CGMAccretion: circumgalactic_medium/accretion.F90:517-593 with accretionHaloSimple
       (accretion/halo/simple.F90:281-378,592-613); IGM metallicity zero */
    if (P->operatorMask & GLC_OP_CGM_ACCRETION) {
        double rate_hot = 0.0, rate_failed = 0.0, rate_j = 0.0;
        const double fb = P->OmegaBaryon / P->OmegaMatter;
        const int hh = has(&w, GLC_F_HAS_HOTHALO);
        if (!(c->flags & GLC_F_IS_SATELLITE)) {
            double failed = (time > P->timeReionization && w.vvir < P->velocitySuppressionReionization) ? 1.0 : 0.0;
            double unaccreted = hh ? p[GLC_P_HH_UNACCRETED_MASS] : 0.0;
            double growth = p[GLC_P_MASS_RATE] / p[GLC_P_BASIC_MASS];
            rate_hot = fb * p[GLC_P_MASS_RATE] * (1.0 - failed) + unaccreted * growth * (1.0 - failed);
            rate_failed = fb * p[GLC_P_MASS_RATE] * failed - unaccreted * growth * (1.0 - failed);
        }
        if (p[GLC_P_MASS_RATE] != 0.0) rate_j = p[GLC_P_SPIN_RATE] * rate_hot / p[GLC_P_MASS_RATE];
        if (rate_hot > 0.0 || (hh && p[GLC_P_HH_MASS] > 0.0) || P->allowNegativeCGMMass) HHC(GLC_P_HH_MASS, rate_hot);
        if (rate_failed > 0.0 || (hh && p[GLC_P_HH_MASS] > 0.0) || P->allowNegativeCGMMass)
            HHC(GLC_P_HH_UNACCRETED_MASS, rate_failed);
        HHC(GLC_P_HH_ANGMOM, rate_j);
    }

    /* CGMOutflowReincorporation: outflow_reincorporation.F90:272-349 (includeSatellites = true) */
    if ((P->operatorMask & GLC_OP_CGM_OUTFLOW_REINCORPORATION) && has(&w, GLC_F_HAS_HOTHALO)) {
        double mo = p[GLC_P_HH_OUTFLOWED_MASS];
        double ret = mo * P->reincorporationMultiplier / w.tdyn; /* halo_dynamical_time.F90:113-128 */
        if (mo > 0.0) {
            double rj = p[GLC_P_HH_OUTFLOWED_ANGMOM] * (ret / mo);
            double rz = p[GLC_P_HH_OUTFLOWED_ABUND] * (ret / mo);
            HH(GLC_P_HH_OUTFLOWED_MASS, -ret);
            HH(GLC_P_HH_OUTFLOWED_ANGMOM, -rj);
            HH(GLC_P_HH_OUTFLOWED_ABUND, -rz);
            HH(GLC_P_HH_MASS, ret);
            HH(GLC_P_HH_ANGMOM, rj);
            HH(GLC_P_HH_ABUND, rz);
        }
    }

    /* CGMCoolingHeating: cooling_heating.F90:216-381 (component=disk, coolingFrom=currentNode) */
    if ((P->operatorMask & GLC_OP_CGM_COOLING_HEATING) && has(&w, GLC_F_HAS_HOTHALO) && p[GLC_P_HH_MASS] > 0.0 &&
        !(p[GLC_P_HH_ANGMOM] <= 0.0 || hh_outer_radius(&w) <= 0.0)) {
        double cool = cooling_rate(&w);
        /* circumgalacticMediumHeatingAGNFeedback (circumgalactic_medium/heating/AGN_feedback.F90:103-126) over
           blackHoleCGMHeatingJetPower (black_holes/CGM_heating/jet_power.F90:120-138) */
        double heat = ((bh.on ? P->bhEfficiencyRadioMode * disk_power_jet(&w, &bh, bh.acc) : 0.0)) / (w.vvir * w.vvir);
        if (heat > cool) {
            if (P->excessHeatDrivesOutflow) {
                double out = fmin(heat - cool, P->rateMaximumExpulsion * p[GLC_P_HH_MASS] / w.tdyn);
                double rz = p[GLC_P_HH_ABUND] * (out / p[GLC_P_HH_MASS]);
                double rj = p[GLC_P_HH_ANGMOM] * (out / p[GLC_P_HH_MASS]);
                HH(GLC_P_HH_MASS, -out);
                HH(GLC_P_HH_ABUND, -rz);
                HH(GLC_P_HH_ANGMOM, -rj);
                if (c->flags & GLC_F_IS_SATELLITE) {
                    HH(GLC_P_HH_STRIPPED_MASS, out);
                    HH(GLC_P_HH_STRIPPED_ABUND, rz);
                }
            }
        } else if (cool > heat) {
            double rinfall, rj, rz;
            cool = fmax(0.0, cool - heat);
            rinfall = cooling_radius(&w);
            rj = cool * cooling_specific_angular_momentum(&w, rinfall);
            rz = cool * p[GLC_P_HH_ABUND] / p[GLC_P_HH_MASS];
            HH(GLC_P_HH_MASS, -cool);
            HH(GLC_P_HH_ANGMOM, -rj);
            HH(GLC_P_HH_ABUND, -rz);
            DKC(GLC_P_DISK_MASS_GAS, cool);
            DKC(GLC_P_DISK_ABUND_GAS, rz);
            DKC(GLC_P_DISK_ANGMOM, rj * (1.0 - P->fractionLossAngularMomentum));
        }
    }

    /* CGMOuterRadiusRamPressureStripping: outer_radius/ram_pressure_stripping.F90:151-313 with
       hotHaloRamPressureStripping=virialRadius (radiusStripped == virial radius => no stripping term) */
    if ((P->operatorMask & GLC_OP_CGM_OUTER_RADIUS) && has(&w, GLC_F_HAS_HOTHALO)) {
        double ret = p[GLC_P_HH_OUTFLOWED_MASS] * P->reincorporationMultiplier / w.tdyn;
        double router = hh_outer_radius(&w);
        if (router < w.rvir) {
            double rho = hh_density(&w, router);
            if (router > 0.0 && rho > 0.0) {
                double rho_min = P->OmegaBaryon / P->OmegaMatter * p[GLC_P_BASIC_MASS] / (w.rvir * w.rvir * w.rvir) / 4.0 / ORC_PI;
                HH(GLC_P_HH_OUTER_RADIUS, ret / 4.0 / ORC_PI / (router * router) / fmax(rho, rho_min));
            } else if (ret > 0.0) {
                HH(GLC_P_HH_OUTER_RADIUS, ret / p[GLC_P_BASIC_MASS] * w.rvir);
            }
        }
        if (!(c->flags & GLC_F_IS_SATELLITE)) HH(GLC_P_HH_OUTER_RADIUS, rvir_growth_rate(&w, dlnrho_dt));
    }
    return rcs.interrupt;
}

void orc_std_post_step(orc_evolve_ctx *c, int *status) {
    double *p = c->p;
    /* Node_Component_Disk_Standard_Post_Step, disk/standard/_class.F90:473-677
       (diskNegativeAngularMomentumAllowed = true) */
    if (c->flags & GLC_F_HAS_DISK) {
        if (p[GLC_P_DISK_MASS_GAS] < 0.0) {
            double m = p[GLC_P_DISK_MASS_GAS] + p[GLC_P_DISK_MASS_STELLAR], j;
            if (m == 0.0) {
                j = 0.0;
                p[GLC_P_DISK_MASS_STELLAR] = 0.0;
                p[GLC_P_DISK_ABUND_STELLAR] = 0.0;
            } else {
                j = p[GLC_P_DISK_ANGMOM] / m;
                if (j < 0.0) j = p[GLC_P_DISK_RADIUS] * p[GLC_P_DISK_VELOCITY];
            }
            p[GLC_P_DISK_MASS_GAS] = 0.0;
            p[GLC_P_DISK_ABUND_GAS] = 0.0;
            p[GLC_P_DISK_ANGMOM] = j * p[GLC_P_DISK_MASS_STELLAR];
            if (*status == ORC_GSL_SUCCESS) *status = ORC_GSL_CONTINUE;
        }
        if (p[GLC_P_DISK_MASS_STELLAR] < 0.0) {
            double m = p[GLC_P_DISK_MASS_GAS] + p[GLC_P_DISK_MASS_STELLAR], j;
            if (m == 0.0) {
                j = 0.0;
                p[GLC_P_DISK_MASS_GAS] = 0.0;
                p[GLC_P_DISK_ABUND_GAS] = 0.0;
            } else {
                j = p[GLC_P_DISK_ANGMOM] / m;
                if (j < 0.0) j = p[GLC_P_DISK_RADIUS] * p[GLC_P_DISK_VELOCITY];
            }
            p[GLC_P_DISK_MASS_STELLAR] = 0.0;
            p[GLC_P_DISK_ABUND_STELLAR] = 0.0;
            p[GLC_P_DISK_ANGMOM] = j * p[GLC_P_DISK_MASS_GAS];
            if (*status == ORC_GSL_SUCCESS) *status = ORC_GSL_CONTINUE;
        }
        if (p[GLC_P_DISK_ANGMOM] < 0.0) {
            if (p[GLC_P_DISK_MASS_STELLAR] + p[GLC_P_DISK_MASS_GAS] <= 0.0) p[GLC_P_DISK_ANGMOM] = 0.0;
            if (*status == ORC_GSL_SUCCESS) *status = ORC_GSL_CONTINUE;
        }
    }
    /* Node_Component_Hot_Halo_Standard_Post_Step, hot_halo/standard/_class.F90:455-530 */
    if (c->flags & GLC_F_HAS_HOTHALO) {
        if (p[GLC_P_HH_MASS] < 0.0) {
            p[GLC_P_HH_MASS] = 0.0;
            if (*status == ORC_GSL_SUCCESS) *status = ORC_GSL_CONTINUE;
        }
        /* NB: the outerRadius check (:484-488) goes through the deferred getter, which is clipped to
           >= scaleRadiusRelative * r_vir and therefore never negative */
    }
    /* Node_Component_Spheroid_Standard_Post_Step, spheroid/standard/_class.F90:464-636 */
    if (c->flags & GLC_F_HAS_SPHEROID) {
        if (p[GLC_P_SPH_MASS_GAS] < 0.0) {
            double m = p[GLC_P_SPH_MASS_GAS] + p[GLC_P_SPH_MASS_STELLAR], j;
            if (m == 0.0) {
                j = 0.0;
                p[GLC_P_SPH_MASS_STELLAR] = 0.0;
                p[GLC_P_SPH_ABUND_STELLAR] = 0.0;
            } else
                j = p[GLC_P_SPH_ANGMOM] / m;
            p[GLC_P_SPH_MASS_GAS] = 0.0;
            p[GLC_P_SPH_ABUND_GAS] = 0.0;
            p[GLC_P_SPH_ANGMOM] = j * p[GLC_P_SPH_MASS_STELLAR];
            if (*status == ORC_GSL_SUCCESS) *status = ORC_GSL_CONTINUE;
        }
        if (p[GLC_P_SPH_MASS_STELLAR] < 0.0) {
            double m = p[GLC_P_SPH_MASS_GAS] + p[GLC_P_SPH_MASS_STELLAR], j;
            if (m == 0.0) {
                j = 0.0;
                p[GLC_P_SPH_MASS_GAS] = 0.0;
                p[GLC_P_SPH_ABUND_GAS] = 0.0;
            } else
                j = p[GLC_P_SPH_ANGMOM] / m;
            p[GLC_P_SPH_MASS_STELLAR] = 0.0;
            p[GLC_P_SPH_ABUND_STELLAR] = 0.0;
            p[GLC_P_SPH_ANGMOM] = j * p[GLC_P_SPH_MASS_GAS];
            if (*status == ORC_GSL_SUCCESS) *status = ORC_GSL_CONTINUE;
        }
        if (p[GLC_P_SPH_ANGMOM] < 0.0) {
            double j = p[GLC_P_SPH_RADIUS] * p[GLC_P_SPH_VELOCITY] / c->P->spheroidRatioAngularMomentumScaleRadius;
            p[GLC_P_SPH_ANGMOM] = j * (p[GLC_P_SPH_MASS_GAS] + p[GLC_P_SPH_MASS_STELLAR]);
            if (*status == ORC_GSL_SUCCESS) *status = ORC_GSL_CONTINUE;
        }
    }
    /* Node_Component_Black_Hole_Standard_Post_Evolve, black_hole/standard/_class.F90:422-462 */
    if (c->flags & GLC_F_HAS_BH) {
        if (p[GLC_P_BH_SPIN] > 0.9999 || p[GLC_P_BH_SPIN] < 0.0) {
            p[GLC_P_BH_SPIN] = fmax(fmin(p[GLC_P_BH_SPIN], 0.9999), 0.0);
            if (*status == ORC_GSL_SUCCESS) *status = ORC_GSL_CONTINUE;
        }
        if (p[GLC_P_BH_MASS] < 0.0) {
            p[GLC_P_BH_MASS] = c->P->bhSeedMass;
            if (*status == ORC_GSL_SUCCESS) *status = ORC_GSL_CONTINUE;
        }
    }
}

void orc_std_post_evolve(orc_evolve_ctx *c) {
    /* <eventHook postEvolve>: galacticStructureSolverEquilibrium::solve at the final state
       (equilibrium.F90:172,197-217).  The satellite stripped-mass hand-off of the hot halo's postEvolve
       (hot_halo/standard/_class.F90:532-596) touches the host node and is done by the tree-level host. */
    std_work w;
    work_init(&w, c, c->p[GLC_P_TIME]);
    structure_solve(&w);
}

/* debugging aid for the bit-exact parity work: selected intermediates of one RHS evaluation */
void orc_probe_node(const glc_params *P, const orc_tables *T, double *props, int flags, double *out) {
    orc_evolve_ctx c;
    std_work w;
    double r0;
    memset(&c, 0, sizeof(c));
    c.P = P;
    c.T = T;
    c.p = props;
    c.flags = flags;
    orc_std_solve_analytics(&c, props[GLC_P_TIME]);
    work_init(&w, &c, props[GLC_P_TIME]);
    halo_scales(&w);
    hh_profile(&w);
    r0 = props[GLC_P_DISK_RADIUS] > 0.0 ? props[GLC_P_DISK_RADIUS] : 0.01 * w.rvir;
    out[0] = w.rvir;
    out[1] = w.vvir;
    out[2] = w.tvir;
    out[3] = w.hh_rho0;
    out[4] = nfw_mass_enclosed(&w, r0);
    out[5] = ac_orbital_mean(&w, r0);
    out[6] = baryonic_vc2(&w, r0);
    out[7] = dark_matter_mass_enclosed(&w, r0);
    out[8] = disk_bessel_factor(&w, 0.37);
    out[9] = hh_mass_enclosed(&w, r0);
    out[10] = orc_fast_exponentiate(1.0e-3, 1.0, 0.7, 1.0e4, 0.0123);
    out[11] = nfw_radius_from_j(&w, 0.3 * w.rvir * w.vvir);
    out[12] = has(&w, GLC_F_HAS_DISK) ? sfr_disk(&w) : 0.0;
    out[13] = has(&w, GLC_F_HAS_HOTHALO) && props[GLC_P_HH_MASS] > 0 ? cooling_radius(&w) : 0.0;
    out[14] = sqrt(ORC_G_INTERNAL * out[7] / r0 + out[6]);
    out[15] = dm_log(out[14] / r0);
}

/* rotation curve at `radius` as nodePropertyExtractorRotationCurve reports it for the radius specifiers of
 * testSuite/parameters/reproducibility/adiabaticContraction.xml: out[0] = all components / all mass, out[1] = dark halo /
 * dark (the contracted profile), out[2] = baryonic; km/s; out[3] = r_vir, out[4] = V_vir */
void orc_rotation_curve_probe(const glc_params *P, const orc_tables *T, double *props, int flags, double radius, double *out) {
    orc_evolve_ctx c;
    std_work w;
    double vdm2, vb2;
    memset(&c, 0, sizeof(c));
    c.P = P;
    c.T = T;
    c.p = props;
    c.flags = flags;
    orc_std_solve_analytics(&c, props[GLC_P_TIME]);
    work_init(&w, &c, props[GLC_P_TIME]);
    halo_scales(&w);
    hh_profile(&w);
    vdm2 = ORC_G_INTERNAL * dark_matter_mass_enclosed(&w, radius) / radius;
    vb2 = baryonic_vc2(&w, radius);
    out[0] = sqrt(vdm2 + vb2);
    out[1] = sqrt(vdm2);
    out[2] = sqrt(vb2);
    out[3] = w.rvir;
    out[4] = w.vvir;
}

/* known-answer access to the mass distributions the path uses (tests/test_oracle_mass_distributions.py pins them to
 * source/tests/mass_distributions.F90): beta profile (beta = 2/3) of total `mass` inside `router` with core radius `rcore`:
 * out[0] = M(<radius), out[1] = rho(radius), out[2], out[3] = the scale-free radial moments m = 2, 3 from 0 to radius/rcore
 * (radialMomentTwoThirds), out[4] = rho_0; Hernquist sphere of unit mass and scale length: out[5] = M(<radius) via the
 * rotation-curve term of baryonic_vc2 (V^2 r / G); out[6] = NFW scale-free enclosed mass m(radius). */
void orc_mass_distribution_probe(const glc_params *P, double mass, double rcore, double router, double radius, double *out) {
    orc_evolve_ctx c;
    std_work w;
    double props[GLC_NPROP];
    memset(&c, 0, sizeof(c));
    memset(props, 0, sizeof(props));
    c.P = P;
    c.p = props;
    c.flags = GLC_F_HAS_HOTHALO | GLC_F_HAS_SPHEROID;
    work_init(&w, &c, 1.0);
    /* fill the memoised profile directly (hh_profile would derive the radii from the halo scales) */
    w.halo_done = 1;
    w.hh_done = 1;
    w.hh_valid = 1;
    w.hh_router = router;
    w.hh_rcore = rcore;
    w.hh_mass = mass;
    {
        const double r = router / rcore;
        const double nf = (r < 1.0e-6) ? 3.0 / (r * r * r) + 9.0 / 5.0 / r - 36.0 * r / 175.0 : 1.0 / (r - dm_atan(r));
        w.hh_rho0 = mass / 4.0 / ORC_PI / (rcore * rcore * rcore) * nf;
        if (!beta_is_two_thirds(&w)) w.hh_rho0 = mass / 4.0 / ORC_PI / (rcore * rcore * rcore) / dm_beta_moment(2, r, P->hotHaloBeta);
    }
    out[0] = hh_mass_enclosed(&w, radius);
    out[1] = hh_density(&w, radius);
    out[2] = beta_is_two_thirds(&w) ? hh_radial_moment23(2, radius / rcore) : dm_beta_moment(2, radius / rcore, P->hotHaloBeta);
    out[3] = beta_is_two_thirds(&w) ? hh_radial_moment23(3, radius / rcore) : dm_beta_moment(3, radius / rcore, P->hotHaloBeta);
    out[4] = w.hh_rho0;
    /* Hernquist: unit mass in stars, unit scale length, no disk, no hot halo contribution */
    props[GLC_P_SPH_MASS_STELLAR] = 1.0;
    props[GLC_P_SPH_RADIUS] = 1.0;
    c.flags = GLC_F_HAS_SPHEROID;
    work_init(&w, &c, 1.0);
    w.halo_done = 1;
    out[5] = baryonic_vc2(&w, radius) * radius / ORC_G_INTERNAL;
    out[6] = nfw_mass_scale_free(radius);
}

/* known-answer access to the black-hole helper functions (tests/test_oracle_black_holes.py):
 * out[0..3] = ISCO radius, specific energy, specific angular momentum (gravitational units, prograde), Eddington
 * accretion rate (Msun/Gyr); out[4] = Bondi-Hoyle-Lyttleton radius (Mpc) and out[5] = rate (Msun/Gyr) for unit
 * density (1 Msun/Mpc^3) at `temperature`; out[6] = ideal-gas sound speed (km/s); out[7] = Jeans length at unit density */
void orc_bh_probe(double mass, double spin, double temperature, double *out) {
    const double r = bh_isco_radius(spin);
    out[0] = r;
    out[1] = bh_isco_energy(spin, r);
    out[2] = bh_isco_angular_momentum(spin, r);
    out[3] = 4.0 * ORC_PI * ORC_GRAVITATIONAL_CONSTANT * mass * ORC_MASS_HYDROGEN_ATOM * ORC_GIGAYEAR /
             ORC_THOMSON_CROSS_SECTION / ORC_SPEED_LIGHT;
    out[4] = bhl_radius(mass, temperature);
    out[5] = bhl_rate(mass, 1.0, 0.0, temperature, 0, 0.0);
    out[6] = ideal_gas_sound_speed(temperature);
    out[7] = ideal_gas_sound_speed(temperature) / sqrt(ORC_G_INTERNAL) / sqrt(1.0);
}

/* known answers of source/tests/dark_matter_profiles.F90:73-410 for the NFW profile ("Zhao1996 (1,3,1)" and NFW groups hold the
 * same vectors): a halo of `mass` at `time` with concentration `conc`; out[0] = M(<x r_s) / M_vir, out[1] = the radius, in units
 * of r_s, that massDistribution_%radiusFromSpecificAngularMomentum returns for j = sqrt(G M(<r) r) (the inverse tabulation of
 * NFW.F90:589-639: recovers x to 2e-4 in the reference's test) */
void orc_nfw_probe(const glc_params *P, const orc_tables *T, double mass, double time, double conc, double x, double *out) {
    orc_evolve_ctx c;
    std_work w;
    double props[GLC_NPROP], rs, r, m;
    memset(&c, 0, sizeof(c));
    memset(props, 0, sizeof(props));
    c.P = P;
    c.T = T;
    c.p = props;
    props[GLC_P_BASIC_MASS] = mass;
    props[GLC_P_TIME] = time;
    props[GLC_P_DMSCALE] = 1.0;
    work_init(&w, &c, time);
    halo_scales(&w);
    rs = w.rvir / conc;
    props[GLC_P_DMSCALE] = rs;
    r = x * rs;
    m = nfw_mass_enclosed(&w, r);
    out[0] = m / mass;
    out[1] = nfw_radius_from_j(&w, sqrt(ORC_G_INTERNAL * m * r)) / rs;
}
