// glc_model_standard.cuh -- device rate functions for the operator set of parameters/quickTest.xml:
// the RHS evaluated by standardDerivativesCompute (source/merger_trees/node_evolver/standard.F90:1019-1061)
// through nodeOperatorMulti (source/nodes/operators/multi.F90:313-332), and the scale-set / pre-evolve /
// post-step / post-evolve hooks of the standard components.  Citations are relative to
// /root/reference/source.
//
// GPU layout of the computation: one thread = one node.  Everything a reference
// "Calculations_Reset" memoises per RHS call (halo scales, beta-profile normalisation, cooling
// function at (T_vir, Z_hot), Krumholz factors, disk SFR) is computed once per call into a small
// register-resident struct (Work) and re-used by the operators, which are fused into one pass that
// accumulates directly into the rate vector.  Table look-ups go through the read-only path; the
// nested adaptive numerics (structure fixed point, Brent, QAG) are single-call-site state machines
// (glc_numerics.cuh).
//
// Documented deviations from quickTest.xml ("out of scope / next"):
// hotHaloRamPressureStripping=virialRadius; ADAF tabulations supplied by the host (GLC_TABLE_ADAF).
#pragma once

#include "glc_common.cuh"
#include "glc_numerics.cuh"

namespace glc {

struct Work {
    // halo scales
    double rhoMean, rvir, vvir, tdyn, tvir, dlnrhoDt;
    // hot-halo beta profile
    double hhRouter, hhRcore, hhRho0;
    bool hhValid;
    // cooling: CIE table values at the node's (T_vir, Z_hot), looked up once per RHS call
    double coolLambda, coolEfrac, coolXH, coolFHn, coolTavail;
    bool plausible, solvable;
    // dark-matter-only profile (darkMatterProfileDMO): M(<r) = dmoNorm * m(r / dmoScale), see dmo_prepare
    double dmoNorm, dmoScale;
};

GLC_DEVICE_INLINE double mass_to_fraction(double ab, double mass) {
    // Abundances_Mass_To_Mass_Fraction, objects/abundances.F90:811-828
    return (ab > mass) ? 1.0 : ((ab <= 0.0) ? 0.0 : ab / mass);
}
GLC_DEVICE_INLINE double hydrogen_mass_fraction(double z) {
    // objects/abundances.F90:830-850
    double x = z / kMetallicitySolar * (kHydrogenByMassSolar - kHydrogenByMassPrimordial) + kHydrogenByMassPrimordial;
    return fmin(fmax(x, 0.7), kHydrogenByMassPrimordial);
}
GLC_DEVICE_INLINE double hydrogen_number_fraction(double z) {
    const double nh = hydrogen_mass_fraction(z) / kAtomicMassHydrogen;
    const double nhe = fmin(z / kMetallicitySolar * (kHeliumByMassSolar - kHeliumByMassPrimordial) + kHeliumByMassPrimordial,
                            kHeliumByMassPrimordial) / kAtomicMassHelium;
    return nh / (nh + nhe);
}

// ------------------------------------------------------------------ CIE tables (CIE_file.F90:665-735)
GLC_DEVICE_INLINE int locate(const double *__restrict__ x, int n, double v) {
    int lo = 0, hi = n - 1;
    while (hi > lo + 1) {
        const int mid = (hi + lo) >> 1;
        if (GLC_LDG(x + mid) > v)
            hi = mid;
        else
            lo = mid;
    }
    return lo + 1;
}

struct CieFactors {
    int iT, iZ;
    double hT, hZ;
};

GLC_DEVICE_INLINE CieFactors cie_factors(const DeviceTable2D &t, bool isLog, bool firstZero,
                                                  double firstNonzero, double temperature, double metallicity) {
    CieFactors f;
    double tu = isLog ? dm_log(temperature) : temperature;
    int i = min(max(locate(t.x1, t.n1, tu), 1), t.n1 - 1);
    f.iT = i;
    f.hT = (tu - GLC_LDG(t.x1 + i - 1)) / (GLC_LDG(t.x1 + i) - GLC_LDG(t.x1 + i - 1));
    double zu = fmax(metallicity, 0.0);
    if (firstZero && zu < firstNonzero) {
        f.iZ = 1;
        f.hZ = zu / firstNonzero;
    } else {
        if (isLog) zu = dm_log(zu);
        i = min(max(locate(t.x0, t.n0, zu), 1), t.n0 - 1);
        f.iZ = i;
        f.hZ = (zu - GLC_LDG(t.x0 + i - 1)) / (GLC_LDG(t.x0 + i) - GLC_LDG(t.x0 + i - 1));
    }
    return f;
}
GLC_DEVICE_INLINE double cie_interpolate(const DeviceTable2D &t, bool isLog, const CieFactors &f) {
    const double *a = t.v + (size_t)(f.iZ - 1) * t.n1 + (f.iT - 1);
    const double *b = a + t.n1;
    const double r = GLC_LDG(a) * (1.0 - f.hT) * (1.0 - f.hZ) + GLC_LDG(b) * (1.0 - f.hT) * f.hZ +
                     GLC_LDG(a + 1) * f.hT * (1.0 - f.hZ) + GLC_LDG(b + 1) * f.hT * f.hZ;
    return isLog ? dm_exp(r) : r;
}

struct ModelStandard {
    // ---------------------------------------------------------------- small accessors
    GLC_DEVICE_INLINE bool has(const NodeCtx &c, int f) { return (c.flags & f) != 0; }

    GLC_DEVICE_INLINE uint32_t active_mask(int flags) {
        uint32_t m = 1u << GLC_P_SAT_BOUND_MASS;
        if (flags & GLC_F_HAS_BH) m |= (1u << GLC_P_BH_MASS) | (1u << GLC_P_BH_SPIN);
        if (flags & GLC_F_HAS_DISK) m |= 0x1fu << GLC_P_DISK_MASS_STELLAR;
        if (flags & GLC_F_HAS_HOTHALO) m |= 0x7ffu << GLC_P_HH_MASS;
        if (flags & GLC_F_HAS_SPHEROID) m |= 0x1fu << GLC_P_SPH_MASS_STELLAR;
        return m;
    }

    GLC_DEVICE_INLINE void solve_analytics(NodeCtx &c, double time) {
        // {dmo,dmpScale,haloAngMom}Interpolate...SolveAnalytics (dark_matter_only_mass/interpolate.F90:217-239,
        // dark_matter_profile_scale/interpolate.F90:153-178, halo_angular_momentum_interpolate.F90:147-192)
        if (c.massRate != 0.0) c.basicMass = c.massTarget + c.massRate * (time - c.timeTarget);
        c.dmScale = c.scaleTarget + (time - c.timeTarget) * c.scaleRate;
        c.spinJ = c.spinTarget + (time - c.timeTarget) * c.spinRate;
    }

    // ---------------------------------------------------------------- halo scales
    GLC_DEVICE_INLINE void halo_scales(const NodeCtx &c, double timeNow, Work &w) {
        // dark_matter_halos/scales/virial_density_contrast.F90:195-417
        double time = c.timeLastIsolated;
        if (!has(c, GLC_F_IS_SATELLITE) || time <= 0.0) time = timeNow;
        const DeviceTable2D &t = GLC_TABLES.density;
        const double lnt = dm_log(time);
        const double x = (lnt - GLC_TABLES.density_lnt0) * GLC_TABLES.density_inv_dlnt;
        int i = (int)x;
        if (lnt < GLC_TABLES.density_lnt0) i = 0;
        i = max(min(i, t.n0 - 2), 0);
        const double h = x - (double)i;
        w.rhoMean = GLC_LDG(t.v + 2 * i) * (1.0 - h) + GLC_LDG(t.v + 2 * (i + 1)) * h;
        w.dlnrhoDt = GLC_LDG(t.v + 2 * i + 1) * (1.0 - h) + GLC_LDG(t.v + 2 * (i + 1) + 1) * h;
        w.rvir = dm_cbrt(3.0 * c.basicMass / 4.0 / kPi / w.rhoMean);
        w.vvir = sqrt(kGInternal * c.basicMass / w.rvir);
        w.tdyn = w.rvir / w.vvir * kMpcPerKmPerSToGyr;
        w.tvir = 0.5 * kAtomicMassUnit * kMeanAtomicMassPrimordial * ((kKilo * w.vvir) * (kKilo * w.vvir)) / kBoltzmann;
    }

    // ---------------------------------------------------------------- hot halo beta profile
    // betaIsTwoThirds = Values_Agree(beta, 2/3, relTol = 1e-3) (beta_profile.F90:214) selects the closed forms; any other beta
    // goes through I_m(x) = x^(m+1)/(m+1) 2F1((m+1)/2, 3 beta/2; (m+3)/2; -x^2) (glc_specfun.h)
    GLC_DEVICE_INLINE bool beta_is_two_thirds() {
        const double b = GLC_PARAMS.hotHaloBeta, t = 2.0 / 3.0;
        return fabs(b - t) <= 1.0e-3 * 0.5 * (fabs(b) + fabs(t));
    }
    GLC_DEVICE_INLINE double hh_outer_radius(const Work &w, const double (&y)[NY]) {
        // Node_Component_Hot_Halo_Standard_Outer_Radius, hot_halo/standard/_class.F90:430-450
        return fmax(fmin(y[GLC_P_HH_OUTER_RADIUS], w.rvir), GLC_PARAMS.hotHaloScaleRadiusRelative * w.rvir);
    }
    GLC_DEVICE_INLINE void hh_profile(const NodeCtx &c, const double (&y)[NY], Work &w) {
        // hot_halo/mass_distribution/beta_profile.F90:140-215; mass_distributions/spherical/beta_profile.F90:190-301
        const bool hh = has(c, GLC_F_HAS_HOTHALO);
        w.hhRouter = hh ? hh_outer_radius(w, y) : 0.0;
        const double mass = hh ? y[GLC_P_HH_MASS] : 0.0;
        w.hhRcore = GLC_PARAMS.coreRadiusOverVirialRadius * w.rvir;
        w.hhValid = !(w.hhRouter <= 0.0 || mass <= 0.0);
        w.hhRho0 = 0.0;
        if (!w.hhValid) return;
        const double r = w.hhRouter / w.hhRcore;
        if (beta_is_two_thirds()) {
            const double nf = (r < 1.0e-6) ? 3.0 / (r * r * r) + 9.0 / 5.0 / r - 36.0 * r / 175.0 : 1.0 / (r - dm_atan(r));
            w.hhRho0 = mass / 4.0 / kPi / (w.hhRcore * w.hhRcore * w.hhRcore) * nf;
        } else  // :258: 3 M / (4 pi r_outer^3) / 2F1(3/2, 3 beta/2; 5/2; -r^2), with r^3/3 2F1(...) = I_2(r) (glc_specfun.h)
            w.hhRho0 = mass / 4.0 / kPi / (w.hhRcore * w.hhRcore * w.hhRcore) / dm_beta_moment(2, r, GLC_PARAMS.hotHaloBeta);
    }
    GLC_DEVICE_INLINE double hh_density(const Work &w, double radius) {
        if (!w.hhValid || radius > w.hhRouter) return 0.0;
        const double x = radius / w.hhRcore;
        return w.hhRho0 / dm_pow(1.0 + x * x, 1.5 * GLC_PARAMS.hotHaloBeta);
    }
    GLC_DEVICE_INLINE double hh_mass_enclosed(const Work &w, double radius) {
        if (!w.hhValid) return 0.0;
        if (radius > w.hhRouter) radius = w.hhRouter;
        const double x = radius / w.hhRcore;
        const double rc3 = w.hhRcore * w.hhRcore * w.hhRcore;
        if (!beta_is_two_thirds()) return 4.0 * kPi * w.hhRho0 * dm_beta_moment(2, x, GLC_PARAMS.hotHaloBeta) * rc3;  // :425-436
        if (x < 1.0e-6)
            return 4.0 * kPi * w.hhRho0 * rc3 * (x * x * x) * (1.0 / 3.0 + x * x * (-1.0 / 5.0 + x * x * (1.0 / 7.0)));
        return 4.0 * kPi * w.hhRho0 * (x - dm_atan(x)) * rc3;
    }

    // ---------------------------------------------------------------- cooling
    // CIE table look-ups are done once per RHS call at (T_vir, Z_hot) -- the reference memoises the same
    // way (CIE_file.F90:300-312): only n_H varies along the cooling-radius root find.
    GLC_DEVICE_INLINE void cooling_prepare(const double (&y)[NY], Work &w, double &logSlopeT) {
        const double z = mass_to_fraction(y[GLC_P_HH_ABUND], y[GLC_P_HH_MASS]);
        const DeviceTable2D &tc = GLC_TABLES.cooling;
        const DeviceTable2D &te = GLC_TABLES.electron;
        double tu = w.tvir, zu = z / kMetallicitySolar;
        bool outsideT = false;
        if (tu < GLC_TABLES.cooling_t_min) {
            tu = GLC_TABLES.cooling_t_min;
            outsideT = true;
        }
        if (tu > GLC_TABLES.cooling_t_max) {
            tu = GLC_TABLES.cooling_t_max;
            outsideT = true;
        }
        zu = fmin(fmax(zu, GLC_TABLES.cooling_z_min), GLC_TABLES.cooling_z_max);
        const CieFactors f = cie_factors(tc, GLC_TABLES.cooling_log, GLC_TABLES.cooling_first_z_zero,
                                         GLC_TABLES.cooling_first_nonzero_z, tu, zu);
        const double lambda = cie_interpolate(tc, GLC_TABLES.cooling_log, f);
        if (outsideT)
            logSlopeT = 0.0;
        else {
            // cieFileCoolingFunctionTemperatureLogSlope :410-512
            const double *a = tc.v + (size_t)(f.iZ - 1) * tc.n1 + (f.iT - 1);
            const double *b = a + tc.n1;
            double s = ((GLC_LDG(a + 1) - GLC_LDG(a)) * (1.0 - f.hZ) + (GLC_LDG(b + 1) - GLC_LDG(b)) * f.hZ) /
                       (GLC_LDG(tc.x1 + f.iT) - GLC_LDG(tc.x1 + f.iT - 1));
            if (!GLC_TABLES.cooling_log) s = s * w.tvir / lambda;
            logSlopeT = s;
        }
        double te_t = fmin(fmax(w.tvir, GLC_TABLES.electron_t_min), GLC_TABLES.electron_t_max);
        double te_z = fmin(fmax(z / kMetallicitySolar, GLC_TABLES.electron_z_min), GLC_TABLES.electron_z_max);
        const CieFactors fe = cie_factors(te, GLC_TABLES.electron_log, GLC_TABLES.electron_first_z_zero,
                                          GLC_TABLES.electron_first_nonzero_z, te_t, te_z);
        const double efrac = cie_interpolate(te, GLC_TABLES.electron_log, fe);
        w.coolLambda = lambda;
        w.coolEfrac = efrac;
        w.coolXH = hydrogen_mass_fraction(z);
        w.coolFHn = hydrogen_number_fraction(z);
        w.coolTavail = w.tdyn;  // whiteFrenk1991TimeAvailable, ageFactor = 0 (time_available/White-Frenk.F90:144-146)
    }
    GLC_DEVICE_INLINE double cooling_time(const Work &w, double density) {
        // coolingTimeSimple::time, cooling/cooling_time/simple.F90:128-179
        const double timeLarge = 1.0e10;
        const double nh = density * w.coolXH * kMassSolar / kMassHydrogenAtom / (kHecto * kHecto * kHecto) /
                          (kMegaParsec * kMegaParsec * kMegaParsec);
        const double nall = nh / w.coolFHn + w.coolEfrac * nh;
        const double cf = w.coolLambda * nh * nh;
        if (cf > 0.0) {
            const double e = GLC_PARAMS.coolingDegreesOfFreedom / 2.0 * kBoltzmann * w.tvir * nall / kErgs;
            return e / cf / kGigaYear;
        }
        return timeLarge;
    }
    // coolingRadiusSimple::radius, cooling/cooling_radius/simple.F90:313-387: root of t_cool(r) - t_available on
    // [0, r_outer] with the two shortcuts of :363-381
    GLC_DEVICE_INLINE double cooling_function(const Work &w, double radius) {
        return cooling_time(w, hh_density(w, radius)) - w.coolTavail;
    }
    GLC_DEVICE_INLINE void cooling_setup(const Work &w, double &rootOuter, double &rootZero, double &result, bool &need) {
        const double router = w.hhRouter;
        need = false;
        result = 0.0;
        rootZero = 0.0;
        rootOuter = cooling_function(w, router);
        if (rootOuter < 0.0)
            result = router;
        else {
            rootZero = cooling_function(w, 0.0);
            if (rootZero > 0.0)
                result = 0.0;
            else
                need = true;
        }
    }
    GLC_DEVICE_INLINE RootOptions cooling_root_options() {
        return RootOptions{0.0, 1.0e-6, EXPAND_NONE, 0.0, 0.0, SIGN_NONE, SIGN_NONE};
    }
    GLC_DEVICE_INLINE double cooling_radius(const double (&y)[NY], Work &w, int &bad, bool on) {
        // [warp-synchronous]
        double result = 0.0, rootOuter = 0.0, rootZero = 0.0;
        bool need = false;
        if (on) cooling_setup(w, rootOuter, rootZero, result, need);
        const RootOptions o = cooling_root_options();
        int st;
        const double r = root_find(
            [&](double radius) {
                GLC_COUNT(3);
                return cooling_function(w, radius);
            },
            need, o, 0.0, w.hhRouter, true, rootZero, rootOuter, st);
        if (need) {
            result = r;
            if (st != 0) bad = 1;
        }
        return result;
    }

    // ---------------------------------------------------------------- galactic structure
    GLC_DEVICE_INLINE double disk_mass(const double (&y)[NY]) {
        return fmax(0.0, y[GLC_P_DISK_MASS_STELLAR]) + fmax(0.0, y[GLC_P_DISK_MASS_GAS]);
    }
    GLC_DEVICE_INLINE double sph_mass(const double (&y)[NY]) {
        return fmax(0.0, y[GLC_P_SPH_MASS_STELLAR]) + fmax(0.0, y[GLC_P_SPH_MASS_GAS]);
    }
    GLC_DEVICE_INLINE double disk_bessel_factor(double halfRadius) {
        // exponentialDiskBesselFactorRotationCurve, mass_distributions/cylindrical/exponential_disk.F90:675-733
        const double ln2 = 0.69314718055994530942, euler = 0.57721566490153286061;
        if (halfRadius <= 0.0) return 0.0;
        if (halfRadius < 1.0e-3) return (ln2 - euler - 0.5 - dm_log(halfRadius)) * halfRadius * halfRadius;
        const DeviceTable2D &t = GLC_TABLES.diskrc;
        const double x = (dm_log(halfRadius) - GLC_TABLES.diskrc_lnx0) * GLC_TABLES.diskrc_inv_dlnx;
        const int i = max(min((int)x, t.n0 - 2), 0);
        const double h = x - (double)i;
        return GLC_LDG(t.v + i) * (1.0 - h) + GLC_LDG(t.v + i + 1) * h;
    }
    GLC_DEVICE_INLINE double baryonic_vc2(const NodeCtx &c, const double (&y)[NY], const Work &w,
                                                          double radius) {
        // rotation curve of massType=massTypeBaryonic: disk + spheroid + hot halo (gas + stars)
        double v2 = 0.0;
        if (!(radius > 0.0)) return 0.0;
        if (has(c, GLC_F_HAS_DISK) && c.diskRadius > 0.0) {
            const double rd = c.diskRadius, m = disk_mass(y), r = radius / rd;
            if (r > 30.0)
                v2 += kGInternal * m / radius;  // exponential_disk.F90:532-535
            else
                v2 += kGInternal * 2.0 * (m / rd) * disk_bessel_factor(0.5 * r);
        }
        if (has(c, GLC_F_HAS_SPHEROID) && c.sphRadius > 0.0) {
            const double a = c.sphRadius, m = sph_mass(y);
            v2 += kGInternal * m * radius / ((radius + a) * (radius + a));  // Hernquist
        }
        if (has(c, GLC_F_HAS_HOTHALO)) v2 += kGInternal * hh_mass_enclosed(w, radius) / radius;
        return v2;
    }
    GLC_DEVICE_INLINE double nfw_mass_scale_free(double x) {
        // massEnclosedScaleFree, mass_distributions/spherical/NFW.F90:550-571
        if (x == 1.0) return dm_log(2.0) - 0.5;
        if (x >= 1.0e-6) return dm_log(1.0 + x) - x / (1.0 + x);
        return x * x * (0.5 + x * (-2.0 / 3.0 + x * (0.75 + x * (-0.8))));
    }
    // darkMatterProfileDMO.  NFW (mass_distributions/spherical/NFW.F90:254-255,444-464): M(<r) = norm * m(r / r_s) with
    // norm = M_vir / m(c); isothermal (mass_distributions/spherical/isothermal.F90:150-205,287-302; mass = M_vir,
    // lengthReference = r_vir): M(<r) = 4 pi rho_n L^2 r.
    GLC_DEVICE_INLINE bool dmo_isothermal() { return GLC_PARAMS.darkMatterProfileDMO == GLC_DMO_ISOTHERMAL; }
    GLC_DEVICE_INLINE void dmo_prepare(const NodeCtx &c, Work &w) {
        if (dmo_isothermal()) {
            const double L = w.rvir;
            const double rhoN = c.basicMass / 4.0 / kPi / (L * L * L);
            w.dmoNorm = 4.0 * kPi * rhoN * (L * L);
            w.dmoScale = L;
        } else {
            const double conc = w.rvir / c.dmScale;
            w.dmoNorm = c.basicMass / (dm_log(1.0 + conc) - conc / (1.0 + conc));
            w.dmoScale = c.dmScale;
        }
    }
    GLC_DEVICE_INLINE double dmo_mass(double norm, double scale, double radius) {
        if (dmo_isothermal()) return norm * radius;
        return norm * nfw_mass_scale_free(radius / scale);
    }
    // massDistribution_%rotationCurve(radius) of the dark-matter-only profile
    GLC_DEVICE_INLINE double dmo_rotation_curve(const Work &w, double radius) {
        if (dmo_isothermal()) return sqrt(kGInternal * w.dmoNorm);  // velocityRotation, isothermal.F90:196-204
        return (radius > 0.0) ? sqrt(kGInternal * dmo_mass(w.dmoNorm, w.dmoScale, radius) / radius) : 0.0;
    }
    // massDistribution_%radiusFromSpecificAngularMomentum(j).  NFW.F90:589-625: inverse tabulation of the scale-free
    // specific angular momentum sqrt(4 pi m(x) x) on the octave lattice x_k = 2^(k/30), linear interpolation of x in j
    // (numerical/tabulations_inverse.F90:202-245); the lattice is tabulated once on the host over 2^-40 .. 2^40 (the
    // reference grows it octave by octave on demand; values on a pinned lattice do not depend on its extent).
    // isothermal.F90:350-370: r = j / sqrt(4 pi rho_n L^2) / sqrt(G).
    GLC_DEVICE_INLINE double dmo_radius_from_j(const NodeCtx &c, const Work &w, double j) {
        if (!(j > 0.0)) return 0.0;
        if (dmo_isothermal()) return j / sqrt(w.dmoNorm) / sqrt(kGInternal);
        const double rs = c.dmScale, conc = w.rvir / rs;
        const double rhoN = c.basicMass / 4.0 / kPi / (rs * rs * rs) / (dm_log(1.0 + conc) - conc / (1.0 + conc));
        const double jsf = j / sqrt(kGInternal * rhoN) / (rs * rs);
        const double *__restrict__ xs = GLC_TABLES.nfwJx, *__restrict__ js = GLC_TABLES.nfwJv;
        int lo = 0, hi = GLC_TABLES.nfwJN - 1;
        while (hi > lo + 1) {
            const int mid = (hi + lo) >> 1;
            if (GLC_LDG(js + mid) > jsf)
                hi = mid;
            else
                lo = mid;
        }
        const double x = GLC_LDG(xs + lo) + (jsf - GLC_LDG(js + lo)) / (GLC_LDG(js + lo + 1) - GLC_LDG(js + lo)) *
                                                (GLC_LDG(xs + lo + 1) - GLC_LDG(xs + lo));
        return x * rs;
    }
    // `powAc`: where the caller keeps the x^omega table (the micro-task machine stages it in shared memory; same values)
    // (kStaged: the table pointer is not a global-memory address -- plain loads instead of the read-only path)
    template <bool kStaged = false>
    GLC_DEVICE_INLINE double ac_orbital_mean(const Work &w, double radius, const double *__restrict__ powAc = GLC_TABLES.powAc) {
        // sphericalAdiabaticGnedin2004RadiusOrbitalMean, adiabatic_Gnedin2004.F90:664-687
        return GLC_PARAMS.adiabaticA * w.rvir *
               fast_exponentiate<kStaged>(powAc, GLC_TABLES.powAcN, GLC_TABLES.powAcDx, GLC_TABLES.powAcInvDx, 1.0e-3, 1.0,
                                 GLC_PARAMS.adiabaticOmega, radius / w.rvir);
    }
    GLC_DEVICE_INLINE double baryonic_mass_self(const NodeCtx &c, const double (&y)[NY]) {
        double m = 0.0;
        if (has(c, GLC_F_HAS_DISK)) m += disk_mass(y);
        if (has(c, GLC_F_HAS_SPHEROID)) m += sph_mass(y);
        if (has(c, GLC_F_HAS_HOTHALO)) m += fmax(0.0, y[GLC_P_HH_MASS]) + fmax(0.0, y[GLC_P_HH_OUTFLOWED_MASS]);
        return m;
    }
    // adiabaticGnedin2004 over NFW: mass_distributions/spherical/adiabatic_Gnedin2004.F90:410-530,707-727;
    // dark_matter_profiles/adiabatic_Gnedin2004.F90:302-364.  The initial radius r_i of the shell now at `radius`
    // solves  M_i(rbar(r_i)) (f_i r_i - f_d r) - M_b(rbar(r)) rbar(r) r / G... = 0  (:707-727).
    struct AcProblem {
        double fd, fi, bterm, rup, rInit;
        int need;  // 1: r_i must be found by the root finder on [radius, rup]; 0: rInit is final
    };
    template <bool kStaged = false>
    GLC_DEVICE_INLINE double ac_function(double dmoNorm, double dmoScale, const Work &w, const AcProblem &P, double radius,
                                         double ri, const double *__restrict__ powAc = GLC_TABLES.powAc) {
        return dmo_mass(dmoNorm, dmoScale, ac_orbital_mean<kStaged>(w, ri, powAc)) * (P.fi * ri - P.fd * radius) - P.bterm;
    }
    // set-up for a shell inside the virial radius (radius > 0)
    template <bool kStaged = false>
    GLC_DEVICE_INLINE void ac_setup(const NodeCtx &c, const double (&y)[NY], const Work &w, double radius, AcProblem &P,
                                    const double *__restrict__ powAc = GLC_TABLES.powAc) {
        const double nfwNorm = w.dmoNorm, rs = w.dmoScale;
        const double fDm = 1.0 - GLC_PARAMS.OmegaBaryon / GLC_PARAMS.OmegaMatter;
        P.fd = P.fi = P.bterm = 0.0;
        P.rup = radius;
        P.rInit = radius;
        P.need = 0;
        if (radius >= w.rvir) return;
        const double mSelfRaw = baryonic_mass_self(c, y);
        const double mSelf = fmax(mSelfRaw, 0.0);
        const double mTot = fmax(mSelfRaw + c.massBaryonicSubhalos, 0.0);
        P.fd = fmin(fDm + (mTot - mSelf) / c.basicMass, 1.0);
        P.fi = fmin(fDm + mTot / c.basicMass, 1.0);
        const double rmean = ac_orbital_mean<kStaged>(w, radius, powAc);
        P.bterm = baryonic_vc2(c, y, w, rmean) * rmean * radius / kGInternal;
        const double menc = dmo_mass(nfwNorm, rs, rmean);
        if (menc > 0.0) P.rup = fmax((P.bterm / menc + P.fd * radius) / P.fi, radius);
        // the reference first tests solver(r_vir) < 0 (:463-466)
        const double fVir = ac_function<kStaged>(nfwNorm, rs, w, P, radius, w.rvir, powAc);
        if (fVir < 0.0)
            P.rInit = w.rvir;
        else
            P.need = 1;
    }
    GLC_DEVICE_INLINE RootOptions ac_root_options() {
        return RootOptions{0.0, 1.0e-2, EXPAND_MULTIPLICATIVE, 1.1, 0.9, SIGN_POSITIVE, SIGN_NEGATIVE};
    }
    GLC_DEVICE_INLINE double dark_matter_mass_enclosed(const NodeCtx &c, const double (&y)[NY], const Work &w, double radius,
                                                       int &bad, bool on) {
        // [warp-synchronous]
        const double nfwNorm = w.dmoNorm, rs = w.dmoScale;
        const double fDm = 1.0 - GLC_PARAMS.OmegaBaryon / GLC_PARAMS.OmegaMatter;
        if (!GLC_PARAMS.adiabaticContraction) return on ? dmo_mass(nfwNorm, rs, radius) : 0.0;
        const bool live = on && !(radius <= 0.0);
        AcProblem P;
        P.fd = P.fi = P.bterm = 0.0;
        P.rup = P.rInit = radius;
        P.need = 0;
        if (live) ac_setup(c, y, w, radius, P);
        const RootOptions o = ac_root_options();
        int st = 0;
        const double root = root_find([&](double ri) {
                GLC_COUNT(0);
                return ac_function(nfwNorm, rs, w, P, radius, ri);
            },
            P.need != 0, o, radius, P.rup, false, 0.0, 0.0, st);
        if (P.need) {
            P.rInit = root;
            if (st != 0) bad = 1;
        }
        return live ? fDm * dmo_mass(nfwNorm, rs, P.rInit) : 0.0;
    }
    GLC_DEVICE_INLINE void plausibility(const NodeCtx &c, const double (&y)[NY], double time, Work &w) {
        // basic/standard/_class.F90:105-123; disk/standard/_class.F90:999-1049; spheroid/standard/_class.F90:1249-1296
        w.plausible = true;
        w.solvable = true;
        if (c.basicMass <= 0.0 || time <= 0.0) {
            w.plausible = false;
            w.solvable = false;
            return;
        }
        if (has(c, GLC_F_HAS_DISK)) {
            const double m = y[GLC_P_DISK_MASS_STELLAR] + y[GLC_P_DISK_MASS_GAS], j = y[GLC_P_DISK_ANGMOM];
            const double s = m * w.rvir * w.vvir;
            if (m >= 0.0 && j > 0.0 && (j > 1.0e1 * s || j < 1.0e-6 * s)) w.plausible = false;
        }
        if (w.plausible && has(c, GLC_F_HAS_SPHEROID)) {
            const double m = y[GLC_P_SPH_MASS_STELLAR] + y[GLC_P_SPH_MASS_GAS], j = y[GLC_P_SPH_ANGMOM];
            const double s = m * w.rvir * w.vvir;
            if (m >= 0.0 && j > 0.0 && (j > 1.0e1 * s || j < 1.0e-6 * s)) w.plausible = false;
        }
    }
    GLC_DEVICE_INLINE double component_j(const double (&y)[NY], int comp) {
        // disk/standard/_class.F90:1112-1177 (ratio 1/2 for the exponential disk :345-356);
        // spheroid/standard/_class.F90:1359-1407
        const double j = comp == 0 ? y[GLC_P_DISK_ANGMOM] : y[GLC_P_SPH_ANGMOM];
        const double m = comp == 0 ? y[GLC_P_DISK_MASS_GAS] + y[GLC_P_DISK_MASS_STELLAR]
                                   : y[GLC_P_SPH_MASS_GAS] + y[GLC_P_SPH_MASS_STELLAR];
        if (!(j >= 0.0)) return 0.0;
        const double ratio = comp == 0 ? 0.5 : GLC_PARAMS.spheroidRatioAngularMomentumScaleRadius;
        return ratio * ((m > 0.0) ? j / m : 0.0);
    }
    // galacticStructureSolverEquilibrium::solve, galactic_structure/radius_solver/equilibrium.F90:243-506.
    // The pieces of one (iteration, component) visit, shared by the warp-synchronous solver below and by the
    // micro-task machine (glc_machine.cuh).
    //   first pass (:356-404): previous solution, else a first guess from the dark-matter-only rotation curve
    GLC_DEVICE_INLINE void structure_first_pass(const NodeCtx &c, const Work &w, int comp, double j, double &radius,
                                                double &velocity) {
        radius = comp == 0 ? c.diskRadius : c.sphRadius;
        if (radius <= 0.0) {
            const double radiusLarge = 1.0e10;  // galactic_structure/options.F90
            const double jmax = dmo_rotation_curve(w, radiusLarge) * radiusLarge;
            if (jmax < j)
                radius = w.rvir;
            else
                radius = dmo_radius_from_j(c, w, j);
            velocity = dmo_rotation_curve(w, radius);
        } else
            velocity = comp == 0 ? c.diskVelocity : c.sphVelocity;
        if (GLC_PARAMS.structureVelocityMaximumFactor > 0.0)
            velocity = fmin(velocity, GLC_PARAMS.structureVelocityMaximumFactor * w.vvir);
    }
    //   later passes (:406-481): one fixed-point update in the current potential, with the oscillation breaker
    GLC_DEVICE_INLINE void structure_update(const NodeCtx &c, const double (&y)[NY], const Work &w, double j, double mdm,
                                            int count, double &h0, double &h1, double &fit, int &bad, double &radius,
                                            double &velocity) {
        const double vdm2 = kGInternal * mdm / radius;
        const double vb2 = GLC_PARAMS.includeBaryonGravity ? baryonic_vc2(c, y, w, radius) : 0.0;
        velocity = sqrt(vdm2 + vb2);
        if (GLC_PARAMS.structureVelocityMaximumFactor > 0.0)
            velocity = fmin(velocity, GLC_PARAMS.structureVelocityMaximumFactor * w.vvir);  // equilibrium.F90:429-433
        const double radiusNew = (radius > 0.0) ? sqrt(j / velocity * radius) : j / velocity;
        if (count > 10 && h0 >= 0.0 && h1 >= 0.0 && (h1 - h0) * (h0 - radius) < 0.0) {
            switch (count % 4) {
                case 0: radius = sqrt(radius * h0); break;
                case 1: radius = 0.5 * (radius + h0); break;
                case 2: radius = sqrt(h0 * h1); break;
                default: radius = 0.5 * (h0 + h1); break;
            }
            h0 = h1 = -1.0;
        }
        h1 = h0;
        h0 = radius;
        if (radius > 0.0 && radiusNew > 0.0) fit += fabs(dm_log(radiusNew / radius));
        radius = radiusNew;
        if (!(radius > 0.0)) bad = 1;
    }
    GLC_DEVICE_INLINE void structure_store(NodeCtx &c, int comp, double radius, double velocity) {
        if (comp == 0) {
            c.diskRadius = fmax(radius, 0.0);
            c.diskVelocity = velocity;
        } else {
            c.sphRadius = fmax(radius, 0.0);
            c.sphVelocity = velocity;
        }
    }
    // [warp-synchronous: all lanes iterate together, a lane drops out when its own fixed point has converged]
    GLC_DEVICE_INLINE void structure_solve(NodeCtx &c, const double (&y)[NY], double time, Work &w, int &bad, bool on) {
        w.plausible = false;
        w.solvable = false;
        if (on) plausibility(c, y, time, w);
        const double tolerance = GLC_PARAMS.structureSolutionTolerance;
        bool looping = on && w.plausible;
        w.dmoNorm = w.dmoScale = 0.0;
        if (looping) dmo_prepare(c, w);
        double hist00 = -1.0, hist01 = -1.0, hist10 = -1.0, hist11 = -1.0;
        double fit = 2.0 * tolerance;
        int count = 0;
        while (GLC_ANY_OUTER(looping)) {
            int active = 0;
            if (looping) {
                GLC_COUNT(1);
                count++;
                if (count > 1) fit = 0.0;
            }
#pragma unroll 1
            for (int comp = 0; comp < 2; comp++) {
                const bool compOn = looping && has(c, comp == 0 ? GLC_F_HAS_DISK : GLC_F_HAS_SPHEROID);
                const double j = compOn ? component_j(y, comp) : 0.0;
                double radius = 0.0, velocity = 0.0;
                if (compOn) active++;
                const bool first = compOn && count == 1;
                if (first) structure_first_pass(c, w, comp, j, radius, velocity);
                const bool later = compOn && count > 1 && !(j <= 0.0);
                if (later) radius = comp == 0 ? c.diskRadius : c.sphRadius;
                const double mdm = dark_matter_mass_enclosed(c, y, w, radius, bad, later);
                if (later)
                    structure_update(c, y, w, j, mdm, count, comp == 0 ? hist00 : hist10, comp == 0 ? hist01 : hist11, fit,
                                     bad, radius, velocity);
                if (first || later) structure_store(c, comp, radius, velocity);
            }
            if (looping) {
                if (active == 0) {
                    fit = 0.0;
                    looping = false;
                } else {
                    fit /= (double)active;
                    looping = count <= 1 || (fit > tolerance && count < 100);
                }
            }
        }
    }

    // ---------------------------------------------------------------- star formation in disks
    GLC_DEVICE_INLINE double kmt_fh2_fast(double s) {
        return (s < 2.0) ? 1.0 - 0.75 * s / (1.0 + 0.25 * s) : 0.0;  // Krumholz2009.F90:462-476
    }
    struct Kmt {
        double xh, zsolar, sigmaNorm, sNorm, sigmaTrunc, sigma0, rdisk;  // sigma0 = M_gas/(2 pi Rd^2)
    };
    GLC_DEVICE_INLINE double kmt_rate(const Kmt &k, double radius) {
        // krumholz2009Rate :360-414 with the exponential-disk surface density (exponential_disk.F90:484-499)
        const double sg = k.sigma0 * dm_exp(-radius / k.rdisk);
        const double sgd = k.xh * sg / 85.0e12;
        if (sg <= 1.0e-100) return 0.0;
        const double s = k.sNorm / (k.sigmaNorm * sg);
        const double fh2 = (s > 10.0) ? kmt_fh2_fast(s)
                                      : linear_table_eval([](double t) { return kmt_fh2_fast(t); }, 0.0, 10.0, 1000, s, true);
        double factor;
        if (sgd <= 0.0)
            factor = 0.0;
        else
            factor = fast_exponentiate(GLC_TABLES.powKmt, GLC_TABLES.powKmtN, GLC_TABLES.powKmtDx, GLC_TABLES.powKmtInvDx, 1.0,
                                       1000.0, 0.33, (sgd < 1.0) ? 1.0 / sgd : sgd);
        return GLC_PARAMS.frequencyStarFormation * sg * factor * fh2;
    }
    // starFormationRateDisksIntgrtdSurfaceDensity::rate (rates/disks/integrated_surface_density.F90:131-190)
    // with krumholz2009Intervals (rate_surface_density/disks/Krumholz2009.F90:478-587)
    // starFormationRateDisksIntgrtdSurfaceDensity::rate with krumholz2009Intervals: the pieces, shared by the
    // warp-synchronous sfr_disk below and by the micro-task machine
    struct SfrProblem {
        Kmt k;
        double rOut, rMax, sgdIn, sgd;
        int live, needRmax;
    };
    GLC_DEVICE_INLINE double sfr_sigma(const Kmt &k, double r) { return k.sigma0 * dm_exp(-r / k.rdisk); }
    GLC_DEVICE_INLINE void sfr_setup(const NodeCtx &c, const double (&y)[NY], bool on, SfrProblem &S) {
        const double mgas = y[GLC_P_DISK_MASS_GAS], rdisk = c.diskRadius;
        bool live = on && !(mgas <= 0.0 || rdisk <= 0.0);
        Kmt &k = S.k;
        k.xh = k.zsolar = k.sigmaNorm = k.sNorm = k.sigmaTrunc = k.sigma0 = 0.0;
        k.rdisk = 1.0;
        if (live) {
            const double z = mass_to_fraction(y[GLC_P_DISK_ABUND_GAS], mgas);
            k.xh = hydrogen_mass_fraction(z);
            k.zsolar = z / kMetallicitySolar;
            k.rdisk = rdisk;
            k.sigma0 = fmax(0.0, mgas) / (2.0 * kPi * rdisk * rdisk);
            if (!(k.zsolar > 0.0)) live = false;
        }
        if (live) {
            const double chi = 0.77 * (1.0 + 3.1 * dm_pow(k.zsolar, 0.365));
            k.sigmaNorm = k.xh * GLC_PARAMS.clumpingFactorMolecularComplex / (kMega * kMega);
            k.sNorm = dm_log(1.0 + 0.6 * chi + 0.01 * chi * chi) / (0.04 * k.zsolar);
            if (!(k.sigmaNorm > 0.0)) live = false;
        }
        S.rOut = 10.0 * rdisk;
        S.rMax = S.rOut;
        S.sgdIn = S.sgd = 0.0;
        S.needRmax = 0;
        if (live) {
            k.sigmaTrunc = k.sNorm / k.sigmaNorm / GLC_PARAMS.krumholzSTruncation;
            double sg = sfr_sigma(k, 0.0);
            S.sgdIn = k.xh * sg / 85.0e12;
            if (sg <= k.sigmaTrunc)
                live = false;
            else {
                sg = sfr_sigma(k, S.rOut);
                S.sgd = k.xh * sg / 85.0e12;
                S.needRmax = sg <= k.sigmaTrunc;
            }
        }
        S.live = live;
    }
    GLC_DEVICE_INLINE RootOptions sfr_root_options() {
        return RootOptions{0.0, 1.0e-4, EXPAND_MULTIPLICATIVE, 2.0, 0.5, SIGN_NEGATIVE, SIGN_POSITIVE};
    }
    GLC_DEVICE_INLINE double sfr_trunc_function(const Kmt &k, double r) { return sfr_sigma(k, r) - k.sigmaTrunc; }
    GLC_DEVICE_INLINE double sfr_crit_function(const Kmt &k, double r) { return k.xh * sfr_sigma(k, r) / 85.0e12 - 1.0; }
    // after the truncation-radius root find; returns whether the critical-density radius must be found too
    GLC_DEVICE_INLINE bool sfr_after_trunc(SfrProblem &S, double rTrunc, int st, int &bad) {
        if (S.needRmax) {
            S.rMax = rTrunc;
            if (st != 0) bad = 1;
            S.sgd = S.k.xh * sfr_sigma(S.k, S.rMax) / 85.0e12;
        }
        return S.live && !(S.sgdIn <= 1.0 || S.sgd >= 1.0);
    }
    GLC_DEVICE_INLINE double sfr_integrand(const Kmt &k, double r) { return r * kmt_rate(k, r); }
    GLC_DEVICE_INLINE double sfr_disk(const NodeCtx &c, const double (&y)[NY], int &bad, bool on) {
        // [warp-synchronous]
        SfrProblem S;
        sfr_setup(c, y, on, S);
        const Kmt &k = S.k;
        const double rIn = 0.0;
        const RootOptions o = sfr_root_options();
        int st;
        const double rTrunc =
            root_find([&](double r) { return sfr_trunc_function(k, r); }, S.needRmax != 0, o, rIn, S.rOut, false, 0.0, 0.0, st);
        const bool two = sfr_after_trunc(S, rTrunc, st, bad);
        const double rCrit = root_find([&](double r) { return sfr_crit_function(k, r); }, two, o, rIn, S.rMax, false, 0.0, 0.0, st);
        if (two && st != 0) bad = 1;
        double lo[2] = {rIn, rCrit}, hi[2] = {two ? rCrit : S.rMax, S.rMax};
        const int nIv = two ? 2 : 1;
        double total = 0.0;
#pragma unroll 1
        for (int i = 0; i < 2; i++) {
            const bool ion = S.live && i < nIv;
#if defined(__CUDACC__) && !defined(GLC_NO_COOP_QAG)
            const double v = qag15_coop(
                k, [](const Kmt &kk, double r) { return sfr_integrand(kk, r); }, ion, lo[i], hi[i], 1.0e-12,
                GLC_PARAMS.sfrIntegrationTolerance, st);
#else
            const double v = qag15(
                [&](double r) {
                    GLC_COUNT(2);
                    return sfr_integrand(k, r);
                },
                ion, lo[i], hi[i], 1.0e-12, GLC_PARAMS.sfrIntegrationTolerance, st);
#endif
            if (ion) {
                total += v;
                if (st == 11) bad = 1;
            }
        }
        return S.live ? 2.0 * kPi * total : 0.0;
    }
    GLC_DEVICE_INLINE double sfr_spheroid(const NodeCtx &c, const double (&y)[NY]) {
        // rates/spheroids/timescale.F90:109-130 + timescales/dynamical_time.F90:121-189
        const double v = c.sphVelocity, r = c.sphRadius;
        if (v <= 0.0 || GLC_PARAMS.sfSpheroidEfficiency == 0.0) return 0.0;
        const double tau = fmax(kMpcPerKmPerSToGyr * r / v * dm_pow(v / 200.0, GLC_PARAMS.sfSpheroidExponentVelocity) /
                                    GLC_PARAMS.sfSpheroidEfficiency,
                                GLC_PARAMS.sfSpheroidTimescaleMinimum);
        return (tau > 0.0) ? y[GLC_P_SPH_MASS_GAS] / tau : 0.0;
    }

    // ---------------------------------------------------------------- black holes
    // Everything here is straight-line arithmetic once the cooling radius is known (standardHotModeFraction).
    struct Bh {
        bool on;            // a black hole of positive mass exists and a black-hole consumer is enabled
        double mass, spin;
        double eIsco, lIsco;  // ISCO specific energy / angular momentum (gravitational units, prograde)
        double eddington;
        double accSph, accHot, acc;  // blackHoleAccretionRateStandard::rateAccretion
        // one-entry memos keyed on the exact accretion rate (the operators ask for the same quantities at the same rate
        // several times per evaluation: accretion, winds, CGM heating); the functions are pure, so a hit returns the very
        // bits a fresh evaluation would
        double fKey, fVal, effKey, effVal, jetKey, jetVal;
        int fOk, effOk, jetOk;
    };
    GLC_DEVICE_INLINE double ideal_gas_sound_speed(double temperature) {
        // Ideal_Gas_Sound_Speed, thermodynamics/ideal_gases.F90:46-69 (primordial mean atomic mass)
        return sqrt(5.0 * kBoltzmann * temperature / 3.0 / kMeanAtomicMassPrimordial / kAtomicMassUnit) / kKilo;
    }
    GLC_DEVICE_INLINE double bhl_radius(double mass, double temperature) {
        // Bondi_Hoyle_Lyttleton_Accretion_Radius, accretion/Bondi_Hoyle_Lyttleton.F90:61-78
        if (!(temperature > 0.0)) return DBL_MAX;
        const double cs = ideal_gas_sound_speed(temperature);
        return kGInternal * mass / (cs * cs);
    }
    GLC_DEVICE_INLINE double bhl_rate(double mass, double density, double velocity, double temperature, bool withRadius,
                                      double radius) {
        // Bondi_Hoyle_Lyttleton_Accretion_Rate, accretion/Bondi_Hoyle_Lyttleton.F90:34-59
        const double cs = ideal_gas_sound_speed(temperature);
        const double gm = kGInternal * mass;
        if (withRadius)
            return (kKilo * kGigaYear / kMegaParsec) * 4.0 * kPi * (radius * radius) * density *
                   sqrt(cs * cs + velocity * velocity);
        return (kKilo * kGigaYear / kMegaParsec) * 4.0 * kPi * (gm * gm) * density /
               dm_pow(cs * cs + velocity * velocity, 1.5);
    }
    GLC_DEVICE_INLINE double bh_isco_radius(double j) {
        // Black_Hole_ISCO_Radius_Spin (prograde), black_holes/fundamentals.F90:78-121; A1, A2 :553-573
        const double third = 1.0 / 3.0;
        const double a1 = 1.0 + dm_pow(1.0 - j * j, third) * (dm_pow(1.0 + j, third) + dm_pow(1.0 - j, third));
        const double a2 = sqrt(3.0 * (j * j) + a1 * a1);
        return 3.0 + a2 - sqrt((3.0 - a1) * (3.0 + a1 + 2.0 * a2));
    }
    GLC_DEVICE_INLINE double bh_isco_energy(double j, double r) {
        // Black_Hole_ISCO_Specific_Energy_Spin, black_holes/fundamentals.F90:207-233
        if (j >= 0.99999) return 0.5773502693 + 0.9164864242 * dm_pow(1.0 - j, 1.0 / 3.0);
        return (r * r - 2.0 * r + j * sqrt(r)) / r / sqrt(r * r - 3.0 * r + 2.0 * j * sqrt(r));
    }
    GLC_DEVICE_INLINE double bh_isco_angular_momentum(double j, double r) {
        // Black_Hole_ISCO_Specific_Angular_Momentum (gravitational units), black_holes/fundamentals.F90:235-282
        if (j > 0.99999) return 1.154700538 + 1.832972849 * dm_pow(1.0 - j, 1.0 / 3.0);
        return sqrt(r) * (r * r - 2.0 * j * sqrt(r) + j * j) / r / sqrt(r * r - 3.0 * r + 2.0 * j * sqrt(r));
    }
    GLC_DEVICE_INLINE double adaf_table(double spin, int column) {
        // table1DLogarithmicLinear::interpolate, extrapolationTypeFix (objects/tables/_module.F90:1360-1405,1518-1533,
        // 2505-2555) of the ADAF tabulations in 1-j (accretion_disks/ADAF.F90:394-447,481-523); 1-j <= 0 (where the
        // reference would take the logarithm of a non-positive number) is treated as below the table
        const DeviceTable2D &t = GLC_TABLES.adaf;
        const double xinv = 1.0 - spin;
        const double lx0 = GLC_LDG(t.x0), lxn = GLC_LDG(t.x0 + t.n0 - 1);
        double xe = (xinv > 0.0) ? dm_log(xinv) : lx0;
        if (xe < lx0) xe = lx0;
        if (xe > lxn) xe = lxn;
        int i;
        if (xe >= lxn)
            i = t.n0 - 2;
        else {
            i = (int)((xe - lx0) * GLC_TABLES.adaf_inv_dlnx);
            i = max(min(i, t.n0 - 2), 0);
        }
        const double h = (xe - GLC_LDG(t.x0 + i)) * GLC_TABLES.adaf_inv_dlnx;
        return GLC_LDG(t.v + 2 * i + column) * (1.0 - h) + GLC_LDG(t.v + 2 * (i + 1) + column) * h;
    }
    GLC_DEVICE_INLINE double disk_fraction_adaf(Bh &b, double mdot) {
        // switchedFractionADAF, accretion_disks/switched.F90:259-297
        double f = 0.0;
        if (!(b.eddington > 0.0 && mdot > 0.0)) return 0.0;
        if (b.fOk && mdot == b.fKey) return b.fVal;
        const double lm = dm_log(mdot / b.eddington);
        if (GLC_PARAMS.accretionRateThinDiskMinimum > 0.0) {
            const double arg = fmin(+(lm - GLC_TABLES.lnThinDiskMin) / GLC_PARAMS.accretionRateTransitionWidth, 60.0);
            f = f + 1.0 / (1.0 + dm_exp(arg));
        }
        if (GLC_PARAMS.accretionRateThinDiskMaximum < DBL_MAX) {
            const double arg = fmin(-(lm - GLC_TABLES.lnThinDiskMax) / GLC_PARAMS.accretionRateTransitionWidth, 60.0);
            f = f + 1.0 / (1.0 + dm_exp(arg));
        }
        b.fKey = mdot;
        b.fVal = f;
        b.fOk = 1;
        return f;
    }
    GLC_DEVICE_INLINE double disk_efficiency_radiative(Bh &b, double mdot) {
        // switchedEfficiencyRadiative :199-226 over shakuraSunyaevEfficiencyRadiative (Shakura_Sunyaev.F90:69-93) and
        // adafEfficiencyRadiative (ADAF.F90:453-479) with switchedEfficiencyRadiativeScalingADAF :299-331
        if (b.effOk && mdot == b.effKey) return b.effVal;
        const double f = disk_fraction_adaf(b, mdot);
        const double effThin = 1.0 - b.eIsco;
        double effAdaf = GLC_PARAMS.adafEfficiencyRadiationTypeThinDisk ? effThin : GLC_PARAMS.adafEfficiencyRadiation;
        if (GLC_PARAMS.scaleADAFRadiativeEfficiency) {
            double scaling = 1.0;
            if (b.eddington > 0.0 && mdot > 0.0) {
                const double md = mdot / b.eddington;
                if (GLC_PARAMS.accretionRateThinDiskMinimum > 0.0 && md < GLC_PARAMS.accretionRateThinDiskMinimum)
                    scaling = md / GLC_PARAMS.accretionRateThinDiskMinimum;
            }
            effAdaf = effAdaf * scaling;
        }
        double eff = 0.0;
        eff = eff + f * effAdaf;
        eff = eff + (1.0 - f) * effThin;
        b.effKey = mdot;
        b.effVal = eff;
        b.effOk = 1;
        return eff;
    }
    GLC_DEVICE_INLINE double disk_power_jet(Bh &b, double mdot) {
        // switchedPowerJet :228-242; shakuraSunyaevPowerJet (Shakura_Sunyaev.F90:95-155); adafPowerJet (ADAF.F90:481-499).
        // 10**42.7 and 10**41.7 are compile-time constants of the reference.
        const double normKerr = 5.011872336272756e+42 * kErgs * kGigaYear / kMassSolar / (kKilo * kKilo);
        const double normSchw = 5.011872336272755e+41 * kErgs * kGigaYear / kMassSolar / (kKilo * kKilo);
        if (b.jetOk && mdot == b.jetKey) return b.jetVal;
        const double f = disk_fraction_adaf(b, mdot);
        double thin = 0.0;
        if (mdot > 0.0) {
            const double md = mdot / b.eddington, mb = b.mass / 1.0e9;
            if (mb > 0.0 && md > 0.0) {
                if (b.spin > 0.8)
                    thin = normKerr * dm_pow(mb, 0.9) * dm_pow(md, 1.2) / 1.0 * (1.0 + 1.1 * b.spin + 0.29 * (b.spin * b.spin));
                else
                    thin = normSchw * dm_pow(mb, 0.9) * dm_pow(md, 1.2) / 1.0 * dm_exp(3.785 * b.spin);
            }
        }
        const double adaf = mdot * adaf_table(b.spin, 0);
        const double power = (1.0 - f) * thin + f * adaf;
        b.jetKey = mdot;
        b.jetVal = power;
        b.jetOk = 1;
        return power;
    }
    GLC_DEVICE_INLINE double disk_rate_spin_up(Bh &b, double mdot) {
        // switchedRateSpinUp :244-257; shakuraSunyaevRateSpinUp (Shakura_Sunyaev.F90:157-179); adafRateSpinUp (ADAF.F90:501-523)
        const double f = disk_fraction_adaf(b, mdot);
        double thin = 0.0;
        if (mdot != 0.0) thin = (b.lIsco - 2.0 * b.spin * b.eIsco) * mdot / b.mass;
        const double adaf = adaf_table(b.spin, 1) * mdot / b.mass;
        return (1.0 - f) * thin + f * adaf;
    }
    GLC_DEVICE_INLINE double sph_gas_density(const NodeCtx &c, const double (&y)[NY], double radius) {
        // node%massDistribution(spheroid, gaseous)%density: Hernquist profile inside a spherical scaler
        // (spheroid/standard/bound_functions.Inc; mass_distributions/spherical/{scaler,Hernquist}.F90)
        if (!has(c, GLC_F_HAS_SPHEROID)) return 0.0;
        const double a = c.sphRadius;
        const double m = fmax(0.0, y[GLC_P_SPH_MASS_GAS]);
        if (a <= 0.0 || !(m > 0.0)) return 0.0;
        const double x = radius * (1.0 / a);
        return 0.5 / kPi / x / ((1.0 + x) * (1.0 + x) * (1.0 + x)) * m / (a * a * a);
    }
    GLC_DEVICE_INLINE bool bh_on(const NodeCtx &c, const double (&y)[NY], bool go) {
        return go && has(c, GLC_F_HAS_BH) && y[GLC_P_BH_MASS] > 0.0 &&
               (GLC_PARAMS.operatorMask & (GLC_OP_BLACK_HOLES_ACCRETION | GLC_OP_BLACK_HOLES_WINDS | GLC_OP_CGM_COOLING_HEATING));
    }
    // standardHotModeFraction (accretion_rates/standard.F90:442-466) asks for the cooling radius whenever the hot-halo
    // density can be non-zero
    GLC_DEVICE_INLINE bool bh_radius_on(const NodeCtx &c, const double (&y)[NY], const Work &w, bool go) {
        return bh_on(c, y, go) && w.hhValid && GLC_PARAMS.bondiHoyleAccretionHotModeOnly;
    }
    GLC_DEVICE_INLINE void bh_accretion(const NodeCtx &c, const double (&y)[NY], const Work &w, bool go, double rcool, Bh &b) {
        // blackHoleAccretionRateStandard::rateAccretion, black_holes/accretion_rates/standard.F90:221-440 (no nuclear
        // star cluster; blackHoleBinarySeparationGrowthRate "zero" => velocityRelative = 0; radialPosition = 0 for the
        // central black hole; cold mode not tracked)
        const double densityGasMinimum = 1.0;
        const double velocity = 0.0 * kMpcPerKmPerSToGyr;
        b.on = bh_on(c, y, go);
        b.mass = b.spin = b.eIsco = b.lIsco = b.eddington = b.accSph = b.accHot = b.acc = 0.0;
        b.fKey = b.fVal = b.effKey = b.effVal = b.jetKey = b.jetVal = 0.0;
        b.fOk = b.effOk = b.jetOk = 0;
        if (!b.on) return;
        b.mass = y[GLC_P_BH_MASS];
        b.spin = y[GLC_P_BH_SPIN];
        // a trial RK stage can carry the spin out of the range in which the Kerr expressions of the reference are finite
        // (j > 1: cube roots of negative numbers; j < ~-0.63: negative radicand in the ISCO energy of the prograde formula).
        // The ISCO quantities are evaluated at the spin clamped to [-0.5, 1]; states reached by accepted steps are in
        // [0, 0.9999] (post-step clamp), so this only replaces NaNs of rejected trial stages.
        const double jk = fmax(fmin(b.spin, 1.0), -0.5);
        const double riso = bh_isco_radius(jk);
        b.eIsco = bh_isco_energy(jk, riso);
        b.lIsco = bh_isco_angular_momentum(jk, riso);
        // Black_Hole_Eddington_Accretion_Rate, fundamentals.F90:123-138
        b.eddington = 4.0 * kPi * kGravitationalConstant * b.mass * kMassHydrogenAtom * kGigaYear / kThomsonCrossSection / kSpeedLight;
        // spheroid
        double rAcc = fmax(bhl_radius(b.mass, GLC_PARAMS.bondiHoyleAccretionTemperatureSpheroid), 0.0);
        double rho = sph_gas_density(c, y, rAcc);
        if (rho > densityGasMinimum) {
            double lj = ideal_gas_sound_speed(GLC_PARAMS.bondiHoyleAccretionTemperatureSpheroid) / sqrt(kGInternal) / sqrt(rho);
            lj = fmin(lj, c.sphRadius);
            if (lj > rAcc) rho = sph_gas_density(c, y, lj);
            b.accSph = fmax(GLC_PARAMS.bondiHoyleAccretionEnhancementSpheroid *
                                bhl_rate(b.mass, rho, velocity, GLC_PARAMS.bondiHoyleAccretionTemperatureSpheroid, false, 0.0),
                            0.0);
            const double eff = disk_efficiency_radiative(b, b.accSph);
            if (eff > 0.0) b.accSph = fmin(b.accSph, b.eddington / eff);
        }
        // hot halo
        if (w.hhValid) {
            const double tHot = w.tvir;  // hotHaloTemperatureProfile virial
            double fHot = 1.0;
            rAcc = bhl_radius(b.mass, tHot);
            rAcc = fmin(rAcc, w.hhRouter);
            if (GLC_PARAMS.bondiHoyleAccretionHotModeOnly) {
                const double xf = rcool / w.rvir;
                if (xf < 0.9)
                    fHot = 1.0;
                else if (xf > 1.0)
                    fHot = 0.0;
                else {
                    const double x = (xf - 0.9) / (1.0 - 0.9);
                    fHot = x * x * (2.0 * x - 3.0) + 1.0;
                }
            }
            rho = fHot * hh_density(w, rAcc);
            if (rho > densityGasMinimum) {
                b.accHot = fmax(GLC_PARAMS.bondiHoyleAccretionEnhancementHotHalo * bhl_rate(b.mass, rho, velocity, tHot, true, rAcc), 0.0);
                const double rateMax = fmax(y[GLC_P_HH_MASS] / (w.hhRouter / ideal_gas_sound_speed(tHot) * kMpcPerKmPerSToGyr), 0.0);
                b.accHot = fmin(b.accHot, rateMax);
                const double eff = disk_efficiency_radiative(b, b.accHot);
                if (eff > 0.0) b.accHot = fmin(b.accHot, b.eddington / eff);
            }
        }
        b.acc = b.accSph + b.accHot;
    }
    GLC_DEVICE_INLINE double bh_wind_power(const NodeCtx &c, const double (&y)[NY], Bh &b) {
        // blackHoleWindCiotti2009::power, black_holes/winds/Ciotti2009.F90:153-246
        const double velocityWind = 1.0e4, temperatureISM = 1.0e4;
        double eff = GLC_PARAMS.bhEfficiencyWind, coupled = 0.0;
        if (GLC_PARAMS.bhEfficiencyWindScalesWithEfficiencyRadiative) eff = eff * disk_efficiency_radiative(b, b.acc);
        if (b.acc <= 0.0 || eff <= 0.0) return 0.0;
        const double mgas = has(c, GLC_F_HAS_SPHEROID) ? y[GLC_P_SPH_MASS_GAS] : 0.0;
        if (mgas > 0.0) {
            const double rs = c.sphRadius;
            if (rs > 0.0) {
                const double c2 = kSpeedLight * kSpeedLight;
                const double pWind = eff * b.acc * kMassSolar / kGigaYear * c2 / 4.0 / kPi / velocityWind / kKilo / (rs * rs) /
                                     (kMegaParsec * kMegaParsec);
                const double pIsm = 3.0 / 4.0 / kPi * mgas * kMassSolar / (rs * rs * rs) / (kMegaParsec * kMegaParsec * kMegaParsec) /
                                    kMassHydrogenAtom * 3.0 / 2.0 * kBoltzmann * temperatureISM;
                const double x = pIsm / pWind - 0.50;
                if (x <= 0.0)
                    coupled = 0.0;
                else if (x >= 1.0)
                    coupled = 1.0;
                else
                    coupled = 3.0 * (x * x) - 2.0 * (x * x * x);
            }
        }
        eff = eff * coupled;
        return eff * b.acc * (kSpeedLight * kSpeedLight) / (kKilo * kKilo);
    }

    // ---------------------------------------------------------------- scales (scaleSetTask hooks)
    GLC_DEVICE_INLINE void scales(const NodeCtx &c, const double (&y)[NY], double (&s)[NY]) {
        Work w;
        halo_scales(c, c.timeNode, w);
        s[GLC_P_SAT_BOUND_MASS] = 1.0e-6 * c.basicMass;  // satellite/standard.F90:209-232
        if (c.flags & (GLC_F_HAS_DISK | GLC_F_HAS_SPHEROID)) {
            // disk/standard/_class.F90:735-824 and spheroid/standard/_class.F90:804-897
            const bool hd = has(c, GLC_F_HAS_DISK), hs = has(c, GLC_F_HAS_SPHEROID);
            const double jd = hd ? y[GLC_P_DISK_ANGMOM] : 0.0, js = hs ? y[GLC_P_SPH_ANGMOM] : 0.0;
            const double mgd = hd ? y[GLC_P_DISK_MASS_GAS] : 0.0, msd = hd ? y[GLC_P_DISK_MASS_STELLAR] : 0.0;
            const double mgs = hs ? y[GLC_P_SPH_MASS_GAS] : 0.0, mss = hs ? y[GLC_P_SPH_MASS_STELLAR] : 0.0;
            const double zgd = hd ? y[GLC_P_DISK_ABUND_GAS] : 0.0, zsd = hd ? y[GLC_P_DISK_ABUND_STELLAR] : 0.0;
            const double zgs = hs ? y[GLC_P_SPH_ABUND_GAS] : 0.0, zss = hs ? y[GLC_P_SPH_ABUND_STELLAR] : 0.0;
            const double sj = fmax(fabs(jd) + fabs(js), 0.1);
            const double sm = fmax(fabs(mgd) + fabs(mgs) + fabs(msd) + fabs(mss), 1.0);
            const double sz = fmax(fabs(zgd) + fabs(zsd) + fabs(zgs) + fabs(zss), fmax(sm * 1.0e-4, 1.0));
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
        if (has(c, GLC_F_HAS_HOTHALO)) {
            // hot_halo/standard/_class.F90:804-850
            const double sm = c.basicMass * GLC_PARAMS.hotHaloScaleMassRelative;
            const double sj = c.basicMass * w.rvir * w.vvir * GLC_PARAMS.hotHaloScaleMassRelative;
            s[GLC_P_HH_MASS] = s[GLC_P_HH_OUTFLOWED_MASS] = s[GLC_P_HH_UNACCRETED_MASS] = sm;
            s[GLC_P_HH_ABUND] = s[GLC_P_HH_UNACCRETED_ABUND] = s[GLC_P_HH_OUTFLOWED_ABUND] = sm;
            s[GLC_P_HH_ANGMOM] = s[GLC_P_HH_OUTFLOWED_ANGMOM] = sj;
            s[GLC_P_HH_OUTER_RADIUS] = w.rvir * GLC_PARAMS.hotHaloScaleRadiusRelative;
            const double ss = has(c, GLC_F_IS_SATELLITE) ? sm : 1.0;
            s[GLC_P_HH_STRIPPED_MASS] = s[GLC_P_HH_STRIPPED_ABUND] = ss;
        }
        if (has(c, GLC_F_HAS_BH)) {
            // Node_Component_Black_Hole_Standard_Scale_Set, black_hole/standard/_class.F90:193-257 (no nuclear star cluster)
            double mstar = has(c, GLC_F_HAS_SPHEROID) ? y[GLC_P_SPH_MASS_STELLAR] : 0.0;
            if (!(mstar > 0.0)) mstar = 0.0;
            s[GLC_P_BH_MASS] = fmax(fmax(1.0, 1.0e-4 * mstar), y[GLC_P_BH_MASS]);
            s[GLC_P_BH_SPIN] = 1.0;
        }
    }

    GLC_DEVICE_INLINE void pre_evolve(NodeCtx &c, double (&y)[NY]) {
        // Node_Component_Hot_Halo_Standard_Pre_Evolve -> Initializor (hot_halo/standard/_class.F90:725-747,871-891)
        if (has(c, GLC_F_HAS_HOTHALO) && !has(c, GLC_F_HH_INITIALIZED)) {
            Work w;
            solve_analytics(c, c.timeNode);
            halo_scales(c, c.timeNode, w);
            y[GLC_P_HH_OUTER_RADIUS] = w.rvir;
            c.flags |= GLC_F_HH_INITIALIZED;
        }
    }

    // ---------------------------------------------------------------- the RHS
    // rate accumulation with the semantics of the generated <prop>Rate functions
    // (python/Galacticus/Build/Components/Properties/Evolve.py:202-493)
    struct Acc {
        int flags, interrupt;
        GLC_DEVICE_METHOD void add(double &slot, int compFlag, double v) const {
            if (flags & compFlag) slot += v;
        }
        GLC_DEVICE_METHOD void addCreate(double &slot, int compFlag, int code, double v) {
            if (flags & compFlag)
                slot += v;
            else if (v != 0.0)
                interrupt = code;
        }
    };

    GLC_DEVICE_INLINE void hh_outflowing(Acc &a, const NodeCtx &c, const Work &w, double (&rate)[NY],
                                                         double mass, double angmom, double abund) {
        // hot_halo/standard/_class.F90:598-722 with hotHaloOutflowStrippingStandard (outflow_stripping/standard.F90:133-173)
        if (!has(c, GLC_F_HAS_HOTHALO)) return;
        double fs = 0.0;
        if (has(c, GLC_F_IS_SATELLITE)) {
            const double mo = hh_mass_enclosed(w, w.hhRouter), mv = hh_mass_enclosed(w, w.rvir);
            fs = (mv > 0.0) ? GLC_PARAMS.outflowStrippingEfficiency * (1.0 - mo / mv) : GLC_PARAMS.outflowStrippingEfficiency;
        }
        rate[GLC_P_HH_STRIPPED_MASS] += mass * fs;
        rate[GLC_P_HH_OUTFLOWED_MASS] += mass * (1.0 - fs);
        rate[GLC_P_HH_OUTFLOWED_ANGMOM] += angmom * (1.0 - fs) / (1.0 - GLC_PARAMS.fractionLossAngularMomentum);
        rate[GLC_P_HH_OUTFLOWED_ABUND] += abund * (1.0 - fs);
        (void)a;
    }

    template <bool IS_DISK>
    GLC_DEVICE_INLINE void star_formation_and_feedback(Acc &a, const NodeCtx &c, const Work &w,
                                                                       const double (&y)[NY], double (&rate)[NY],
                                                                       double psi, bool doSf, bool doFb) {
        constexpr int pm = IS_DISK ? GLC_P_DISK_MASS_GAS : GLC_P_SPH_MASS_GAS;
        constexpr int pa = IS_DISK ? GLC_P_DISK_ABUND_GAS : GLC_P_SPH_ABUND_GAS;
        constexpr int ps = IS_DISK ? GLC_P_DISK_MASS_STELLAR : GLC_P_SPH_MASS_STELLAR;
        constexpr int pz = IS_DISK ? GLC_P_DISK_ABUND_STELLAR : GLC_P_SPH_ABUND_STELLAR;
        constexpr int pj = IS_DISK ? GLC_P_DISK_ANGMOM : GLC_P_SPH_ANGMOM;
        const double massGas = y[pm];
        const double z = mass_to_fraction(y[pa], massGas);
        if (doSf) {
            // instantaneousRates, stellar_populations/properties/instantaneous.F90:173-182
            const double rateStellar = (1.0 - GLC_PARAMS.recycledFraction) * psi;
            const double rateZStellar = z * rateStellar;
            rate[ps] += rateStellar;
            rate[pm] += -rateStellar;
            rate[pz] += rateZStellar;
            rate[pa] += -rateZStellar + GLC_PARAMS.metalYield * psi;
        }
        if (doFb) {
            // stellar_feedback/{disks,spheroids}.F90:116-206; outflows/power_law/_class.F90:118-164;
            // outflows/rate_limit.F90:111-170
            const double radius = IS_DISK ? c.diskRadius : c.sphRadius;
            const double velocity = IS_DISK ? c.diskVelocity : c.sphVelocity;
            const double vchar = IS_DISK ? GLC_PARAMS.fbDiskVelocityCharacteristic : GLC_PARAMS.fbSpheroidVelocityCharacteristic;
            const double expo = IS_DISK ? GLC_PARAMS.fbDiskExponent : GLC_PARAMS.fbSpheroidExponent;
            const double energy = kFeedbackEnergyInputAtInfinityCanonical * psi;
            double outflow = (velocity <= 0.0) ? 0.0 : dm_pow(vchar / velocity, expo) * energy / kFeedbackEnergyInputAtInfinityCanonical;
            const double tdyn = (velocity <= 0.0 || radius <= 0.0) ? 1.0 : kMpcPerKmPerSToGyr * radius / velocity;
            const double outflowMax = fmax(massGas / tdyn / GLC_PARAMS.fbTimescaleOutflowFractionalMinimum, 0.0);
            if (outflow > outflowMax) outflow = outflow * outflowMax / outflow;
            if (outflow > 0.0) {
                const double massComp = massGas + y[ps];
                const double jOut = (massComp > 0.0) ? y[pj] * (outflow / massComp) : 0.0;
                const double zOut = (massGas > 0.0) ? z * outflow : 0.0;
                hh_outflowing(a, c, w, rate, outflow, jOut, zOut);
                rate[pm] += -outflow;
                rate[pj] += -jOut;
                rate[pa] += -zOut;
            }
        }
    }

    static constexpr bool kHasPostEvolve = true;

    GLC_DEVICE_INLINE void work_clear(Work &w) {
        w.hhRouter = w.hhRcore = w.hhRho0 = w.rvir = w.vvir = w.tdyn = w.tvir = w.rhoMean = w.dlnrhoDt = 0.0;
        w.hhValid = false;
        w.coolLambda = w.coolEfrac = w.coolXH = w.coolFHn = w.coolTavail = 0.0;
        w.plausible = w.solvable = false;
        w.dmoNorm = w.dmoScale = 0.0;
    }
    // which of the nested solvers an evaluation needs (operator gates of nodeOperatorMulti)
    GLC_DEVICE_INLINE bool disk_sfr_on(const NodeCtx &c, const double (&y)[NY], bool go) {
        return go && has(c, GLC_F_HAS_DISK) &&
               (GLC_PARAMS.operatorMask & (GLC_OP_STAR_FORMATION_DISKS | GLC_OP_STELLAR_FEEDBACK_DISKS)) &&
               !(y[GLC_P_DISK_ANGMOM] < 0.0 || c.diskRadius < 0.0 || y[GLC_P_DISK_MASS_GAS] < 0.0);
    }
    GLC_DEVICE_INLINE bool cooling_on(const NodeCtx &c, const double (&y)[NY], const Work &w, bool go) {
        return go && (GLC_PARAMS.operatorMask & GLC_OP_CGM_COOLING_HEATING) && has(c, GLC_F_HAS_HOTHALO) &&
               y[GLC_P_HH_MASS] > 0.0 && !(y[GLC_P_HH_ANGMOM] <= 0.0 || w.hhRouter <= 0.0);
    }
    // coolingRateWhiteFrenk1991::rate needs the cooling radius unless the halo is above the velocity cut-off
    GLC_DEVICE_INLINE bool cooling_rate_needs_radius(const Work &w, bool coolOn) {
        return coolOn && !(w.vvir > GLC_PARAMS.coolingVelocityCutOff);
    }
    // ... and so does the hot-mode fraction of the black-hole accretion rate
    GLC_DEVICE_INLINE bool cooling_radius_on(const NodeCtx &c, const double (&y)[NY], const Work &w, bool coolOn, bool go) {
        return cooling_rate_needs_radius(w, coolOn) || bh_radius_on(c, y, w, go);
    }

    // Everything of the RHS that is straight-line once the nested solvers have delivered the disk star formation
    // rate (psiDisk) and the cooling radius (rinfallSolved): accumulates the operators' rates in the order of
    // nodeOperatorMulti (multi.F90:313-332) and returns the interrupt code.
    GLC_DEVICE_INLINE int rates_accumulate(NodeCtx &c, double time, const double (&y)[NY], double (&rate)[NY], const Work &w,
                                           bool go, bool dOn, double psiDisk, bool coolOn, bool radiusOn,
                                           double rinfallSolved, double logSlopeT, int bad) {
        Acc a{c.flags, GLC_INT_NONE};
        const uint32_t ops = GLC_PARAMS.operatorMask;
        const bool hh = has(c, GLC_F_HAS_HOTHALO), hd = has(c, GLC_F_HAS_DISK), hs = has(c, GLC_F_HAS_SPHEROID);
        const bool sat = has(c, GLC_F_IS_SATELLITE);
        const bool coolRadiusOn = cooling_rate_needs_radius(w, coolOn);  // radiusOn may also be set for the black hole
        // satelliteMassLoss (satellite/mass_loss/_class.F90:230-257; darkMatterHaloMassLossRate "zero")
        if (go && (ops & GLC_OP_SATELLITE_MASS_LOSS)) rate[GLC_P_SAT_BOUND_MASS] += sat ? 0.0 : c.massRate;

        // star formation + stellar feedback, disks (star_formation/disks.F90:200-284, stellar_feedback/disks.F90:116-206)
        if (dOn)
            star_formation_and_feedback<true>(a, c, w, y, rate, psiDisk, (ops & GLC_OP_STAR_FORMATION_DISKS) != 0,
                                              (ops & GLC_OP_STELLAR_FEEDBACK_DISKS) != 0);
        // spheroids (star_formation/spheroids.F90:152-240, stellar_feedback/spheroids.F90:116-208)
        if (go && hs && (ops & (GLC_OP_STAR_FORMATION_SPHEROIDS | GLC_OP_STELLAR_FEEDBACK_SPHEROIDS)) &&
            !(y[GLC_P_SPH_ANGMOM] < 1.0e-20 || c.sphRadius < 1.0e-12 || y[GLC_P_SPH_MASS_GAS] < 1.0e-6)) {
            const double psi = sfr_spheroid(c, y);
            star_formation_and_feedback<false>(a, c, w, y, rate, psi, (ops & GLC_OP_STAR_FORMATION_SPHEROIDS) != 0,
                                               (ops & GLC_OP_STELLAR_FEEDBACK_SPHEROIDS) != 0);
        }

        // barInstability (bar_instability.F90:145-249; galactic_dynamics/bar_instability/Efstathiou1982.F90:153-258)
        if (go && (ops & GLC_OP_BAR_INSTABILITY) && hd &&
            !(y[GLC_P_DISK_ANGMOM] < 0.0 || c.diskRadius < 0.0 || y[GLC_P_DISK_MASS_GAS] < 0.0)) {
            double timescale = -1.0;
            if (w.plausible && y[GLC_P_DISK_ANGMOM] > 0.0 && c.diskVelocity > 0.0 && c.diskRadius > 0.0) {
                const double stabilityIsolated = 0.6221297315, boost = 1.1800237580;
                const double md = y[GLC_P_DISK_MASS_GAS] + y[GLC_P_DISK_MASS_STELLAR];
                const double fgas = y[GLC_P_DISK_MASS_GAS] / md;
                const double thr = GLC_PARAMS.barStabilityThresholdStellar * (1.0 - fgas) + GLC_PARAMS.barStabilityThresholdGaseous * fgas;
                double est = DBL_MAX;
                if (md >= 0.0) {
                    const double vself = sqrt(kGInternal * md / c.diskRadius);
                    if (vself > 0.0) est = fmax(stabilityIsolated, boost * c.diskVelocity / vself);
                }
                if (est < thr) {
                    const double tdyn = kMpcPerKmPerSToGyr * c.diskRadius / fmin(c.diskVelocity, kSpeedLight / kKilo);
                    const double aa = thr - stabilityIsolated, bb = thr - est;
                    const double tdim = (aa > 1.0e10 * bb) ? 1.0e10 : (aa / bb) * (aa / bb);
                    timescale = fmax(tdyn, 1.0e-9) * tdim;
                }
            }
            if (!(timescale < 0.0)) {
                double tr = fmax(0.0, y[GLC_P_DISK_MASS_GAS]) / timescale;
                rate[GLC_P_DISK_MASS_GAS] += -tr;
                a.addCreate(rate[GLC_P_SPH_MASS_GAS], GLC_F_HAS_SPHEROID, GLC_INT_SPHEROID_CREATE, tr);
                tr = fmax(0.0, y[GLC_P_DISK_MASS_STELLAR]) / timescale;
                rate[GLC_P_DISK_MASS_STELLAR] += -tr;
                a.addCreate(rate[GLC_P_SPH_MASS_STELLAR], GLC_F_HAS_SPHEROID, GLC_INT_SPHEROID_CREATE, tr);
                tr = fmax(0.0, y[GLC_P_DISK_ANGMOM]) / timescale;
                rate[GLC_P_DISK_ANGMOM] += -(1.0 - 1.0) * tr;  // fractionAngularMomentumRetainedDisk = 1
                a.addCreate(rate[GLC_P_SPH_ANGMOM], GLC_F_HAS_SPHEROID, GLC_INT_SPHEROID_CREATE, 1.0 * tr);
                tr = fmax(0.0, y[GLC_P_DISK_ABUND_GAS]) / timescale;
                rate[GLC_P_DISK_ABUND_GAS] += -tr;
                a.addCreate(rate[GLC_P_SPH_ABUND_GAS], GLC_F_HAS_SPHEROID, GLC_INT_SPHEROID_CREATE, tr);
                tr = fmax(0.0, y[GLC_P_DISK_ABUND_STELLAR]) / timescale;
                rate[GLC_P_DISK_ABUND_STELLAR] += -tr;
                a.addCreate(rate[GLC_P_SPH_ABUND_STELLAR], GLC_F_HAS_SPHEROID, GLC_INT_SPHEROID_CREATE, tr);
            }
        }

        // blackHolesSeed (black_holes/seed.F90:153-182, blackHoleSeeds fixed): create the seed by interrupt
        if (go && (ops & GLC_OP_BLACK_HOLES_SEED) && !has(c, GLC_F_HAS_BH) && GLC_PARAMS.bhSeedMass > 0.0)
            a.interrupt = GLC_INT_BH_CREATE;
        Bh bh;
        bh_accretion(c, y, w, go, radiusOn ? rinfallSolved : 0.0, bh);
        // blackHolesAccretion (black_holes/accretion.F90:113-173)
        if (bh.on && (ops & GLC_OP_BLACK_HOLES_ACCRETION) && bh.acc > 0.0) {
            const double effRad = disk_efficiency_radiative(bh, bh.acc);
            const double effJet = disk_power_jet(bh, bh.acc) / bh.acc / (kSpeedLight * kSpeedLight) / (kKilo * kKilo);
            const double reduced = bh.acc * (1.0 - effRad - effJet);
            const double spinUp = disk_rate_spin_up(bh, bh.acc);
            rate[GLC_P_BH_MASS] += reduced;
            // Node_Component_Spheroid_Standard_Mass_Gas_Sink_Rate, spheroid/standard/_class.F90:638-671
            if (hs && -bh.accSph != 0.0) {
                const double mg = y[GLC_P_SPH_MASS_GAS], ms = y[GLC_P_SPH_MASS_STELLAR], r = -bh.accSph;
                if (mg > 0.0 && mg + ms > 0.0) {
                    rate[GLC_P_SPH_MASS_GAS] += r;
                    rate[GLC_P_SPH_ANGMOM] += (r / (mg + ms)) * y[GLC_P_SPH_ANGMOM];
                    rate[GLC_P_SPH_ABUND_GAS] += (r / mg) * y[GLC_P_SPH_ABUND_GAS];
                }
            }
            // Node_Component_Hot_Halo_Standard_Mass_Sink -> Hot_Gas_All_Rate, hot_halo/standard/_class.F90:749-800
            if (hh && -bh.accHot != 0.0) {
                const double mg = y[GLC_P_HH_MASS], r = -bh.accHot;
                if (mg > 0.0) {
                    rate[GLC_P_HH_MASS] += r;
                    rate[GLC_P_HH_ANGMOM] += y[GLC_P_HH_ANGMOM] * (r / mg);
                    rate[GLC_P_HH_ABUND] += y[GLC_P_HH_ABUND] * (r / mg);
                }
            }
            rate[GLC_P_BH_SPIN] += spinUp;
        }
        // blackHolesWinds (black_holes/winds.F90:103-134) -> Node_Component_Spheroid_Standard_Energy_Gas_Input_Rate,
        // spheroid/standard/_class.F90:673-725
        if (bh.on && (ops & GLC_OP_BLACK_HOLES_WINDS)) {
            const double power = bh_wind_power(c, y, bh);
            if (power != 0.0 && hs) {
                const double mg = y[GLC_P_SPH_MASS_GAS], ms = y[GLC_P_SPH_MASS_STELLAR], vs = c.sphVelocity;
                if (mg > 0.0 && mg + ms > 0.0 && vs > 0.0) {
                    const double out = GLC_PARAMS.spheroidEfficiencyEnergeticOutflow * power / (vs * vs);
                    const double jout = (out / (mg + ms)) * y[GLC_P_SPH_ANGMOM];
                    const double zout = (out / mg) * y[GLC_P_SPH_ABUND_GAS];
                    rate[GLC_P_SPH_MASS_GAS] += -out;
                    rate[GLC_P_SPH_ANGMOM] += -jout;
                    rate[GLC_P_SPH_ABUND_GAS] += -zout;
                    hh_outflowing(a, c, w, rate, out, jout, zout);
                }
            }
        }

        // CGMAccretion (circumgalactic_medium/accretion.F90:517-593; accretion/halo/simple.F90:281-378,592-613)
        if (go && (ops & GLC_OP_CGM_ACCRETION)) {
            double rateHot = 0.0, rateFailed = 0.0, rateJ = 0.0;
            const double fb = GLC_PARAMS.OmegaBaryon / GLC_PARAMS.OmegaMatter;
            if (!sat) {
                const double failed = (time > GLC_PARAMS.timeReionization && w.vvir < GLC_PARAMS.velocitySuppressionReionization) ? 1.0 : 0.0;
                const double unaccreted = hh ? y[GLC_P_HH_UNACCRETED_MASS] : 0.0;
                const double growth = c.massRate / c.basicMass;
                rateHot = fb * c.massRate * (1.0 - failed) + unaccreted * growth * (1.0 - failed);
                rateFailed = fb * c.massRate * failed - unaccreted * growth * (1.0 - failed);
            }
            if (c.massRate != 0.0) rateJ = c.spinRate * rateHot / c.massRate;
            const bool hotPositive = hh && y[GLC_P_HH_MASS] > 0.0;
            if (rateHot > 0.0 || hotPositive || GLC_PARAMS.allowNegativeCGMMass)
                a.addCreate(rate[GLC_P_HH_MASS], GLC_F_HAS_HOTHALO, GLC_INT_HOTHALO_CREATE, rateHot);
            if (rateFailed > 0.0 || hotPositive || GLC_PARAMS.allowNegativeCGMMass)
                a.addCreate(rate[GLC_P_HH_UNACCRETED_MASS], GLC_F_HAS_HOTHALO, GLC_INT_HOTHALO_CREATE, rateFailed);
            a.addCreate(rate[GLC_P_HH_ANGMOM], GLC_F_HAS_HOTHALO, GLC_INT_HOTHALO_CREATE, rateJ);
        }

        // CGMOutflowReincorporation (outflow_reincorporation.F90:272-349; halo_dynamical_time.F90:113-128)
        const double massReturnRate = hh ? y[GLC_P_HH_OUTFLOWED_MASS] * GLC_PARAMS.reincorporationMultiplier / w.tdyn : 0.0;
        if (go && (ops & GLC_OP_CGM_OUTFLOW_REINCORPORATION) && hh && y[GLC_P_HH_OUTFLOWED_MASS] > 0.0) {
            const double mo = y[GLC_P_HH_OUTFLOWED_MASS];
            const double rj = y[GLC_P_HH_OUTFLOWED_ANGMOM] * (massReturnRate / mo);
            const double rz = y[GLC_P_HH_OUTFLOWED_ABUND] * (massReturnRate / mo);
            rate[GLC_P_HH_OUTFLOWED_MASS] += -massReturnRate;
            rate[GLC_P_HH_OUTFLOWED_ANGMOM] += -rj;
            rate[GLC_P_HH_OUTFLOWED_ABUND] += -rz;
            rate[GLC_P_HH_MASS] += massReturnRate;
            rate[GLC_P_HH_ANGMOM] += rj;
            rate[GLC_P_HH_ABUND] += rz;
        }

        // CGMCoolingHeating (cooling_heating.F90:216-381; component=disk, coolingFrom=currentNode)
        if (coolOn) {
            double cool = 0.0, rinfall = 0.0;
            if (coolRadiusOn) {
                rinfall = rinfallSolved;
                if (rinfall >= w.hhRouter)
                    cool = y[GLC_P_HH_MASS] / w.tdyn;
                else {
                    // coolingRadiusSimple::radiusGrowthRate :229-311 (isothermal: temperature slope 0)
                    const double x = rinfall / w.hhRcore;
                    const double densityLogSlope = -3.0 * GLC_PARAMS.hotHaloBeta * x * x / (x * x + 1.0);
                    double growth = 0.0;
                    if (rinfall > 0.0) {
                        const double slope = densityLogSlope * (1.0 - 2.0) + 0.0 * (-logSlopeT);
                        if (slope != 0.0) growth = rinfall / w.coolTavail * 1.0 / slope;
                    }
                    cool = 4.0 * kPi * rinfall * rinfall * hh_density(w, rinfall) * growth;
                }
            }
            // circumgalacticMediumHeatingAGNFeedback (circumgalactic_medium/heating/AGN_feedback.F90:103-126) over
            // blackHoleCGMHeatingJetPower (black_holes/CGM_heating/jet_power.F90:120-138)
            const double heat = (bh.on ? GLC_PARAMS.bhEfficiencyRadioMode * disk_power_jet(bh, bh.acc) : 0.0) / (w.vvir * w.vvir);
            if (heat > cool) {
                if (GLC_PARAMS.excessHeatDrivesOutflow) {
                    const double out = fmin(heat - cool, GLC_PARAMS.rateMaximumExpulsion * y[GLC_P_HH_MASS] / w.tdyn);
                    const double rz = y[GLC_P_HH_ABUND] * (out / y[GLC_P_HH_MASS]);
                    const double rj = y[GLC_P_HH_ANGMOM] * (out / y[GLC_P_HH_MASS]);
                    rate[GLC_P_HH_MASS] += -out;
                    rate[GLC_P_HH_ABUND] += -rz;
                    rate[GLC_P_HH_ANGMOM] += -rj;
                    if (sat) {
                        rate[GLC_P_HH_STRIPPED_MASS] += out;
                        rate[GLC_P_HH_STRIPPED_ABUND] += rz;
                    }
                }
            } else if (cool > heat) {
                cool = fmax(0.0, cool - heat);
                // coolingSpecificAngularMomentumConstantRotation (hotGas, hotGas), constant_rotation.F90:198-286
                double jSpecific = 0.0;
                if (rinfall > 0.0) {
                    const double x = w.hhRouter / w.hhRcore;
                    const bool b23 = beta_is_two_thirds();  // else the moments through 2F1 (:600-612) = I_m(x)
                    const double m2 = !b23 ? dm_beta_moment(2, x, GLC_PARAMS.hotHaloBeta)
                                           : ((x < 1.0e-6) ? x * x * x * (1.0 / 3.0 - x * x / 5.0) : x - dm_atan(x));
                    const double m3 = !b23 ? dm_beta_moment(3, x, GLC_PARAMS.hotHaloBeta)
                                           : ((x < 1.0e-6) ? x * x * x * x * (1.0 / 4.0 - x * x / 6.0) : 0.5 * (x * x - dm_log(1.0 + x * x)));
                    const double rc = w.hhRcore;
                    const double norm = (m2 * w.hhRho0 * (rc * rc * rc)) / (m3 * w.hhRho0 * (rc * rc * rc * rc));
                    jSpecific = norm * (y[GLC_P_HH_ANGMOM] / y[GLC_P_HH_MASS]) * rinfall;
                }
                const double rj = cool * jSpecific;
                const double rz = cool * y[GLC_P_HH_ABUND] / y[GLC_P_HH_MASS];
                rate[GLC_P_HH_MASS] += -cool;
                rate[GLC_P_HH_ANGMOM] += -rj;
                rate[GLC_P_HH_ABUND] += -rz;
                a.addCreate(rate[GLC_P_DISK_MASS_GAS], GLC_F_HAS_DISK, GLC_INT_DISK_CREATE, cool);
                a.addCreate(rate[GLC_P_DISK_ABUND_GAS], GLC_F_HAS_DISK, GLC_INT_DISK_CREATE, rz);
                a.addCreate(rate[GLC_P_DISK_ANGMOM], GLC_F_HAS_DISK, GLC_INT_DISK_CREATE,
                            rj * (1.0 - GLC_PARAMS.fractionLossAngularMomentum));
            }
        }

        // CGMOuterRadiusRamPressureStripping (outer_radius/ram_pressure_stripping.F90:151-313) with
        // hotHaloRamPressureStripping=virialRadius
        if (go && (ops & GLC_OP_CGM_OUTER_RADIUS) && hh) {
            const double router = w.hhRouter;
            if (router < w.rvir) {
                const double rho = hh_density(w, router);
                if (router > 0.0 && rho > 0.0) {
                    const double rhoMin = GLC_PARAMS.OmegaBaryon / GLC_PARAMS.OmegaMatter * c.basicMass / (w.rvir * w.rvir * w.rvir) / 4.0 / kPi;
                    rate[GLC_P_HH_OUTER_RADIUS] += massReturnRate / 4.0 / kPi / (router * router) / fmax(rho, rhoMin);
                } else if (massReturnRate > 0.0) {
                    rate[GLC_P_HH_OUTER_RADIUS] += massReturnRate / c.basicMass * w.rvir;
                }
            }
            if (!sat) {
                // virialDensityContrastDefinitionVirialRadiusGrowthRate :338-354
                rate[GLC_P_HH_OUTER_RADIUS] += (1.0 / 3.0) * w.rvir * (c.massRate / c.basicMass - w.dlnrhoDt);
            }
        }
        if (bad) c.numericsFailed = 1;
        return a.interrupt;
    }

    // structureOnly: the <eventHook postEvolve> call -- galactic structure solve at the final state
    // (equilibrium.F90:172,197-217) -- shares this entry point so that the kernel has ONE heavy call site.
    GLC_DEVICE_INLINE int rates(NodeCtx &c, double time, const double (&y)[NY], double (&rate)[NY],
                                bool structureOnly, bool on) {
        // [warp-synchronous: called by every lane of the warp; `on` = this lane wants an evaluation]
        Work w;
        int bad = 0;
        work_clear(w);
        if (on) {
            halo_scales(c, time, w);
            hh_profile(c, y, w);
        }
        // <eventHook preDerivative>: galactic structure solve (standard.F90:1045)
        structure_solve(c, y, time, w, bad, on);
        GLC_PHASE_SYNC();
        const bool go = on && !structureOnly && w.solvable;
        const bool dOn = disk_sfr_on(c, y, go);
        const double psiDisk = sfr_disk(c, y, bad, dOn);
        GLC_PHASE_SYNC();
        const bool coolOn = cooling_on(c, y, w, go);
        const bool radiusOn = cooling_radius_on(c, y, w, coolOn, go);
        double logSlopeT = 0.0;
        if (radiusOn) cooling_prepare(y, w, logSlopeT);
        const double rinfallSolved = cooling_radius(y, w, bad, radiusOn);
        GLC_PHASE_SYNC();
        return rates_accumulate(c, time, y, rate, w, go, dOn, psiDisk, coolOn, radiusOn, rinfallSolved, logSlopeT, bad);
    }

    // ---------------------------------------------------------------- post-step clamps
    GLC_DEVICE_INLINE int post_step(NodeCtx &c, double (&y)[NY]) {
        int status = kGslSuccess;
        // Node_Component_Disk_Standard_Post_Step, disk/standard/_class.F90:473-677
        if (has(c, GLC_F_HAS_DISK)) {
            if (y[GLC_P_DISK_MASS_GAS] < 0.0) {
                const double m = y[GLC_P_DISK_MASS_GAS] + y[GLC_P_DISK_MASS_STELLAR];
                double j;
                if (m == 0.0) {
                    j = 0.0;
                    y[GLC_P_DISK_MASS_STELLAR] = 0.0;
                    y[GLC_P_DISK_ABUND_STELLAR] = 0.0;
                } else {
                    j = y[GLC_P_DISK_ANGMOM] / m;
                    if (j < 0.0) j = c.diskRadius * c.diskVelocity;
                }
                y[GLC_P_DISK_MASS_GAS] = 0.0;
                y[GLC_P_DISK_ABUND_GAS] = 0.0;
                y[GLC_P_DISK_ANGMOM] = j * y[GLC_P_DISK_MASS_STELLAR];
                status = kGslContinue;
            }
            if (y[GLC_P_DISK_MASS_STELLAR] < 0.0) {
                const double m = y[GLC_P_DISK_MASS_GAS] + y[GLC_P_DISK_MASS_STELLAR];
                double j;
                if (m == 0.0) {
                    j = 0.0;
                    y[GLC_P_DISK_MASS_GAS] = 0.0;
                    y[GLC_P_DISK_ABUND_GAS] = 0.0;
                } else {
                    j = y[GLC_P_DISK_ANGMOM] / m;
                    if (j < 0.0) j = c.diskRadius * c.diskVelocity;
                }
                y[GLC_P_DISK_MASS_STELLAR] = 0.0;
                y[GLC_P_DISK_ABUND_STELLAR] = 0.0;
                y[GLC_P_DISK_ANGMOM] = j * y[GLC_P_DISK_MASS_GAS];
                status = kGslContinue;
            }
            if (y[GLC_P_DISK_ANGMOM] < 0.0) {
                if (y[GLC_P_DISK_MASS_STELLAR] + y[GLC_P_DISK_MASS_GAS] <= 0.0) y[GLC_P_DISK_ANGMOM] = 0.0;
                status = kGslContinue;
            }
        }
        // Node_Component_Hot_Halo_Standard_Post_Step, hot_halo/standard/_class.F90:455-530
        if (has(c, GLC_F_HAS_HOTHALO) && y[GLC_P_HH_MASS] < 0.0) {
            y[GLC_P_HH_MASS] = 0.0;
            status = kGslContinue;
        }
        // Node_Component_Spheroid_Standard_Post_Step, spheroid/standard/_class.F90:464-636
        if (has(c, GLC_F_HAS_SPHEROID)) {
            if (y[GLC_P_SPH_MASS_GAS] < 0.0) {
                const double m = y[GLC_P_SPH_MASS_GAS] + y[GLC_P_SPH_MASS_STELLAR];
                double j;
                if (m == 0.0) {
                    j = 0.0;
                    y[GLC_P_SPH_MASS_STELLAR] = 0.0;
                    y[GLC_P_SPH_ABUND_STELLAR] = 0.0;
                } else
                    j = y[GLC_P_SPH_ANGMOM] / m;
                y[GLC_P_SPH_MASS_GAS] = 0.0;
                y[GLC_P_SPH_ABUND_GAS] = 0.0;
                y[GLC_P_SPH_ANGMOM] = j * y[GLC_P_SPH_MASS_STELLAR];
                status = kGslContinue;
            }
            if (y[GLC_P_SPH_MASS_STELLAR] < 0.0) {
                const double m = y[GLC_P_SPH_MASS_GAS] + y[GLC_P_SPH_MASS_STELLAR];
                double j;
                if (m == 0.0) {
                    j = 0.0;
                    y[GLC_P_SPH_MASS_GAS] = 0.0;
                    y[GLC_P_SPH_ABUND_GAS] = 0.0;
                } else
                    j = y[GLC_P_SPH_ANGMOM] / m;
                y[GLC_P_SPH_MASS_STELLAR] = 0.0;
                y[GLC_P_SPH_ABUND_STELLAR] = 0.0;
                y[GLC_P_SPH_ANGMOM] = j * y[GLC_P_SPH_MASS_GAS];
                status = kGslContinue;
            }
            if (y[GLC_P_SPH_ANGMOM] < 0.0) {
                const double j = c.sphRadius * c.sphVelocity / GLC_PARAMS.spheroidRatioAngularMomentumScaleRadius;
                y[GLC_P_SPH_ANGMOM] = j * (y[GLC_P_SPH_MASS_GAS] + y[GLC_P_SPH_MASS_STELLAR]);
                status = kGslContinue;
            }
        }
        // Node_Component_Black_Hole_Standard_Post_Evolve, black_hole/standard/_class.F90:422-462
        if (has(c, GLC_F_HAS_BH)) {
            if (y[GLC_P_BH_SPIN] > 0.9999 || y[GLC_P_BH_SPIN] < 0.0) {
                y[GLC_P_BH_SPIN] = fmax(fmin(y[GLC_P_BH_SPIN], 0.9999), 0.0);
                status = kGslContinue;
            }
            if (y[GLC_P_BH_MASS] < 0.0) {
                y[GLC_P_BH_MASS] = GLC_PARAMS.bhSeedMass;
                status = kGslContinue;
            }
        }
        return status;
    }
};

}  // namespace glc
