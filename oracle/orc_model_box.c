/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.
 * "Box" model: the operator set of testSuite/parameters/reproducibility/closedBox.xml and
 * leakyBox.xml (disk verySimple + [hot halo verySimple], starFormationRateDisks=timescale
 * with starFormationTimescale=fixed, stellarPopulationProperties=instantaneous,
 * stellarFeedbackOutflows=fixed inside rateLimit).
 *
 * Pinned by the reference's golden values testSuite/test-reproducibility.py:46-67
 * (tests/test_oracle_golden.py).
 *
 * Not restated (inert for these parameter files): the rateLimit cap of
 * stellar_feedback/outflows/rate_limit.F90:158-168 (M_gas/(1e-3 tau_dyn) is ~1e4x the fixed
 * outflow rate), unaccretedMass/outerRadius of the verySimple hot halo (zero rates with
 * accretionHalo=zero), CGMStarvation (no satellites in the test trees).
 */
#include <math.h>

#include "orc_constants.h"
#include "orc_node.h"

int orc_box_active_list(const orc_evolve_ctx *c, int *active) {
    int n = 0;
    if (c->flags & GLC_F_HAS_DISK) {
        active[n++] = GLC_P_DISK_MASS_STELLAR;
        active[n++] = GLC_P_DISK_ABUND_STELLAR;
        active[n++] = GLC_P_DISK_MASS_GAS;
        active[n++] = GLC_P_DISK_ABUND_GAS;
    }
    if (c->flags & GLC_F_HAS_HOTHALO) {
        active[n++] = GLC_P_HH_MASS;
        active[n++] = GLC_P_HH_ABUND;
    }
    return n;
}

/* objects/nodes/components/disk/very_simple/_class.F90:371-419 (scaleAbsoluteMass=100, :137-138);
 * objects/nodes/components/hot_halo/very_simple/_class.F90:149-175 (scaleMassRelative=1e-2) */
void orc_box_scales(orc_evolve_ctx *c, double *s) {
    const double scale_absolute_mass = 100.0;
    const double *p = c->p;
    if (c->flags & GLC_F_HAS_DISK) {
        double mass = p[GLC_P_DISK_MASS_GAS] + p[GLC_P_DISK_MASS_STELLAR];
        double ab = p[GLC_P_DISK_ABUND_GAS] + p[GLC_P_DISK_ABUND_STELLAR];
        s[GLC_P_DISK_MASS_GAS] = fmax(mass, scale_absolute_mass);
        s[GLC_P_DISK_MASS_STELLAR] = fmax(mass, scale_absolute_mass);
        s[GLC_P_DISK_ABUND_GAS] = fmax(ab, scale_absolute_mass);
        s[GLC_P_DISK_ABUND_STELLAR] = fmax(ab, scale_absolute_mass);
    }
    if (c->flags & GLC_F_HAS_HOTHALO) {
        s[GLC_P_HH_MASS] = p[GLC_P_BASIC_MASS] * 1.0e-2;
        s[GLC_P_HH_ABUND] = p[GLC_P_BASIC_MASS] * 1.0e-2;
    }
}

void orc_box_solve_analytics(orc_evolve_ctx *c, double time) {
    /* dmoInterpolateDifferentialEvolutionSolveAnalytics, dark_matter_only_mass/interpolate.F90:217-239 */
    double *p = c->p;
    if (p[GLC_P_MASS_RATE] != 0.0)
        p[GLC_P_BASIC_MASS] = p[GLC_P_MASS_TARGET] + p[GLC_P_MASS_RATE] * (time - p[GLC_P_TIME_TARGET]);
}

int orc_box_rates(orc_evolve_ctx *c, double time, double *rate) {
    const glc_params *P = c->P;
    const double *p = c->p;
    (void)time;
    if (!(c->flags & GLC_F_HAS_DISK)) return GLC_INT_NONE;
    /* nodeOperatorStarFormationDisks, star_formation/disks.F90:200-284 with
       starFormationRateDisksTimescale::rate, star_formation/rates/disks/timescale.F90:109-130 */
    {
        double mass_gas = p[GLC_P_DISK_MASS_GAS];
        double psi, z_fuel, rate_mass_stellar, rate_metals_stellar, rate_metals_fuel;
        if (mass_gas < 0.0) return GLC_INT_NONE;
        psi = (P->box_timescaleStarFormation > 0.0) ? mass_gas / P->box_timescaleStarFormation : 0.0;
        /* abundancesFuel%massToMassFraction, objects/abundances.F90:811-828 */
        z_fuel = p[GLC_P_DISK_ABUND_GAS];
        if (z_fuel > mass_gas)
            z_fuel = 1.0;
        else if (z_fuel <= 0.0)
            z_fuel = 0.0;
        else
            z_fuel = z_fuel / mass_gas;
        /* instantaneousRates, stellar_populations/properties/instantaneous.F90:173-182 */
        rate_mass_stellar = (1.0 - P->recycledFraction) * psi;
        rate_metals_stellar = z_fuel * rate_mass_stellar;
        rate_metals_fuel = -rate_metals_stellar + P->metalYield * psi;
        rate[GLC_P_DISK_MASS_STELLAR] += rate_mass_stellar;
        rate[GLC_P_DISK_MASS_GAS] += -rate_mass_stellar;
        rate[GLC_P_DISK_ABUND_STELLAR] += rate_metals_stellar;
        rate[GLC_P_DISK_ABUND_GAS] += rate_metals_fuel;
        /* nodeOperatorStellarFeedbackDisks (stellar_feedback/disks.F90:116-205) with
           stellarFeedbackOutflowsFixed: ejective outflow = fraction * energy input / canonical */
        if (P->box_fractionOutflow > 0.0 && (c->flags & GLC_F_HAS_HOTHALO)) {
            double rate_energy = ORC_FEEDBACK_ENERGY_INPUT_AT_INFINITY_CANONICAL * psi;
            double outflow = P->box_fractionOutflow * rate_energy / ORC_FEEDBACK_ENERGY_INPUT_AT_INFINITY_CANONICAL;
            if (outflow > 0.0) {
                double ab_out = (mass_gas > 0.0) ? z_fuel * outflow : 0.0;
                rate[GLC_P_HH_MASS] += outflow;
                rate[GLC_P_DISK_MASS_GAS] -= outflow;
                rate[GLC_P_HH_ABUND] += ab_out;
                rate[GLC_P_DISK_ABUND_GAS] -= ab_out;
            }
        }
    }
    return GLC_INT_NONE;
}

/* Node_Component_Disk_Very_Simple_Post_Step, disk/very_simple/_class.F90:257-335 */
void orc_box_post_step(orc_evolve_ctx *c, int *status) {
    double *p = c->p;
    if (!(c->flags & GLC_F_HAS_DISK)) return;
    if (p[GLC_P_DISK_MASS_GAS] < 0.0) {
        double mass_disk = p[GLC_P_DISK_MASS_GAS] + p[GLC_P_DISK_MASS_STELLAR];
        if (mass_disk == 0.0) {
            p[GLC_P_DISK_MASS_STELLAR] = 0.0;
            p[GLC_P_DISK_ABUND_STELLAR] = 0.0;
        }
        p[GLC_P_DISK_MASS_GAS] = 0.0;
        p[GLC_P_DISK_ABUND_GAS] = 0.0;
        if (*status == ORC_GSL_SUCCESS) *status = ORC_GSL_CONTINUE;
    }
}
