// glc_model_box.cuh -- device rate functions for the "box" operator set
// (testSuite/parameters/reproducibility/{closedBox,leakyBox}.xml):
//   nodeOperatorStarFormationDisks  source/nodes/operators/physics/star_formation/disks.F90:200-284
//     + starFormationRateDisksTimescale  source/star_formation/rates/disks/timescale.F90:109-130
//     + stellarPopulationPropertiesInstantaneous  source/stellar_populations/properties/instantaneous.F90:173-182
//   nodeOperatorStellarFeedbackDisks source/nodes/operators/physics/stellar_feedback/disks.F90:116-205
//     + stellarFeedbackOutflowsFixed
//   disk/verySimple scales + post-step  source/objects/nodes/components/disk/very_simple/_class.F90:257-419
#pragma once

#include "glc_common.cuh"

namespace glc {

struct ModelBox {
    GLC_DEVICE_INLINE uint32_t active_mask(int flags) {
        uint32_t m = 0;
        if (flags & GLC_F_HAS_DISK)
            m |= (1u << GLC_P_DISK_MASS_STELLAR) | (1u << GLC_P_DISK_ABUND_STELLAR) |
                 (1u << GLC_P_DISK_MASS_GAS) | (1u << GLC_P_DISK_ABUND_GAS);
        if (flags & GLC_F_HAS_HOTHALO) m |= (1u << GLC_P_HH_MASS) | (1u << GLC_P_HH_ABUND);
        return m;
    }

    GLC_DEVICE_INLINE void solve_analytics(NodeCtx &c, double time) {
        // dmoInterpolateDifferentialEvolutionSolveAnalytics, dark_matter_only_mass/interpolate.F90:217-239
        if (c.massRate != 0.0) c.basicMass = c.massTarget + c.massRate * (time - c.timeTarget);
    }

    GLC_DEVICE_INLINE void scales(const NodeCtx &c, const double (&y)[NY],
                                                  double (&s)[NY]) {
        const double scaleAbsoluteMass = 100.0;  // disk/very_simple/_class.F90:137-138
        if (c.flags & GLC_F_HAS_DISK) {
            const double mass = y[GLC_P_DISK_MASS_GAS] + y[GLC_P_DISK_MASS_STELLAR];
            const double ab = y[GLC_P_DISK_ABUND_GAS] + y[GLC_P_DISK_ABUND_STELLAR];
            s[GLC_P_DISK_MASS_GAS] = fmax(mass, scaleAbsoluteMass);
            s[GLC_P_DISK_MASS_STELLAR] = fmax(mass, scaleAbsoluteMass);
            s[GLC_P_DISK_ABUND_GAS] = fmax(ab, scaleAbsoluteMass);
            s[GLC_P_DISK_ABUND_STELLAR] = fmax(ab, scaleAbsoluteMass);
        }
        if (c.flags & GLC_F_HAS_HOTHALO) {
            // hot_halo/very_simple/_class.F90:149-175, scaleMassRelative = 1e-2
            s[GLC_P_HH_MASS] = c.basicMass * 1.0e-2;
            s[GLC_P_HH_ABUND] = c.basicMass * 1.0e-2;
        }
    }

    // returns an interrupt code (GLC_INT_NONE here: the box trees never create components)
    static constexpr bool kHasPostEvolve = false;
    GLC_DEVICE_INLINE int rates(NodeCtx &c, double /*time*/, const double (&y)[NY], double (&rate)[NY],
                                bool /*structureOnly*/, bool on) {
        if (!on || !(c.flags & GLC_F_HAS_DISK)) return GLC_INT_NONE;
        const double massGas = y[GLC_P_DISK_MASS_GAS];
        if (massGas < 0.0) return GLC_INT_NONE;
        const double tau = GLC_PARAMS.box_timescaleStarFormation;
        const double psi = (tau > 0.0) ? massGas / tau : 0.0;
        // abundances%massToMassFraction, objects/abundances.F90:811-828
        double zFuel = y[GLC_P_DISK_ABUND_GAS];
        zFuel = (zFuel > massGas) ? 1.0 : ((zFuel <= 0.0) ? 0.0 : zFuel / massGas);
        const double rateMassStellar = (1.0 - GLC_PARAMS.recycledFraction) * psi;
        const double rateMetalsStellar = zFuel * rateMassStellar;
        const double rateMetalsFuel = -rateMetalsStellar + GLC_PARAMS.metalYield * psi;
        rate[GLC_P_DISK_MASS_STELLAR] += rateMassStellar;
        rate[GLC_P_DISK_MASS_GAS] += -rateMassStellar;
        rate[GLC_P_DISK_ABUND_STELLAR] += rateMetalsStellar;
        rate[GLC_P_DISK_ABUND_GAS] += rateMetalsFuel;
        if (GLC_PARAMS.box_fractionOutflow > 0.0 && (c.flags & GLC_F_HAS_HOTHALO)) {
            const double rateEnergy = kFeedbackEnergyInputAtInfinityCanonical * psi;
            const double outflow =
                GLC_PARAMS.box_fractionOutflow * rateEnergy / kFeedbackEnergyInputAtInfinityCanonical;
            if (outflow > 0.0) {
                const double abOut = (massGas > 0.0) ? zFuel * outflow : 0.0;
                rate[GLC_P_HH_MASS] += outflow;
                rate[GLC_P_DISK_MASS_GAS] -= outflow;
                rate[GLC_P_HH_ABUND] += abOut;
                rate[GLC_P_DISK_ABUND_GAS] -= abOut;
            }
        }
        return GLC_INT_NONE;
    }

    // Node_Component_Disk_Very_Simple_Post_Step; returns GSL status (Success / Continue / Failure)
    GLC_DEVICE_INLINE int post_step(NodeCtx &c, double (&y)[NY]) {
        int status = kGslSuccess;
        if ((c.flags & GLC_F_HAS_DISK) && y[GLC_P_DISK_MASS_GAS] < 0.0) {
            const double massDisk = y[GLC_P_DISK_MASS_GAS] + y[GLC_P_DISK_MASS_STELLAR];
            if (massDisk == 0.0) {
                y[GLC_P_DISK_MASS_STELLAR] = 0.0;
                y[GLC_P_DISK_ABUND_STELLAR] = 0.0;
            }
            y[GLC_P_DISK_MASS_GAS] = 0.0;
            y[GLC_P_DISK_ABUND_GAS] = 0.0;
            status = kGslContinue;
        }
        return status;
    }

    GLC_DEVICE_INLINE void pre_evolve(NodeCtx &, double (&)[NY]) {}
};

}  // namespace glc
