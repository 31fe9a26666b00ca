// glc_params.cpp -- default parameter sets (host side of the C-ABI).
// Values are those of parameters/quickTest.xml (model STANDARD) and
// testSuite/parameters/reproducibility/closedBox.xml (model BOX); class defaults are cited inline.
#include <cstring>

#include "../../include/glc_b200.h"

extern "C" int glc_params_default(glc_params *P, int32_t model) {
    if (!P) return -1;
    if (model != GLC_MODEL_BOX && model != GLC_MODEL_STANDARD) return -7;
    std::memset(P, 0, sizeof(*P));
    P->abi_version = GLC_ABI_VERSION;
    P->model = model;
    P->reuseODEStepSize = 1;          // node_evolver/standard.F90:239-240
    P->enforceNonNegativity = 0;      // node_evolver/standard.F90:247-248
    P->resolveInterruptsOnDevice = 1;
    if (model == GLC_MODEL_BOX) {
        P->odeToleranceAbsolute = 1.0e-6;  // closedBox.xml mergerTreeNodeEvolver
        P->odeToleranceRelative = 1.0e-3;
        P->OmegaMatter = 0.3;
        P->OmegaBaryon = 0.05;
        P->HubbleConstant = 70.0;
        P->recycledFraction = 0.4;
        P->metalYield = 0.025;
        P->box_timescaleStarFormation = 0.25;
        P->box_fractionOutflow = 0.0;
        P->operatorMask = GLC_OP_STAR_FORMATION_DISKS | GLC_OP_STELLAR_FEEDBACK_DISKS;
        return 0;
    }
    P->odeToleranceAbsolute = 0.01;   // quickTest.xml:292-295
    P->odeToleranceRelative = 0.01;
    P->OmegaMatter = 0.2725;          // quickTest.xml:37-43
    P->OmegaBaryon = 0.0455;
    P->HubbleConstant = 70.2;
    P->recycledFraction = 0.46;       // quickTest.xml:168-171
    P->metalYield = 0.035;
    P->timeReionization = 0.0;        // host converts redshiftReionization=10.5 (quickTest.xml:101-104)
    P->velocitySuppressionReionization = 35.0;
    P->hotHaloBeta = 2.0 / 3.0;       // hot_halo/mass_distribution/beta_profile.F90:94
    P->coreRadiusOverVirialRadius = 0.3;
    P->hotHaloScaleMassRelative = 1.0e-3;   // hot_halo/standard/_class.F90:226-227
    P->hotHaloScaleRadiusRelative = 1.0e-1;
    P->outflowStrippingEfficiency = 0.1;
    P->reincorporationMultiplier = 5.0;
    P->fractionLossAngularMomentum = 0.3;
    P->coolingVelocityCutOff = 10000.0;
    P->coolingDegreesOfFreedom = 3.0;
    P->rateMaximumExpulsion = 1.0;
    P->excessHeatDrivesOutflow = 1;
    P->allowNegativeCGMMass = 1;
    P->frequencyStarFormation = 0.385;
    P->clumpingFactorMolecularComplex = 5.0;
    P->sfrIntegrationTolerance = 1.0e-3;
    P->krumholzSTruncation = 2.0 - 2.0e-10;  /* exact root of f_H2(s) = 1e-10 for the fast fit; host may refine */    // star_formation/rates/disks/integrated_surface_density.F90:81-82
    P->sfSpheroidEfficiency = 0.04;
    P->sfSpheroidExponentVelocity = 2.0;
    P->sfSpheroidTimescaleMinimum = 0.001;
    P->fbDiskVelocityCharacteristic = 250.0;
    P->fbDiskExponent = 3.5;
    P->fbSpheroidVelocityCharacteristic = 100.0;
    P->fbSpheroidExponent = 3.5;
    P->fbTimescaleOutflowFractionalMinimum = 0.001;
    P->diskToleranceAbsoluteMass = 1.0e-6;
    P->spheroidToleranceAbsoluteMass = 1.0e-6;
    P->spheroidRatioAngularMomentumScaleRadius = 0.5;
    P->spheroidEfficiencyEnergeticOutflow = 1.0e-2;
    P->structureSolutionTolerance = 1.0e-2;  // galactic_structure/radius_solver/equilibrium.F90:117-118
    P->adiabaticA = 0.73;
    P->adiabaticOmega = 0.7;
    P->includeBaryonGravity = 1;
    P->adiabaticContraction = 1;
    P->barStabilityThresholdGaseous = 0.7;
    P->barStabilityThresholdStellar = 1.1;
    P->bhSeedMass = 100.0;
    P->bhSeedSpin = 0.0;
    P->bondiHoyleAccretionEnhancementSpheroid = 5.0;
    P->bondiHoyleAccretionEnhancementHotHalo = 6.0;
    P->bondiHoyleAccretionTemperatureSpheroid = 100.0;
    P->bondiHoyleAccretionHotModeOnly = 1;
    P->bhEfficiencyWind = 0.0024;
    P->bhEfficiencyRadioMode = 1.0;
    P->accretionRateThinDiskMaximum = 0.30;
    P->accretionRateThinDiskMinimum = 0.01;
    P->adafEfficiencyRadiation = 0.01;
    P->adafAdiabaticIndex = 1.444;
    P->accretionRateTransitionWidth = 0.1;  /* accretion_disks/switched.F90:125-127 */
    P->scaleADAFRadiativeEfficiency = 1;
    P->bhEfficiencyWindScalesWithEfficiencyRadiative = 1;
    P->adafEfficiencyRadiationTypeThinDisk = 1;
    P->operatorMask = GLC_OP_ALL;
    P->darkMatterProfileDMO = GLC_DMO_NFW;        // quickTest.xml:88
    P->structureVelocityMaximumFactor = 0.0;      // equilibrium.F90:124-128
    P->timestepHostRelative = 0.1;                // quickTest.xml:297-300
    P->timestepHostAbsolute = 1.0;
    P->timestepSimpleRelative = 0.1;              // merger_trees/evolve/timesteps/simple.F90 defaults
    P->timestepSimpleAbsolute = 1.0;
    P->wallClockMaximumSeconds = 0.0;
    P->profileOdeEvolver = 0;                     // node_evolver/standard.F90:249-253
    P->profilerTimeStepPointsPerDecade = 3;       // merger_trees/evolve/profiler/simple.F90:84-104
    P->profilerTimeStepMinimum = 1.0e-6;
    P->profilerTimeStepMaximum = 1.0e+1;
    return 0;
}
