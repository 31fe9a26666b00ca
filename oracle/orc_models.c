/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.
 * Model dispatch, default parameters (parameters/quickTest.xml, reproducibility/closedBox.xml),
 * table storage and the host-side interrupt procedures (component creation).
 */
#include <math.h>

#include "../galacticus_b200/csrc/glc_detmath.h"
#include <stdlib.h>
#include <string.h>

#include "orc_node.h"

/* box model */
int orc_box_active_list(const orc_evolve_ctx *c, int *active);
void orc_box_scales(orc_evolve_ctx *c, double *s);
void orc_box_solve_analytics(orc_evolve_ctx *c, double time);
int orc_box_rates(orc_evolve_ctx *c, double time, double *rate);
void orc_box_post_step(orc_evolve_ctx *c, int *status);
/* standard model */
int orc_std_active_list(const orc_evolve_ctx *c, int *active);
void orc_std_scales(orc_evolve_ctx *c, double *s);
void orc_std_solve_analytics(orc_evolve_ctx *c, double time);
int orc_std_rates(orc_evolve_ctx *c, double time, double *rate);
void orc_std_post_step(orc_evolve_ctx *c, int *status);
void orc_std_post_evolve(orc_evolve_ctx *c);
void orc_std_pre_evolve(orc_evolve_ctx *c);

int orc_model_active_list(const orc_evolve_ctx *c, int *active) {
    return c->P->model == GLC_MODEL_BOX ? orc_box_active_list(c, active) : orc_std_active_list(c, active);
}
void orc_model_scales(orc_evolve_ctx *c, double *s) {
    if (c->P->model == GLC_MODEL_BOX)
        orc_box_scales(c, s);
    else
        orc_std_scales(c, s);
}
void orc_model_solve_analytics(orc_evolve_ctx *c, double time) {
    if (c->P->model == GLC_MODEL_BOX)
        orc_box_solve_analytics(c, time);
    else
        orc_std_solve_analytics(c, time);
}
int orc_model_rates(orc_evolve_ctx *c, double time, double *rate) {
    return c->P->model == GLC_MODEL_BOX ? orc_box_rates(c, time, rate) : orc_std_rates(c, time, rate);
}
void orc_model_post_step(orc_evolve_ctx *c, int *status) {
    if (c->P->model == GLC_MODEL_BOX)
        orc_box_post_step(c, status);
    else
        orc_std_post_step(c, status);
}
void orc_model_post_evolve(orc_evolve_ctx *c) {
    if (c->P->model != GLC_MODEL_BOX) orc_std_post_evolve(c);
}
void orc_model_pre_evolve(orc_evolve_ctx *c) {
    if (c->P->model != GLC_MODEL_BOX) orc_std_pre_evolve(c);
}

/* functionInterrupt procedures: <class>CreateByInterrupt (Properties/Evolve.py:486-493) create the
 * component with default (zero) property values; blackHoleCreate (black_holes/seed.F90:179-237)
 * seeds mass and spin from blackHoleSeeds=fixed. */
void orc_apply_interrupt(const glc_params *P, double *p, int *flags, int code) {
    switch (code) {
    case GLC_INT_HOTHALO_CREATE:
        *flags |= GLC_F_HAS_HOTHALO;
        break;
    case GLC_INT_DISK_CREATE:
        *flags |= GLC_F_HAS_DISK;
        break;
    case GLC_INT_SPHEROID_CREATE:
        *flags |= GLC_F_HAS_SPHEROID;
        break;
    case GLC_INT_BH_CREATE:
        *flags |= GLC_F_HAS_BH;
        p[GLC_P_BH_MASS] = P->bhSeedMass;
        p[GLC_P_BH_SPIN] = P->bhSeedSpin;
        break;
    default:
        break;
    }
}

void orc_params_default(glc_params *P, int model) {
    memset(P, 0, sizeof(*P));
    P->abi_version = GLC_ABI_VERSION;
    P->model = model;
    P->reuseODEStepSize = 1;      /* node_evolver/standard.F90:239-240 */
    P->enforceNonNegativity = 0;  /* :247-248 */
    P->resolveInterruptsOnDevice = 1;
    if (model == GLC_MODEL_BOX) {
        /* testSuite/parameters/reproducibility/closedBox.xml */
        P->odeToleranceAbsolute = 1.0e-6;
        P->odeToleranceRelative = 1.0e-3;
        P->OmegaMatter = 0.3;
        P->OmegaBaryon = 0.05;
        P->HubbleConstant = 70.0;
        P->recycledFraction = 0.4;
        P->metalYield = 0.025;
        P->box_timescaleStarFormation = 0.25;
        P->box_fractionOutflow = 0.0;
        P->operatorMask = GLC_OP_STAR_FORMATION_DISKS | GLC_OP_STELLAR_FEEDBACK_DISKS;
        return;
    }
    /* parameters/quickTest.xml */
    P->odeToleranceAbsolute = 0.01;
    P->odeToleranceRelative = 0.01;
    P->OmegaMatter = 0.2725;
    P->OmegaBaryon = 0.0455;
    P->HubbleConstant = 70.2;
    P->recycledFraction = 0.46;
    P->metalYield = 0.035;
    P->timeReionization = 0.0; /* filled by the host from redshiftReionization=10.5 */
    P->velocitySuppressionReionization = 35.0;
    P->hotHaloBeta = 2.0 / 3.0;
    P->coreRadiusOverVirialRadius = 0.3;
    P->hotHaloScaleMassRelative = 1.0e-3;
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
    P->krumholzSTruncation = 2.0 - 2.0e-10;  /* exact root of f_H2(s) = 1e-10 for the fast fit; host may refine */
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
    P->structureSolutionTolerance = 1.0e-2;
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
    P->darkMatterProfileDMO = GLC_DMO_NFW;   /* quickTest.xml:88 */
    P->structureVelocityMaximumFactor = 0.0; /* equilibrium.F90:124-128 */
    P->timestepHostRelative = 0.1;           /* quickTest.xml:297-300 */
    P->timestepHostAbsolute = 1.0;
    P->timestepSimpleRelative = 0.1;         /* merger_trees/evolve/timesteps/simple.F90 */
    P->timestepSimpleAbsolute = 1.0;
    P->wallClockMaximumSeconds = 0.0;
    P->profileOdeEvolver = 0;                     /* node_evolver/standard.F90:249-253 */
    P->profilerTimeStepPointsPerDecade = 3;       /* merger_trees/evolve/profiler/simple.F90:84-104 */
    P->profilerTimeStepMinimum = 1.0e-6;
    P->profilerTimeStepMaximum = 1.0e+1;
}

/* ---- tables -------------------------------------------------------------------------- */
orc_tables *orc_tables_create(void) {
    orc_tables *T = (orc_tables *)calloc(1, sizeof(orc_tables));
    return T;
}

static void free2d(orc_table2d *t) {
    free(t->x0);
    free(t->x1);
    free(t->v);
    memset(t, 0, sizeof(*t));
}

void orc_tables_destroy(orc_tables *T) {
    int i;
    if (!T) return;
    for (i = 0; i < GLC_NTABLES; i++) free2d(&T->t[i]);
    free(T->cooling_lnZ);
    free(T->cooling_lnT);
    free(T->cooling_lnL);
    free(T->electron_lnZ);
    free(T->electron_lnT);
    free(T->electron_lnV);
    free(T);
}

static double *dupd(const double *s, size_t n) {
    double *d = (double *)malloc(n * sizeof(double));
    memcpy(d, s, n * sizeof(double));
    return d;
}

/* cieFileReadFile post-processing, cooling/cooling_function/CIE_file.F90:627-659:
 * tables whose entries are all > 0 are stored as logarithms; a zero first metallicity is
 * flagged and stored as metallicityLogarithmicZero. */
static void cie_prepare(const orc_table2d *t, int *is_log, int *first_zero, double *first_nonzero,
                        double **lnZ, double **lnT, double **lnV) {
    int i, n = t->n0 * t->n1;
    const double metallicity_logarithmic_zero = -999.0;
    *is_log = 1;
    for (i = 0; i < n; i++)
        if (!(t->v[i] > 0.0)) *is_log = 0;
    free(*lnZ);
    free(*lnT);
    free(*lnV);
    *lnZ = dupd(t->x0, (size_t)t->n0);
    *lnT = dupd(t->x1, (size_t)t->n1);
    *lnV = dupd(t->v, (size_t)n);
    *first_zero = 0;
    *first_nonzero = 0.0;
    if (*is_log) {
        *first_zero = (t->x0[0] == 0.0);
        if (*first_zero) *first_nonzero = t->x0[1];
        for (i = 0; i < t->n0; i++)
            (*lnZ)[i] = (t->x0[i] > 0.0) ? dm_log(t->x0[i]) : metallicity_logarithmic_zero;
        for (i = 0; i < t->n1; i++) (*lnT)[i] = dm_log(t->x1[i]);
        for (i = 0; i < n; i++) (*lnV)[i] = dm_log(t->v[i]);
    }
}

int orc_tables_set(orc_tables *T, int id, int n0, int n1, const double *x0, const double *x1,
                   const double *values) {
    orc_table2d *t;
    if (id < 0 || id >= GLC_NTABLES) return -1;
    t = &T->t[id];
    free2d(t);
    t->n0 = n0;
    t->n1 = n1;
    t->x0 = dupd(x0, (size_t)n0);
    t->x1 = x1 ? dupd(x1, (size_t)n1) : NULL;
    t->v = dupd(values, (size_t)n0 * (size_t)n1);
    if (id == GLC_TABLE_COOLING_FUNCTION)
        cie_prepare(t, &T->cooling_log, &T->cooling_first_z_zero, &T->cooling_first_nonzero_z,
                    &T->cooling_lnZ, &T->cooling_lnT, &T->cooling_lnL);
    if (id == GLC_TABLE_ELECTRON_FRACTION)
        cie_prepare(t, &T->electron_log, &T->electron_first_z_zero, &T->electron_first_nonzero_z,
                    &T->electron_lnZ, &T->electron_lnT, &T->electron_lnV);
    return 0;
}
