/*
 * glc_b200.h -- C-ABI of the B200-native node-ODE evolver for Galacticus.
 *
 * This is the drop-in boundary for ONE path of the reference: the per-node ODE
 * integration performed by mergerTreeNodeEvolverStandard
 *   (source/merger_trees/node_evolver/standard.F90:385-755, "standardEvolve")
 * including its RHS (standardODEs :831-946, standardDerivativesCompute :1019-1061),
 * post-step hook (:1160-1185), the GSL driver chain below it
 *   (source/numerical/ODE_solver/solver.F90:492-636,
 *    source/external/gslODEInitVal2/driver2.c:148-250, cscal2.c:93-169,
 *    libgsl-2.6 rkck.c / evolve.c)
 * and the standard-component rate functions reached from nodeOperatorMulti
 *   (source/nodes/operators/multi.F90:313-332).
 *
 * All entry points are extern "C", take plain pointers/sizes and are callable from
 * Fortran through ISO_C_BINDING (see INTEGRATION.md for the bind(C) interface block
 * and the mergerTreeNodeEvolverB200 shim a maintainer would add).  No torch types.
 *
 * Units follow Galacticus: Msun, Mpc, km/s, Gyr.
 */
#ifndef GLC_B200_H
#define GLC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLC_ABI_VERSION 3

/* ---------------------------------------------------------------------------------
 * Node record.  One row of GLC_NPROP doubles per node, node-major ("what
 * treeNode%serializeValues would write", python/Galacticus/Build/Components/
 * TreeNodes/ODESolver.py:95-138, extended by the analytic / non-evolved properties
 * the RHS reads).  Component class order and property order follow the generated
 * serialization (component classes in the order of the generated treeNodeSerializeValuesToArray, XML property order;
 * checked against the output of the reference generators by tests/test_layout.py).
 * ------------------------------------------------------------------------------- */
enum glc_prop {
    /* --- numerically integrated properties (the ODE state vector y) ------------- */
    /* blackHole/standard  objects/nodes/components/black_hole/standard/_class.F90:42 */
    GLC_P_BH_MASS = 0,
    GLC_P_BH_SPIN,
    /* disk/standard       objects/nodes/components/disk/standard/_class.F90:42-130 */
    GLC_P_DISK_MASS_STELLAR,
    GLC_P_DISK_ABUND_STELLAR, /* abundancesStellar (single "metals" scalar)  */
    GLC_P_DISK_MASS_GAS,
    GLC_P_DISK_ABUND_GAS,
    GLC_P_DISK_ANGMOM,
    /* hotHalo/standard    objects/nodes/components/hot_halo/standard/_class.F90:50-180 */
    GLC_P_HH_MASS,
    GLC_P_HH_ABUND,
    GLC_P_HH_ANGMOM,
    GLC_P_HH_OUTFLOWED_MASS,
    GLC_P_HH_OUTFLOWED_ANGMOM,
    GLC_P_HH_OUTFLOWED_ABUND,
    GLC_P_HH_UNACCRETED_MASS,
    GLC_P_HH_UNACCRETED_ABUND,
    GLC_P_HH_OUTER_RADIUS,
    GLC_P_HH_STRIPPED_MASS,
    GLC_P_HH_STRIPPED_ABUND,
    /* satellite/standard  objects/nodes/components/satellite/standard.F90:33 */
    GLC_P_SAT_BOUND_MASS,
    /* spheroid/standard   objects/nodes/components/spheroid/standard/_class.F90:44 */
    GLC_P_SPH_MASS_STELLAR,
    GLC_P_SPH_ABUND_STELLAR,
    GLC_P_SPH_MASS_GAS,
    GLC_P_SPH_ABUND_GAS,
    GLC_P_SPH_ANGMOM,
    GLC_NY, /* = 24: size of the ODE state */

    /* --- analytic / non-evolved properties read (and some written) by the path --- */
    GLC_P_TIME = GLC_NY,     /* basic%time(): in = start, out = time reached            */
    GLC_P_TIME_STEP,         /* node%timeStep(): in = step guess (<=0: none), out = last */
    GLC_P_MASS_TARGET,       /* basic mass at GLC_P_TIME_TARGET (massDMOTarget meta-property,
                                nodes/operators/physics/dark_matter_only_mass/interpolate.F90:216-238) */
    GLC_P_MASS_RATE,         /* basic%accretionRate()                                   */
    GLC_P_TIME_TARGET,       /* parent%basic%time()                                     */
    GLC_P_DMSCALE_TARGET,    /* darkMatterProfile%scale() target (…/dark_matter_profile_scale/interpolate.F90) */
    GLC_P_DMSCALE_RATE,
    GLC_P_SPIN_TARGET,       /* spin%angularMomentum() target (…/halo_angular_momentum_interpolate.F90) */
    GLC_P_SPIN_RATE,
    GLC_P_TIME_LAST_ISOLATED,/* basic%timeLastIsolated(); <=0 -> use time               */
    GLC_P_DISK_RADIUS,       /* disk%radius()   (in = warm start, out = solved)         */
    GLC_P_DISK_VELOCITY,
    GLC_P_SPH_RADIUS,
    GLC_P_SPH_VELOCITY,
    GLC_P_BASIC_MASS,        /* basic%mass(): in = value at GLC_P_TIME, out = value at the time reached */
    GLC_P_DMSCALE,           /* darkMatterProfile%scale() current value (same convention)  */
    GLC_P_SPIN,              /* spin%angularMomentum() current value (same convention)     */
    GLC_P_MASS_BARYONIC_SUBHALOS, /* in: sum of massBaryonic() over all sub-satellites, frozen for the
                                call (dark_matter_profiles/adiabatic_Gnedin2004.F90:327-347)   */
    GLC_NPROP
};

/* component-existence / status bits (int32 per node) */
enum glc_flag {
    GLC_F_HAS_HOTHALO  = 1 << 0,
    GLC_F_HAS_DISK     = 1 << 1,
    GLC_F_HAS_SPHEROID = 1 << 2,
    GLC_F_HAS_BH       = 1 << 3,
    GLC_F_IS_SATELLITE = 1 << 4,
    GLC_F_HH_INITIALIZED = 1 << 5 /* hotHalo%isInitialized(): outerRadius has been set to the virial
                                     radius by the pre-evolve task (hot_halo/standard/_class.F90:871-891) */
};

/* per-node status.  Numerically EQUAL to the reference's errorStatus* constants (source/error/_module.F90:66-75, which
 * copy the GSL error codes of gsl_errno.h) as used at node_evolver/standard.F90:392-393,697-720, so that the Fortran shim
 * hands them to the tree evolver unchanged; library-specific codes are above 1024 like the reference's own additions. */
enum glc_status {
    GLC_STATUS_SUCCESS   = 0,    /* errorStatusSuccess   = GSL_SUCCESS                                          */
    GLC_STATUS_FAIL      = -1,   /* errorStatusFail      = GSL_FAILURE                                          */
    GLC_STATUS_UNDERFLOW = 15,   /* errorStatusUnderflow = GSL_EUNDRFLW: 8 trials exhausted (standard.F90:707-722) */
    GLC_STATUS_XCPU      = 1025, /* errorStatusXCPU: systemClockMaximum exceeded (standard.F90:694-705,861-867)  */
    GLC_STATUS_NONFINITE = 1027, /* device flagged NaN/Inf where the reference would trap (-ffpe-trap, Makefile:100) */
    GLC_STATUS_PENDING   = -2139062144 /* 0x80808080: not evolved (yet) -- streaming sessions, and nodes of a batch call
                                    that returned an error */
};

/* return codes of the entry points: 0 = success, negative = error (positive: completed with a warning).  -1..-12: argument / state errors (see
 * glc_last_error), <= -1000: -(cudaError_t) - 1000. */
enum glc_error {
    GLC_ERR_STALLED        = -20, /* the device made no progress on a batch (a defect, never expected): nothing was lost,
                                     the nodes not evolved keep GLC_STATUS_PENDING */
    GLC_ERR_BUSY           = -21, /* batch call while a streaming session is active, or stream call out of sequence */
    GLC_ERR_BAD_FOREST     = -22, /* parent array is not a forest (index out of range or a cycle) */
    GLC_WARN_EVOLVE_FAILED = 1,   /* glc_forest_evolve completed, but failed_evolves node evolves came back with a status other
                                     than success (the reference aborts the run or, with tolerateFailures, drops the tree there:
                                     tasks/evolve_forests/_class.F90:887-897); the only positive return code */
    GLC_ERR_DEADLOCK       = -24  /* glc_forest_evolve: trees not at their final time although no node can move
                                     (merger_trees/evolver/standard.F90:606-625) */
};

/* interrupt codes = which functionInterrupt the host must call
 * (python/Galacticus/Build/Components/Properties/Evolve.py:486-493; black_holes/seed.F90:179-180) */
enum glc_interrupt {
    GLC_INT_NONE = 0,
    GLC_INT_HOTHALO_CREATE,  /* hotHaloCreateByInterrupt  */
    GLC_INT_DISK_CREATE,     /* diskCreateByInterrupt     */
    GLC_INT_SPHEROID_CREATE, /* spheroidCreateByInterrupt */
    GLC_INT_BH_CREATE        /* blackHoleCreate           */
};

enum glc_model {
    GLC_MODEL_BOX = 0,      /* testSuite/parameters/reproducibility/{closedBox,leakyBox}.xml physics */
    GLC_MODEL_STANDARD = 1  /* parameters/quickTest.xml operator set                                  */
};

/* ---------------------------------------------------------------------------------
 * Parameters: flat POD filled once by the Fortran shim from the already-constructed
 * objects.  Field names follow the XML parameter names.
 * ------------------------------------------------------------------------------- */
typedef struct glc_params {
    int32_t abi_version; /* GLC_ABI_VERSION */
    int32_t model;       /* enum glc_model  */
    /* mergerTreeNodeEvolverStandard, node_evolver/standard.F90:166-253 */
    double odeToleranceAbsolute;
    double odeToleranceRelative;
    int32_t reuseODEStepSize;
    int32_t enforceNonNegativity;
    int32_t resolveInterruptsOnDevice; /* 1: component-creation interrupts are applied in the
                                          kernel and the segment restarted (equivalent to the
                                          host loop evolver/standard.F90:425-476); 0: return to host */
    int32_t pad0;
    /* cosmologyParameters */
    double OmegaMatter, OmegaBaryon, HubbleConstant;
    /* stellarPopulation standard / instantaneous (stellar_populations/properties/instantaneous.F90) */
    double recycledFraction, metalYield;
    /* box model: starFormationRateDisks=timescale(fixed), stellarFeedbackOutflows=fixed */
    double box_timescaleStarFormation;
    double box_fractionOutflow;
    /* accretionHalo simple */
    double timeReionization, velocitySuppressionReionization;
    /* hot halo */
    double hotHaloBeta;                /* hotHaloMassDistribution betaProfile [beta] */
    double coreRadiusOverVirialRadius; /* hotHaloMassDistributionCoreRadius virialFraction */
    double hotHaloScaleMassRelative, hotHaloScaleRadiusRelative; /* hot_halo/standard/_class.F90:226-227 */
    double outflowStrippingEfficiency; /* hotHaloOutflowStripping standard [efficiency] */
    double reincorporationMultiplier;  /* hotHaloOutflowReincorporation haloDynamicalTime [multiplier] */
    double fractionLossAngularMomentum;/* coolingInfallTorque fixed */
    double coolingVelocityCutOff;      /* coolingRate whiteFrenk1991 [velocityCutOff] */
    double coolingDegreesOfFreedom;    /* coolingTime simple [degreesOfFreedom] */
    double rateMaximumExpulsion;       /* nodeOperator CGMCoolingHeating */
    int32_t excessHeatDrivesOutflow;
    int32_t allowNegativeCGMMass;      /* nodeOperator CGMAccretion */
    /* star formation: krumholz2009 + intgrtdSurfaceDensity */
    double frequencyStarFormation, clumpingFactorMolecularComplex;
    double sfrIntegrationTolerance;    /* starFormationRateDisks intgrtdSurfaceDensity [tolerance] */
    double krumholzSTruncation;        /* s at which f_H2 drops to 1e-10: a constructor-time constant of
                                          starFormationRateSurfaceDensityDisksKrumholz2009
                                          (Krumholz2009.F90:236-241), computed by the host */
    /* star formation spheroids: timescale dynamicalTime */
    double sfSpheroidEfficiency, sfSpheroidExponentVelocity, sfSpheroidTimescaleMinimum;
    /* stellar feedback (power law inside rate limit) */
    double fbDiskVelocityCharacteristic, fbDiskExponent;
    double fbSpheroidVelocityCharacteristic, fbSpheroidExponent;
    double fbTimescaleOutflowFractionalMinimum;
    /* disk / spheroid components */
    double diskToleranceAbsoluteMass, spheroidToleranceAbsoluteMass;
    double spheroidRatioAngularMomentumScaleRadius, spheroidEfficiencyEnergeticOutflow;
    /* galactic structure */
    double structureSolutionTolerance; /* galacticStructureSolverEquilibrium [solutionTolerance] */
    double adiabaticA, adiabaticOmega;  /* darkMatterProfile adiabaticGnedin2004 */
    int32_t includeBaryonGravity;
    int32_t adiabaticContraction;       /* 1: adiabaticGnedin2004, 0: darkMatterOnly */
    /* bar instability efstathiou1982 */
    double barStabilityThresholdGaseous, barStabilityThresholdStellar;
    /* black holes */
    double bhSeedMass, bhSeedSpin;
    double bondiHoyleAccretionEnhancementSpheroid, bondiHoyleAccretionEnhancementHotHalo;
    double bondiHoyleAccretionTemperatureSpheroid;
    int32_t bondiHoyleAccretionHotModeOnly;
    int32_t pad1;
    double bhEfficiencyWind, bhEfficiencyRadioMode;
    double accretionRateThinDiskMaximum, accretionRateThinDiskMinimum;
    double adafEfficiencyRadiation, adafAdiabaticIndex;
    double accretionRateTransitionWidth;      /* accretionDisks switched [accretionRateTransitionWidth] (switched.F90:125-127) */
    int32_t scaleADAFRadiativeEfficiency;     /* accretionDisks switched */
    int32_t bhEfficiencyWindScalesWithEfficiencyRadiative; /* blackHoleWind ciotti2009 */
    int32_t adafEfficiencyRadiationTypeThinDisk; /* accretionDisksADAF [efficiencyRadiationType]: 1 thinDisk, 0 fixed */
    /* operator enable mask (bit i = operator i of enum glc_operator); lets a host run
       reduced operator sets exactly as a reduced <nodeOperator value="multi"> would */
    uint32_t operatorMask;
    /* darkMatterProfileDMO: the dark-matter-only profile under the contraction (enum glc_dmo_profile) */
    int32_t darkMatterProfileDMO;
    /* galacticStructureSolverEquilibrium [velocityMaximumFactor] (equilibrium.F90:124-128): 0 = no cap */
    double structureVelocityMaximumFactor;
    /* tree level (glc_forest_evolve): mergerTreeEvolverStandard [timestepHostRelative/Absolute]
       (merger_trees/evolver/standard.F90:942-968) and mergerTreeEvolveTimestepSimple [timeStepRelative/Absolute]
       (merger_trees/evolve/timesteps/simple.F90) */
    double timestepHostRelative, timestepHostAbsolute;
    double timestepSimpleRelative, timestepSimpleAbsolute;
    /* systemClockMaximum of mergerTreeNodeEvolver%evolve (node_evolver/standard.F90:694-705,861-867) as a wall-clock
       budget in seconds for ONE batched call (0 = none): nodes not finished when it expires come back with
       GLC_STATUS_XCPU (errorStatusXCPU) */
    double wallClockMaximumSeconds;
    /* mergerTreeNodeEvolverStandard [profileOdeEvolver] (node_evolver/standard.F90:249-253): feed every successful step to the
       profiler (standardStepErrorAnalyzer :1187-1239 -> mergerTreeEvolveProfilerSimple, merger_trees/evolve/profiler/
       simple.F90 [timeStepMinimum, timeStepMaximum, timeStepPointsPerDecade]); read back with glc_profiler_read */
    int32_t profileOdeEvolver;
    int32_t profilerTimeStepPointsPerDecade;
    double profilerTimeStepMinimum, profilerTimeStepMaximum;
} glc_params;

enum glc_dmo_profile {
    GLC_DMO_NFW = 0,        /* darkMatterProfileDMO value="NFW" (quickTest.xml:88) */
    GLC_DMO_ISOTHERMAL = 1  /* value="isothermal" (testSuite/parameters/reproducibility/adiabaticContraction.xml) */
};

enum glc_operator {
    GLC_OP_STAR_FORMATION_DISKS      = 1u << 0,
    GLC_OP_STAR_FORMATION_SPHEROIDS  = 1u << 1,
    GLC_OP_STELLAR_FEEDBACK_DISKS    = 1u << 2,
    GLC_OP_STELLAR_FEEDBACK_SPHEROIDS= 1u << 3,
    GLC_OP_BAR_INSTABILITY           = 1u << 4,
    GLC_OP_BLACK_HOLES_SEED          = 1u << 5,
    GLC_OP_BLACK_HOLES_ACCRETION     = 1u << 6,
    GLC_OP_BLACK_HOLES_WINDS         = 1u << 7,
    GLC_OP_CGM_ACCRETION             = 1u << 8,
    GLC_OP_CGM_OUTFLOW_REINCORPORATION = 1u << 9,
    GLC_OP_CGM_COOLING_HEATING       = 1u << 10,
    GLC_OP_CGM_OUTER_RADIUS          = 1u << 11,
    GLC_OP_SATELLITE_MASS_LOSS       = 1u << 12,
    GLC_OP_ALL                       = 0x1fffu
};

/* ---------------------------------------------------------------------------------
 * Tables (read-only inputs uploaded once).
 * ------------------------------------------------------------------------------- */
enum glc_table {
    /* cooling function: CIE table, format of cooling/cooling_function/CIE_file.F90:30-134.
       x0 = metallicities[n0] (linear, Solar units; first may be 0), x1 = temperatures[n1] (K),
       values[n0][n1] = Lambda (erg cm^3 s^-1, i.e. cooling function / n_H^2). */
    GLC_TABLE_COOLING_FUNCTION = 0,
    /* electron density / n_H on the same kind of grid (chemical/state/CIE_file.F90) */
    GLC_TABLE_ELECTRON_FRACTION = 1,
    /* halo mean density table, linear in ln t (dark_matter_halos/scales/
       virial_density_contrast.F90:356-417): x0 = times[n0] (Gyr, log-uniform), n1 = 2:
       values[n0][0] = mean virial density (Msun/Mpc^3), values[n0][1] = its logarithmic
       growth rate d ln rho / dt = (d Delta/dt)/Delta - 3 H  (1/Gyr)  (:419-459). */
    GLC_TABLE_HALO_MEAN_DENSITY = 2,
    /* exponential-disk rotation-curve factor x^2 [I0(x)K0(x) - I1(x)K1(x)] tabulated on a log-uniform
       lattice of half-radii x = R/2R_d, 100 points per decade (table1DLogarithmicLinear,
       mass_distributions/cylindrical/exponential_disk.F90:675-733): x0 = half-radii[n0], n1 = 1. */
    GLC_TABLE_DISK_ROTATION_CURVE = 3,
    /* ADAF tabulations built by the accretionDisksADAF constructor (accretion_disks/ADAF.F90:394-447):
       table1DLogarithmicLinear in the "inverse spin" 1-j on [1e-6,1] (countTable = 10000 there), extrapolation
       "fix" at both ends: x0 = (1-j)[n0] (log-uniform), n1 = 2: values[n0][0] = jet power per unit accretion
       rate ((km/s)^2, adafTablePowerJet), values[n0][1] = spin-up to mass-growth ratio (adafTableRateSpinUp). */
    GLC_TABLE_ADAF = 4,
    GLC_NTABLES
};

typedef struct glc_counters {
    uint64_t steps_accepted; /* successful iterations of driver2.c:190 loop (the metric's "node-ODE steps") */
    uint64_t steps_rejected; /* HADJ_DEC retries inside gsl_odeiv2_evolve_apply */
    uint64_t rhs_evaluations;
    uint64_t segments;       /* calls of standardEvolve (one per interrupt-delimited segment) */
    uint64_t trials_failed;  /* trialCount increments (standard.F90:686-692) */
    uint64_t nodes;
} glc_counters;

/* what mergerTreeEvolveProfilerSimple accumulates (merger_trees/evolve/profiler/simple.F90:250-304): per bin of the step
 * size (bin i holds time_step[i] <= h < time_step[i+1], searchArray) the number of successful steps and of solver
 * evaluations (evolve_apply calls up to and including the successful one), the same for steps taken after an interrupt
 * was found, the number of steps limited by each property (propertyHits; index = enum glc_prop) and the smallest step */
#define GLC_PROFILE_BINS 32
typedef struct glc_profile {
    int32_t n_bins, pad;
    double time_step[GLC_PROFILE_BINS];
    uint64_t time_step_count[GLC_PROFILE_BINS], evaluation_count[GLC_PROFILE_BINS];
    uint64_t time_step_count_interrupted[GLC_PROFILE_BINS], evaluation_count_interrupted[GLC_PROFILE_BINS];
    uint64_t property_hits[GLC_NY];
    uint64_t property_hits_unknown; /* steps with zero error in every property ("unknown") */
    double time_step_smallest;
} glc_profile;

typedef struct glc_evolver glc_evolver; /* opaque; owns device arena, tables, stream */

/* life cycle.  Return 0 on success, negative glc error code otherwise; never abort. */
int glc_evolver_create(glc_evolver **out, int32_t device_ordinal);
int glc_evolver_destroy(glc_evolver *ev);
const char *glc_last_error(const glc_evolver *ev);
int glc_abi_version(void);

/* replaces: parameter hand-off done by the objectBuilder / constructor of
 * mergerTreeNodeEvolverStandard (standard.F90:146-354) and of every physics class */
int glc_evolver_set_params(glc_evolver *ev, const glc_params *params);
/* fills *params with the defaults of parameters/quickTest.xml (model STANDARD) or
 * testSuite/parameters/reproducibility/closedBox.xml (model BOX) */
int glc_params_default(glc_params *params, int32_t model);

/* replaces: the read of the tabulated inputs (CIE_file.F90:535-663 etc.) */
int glc_evolver_set_table(glc_evolver *ev, int32_t table_id, int32_t n0, int32_t n1,
                          const double *x0, const double *x1, const double *values);

/*
 * glc_evolve_batch -- batched mergerTreeNodeEvolver%evolve.
 * replaces: n calls of standardEvolve(self,tree,node,timeEnd,interrupted,functionInterrupt,...)
 *           (node_evolver/standard.F90:385) issued by evolver/standard.F90:452.
 *   props      [n][GLC_NPROP] host, in/out   node records
 *   flags      [n]            host, in/out   component bits (creation interrupts set bits)
 *   time_end   [n]            host, in       timeEnd per node
 *   status     [n]            host, out      enum glc_status
 *   interrupt  [n]            host, out      enum glc_interrupt (GLC_INT_NONE if time_end reached)
 *   counters                  host, out      accumulated over the batch (may be NULL)
 */
int glc_evolve_batch(glc_evolver *ev, int64_t n, double *props, int32_t *flags,
                     const double *time_end, int32_t *status, int32_t *interrupt,
                     glc_counters *counters);

/*
 * Device-resident variant: the arena is SoA in HBM, [GLC_NPROP][capacity] doubles.
 * glc_arena_* move node records between host (node-major) and the arena;
 * glc_evolve_arena runs the hot path on records already resident in HBM.
 */
int glc_arena_reserve(glc_evolver *ev, int64_t capacity);
int glc_arena_upload(glc_evolver *ev, int64_t n, const double *props, const int32_t *flags,
                     const double *time_end);
int glc_arena_download(glc_evolver *ev, int64_t n, double *props, int32_t *flags, int32_t *status,
                       int32_t *interrupt);
int glc_evolve_arena(glc_evolver *ev, int64_t n, glc_counters *counters);
/* keep / restore a device-side copy of the first n arena records (device-to-device), so that a resident
 * batch can be evolved repeatedly from the same initial state */
int glc_arena_snapshot(glc_evolver *ev, int64_t n);
int glc_arena_restore(glc_evolver *ev, int64_t n);
/* duration (ms, CUDA events on the evolver's stream) of the last glc_evolve_arena kernel */
float glc_last_kernel_ms(const glc_evolver *ev);
/* how the last micro-task-machine batch divided into its two kernels (measurement aid): out8 = {device ms of the
 * machine_kernel slices, device ms of the drain_kernel passes, rate-function evaluations of each, accepted steps of each,
 * nodes written back by each} */
int glc_last_phase_stats(const glc_evolver *ev, double *out8);
/* raw device pointers for zero-copy interop (e.g. torch.distributed/NCCL reductions) */
void *glc_arena_device_props(glc_evolver *ev);
int64_t glc_arena_capacity(const glc_evolver *ev);
/* execution options (not physics): */
enum glc_option {
    GLC_OPT_SLICE_BUDGET = 0, /* rate-function evaluations per lane per kernel launch (time slice); lanes park
                                 their solver state in HBM between slices.  0 = one launch runs to completion */
    GLC_OPT_SORT_QUEUE = 1,   /* 1 (default): hand nodes to lanes in component-sorted order */
    GLC_OPT_MICROTASK_MACHINE = 2, /* standard model: 1 = micro-task machine kernel (units of the rate function re-grouped
                                 across lanes every iteration), 0 = warp-synchronous kernel, 2 (default) = machine for
                                 batches large enough to fill its queues, warp-synchronous kernel for small ones */
    GLC_OPT_FOREST_SCHEDULE = 3  /* glc_forest_evolve: 1 (default) = asynchronous -- every host-and-satellites group cycles
                                 on its own over the streaming machine, a slow node delays only its group; 0 = bulk-
                                 synchronous rounds, one batched call per phase.  Same results bit for bit. */
};
int glc_evolver_set_option(glc_evolver *ev, int32_t option, int64_t value);
/* number of time slices (evolve-kernel launches) so far */
int64_t glc_slice_count(const glc_evolver *ev);
/* number of CUDA kernels of this library launched through this evolver since creation */
int64_t glc_kernel_launch_count(const glc_evolver *ev);
/* sustained FP64 FMA throughput of this device (TFLOP/s) from an in-library DFMA-chain microbenchmark;
 * the denominator of the FP64 roofline fraction (MEASURED_PEAKS.json has no FP64 entry) */
double glc_measure_fp64_peak_tflops(glc_evolver *ev);
void *glc_evolver_stream(glc_evolver *ev);

/*
 * Streaming interface: what the batching tree evolver uses in production so that the device never
 * drains between batches.  replaces: the per-node call sequence of evolver/standard.F90:398-577 over a SET of forests.
 *   glc_stream_begin    reserve an arena for `capacity` tickets and reset the node queue
 *   glc_stream_submit   append n node records to the queue (host, node-major as in glc_evolve_batch); *first_ticket = ticket
 *                       of the first one, the others follow consecutively
 *   glc_stream_run      ONE time slice (pops_per_warp unit executions per warp, 0 = default 4096); returns the number of
 *                       nodes finished so far and the accumulated counters
 *   glc_stream_collect  up to max_nodes finished, not yet collected nodes: tickets, records, flags, status, interrupt
 *   glc_stream_finish   run until every submitted node is finished (machine slices + drain hand-over)
 *   glc_stream_end      close the session
 */
int glc_stream_begin(glc_evolver *ev, int64_t capacity);
int glc_stream_submit(glc_evolver *ev, int64_t n, const double *props, const int32_t *flags, const double *time_end,
                      int64_t *first_ticket);
int glc_stream_run(glc_evolver *ev, int32_t pops_per_warp, int64_t *n_finished_total, glc_counters *counters);
int glc_stream_collect(glc_evolver *ev, int64_t max_nodes, int64_t *tickets, double *props, int32_t *flags,
                       int32_t *status, int32_t *interrupt, int64_t *n_out);
int glc_stream_finish(glc_evolver *ev, glc_counters *counters);
int glc_stream_end(glc_evolver *ev);

/* replaces: mergerTreeEvolveProfiler (the metaData/evolverProfiler output of a profileOdeEvolver=true run).  Accumulates over
 * all batches since the last reset. */
int glc_profiler_read(glc_evolver *ev, glc_profile *out);
int glc_profiler_reset(glc_evolver *ev);

/* replaces: standardErrorHandler (node_evolver/standard.F90:1063-1140) with standardODEStepTolerances (:1142-1158): the
 * "ODE system parameters" table the reference prints for a node whose evolution failed on its last trial (the node comes
 * back with GLC_STATUS_UNDERFLOW at its saved values).  Per property: y, dy/dt at the node's time (standardODEs), yScale
 * (propertyScalesActive), yTolerance = odeToleranceRelative |y| + odeToleranceAbsolute yScale, yError and
 * |yError| / yTolerance.  The reference reads yError from the failed solver (solver_%errors); the batched solver keeps no
 * state of a failed node, so yError is the embedded Cash-Karp error estimate of ONE step of size time_step from the
 * node's state after the pre-evolve hooks (pass the step to be diagnosed, e.g. the record's GLC_P_TIME_STEP).
 * active[i] = 0 for properties that are not part of the node's ODE system (their other entries are 0). */
typedef struct glc_error_report {
    double time, time_step;
    int32_t active[GLC_NY];
    double y[GLC_NY], dydt[GLC_NY], scale[GLC_NY], tolerance[GLC_NY], error[GLC_NY], error_scaled[GLC_NY];
    int32_t interrupt, pad; /* functionInterrupt raised by the evaluation at the node's time (enum glc_interrupt), if any */
} glc_error_report;
int glc_error_report_node(glc_evolver *ev, const double *record, int32_t flags, double time_step, glc_error_report *out);

/* one evaluation of the RHS (standardODEs) for each node, for unit-level parity tests:
 *   dydt [n][GLC_NY] host out; props/flags are not modified except radii warm starts. */
int glc_rhs_batch(glc_evolver *ev, int64_t n, double *props, const int32_t *flags,
                  double *dydt, int32_t *interrupt);

/* histogram accumulation on device for the end-of-run reduction
 * (mirrors output/analyses/volume_function_1d.F90:986-987): bins a property of the arena */
int glc_histogram_accumulate(glc_evolver *ev, int64_t n, int32_t prop, double log10_min,
                             double log10_max, int32_t n_bins, double *device_hist);

/*
 * Forest interface: the batching tree evolver ("mergerTreeEvolverB200").
 * replaces: mergerTreeEvolverStandard::evolve (merger_trees/evolver/standard.F90:291-635) over a SET of forests -- the
 *           evolvability test (:723-760), the timeEvolveTo limits (:762-1035), standardPromote / standardMerge
 *           (node_evolver/standard.F90:1241-1356) and the node-operator hooks called there; every round all nodes
 *           allowed to move are evolved by ONE batched call of the node evolver.
 * Forests arrive as flat arrays (node i: parent index or -1 for a root, halo mass [Msun], cosmic time [Gyr], dark-matter
 * scale radius [Mpc], halo angular momentum [Msun Mpc km/s]); the most massive progenitor of a node is its primary.
 *   records [n][GLC_NPROP] host, out   final node records (valid where state is ISOLATED or SATELLITE)
 *   flags   [n]            host, out   component bits
 *   state   [n]            host, out   enum glc_forest_node_state
 * Requires resolveInterruptsOnDevice = 1 and the GLC_TABLE_HALO_MEAN_DENSITY table.
 */
enum glc_forest_node_state {
    GLC_FOREST_NODE_PENDING = 0,   /* still has progenitors (only before/during the run) */
    GLC_FOREST_NODE_ISOLATED = 1,  /* alive, not a satellite (at the end: the root galaxies) */
    GLC_FOREST_NODE_SATELLITE = 2, /* alive, hosted by another node */
    GLC_FOREST_NODE_PROMOTED = 3   /* was promoted into its parent (treeNode destroyed, standardPromote) */
};
typedef struct glc_forest_counters {
    uint64_t trees, nodes;
    uint64_t rounds;       /* batched evolve rounds (satellites, then hosts) */
    uint64_t evolve_calls; /* mergerTreeNodeEvolver%evolve calls the reference would have issued */
    uint64_t promotions, node_mergers; /* integer bookkeeping: must match the reference exactly */
    uint64_t failed_evolves; /* evolve calls that came back with a status other than success (the reference aborts the
                                run there, standard.F90:697-722); the node is moved to its end time and the walk goes on */
} glc_forest_counters;
int glc_forest_evolve(glc_evolver *ev, int64_t n_nodes, const int32_t *parent, const double *mass, const double *time,
                      const double *scale_radius, const double *angular_momentum, double *records, int32_t *flags,
                      int32_t *state, glc_forest_counters *forest_counters, glc_counters *counters);

#ifdef __cplusplus
}
#endif
#endif /* GLC_B200_H */
