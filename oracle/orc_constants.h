/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 * Physical constants as the reference defines them: source/numerical/constants/
 * {astronomical,atomic,physical,units,math}.F90.  Constants the reference takes from
 * GSL (gslSymbol="GSL_CONST_MKSA_*") carry the values of gsl_const_mksa.h in GSL 2.6.
 * Cross-check: gravitationalConstant_internal must equal 4.3011827419096073e-9, the value
 * quoted by testSuite/test-reproducibility.py:15 (asserted in tests/test_oracle_constants.py).
 */
#ifndef ORC_CONSTANTS_H
#define ORC_CONSTANTS_H

#define ORC_PI 3.14159265358979323846
/* GSL_CONST_MKSA_* (GSL 2.6) */
#define ORC_SPEED_LIGHT 2.99792458e8
#define ORC_GRAVITATIONAL_CONSTANT 6.673e-11
#define ORC_PARSEC 3.08567758135e16
#define ORC_MASS_SOLAR 1.98892e30
#define ORC_BOLTZMANN 1.3806504e-23
#define ORC_ATOMIC_MASS_UNIT 1.660538782e-27
#define ORC_THOMSON_CROSS_SECTION 6.65245893699e-29
#define ORC_ELECTRON_VOLT 1.602176487e-19
/* prefixes */
#define ORC_KILO 1.0e3
#define ORC_MEGA 1.0e6
#define ORC_GIGA 1.0e9
#define ORC_HECTO 1.0e2
#define ORC_ERGS 1.0e-7
/* astronomical.F90 */
#define ORC_MEGAPARSEC (ORC_MEGA * ORC_PARSEC)
#define ORC_YEAR 3.15581497635456e7
#define ORC_GIGAYEAR (ORC_GIGA * ORC_YEAR)
#define ORC_G_INTERNAL (ORC_GRAVITATIONAL_CONSTANT * ORC_MASS_SOLAR / (ORC_KILO * ORC_KILO) / ORC_MEGAPARSEC)
#define ORC_MPC_PER_KMS_TO_GYR (ORC_MEGAPARSEC / ORC_KILO / ORC_GIGAYEAR)
#define ORC_HYDROGEN_BY_MASS_SOLAR 0.7070
#define ORC_HELIUM_BY_MASS_SOLAR 0.2740
#define ORC_METALLICITY_SOLAR 0.0188
#define ORC_HYDROGEN_BY_MASS_PRIMORDIAL 0.7514
#define ORC_HELIUM_BY_MASS_PRIMORDIAL 0.2486
/* atomic.F90 */
#define ORC_ATOMIC_MASS_HYDROGEN 1.0078250322
#define ORC_ATOMIC_MASS_HELIUM 4.0026032545
#define ORC_MASS_HYDROGEN_ATOM (ORC_ATOMIC_MASS_HYDROGEN * ORC_ATOMIC_MASS_UNIT)
#define ORC_MEAN_ATOMIC_MASS_PRIMORDIAL                                                   \
    (1.0 / (2.0 * ORC_HYDROGEN_BY_MASS_PRIMORDIAL / ORC_ATOMIC_MASS_HYDROGEN +            \
            3.0 * ORC_HELIUM_BY_MASS_PRIMORDIAL / ORC_ATOMIC_MASS_HELIUM))
/* stellar_astrophysics/feedback/_class.F90:56 */
#define ORC_FEEDBACK_ENERGY_INPUT_AT_INFINITY_CANONICAL 4.517e5

#endif
