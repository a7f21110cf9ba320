// Drop-in for include/ndtpso_slam/config.h of libndtpso_slam: the same configuration surface
// (macro names, struct names, field names, defaults) so that code written against the reference
// compiles unchanged.  Values cited from the reference's config.h:4-25.
#ifndef NDTPSO_B200_SHIM_CONFIG_H
#define NDTPSO_B200_SHIM_CONFIG_H

#define NDT_WINDOW_SIZE 100          // window slots per cell                 (reference config.h:8)
#define NDT_MAX_POINTS_PER_CELL 50   // a slot closes once it holds more      (reference config.h:5)
#define LASER_IGNORE_EPSILON 0.1f    // drop returns closer than 10 cm        (reference config.h:6)
#define BUILD_OCCUPANCY_GRID true    // keeps the reference's ctor/dumpMap arity; the grid itself is out of scope
#define TRANSFORM_POINTS_AT_LOAD true
#define TRANSFORM_POSE_AFTER_ALIGN (!TRANSFORM_POINTS_AT_LOAD)

#define PSO_ITERATIONS 50
#define PSO_POPULATION_SIZE 30
#define PSO_W_DUMPING_COEF 1.
#define PSO_W .8
#define PSO_C1 2.
#define PSO_C2 2.

struct PSOConfig {
  int iterations{PSO_ITERATIONS};
  int populationSize{PSO_POPULATION_SIZE};
  int num_threads{-1};  // accepted, unused: the device path equals the reference run with one thread
  struct {
    double w{PSO_W};
    double c1{PSO_C1};
    double c2{PSO_C2};
    double w_dumping{PSO_W_DUMPING_COEF};
  } coeff;
};

struct NDTPSOConfig {
  PSOConfig psoConfig;
  float laserIgnoreEpsilon{LASER_IGNORE_EPSILON};
};

#endif
