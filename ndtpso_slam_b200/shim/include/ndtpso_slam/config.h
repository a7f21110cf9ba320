// Drop-in for include/ndtpso_slam/config.h of libndtpso_slam.
//
// Code written against the reference (the ROS node, src/ndtpso_slam_node.cpp:30-44) names these macros and the fields of
// these two structs; nothing else of the reference's configuration header is carried over (its logger and
// frontal-point switches configure code this library does not contain).  Every default is the reference's value,
// cited by line of its config.h.
#ifndef NDTPSO_B200_SHIM_CONFIG_H
#define NDTPSO_B200_SHIM_CONFIG_H

// ---- map cells (NDTCell, NDTFrame) ---------------------------------------------------------------------------------
// A cell keeps the partial statistics of its last NDT_WINDOW_SIZE slots; the current slot is closed, and the next one
// opened, by the first point that arrives after it holds more than NDT_MAX_POINTS_PER_CELL points.
#define NDT_WINDOW_SIZE 100         /* config.h:8 */
#define NDT_MAX_POINTS_PER_CELL 50  /* config.h:5 */
// NDTFrame::loadLaser drops returns closer to the sensor than this many metres.
#define LASER_IGNORE_EPSILON 0.1f   /* config.h:6 */
// The scan is moved by the frame's own transform when it is loaded (ndtframe.cpp:151-153), not after the match.
#define TRANSFORM_POINTS_AT_LOAD true            /* config.h:9 */
#define TRANSFORM_POSE_AFTER_ALIGN (!TRANSFORM_POINTS_AT_LOAD)
// Kept so that the constructor and dumpMap keep the reference's arity; the occupancy grid itself is not built here.
#define BUILD_OCCUPANCY_GRID true                /* config.h:12 */

// ---- swarm (pso_optimization, glir_pso_optimization) ------------------------------------------------------------------
#define PSO_POPULATION_SIZE 30  /* config.h:21; the population glir_pso_optimization always uses (core.cpp:134,145) */
#define PSO_ITERATIONS 50       /* config.h:20 */
#define PSO_W .8                /* inertia,                          config.h:23 */
#define PSO_W_DUMPING_COEF 1.   /* inertia decay per iteration,      config.h:22 */
#define PSO_C1 2.               /* pull towards the particle's best, config.h:24 */
#define PSO_C2 2.               /* pull towards the swarm's best,    config.h:25 */

// Field for field the reference's PSOConfig (config.h:27-38): same order, same types, same defaults, 48 bytes.
// include/ndtpso_b200.h's ndtpso_pso_config has this layout (the `variant` selector sits in the padding after num_threads).
struct PSOConfig {
  int iterations{PSO_ITERATIONS};
  int populationSize{PSO_POPULATION_SIZE};
  int num_threads{-1};  // accepted and unused: the device path gives what the reference gives with one thread
  struct Coefficients {
    double w{PSO_W}, c1{PSO_C1}, c2{PSO_C2}, w_dumping{PSO_W_DUMPING_COEF};
  } coeff;
};

// What the node fills from its ROS parameters (config.h:40-45).  NDTFrame::align(guess, frame) ignores psoConfig exactly as
// the reference does (ndtframe.cpp:257); align(guess, frame, psoConfig) is the opt-in that honours it.
struct NDTPSOConfig {
  PSOConfig psoConfig;
  float laserIgnoreEpsilon{LASER_IGNORE_EPSILON};
};

#endif
