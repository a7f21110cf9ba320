// Drop-in for include/ndtpso_slam/core.h: the free functions of libndtpso_slam's scan matcher with
// their reference signatures (core.h:16-19,28-31,40-50).  pso_optimization and cost_function run
// on the GPU through the C ABI of include/ndtpso_b200.h; the inline helpers are host arithmetic.
#ifndef NDTPSO_B200_SHIM_CORE_H
#define NDTPSO_B200_SHIM_CORE_H

#include <cmath>
#include <eigen3/Eigen/Core>
#include <vector>

#include "ndtpso_slam/config.h"
#include "ndtpso_slam/ndtframe.h"

using Eigen::Array2d;
using Eigen::Array3d;
using Eigen::Matrix2d;
using Eigen::Vector2d;
using Eigen::Vector3d;
using std::vector;

// Random numbers: the reference consumes the process-global std::rand() stream (3 + 3P + 6PI draws
// per call, core.cpp:14,58-69,84).  This function draws exactly those values on the host, in the
// same order, and hands them to the device, so a sequence of calls stays in lock-step with a CPU
// build of the reference.
Vector3d pso_optimization(Vector3d initial_guess, NDTFrame* ref_frame, const NDTFrame* const new_frame,
                          const Array3d& deviation = {0, 0, 0}, const PSOConfig& pso_conf = PSOConfig());

// The GLIR-PSO variant with its reference signature (core.h:21-23, core.cpp:118-186; the reference never calls it).  As
// there: population PSO_POPULATION_SIZE, 3(P + 2) + 6PI draws of std::rand(), made on the host in the same order.
Vector3d glir_pso_optimization(Vector3d initial_guess, NDTFrame* ref_frame, NDTFrame* new_frame, unsigned int iters_num = 50,
                               const Array3d& deviation = {0, 0, 0});

double cost_function(Vector3d trans, NDTFrame* const ref_frame, const NDTFrame* const new_frame);

// best cost of the most recent pso_optimization call (the reference only prints it, core.cpp:111-114)
double pso_last_cost();

// Extension (not in the reference, whose CPU path cannot fail): what happens when the device path fails — no usable CUDA
// device, a CUDA error, the pools of a device-resident map exhausted.  Without a handler (the default) the failing call throws
// std::runtime_error: there is no CPU fallback, and a wrong pose must not pass for a match.  With a handler installed it is
// called with the message; if it returns, pso_optimization / glir_pso_optimization / NDTFrame::align return the caller's
// initial guess ("no correction for this scan"), cost_function returns 0 (the cost of a scan that hits no built cell), and
// pso_last_cost() is NaN until the next successful call.  Returns the previous handler.
typedef void (*pso_failure_handler)(const char* what);
pso_failure_handler pso_set_failure_handler(pso_failure_handler handler);

// (x, y) rotated by trans.z() and shifted by (trans.x(), trans.y())
inline Vector2d transform_point(const Vector2d& point, const Vector3d& trans) {
  const double c = std::cos(trans.z()), s = std::sin(trans.z());
  return Vector2d(point.x() * c - point.y() * s + trans.x(), point.x() * s + point.y() * c + trans.y());
}

// beam index -> angle, in float like the sensor message (reference core.h:40-42)
inline float index_to_angle(unsigned int idx, float step, float min_angle) { return idx * step + min_angle; }

// polar -> cartesian in double from float inputs (reference core.h:45-47)
inline Vector2d laser_to_point(float r, float theta) {
  return Vector2d(double(r) * std::cos(double(theta)), double(r) * std::sin(double(theta)));
}

#endif
