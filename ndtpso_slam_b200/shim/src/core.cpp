// pso_optimization / cost_function of the drop-in library: flatten the two frames into the C ABI's
// POD views (zero-copy for the map table) and run on the GPU.  There is no CPU fallback: if the
// device path fails the process is told so and stops, like any other fatal runtime error.
#include "ndtpso_slam/core.h"

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "ndtpso_b200.h"

namespace {

ndtpso_ctx* g_ctx = nullptr;
double g_last_cost = 0.;

ndtpso_ctx* context() {
  if (!g_ctx) {
    const char* dev = std::getenv("NDTPSO_DEVICE");
    const int rc = ndtpso_ctx_create(dev ? std::atoi(dev) : 0, &g_ctx);
    if (rc != NDTPSO_OK) throw std::runtime_error("ndtpso_b200: no usable CUDA device (status " + std::to_string(rc) + "); there is no CPU path");
    std::atexit([]() {
      ndtpso_ctx_destroy(g_ctx);
      g_ctx = nullptr;
    });
  }
  return g_ctx;
}

}  // namespace

namespace ndtpso_b200 {
// the process-wide context of the drop-in library; nullptr (no exception) without a usable CUDA device
ndtpso_ctx* shim_context_or_null() {
  try {
    return context();
  } catch (const std::exception&) {
    return nullptr;
  }
}
void shim_set_last_cost(double c) { g_last_cost = c; }
}  // namespace ndtpso_b200

namespace {

void check(int rc, const char* what) {
  if (rc != NDTPSO_OK) throw std::runtime_error(std::string("ndtpso_b200: ") + what + ": " + ndtpso_last_error(g_ctx));
}

void fill_problem(ndtpso_problem* p, NDTFrame* ref_frame, const NDTFrame* new_frame) {
  if (!ref_frame->built) ref_frame->build();  // the lazy rebuild of cost_function (core.cpp:27-28)
  ref_frame->sparseMapView(&p->map);
  const auto& pts = new_frame->scanPoints();
  static_assert(sizeof(Vector2d) == 2 * sizeof(double), "Vector2d must be two packed doubles");
  p->points_xy = pts.empty() ? nullptr : reinterpret_cast<const double*>(pts.data());
  p->n_points = static_cast<int32_t>(pts.size());
  p->seed = 0;
  p->rand_stream = nullptr;
  p->rand_count = 0;
  for (int k = 0; k < 3; ++k) p->guess[k] = p->deviation[k] = 0.;
}

}  // namespace

namespace {

Vector3d solve(Vector3d initial_guess, NDTFrame* ref_frame, const NDTFrame* new_frame, const Array3d& deviation, const ndtpso_pso_config& cf) {
  ndtpso_ctx* ctx = context();
  ndtpso_problem p;
  fill_problem(&p, ref_frame, new_frame);
  for (int k = 0; k < 3; ++k) {
    p.guess[k] = initial_guess[k];
    p.deviation[k] = deviation[k];
  }
  // the reference's random numbers: the next 3 + 3P + 6PI (GLIR: 3(P + 2) + 6PI) outputs of the process-global std::rand()
  const int64_t n = ndtpso_rand_draws(&cf);
  std::vector<int32_t> stream(static_cast<size_t>(n));
  for (auto& r : stream) r = std::rand();
  p.rand_stream = stream.data();
  p.rand_count = n;
  double pose[3], cost = 0.;
  check(ndtpso_align_batch(ctx, 1, &p, &cf, pose, &cost), "ndtpso_align_batch");
  g_last_cost = cost;
  return Vector3d(pose[0], pose[1], pose[2]);
}

}  // namespace

Vector3d pso_optimization(Vector3d initial_guess, NDTFrame* ref_frame, const NDTFrame* const new_frame, const Array3d& deviation,
                          const PSOConfig& pso_conf) {
  ndtpso_pso_config cf;
  cf.iterations = pso_conf.iterations;
  cf.population = pso_conf.populationSize;
  cf.num_threads = pso_conf.num_threads;
  cf.variant = NDTPSO_VARIANT_PSO;
  cf.w = pso_conf.coeff.w;
  cf.c1 = pso_conf.coeff.c1;
  cf.c2 = pso_conf.coeff.c2;
  cf.w_dumping = pso_conf.coeff.w_dumping;
  return solve(initial_guess, ref_frame, new_frame, deviation, cf);
}

Vector3d glir_pso_optimization(Vector3d initial_guess, NDTFrame* ref_frame, NDTFrame* new_frame, unsigned int iters_num, const Array3d& deviation) {
  ndtpso_pso_config cf;
  ndtpso_pso_config_default(&cf);
  cf.iterations = static_cast<int32_t>(iters_num);
  cf.population = PSO_POPULATION_SIZE;  // core.cpp:134,145
  cf.variant = NDTPSO_VARIANT_GLIR;
  return solve(initial_guess, ref_frame, new_frame, deviation, cf);
}

double cost_function(Vector3d trans, NDTFrame* const ref_frame, const NDTFrame* const new_frame) {
  ndtpso_ctx* ctx = context();
  ndtpso_problem p;
  fill_problem(&p, ref_frame, new_frame);
  const double pose[3] = {trans.x(), trans.y(), trans.z()};
  double cost = 0.;
  check(ndtpso_cost_batch(ctx, 1, &p, 1, pose, &cost), "ndtpso_cost_batch");
  return cost;
}

double pso_last_cost() { return g_last_cost; }
