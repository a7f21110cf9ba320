// TEST INFRASTRUCTURE — not product code.
//
// C-ABI harness around the UNMODIFIED reference library (libndtpso_slam), compiled
// by oracle/Makefile from the sources where they lie under /root/reference into
// oracle/_ref/libndtpso_ref.so.  It exposes NDTFrame construction, loadLaser,
// update, build, align, pso_optimization and cost_function to ctypes, plus
// flattening of the private fields the hot path reads, so that
//   * tests/golden/make_golden.py can generate golden input/output vectors,
//   * tests can pin oracle/ndtpso_oracle.c against the real reference,
//   * bench.py's reference arm can time the reference's own CPU path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.
//
// Reference entry points wrapped (file:line under /root/reference):
//   NDTFrame::NDTFrame            lib/ndtpso_slam/ndtframe.cpp:19-66
//   NDTFrame::loadLaser           lib/ndtpso_slam/ndtframe.cpp:144-185
//   NDTFrame::update              lib/ndtpso_slam/ndtframe.cpp:187-198
//   NDTFrame::build               lib/ndtpso_slam/ndtframe.cpp:68-117
//   NDTFrame::align               lib/ndtpso_slam/ndtframe.cpp:251-266
//   pso_optimization              lib/ndtpso_slam/core.cpp:50-116
//   cost_function                 lib/ndtpso_slam/core.cpp:26-48
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <utility>
#include <vector>

#include <omp.h>

// The hot path reads NDTCell::s_inv_covar and NDTFrame::s_x_min.. which are
// private in the reference headers; open them up for this translation unit only.
#define private public
#include "ndtpso_slam/core.h"
#include "ndtpso_slam/ndtframe.h"
#undef private

extern "C" {

void* ref_frame_new(double tx, double ty, double tth, int width, int height, double cell_side, int init_windows) {
  return new NDTFrame(Vector3d(tx, ty, tth), static_cast<unsigned short>(width), static_cast<unsigned short>(height), cell_side,
                      init_windows != 0);
}

void ref_frame_free(void* f) { delete static_cast<NDTFrame*>(f); }

void ref_frame_load_laser(void* f, const float* ranges, int n, float angle_min, float angle_inc, float range_max) {
  std::vector<float> r(ranges, ranges + n);
  static_cast<NDTFrame*>(f)->loadLaser(r, angle_min, angle_inc, range_max);
}

void ref_frame_update(void* f, const double* pose, void* new_frame) {
  static_cast<NDTFrame*>(f)->update(Vector3d(pose[0], pose[1], pose[2]), static_cast<NDTFrame*>(new_frame));
}

void ref_frame_build(void* f) { static_cast<NDTFrame*>(f)->build(); }

int ref_frame_is_built(void* f) { return static_cast<NDTFrame*>(f)->built ? 1 : 0; }

// geometry: out_i = {widthNumOfCells, heightNumOfCells, numOfCells, width, height}
//           out_d = {cell_side, x_min, x_max, y_min, y_max}
void ref_frame_geometry(void* f, int32_t* out_i, double* out_d) {
  NDTFrame* fr = static_cast<NDTFrame*>(f);
  out_i[0] = fr->widthNumOfCells;
  out_i[1] = fr->heightNumOfCells;
  out_i[2] = static_cast<int32_t>(fr->numOfCells);
  out_i[3] = fr->width;
  out_i[4] = fr->height;
  out_d[0] = fr->cell_side;
  out_d[1] = fr->s_x_min;
  out_d[2] = fr->s_x_max;
  out_d[3] = fr->s_y_min;
  out_d[4] = fr->s_y_max;
}

// Dense (mu, Sigma^-1, built) table in cell-index order.  inv_cov is row-major 00,01,10,11.
// Cells that are not built get zeros (their fields are uninitialised in the reference).
void ref_frame_flatten_map(void* f, double* mean, double* inv_cov, uint8_t* built) {
  NDTFrame* fr = static_cast<NDTFrame*>(f);
  for (unsigned i = 0; i < fr->numOfCells; ++i) {
    const NDTCell& c = fr->cells[i];
    built[i] = c.built ? 1 : 0;
    if (c.built) {
      mean[2 * i + 0] = c.mean[0];
      mean[2 * i + 1] = c.mean[1];
      inv_cov[4 * i + 0] = c.s_inv_covar(0, 0);
      inv_cov[4 * i + 1] = c.s_inv_covar(0, 1);
      inv_cov[4 * i + 2] = c.s_inv_covar(1, 0);
      inv_cov[4 * i + 3] = c.s_inv_covar(1, 1);
    } else {
      mean[2 * i + 0] = mean[2 * i + 1] = 0.;
      inv_cov[4 * i + 0] = inv_cov[4 * i + 1] = inv_cov[4 * i + 2] = inv_cov[4 * i + 3] = 0.;
    }
  }
}

// Points the cost function iterates (core.cpp:33-36): cells in index order, window slot 0.
int ref_frame_count_points(void* f) {
  NDTFrame* fr = static_cast<NDTFrame*>(f);
  size_t n = 0;
  for (auto& c : fr->cells) n += c.points_vector[0].size();
  return static_cast<int>(n);
}

void ref_frame_flatten_points(void* f, double* xy) {
  NDTFrame* fr = static_cast<NDTFrame*>(f);
  size_t k = 0;
  for (auto& c : fr->cells)
    for (auto& p : c.points_vector[0]) {
      xy[2 * k + 0] = p.x();
      xy[2 * k + 1] = p.y();
      ++k;
    }
}

// All points of all windows of all cells (what dumpMap would write), for map-state checks.
int ref_frame_count_all_points(void* f) {
  NDTFrame* fr = static_cast<NDTFrame*>(f);
  size_t n = 0;
  for (auto& c : fr->cells)
    for (auto& v : c.points_vector) n += v.size();
  return static_cast<int>(n);
}

void ref_srand(unsigned seed) { std::srand(seed); }
int ref_rand(void) { return std::rand(); }

double ref_cost(void* ref, void* cur, const double* pose) {
  return cost_function(Vector3d(pose[0], pose[1], pose[2]), static_cast<NDTFrame*>(ref), static_cast<NDTFrame*>(cur));
}

// pso_optimization with an explicit PSOConfig.  use_seed != 0: srand(seed) first.
// Returns wall seconds of the pso_optimization call alone.
double ref_pso(void* ref, void* cur, const double* guess, const double* dev, int population, int iterations, int num_threads, double w,
               double c1, double c2, double w_dumping, int use_seed, unsigned seed, double* out_pose) {
  PSOConfig conf;
  conf.iterations = iterations;
  conf.populationSize = population;
  conf.num_threads = num_threads;
  conf.coeff.w = w;
  conf.coeff.c1 = c1;
  conf.coeff.c2 = c2;
  conf.coeff.w_dumping = w_dumping;
  NDTFrame* r = static_cast<NDTFrame*>(ref);
  if (!r->built) r->build();  // keep the lazy map build out of the timed region
  if (use_seed) std::srand(seed);
  auto t0 = std::chrono::high_resolution_clock::now();
  Vector3d p = pso_optimization(Vector3d(guess[0], guess[1], guess[2]), r, static_cast<NDTFrame*>(cur),
                                Array3d(dev[0], dev[1], dev[2]), conf);
  auto t1 = std::chrono::high_resolution_clock::now();
  out_pose[0] = p.x();
  out_pose[1] = p.y();
  out_pose[2] = p.z();
  return std::chrono::duration<double>(t1 - t0).count();
}

// The production entry: NDTFrame::align (default PSOConfig, process-global rand()).
void ref_align(void* ref, const double* guess, void* cur, double* out_pose) {
  Vector3d p = static_cast<NDTFrame*>(ref)->align(Vector3d(guess[0], guess[1], guess[2]), static_cast<NDTFrame*>(cur));
  out_pose[0] = p.x();
  out_pose[1] = p.y();
  out_pose[2] = p.z();
}

// glir_pso_optimization (core.cpp:118-186, "UNTESTED" in the reference; population fixed to PSO_POPULATION_SIZE).
void ref_glir(void* ref, void* cur, const double* guess, const double* dev, int iterations, int use_seed, unsigned seed, double* out_pose) {
  NDTFrame* r = static_cast<NDTFrame*>(ref);
  if (!r->built) r->build();
  if (use_seed) std::srand(seed);
  Vector3d p = glir_pso_optimization(Vector3d(guess[0], guess[1], guess[2]), r, static_cast<NDTFrame*>(cur),
                                     static_cast<unsigned>(iterations), Array3d(dev[0], dev[1], dev[2]));
  out_pose[0] = p.x();
  out_pose[1] = p.y();
  out_pose[2] = p.z();
}

int ref_omp_max_threads(void) { return omp_get_max_threads(); }
// pso_optimization calls omp_set_num_threads(n) whenever PSOConfig::num_threads limits it (core.cpp:75-79), which lowers what
// omp_get_max_threads() reports from then on: a later "all threads" run needs the team size put back.
void ref_omp_set_num_threads(int n) { omp_set_num_threads(n); }

int ref_sizeof_cell(void) { return static_cast<int>(sizeof(NDTCell)); }

}  // extern "C"
