// pso_optimization / cost_function of the drop-in library: flatten the two frames into the C ABI's
// POD views (zero-copy for the map table) and run on the GPU.  There is no CPU fallback: if the
// device path fails the call throws, unless the application installed a failure handler (core.h).
#include "ndtpso_slam/core.h"

#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "ndtpso_b200.h"

namespace {

ndtpso_ctx* g_ctx = nullptr;
double g_last_cost = 0.;
int g_ctx_status = NDTPSO_OK;
pso_failure_handler g_on_failure = nullptr;

// the process-wide context; nullptr without a usable CUDA device (g_ctx_status says why)
ndtpso_ctx* context_or_null() {
  if (!g_ctx) {
    const char* dev = std::getenv("NDTPSO_DEVICE");
    g_ctx_status = ndtpso_ctx_create(dev ? std::atoi(dev) : 0, &g_ctx);
    if (g_ctx_status != NDTPSO_OK) {
      g_ctx = nullptr;
      return nullptr;
    }
    std::atexit([]() {
      ndtpso_ctx_destroy(g_ctx);
      g_ctx = nullptr;
    });
  }
  return g_ctx;
}

}  // namespace

namespace ndtpso_b200 {
ndtpso_ctx* shim_context_or_null() { return context_or_null(); }
void shim_set_last_cost(double c) { g_last_cost = c; }
// The device path failed: throws unless a failure handler is installed (core.h); if the handler returns, so does this
// function, and the caller hands back its neutral result.
void shim_fail(const std::string& what) {
  g_last_cost = std::nan("");
  if (!g_on_failure) throw std::runtime_error(what);
  g_on_failure(what.c_str());
}
}  // namespace ndtpso_b200

namespace ndtpso_b200 {
// The next n outputs of the process-global std::rand(), which is advanced by exactly n draws — what n calls of rand() do, without
// n trips through glibc's lock (9 093 draws per default align: ~0.15 ms of a 0.7 ms callback).
// glibc keeps rand()'s state in a table the public API hands out: setstate() switches to another table and returns the previous
// one, having saved the generator's position in its first word ((rear index) * 5 + type).  For the default TYPE_3 generator
// (additive feedback, degree 31, separation 3: r[i] = r[i-31] + r[i-3], output = r[i] >> 1) the table is advanced here and
// handed back with setstate().  The first use checks itself against rand() on the same state; anything unexpected (another
// libc, another generator type, NDTPSO_SHIM_FAST_RAND=0) falls back to calling rand() n times.  Like rand() itself this is for
// one drawing thread at a time: a thread that calls rand() during the switch would draw from the scratch table.
void shim_draw_rand(int32_t* out, size_t n) {
#if defined(__GLIBC__)
  static int mode = -1;  // -1 untested, 0 off, 1 on
  alignas(8) static char scratch[128];
  static bool scratch_ready = false;
  if (mode == -1) {
    const char* e = std::getenv("NDTPSO_SHIM_FAST_RAND");
    if (e && std::atoi(e) == 0) mode = 0;
  }
  auto advance = [](int32_t* w, int32_t* dst, size_t m) {  // w = the table (first word: position and type)
    int32_t* st = w + 1;
    int r = w[0] / 5, f = (r + 3) % 31;
    for (size_t i = 0; i < m; ++i) {
      const uint32_t v = static_cast<uint32_t>(st[f]) + static_cast<uint32_t>(st[r]);
      st[f] = static_cast<int32_t>(v);
      dst[i] = static_cast<int32_t>(v >> 1);
      if (++f == 31) f = 0;
      if (++r == 31) r = 0;
    }
    w[0] = 5 * r + 3;
  };
  if (mode != 0 && n >= 4) {  // (shorter requests — a swarm without particles — are not worth the switch, and the self-check draws four)
    char* old = scratch_ready ? setstate(scratch) : initstate(1u, scratch, sizeof scratch);
    scratch_ready = true;
    int32_t* w = reinterpret_cast<int32_t*>(old);
    const bool type3 = old && (w[0] % 5) == 3 && w[0] >= 0 && w[0] / 5 < 31;
    if (type3 && mode == -1) {
      // self-check: four draws from a copy of the table must be the four draws rand() then makes from the table itself
      int32_t copy[32], mine[4];
      std::memcpy(copy, w, sizeof copy);
      advance(copy, mine, 4);
      setstate(old);
      bool same = true;
      for (size_t i = 0; i < 4; ++i) {
        out[i] = std::rand();
        same = same && out[i] == mine[i];
      }
      mode = same ? 1 : 0;
      if (n == 4) return;
      out += 4;
      n -= 4;
      if (mode == 0) {
        for (size_t i = 0; i < n; ++i) out[i] = std::rand();
        return;
      }
      old = setstate(scratch);
      w = reinterpret_cast<int32_t*>(old);
    }
    if (type3 && mode == 1) {
      advance(w, out, n);
      setstate(old);
      return;
    }
    if (old) setstate(old);
    if (!type3) mode = 0;
  }
#endif
  for (size_t i = 0; i < n; ++i) out[i] = std::rand();
}
}  // namespace ndtpso_b200

pso_failure_handler pso_set_failure_handler(pso_failure_handler handler) {
  const pso_failure_handler old = g_on_failure;
  g_on_failure = handler;
  return old;
}

namespace {

// the context, or nullptr after reporting why there is none
ndtpso_ctx* context() {
  ndtpso_ctx* ctx = context_or_null();
  if (!ctx) ndtpso_b200::shim_fail("ndtpso_b200: no usable CUDA device (status " + std::to_string(g_ctx_status) + "); there is no CPU path");
  return ctx;
}

// true when rc is a failure (reported; the caller returns its neutral result)
bool failed(int rc, const char* what) {
  if (rc == NDTPSO_OK) return false;
  ndtpso_b200::shim_fail(std::string("ndtpso_b200: ") + what + ": " + ndtpso_last_error(g_ctx));
  return true;
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
  if (!ctx) return initial_guess;
  ndtpso_problem p;
  fill_problem(&p, ref_frame, new_frame);
  for (int k = 0; k < 3; ++k) {
    p.guess[k] = initial_guess[k];
    p.deviation[k] = deviation[k];
  }
  // the reference's random numbers: the next 3 + 3P + 6PI (GLIR: 3(P + 2) + 6PI) outputs of the process-global std::rand()
  const int64_t n = ndtpso_rand_draws(&cf);
  std::vector<int32_t> stream(static_cast<size_t>(n));
  ndtpso_b200::shim_draw_rand(stream.data(), stream.size());
  p.rand_stream = stream.data();
  p.rand_count = n;
  double pose[3], cost = 0.;
  if (failed(ndtpso_align_batch(ctx, 1, &p, &cf, pose, &cost), "ndtpso_align_batch")) return initial_guess;
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
  if (!ctx) return 0.;
  ndtpso_problem p;
  fill_problem(&p, ref_frame, new_frame);
  const double pose[3] = {trans.x(), trans.y(), trans.z()};
  double cost = 0.;
  if (failed(ndtpso_cost_batch(ctx, 1, &p, 1, pose, &cost), "ndtpso_cost_batch")) return 0.;
  return cost;
}

double pso_last_cost() { return g_last_cost; }
