// C ABI over the drop-in NDTFrame (include/ndtpso_frames.h).  Exceptions never cross the boundary.
#include "ndtpso_frames.h"

#include <exception>
#include <string>

#include "ndtpso_slam/core.h"
#include "ndtpso_slam/ndtframe.h"

struct ndtpso_frame {
  NDTFrame frame;
  ndtpso_frame(const double* t, int w, int h, double side, bool zero)
      : frame(Vector3d(t[0], t[1], t[2]), static_cast<unsigned short>(w), static_cast<unsigned short>(h), side, zero) {}
};

namespace ndtpso_b200 {
void shim_draw_rand(int32_t* out, size_t n);
}
namespace {
std::string g_err;
template <class F>
int guarded(F f) {
  try {
    f();
    return NDTPSO_OK;
  } catch (const std::exception& e) {
    g_err = e.what();
    return NDTPSO_ERR_CUDA;
  } catch (...) {
    g_err = "unknown exception";
    return NDTPSO_ERR_CUDA;
  }
}
}  // namespace

extern "C" {

ndtpso_frame* ndtpso_frame_new(const double* trans, int width_m, int height_m, double cell_side, int calculate_cells_params) {
  static const double zero[3] = {0., 0., 0.};
  try {
    return new ndtpso_frame(trans ? trans : zero, width_m, height_m, cell_side, calculate_cells_params != 0);
  } catch (...) {
    return nullptr;
  }
}
void ndtpso_frame_free(ndtpso_frame* f) { delete f; }

void ndtpso_frame_load_laser(ndtpso_frame* f, const float* ranges, int n, float angle_min, float angle_increment, float range_max) {
  std::vector<float> r(ranges, ranges + n);
  f->frame.loadLaser(r, angle_min, angle_increment, range_max);
}
void ndtpso_frame_update(ndtpso_frame* f, const double* pose, ndtpso_frame* new_frame) {
  f->frame.update(Vector3d(pose[0], pose[1], pose[2]), &new_frame->frame);
}
void ndtpso_frame_build(ndtpso_frame* f) { f->frame.build(); }
void ndtpso_frame_reset_cells(ndtpso_frame* f) { f->frame.resetCells(); }
int ndtpso_frame_is_built(const ndtpso_frame* f) { return f->frame.built ? 1 : 0; }
void ndtpso_frame_map_view(const ndtpso_frame* f, ndtpso_map_view* out) { f->frame.mapView(out); }
void ndtpso_frame_sparse_map_view(const ndtpso_frame* f, ndtpso_map_view* out) { f->frame.sparseMapView(out); }
int ndtpso_frame_scan_points(const ndtpso_frame* f, const double** out_xy) {
  const auto& pts = f->frame.scanPoints();
  if (out_xy) *out_xy = pts.empty() ? nullptr : reinterpret_cast<const double*>(pts.data());
  return static_cast<int>(pts.size());
}
int64_t ndtpso_frame_point_count(const ndtpso_frame* f) { return static_cast<int64_t>(f->frame.pointCount()); }

int ndtpso_frame_align(ndtpso_frame* ref_frame, const double* guess, ndtpso_frame* new_frame, double* out_pose) {
  return guarded([&]() {
    const Vector3d p = ref_frame->frame.align(Vector3d(guess[0], guess[1], guess[2]), &new_frame->frame);
    out_pose[0] = p.x();
    out_pose[1] = p.y();
    out_pose[2] = p.z();
  });
}
int ndtpso_frame_align_conf(ndtpso_frame* ref_frame, const double* guess, ndtpso_frame* new_frame, const ndtpso_pso_config* conf,
                            double* out_pose) {
  if (!conf) return ndtpso_frame_align(ref_frame, guess, new_frame, out_pose);
  return guarded([&]() {
    PSOConfig c;
    c.iterations = conf->iterations;
    c.populationSize = conf->population;
    c.num_threads = conf->num_threads;
    c.coeff.w = conf->w;
    c.coeff.c1 = conf->c1;
    c.coeff.c2 = conf->c2;
    c.coeff.w_dumping = conf->w_dumping;
    const Vector3d p = ref_frame->frame.align(Vector3d(guess[0], guess[1], guess[2]), &new_frame->frame, c);
    out_pose[0] = p.x();
    out_pose[1] = p.y();
    out_pose[2] = p.z();
  });
}
int ndtpso_frame_glir(ndtpso_frame* ref_frame, const double* guess, ndtpso_frame* new_frame, unsigned int iterations, const double* deviation,
                      double* out_pose) {
  return guarded([&]() {
    const Vector3d p = glir_pso_optimization(Vector3d(guess[0], guess[1], guess[2]), &ref_frame->frame, &new_frame->frame, iterations,
                                             Array3d(deviation[0], deviation[1], deviation[2]));
    out_pose[0] = p.x();
    out_pose[1] = p.y();
    out_pose[2] = p.z();
  });
}
int ndtpso_frame_cost(ndtpso_frame* ref_frame, ndtpso_frame* new_frame, const double* pose, double* out_cost) {
  return guarded([&]() { *out_cost = cost_function(Vector3d(pose[0], pose[1], pose[2]), &ref_frame->frame, &new_frame->frame); });
}
void ndtpso_frame_add_pose(ndtpso_frame* f, double timestamp, const double* pose) {
  f->frame.addPose(timestamp, Vector3d(pose[0], pose[1], pose[2]));
}
void ndtpso_frame_dump_map(ndtpso_frame* f, const char* filename) { f->frame.dumpMap(filename, true, true, false); }
int ndtpso_frame_device_resident(const ndtpso_frame* f) { return f->frame.deviceResident() ? 1 : 0; }
void ndtpso_frame_last_h2d_bytes(const ndtpso_frame* f, int64_t* align_bytes, int64_t* update_bytes) {
  if (align_bytes) *align_bytes = static_cast<int64_t>(f->frame.lastAlignH2DBytes());
  if (update_bytes) *update_bytes = static_cast<int64_t>(f->frame.lastUpdateH2DBytes());
}
int ndtpso_frame_download_device_map(ndtpso_frame* f, double* mean, double* inv_cov, uint8_t* built) {
  int ok = 0;
  const int rc = guarded([&]() { ok = f->frame.downloadDeviceMap(mean, inv_cov, built) ? 1 : 0; });
  return rc != NDTPSO_OK ? rc : (ok ? NDTPSO_OK : NDTPSO_ERR_ARG);
}
void ndtpso_frame_draw_rand(int32_t* out, int64_t n) {
  if (out && n > 0) ndtpso_b200::shim_draw_rand(out, static_cast<size_t>(n));
}
void ndtpso_frame_set_failure_mode(int keep_going) {
  // keep_going: a failure of the device path is recorded (ndtpso_frame_last_error) and the call returns its neutral result
  pso_set_failure_handler(keep_going ? +[](const char* what) { g_err = what; } : nullptr);
}
double ndtpso_frame_last_cost(void) { return pso_last_cost(); }
const char* ndtpso_frame_last_error(void) { return g_err.c_str(); }

}  // extern "C"
