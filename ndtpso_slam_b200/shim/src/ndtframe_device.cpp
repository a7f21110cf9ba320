// The device mirror of the drop-in NDTFrame (see ndtframe.h): a map that is filled through update() only is kept in HBM by a
// one-frame ndtpso_dframes object (include/ndtpso_dframes.h), so the reference's per-scan callback
//     current_frame_->loadLaser(...); current_pose_ = ref_frame_->align(previous_pose_, current_frame_); ref_frame_->update(...)
// (src/ndtpso_slam_node.cpp:186-198) moves the scan, the random numbers and the pose over PCIe, never the table.
//
// The mirror is created by the first align(): until then update() only logs (pose, scan points), so frames that are built
// and never matched against (a map builder, a benchmark generator) cost nothing on the device.  The log is replayed into the new
// device frame with the host-computed points, so the device map starts bit-identical to the host's.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "ndtpso_b200.h"
#include "ndtpso_dframes.h"
#include "ndtpso_slam/core.h"
#include "ndtpso_slam/ndtframe.h"

namespace ndtpso_b200 {
ndtpso_ctx* shim_context_or_null();
void shim_set_last_cost(double c);
void shim_fail(const std::string& what);
void shim_draw_rand(int32_t* out, size_t n);
}  // namespace ndtpso_b200

struct NDTFrame::DeviceMirror {
  ndtpso_dframes* df = nullptr;  // null while the updates are only logged
  struct Logged {
    Vector3d trans;
    vector<Vector2d> pts;
  };
  vector<Logged> log;            // every update() since the map was empty
  size_t log_points = 0;
  const NDTFrame* scan_owner = nullptr;  // whose scan the device holds, and which version of it
  unsigned long scan_version = 0;
  int max_beams = 0;
  bool exact_scan = false;       // NDTPSO_SHIM_EXACT_SCAN=1: scans always travel as host-computed points
};

namespace {
constexpr size_t kMaxLoggedUpdates = 256, kMaxLoggedPoints = 1u << 20;

int env_int(const char* name, int dflt) {
  const char* e = std::getenv(name);
  return e ? std::atoi(e) : dflt;
}
}  // namespace

void NDTFrame::dropMirror() {
  if (!dev_) return;
  if (dev_->df) ndtpso_dframes_destroy(dev_->df);
  delete dev_;
  dev_ = nullptr;
}

bool NDTFrame::deviceResident() const { return dev_ && dev_->df && mirror_ok_; }

bool NDTFrame::downloadDeviceMap(double* mean, double* inv_cov, uint8_t* built_out) {
  if (!deviceResident()) return false;
  if (ndtpso_dframes_build(dev_->df) != NDTPSO_OK) return false;
  return ndtpso_dframes_download_map(dev_->df, 0, mean, inv_cov, built_out) == NDTPSO_OK;
}

// the device takes new_frame's scan, unless it already holds exactly that
bool NDTFrame::mirrorSyncScan(const NDTFrame* new_frame, size_t* h2d) {
  DeviceMirror& d = *dev_;
  if (d.scan_owner == new_frame && d.scan_version == new_frame->version_) return true;
  const LaserInput& li = new_frame->laser_;
  int rc;
  if (li.valid && !d.exact_scan && (int)li.ranges.size() <= d.max_beams && new_frame->s_config.laserIgnoreEpsilon == s_config.laserIgnoreEpsilon &&
      new_frame->width == width && new_frame->height == height) {
    // loadLaser redone on the device from the ranges (4 bytes per beam); the points are binned like the caller's scan frame
    const double trans[3] = {new_frame->s_trans.x(), new_frame->s_trans.y(), new_frame->s_trans.z()};
    const double scan_cs = new_frame->numOfCells > 1 ? new_frame->cell_side : 0.;
    rc = ndtpso_dframes_load_laser_binned(d.df, li.ranges.data(), (int32_t)li.ranges.size(), li.min_angle, li.angle_increment, li.max_range, trans, scan_cs);
    *h2d += 4 * li.ranges.size() + 40;
  } else {
    const vector<Vector2d>& pts = new_frame->scanPoints();
    if ((int)pts.size() > d.max_beams) return false;
    const int32_t n = (int32_t)pts.size();
    static const double none[2] = {0., 0.};
    rc = ndtpso_dframes_set_scan_points(d.df, n ? reinterpret_cast<const double*>(pts.data()) : none, &n, n);
    *h2d += 16 * pts.size() + 4;
  }
  if (rc != NDTPSO_OK) return false;
  d.scan_owner = new_frame;
  d.scan_version = new_frame->version_;
  return true;
}

// update(): mirrored on the device, or logged for the mirror the first align() will create.  false = this map cannot be mirrored.
bool NDTFrame::mirrorUpdate(const Vector3d& trans, const NDTFrame* new_frame, bool map_was_empty) {
  if (!dev_) {
    if (!map_was_empty || s_iter != 0) return false;  // points or matches the mirror has not seen
    if (env_int("NDTPSO_SHIM_DEVICE_MAP", 1) == 0) return false;
    dev_ = new DeviceMirror();
  }
  DeviceMirror& d = *dev_;
  if (!d.df) {
    const vector<Vector2d>& pts = new_frame->scanPoints();
    if (d.log.size() >= kMaxLoggedUpdates || d.log_points + pts.size() > kMaxLoggedPoints) return false;
    d.log.push_back(DeviceMirror::Logged{trans, pts});
    d.log_points += pts.size();
    return true;
  }
  size_t h2d = 0;
  if (!mirrorSyncScan(new_frame, &h2d)) return false;
  const double pose[3] = {trans.x(), trans.y(), trans.z()};
  if (ndtpso_dframes_update(d.df, pose) != NDTPSO_OK) return false;
  last_update_h2d_ = h2d + 32;
  return true;
}

// align() through the mirror; creates it (and replays the logged updates) on the first call.  false = use the upload path.
bool NDTFrame::mirrorAlign(const Vector3d& guess, const NDTFrame* new_frame, const PSOConfig& conf, Vector3d* pose_out) {
  DeviceMirror& d = *dev_;
  if (!d.df) {
    if (s_iter != 1) return false;  // an earlier align went another way: the device's copy of the bookkeeping would be behind
    ndtpso_ctx* ctx = ndtpso_b200::shim_context_or_null();
    if (!ctx) return false;  // no device: the upload path reports it
    ndtpso_dframes_config cfg;
    ndtpso_dframes_config_default(&cfg);
    cfg.n_frames = 1;
    cfg.width_m = width;
    cfg.height_m = height;
    cfg.cell_side = cell_side;
    size_t longest = new_frame->scanPoints().size();
    if (new_frame->laser_.valid) longest = std::max(longest, new_frame->laser_.ranges.size());
    for (const auto& l : d.log) longest = std::max(longest, l.pts.size());
    cfg.max_beams = std::max(env_int("NDTPSO_SHIM_MAX_BEAMS", 2048), (int)longest);
    cfg.max_cells = env_int("NDTPSO_SHIM_MAX_CELLS", 4096);
    cfg.window_points = env_int("NDTPSO_SHIM_WINDOW_POINTS", 1024);
    cfg.laser_ignore_epsilon = s_config.laserIgnoreEpsilon;
    if (ndtpso_dframes_create(ctx, &cfg, &d.df) != NDTPSO_OK) {
      d.df = nullptr;
      return false;
    }
    d.max_beams = cfg.max_beams;
    d.exact_scan = env_int("NDTPSO_SHIM_EXACT_SCAN", 0) != 0;
    for (const auto& l : d.log) {  // replay: host-computed points and the host's cos/sin, so the device map equals the host's bit for bit
      const int32_t n = (int32_t)l.pts.size();
      static const double none[2] = {0., 0.};
      const double pose[3] = {l.trans.x(), l.trans.y(), l.trans.z()};
      if (ndtpso_dframes_set_scan_points(d.df, n ? reinterpret_cast<const double*>(l.pts.data()) : none, &n, n) != NDTPSO_OK ||
          ndtpso_dframes_update(d.df, pose) != NDTPSO_OK)
        return false;
    }
    d.log.clear();
    d.log.shrink_to_fit();
    d.scan_owner = nullptr;
  }
  size_t h2d = 0;
  if (!mirrorSyncScan(new_frame, &h2d)) return false;
  ndtpso_pso_config cf;
  cf.iterations = conf.iterations;
  cf.population = conf.populationSize;
  cf.num_threads = conf.num_threads;
  cf.variant = NDTPSO_VARIANT_PSO;
  cf.w = conf.coeff.w;
  cf.c1 = conf.coeff.c1;
  cf.c2 = conf.coeff.c2;
  cf.w_dumping = conf.coeff.w_dumping;
  // the reference's random numbers: the next 3 + 3P + 6PI outputs of the process-global std::rand()
  const int64_t n = ndtpso_rand_draws(&cf);
  std::vector<int32_t> stream(static_cast<size_t>(n));
  ndtpso_b200::shim_draw_rand(stream.data(), stream.size());
  const double g[3] = {guess.x(), guess.y(), guess.z()};
  double pose[3], cost = 0.;
  if (ndtpso_dframes_align_streams(d.df, g, &cf, stream.data(), pose, &cost) != NDTPSO_OK) {
    // the numbers are drawn: hand the call to the upload path would draw them again and leave the stream out of step, so fail loudly
    ndtpso_b200::shim_fail(std::string("ndtpso_b200: device-resident align failed: ") + ndtpso_last_error(ndtpso_b200::shim_context_or_null()));
    *pose_out = guess;  // a failure handler took it: no correction for this scan
    return true;
  }
  int32_t flags = 0;
  if (ndtpso_dframes_status(d.df, &flags) == NDTPSO_OK && (flags & (NDTPSO_DF_CELL_POOL_FULL | NDTPSO_DF_WINDOW_TRUNCATED))) {
    ndtpso_b200::shim_fail("ndtpso_b200: the device-resident map outgrew its pools (raise NDTPSO_SHIM_MAX_CELLS / NDTPSO_SHIM_WINDOW_POINTS, or set NDTPSO_SHIM_DEVICE_MAP=0)");
    *pose_out = guess;
    return true;
  }
  ndtpso_b200::shim_set_last_cost(cost);
  last_align_h2d_ = h2d + 4 * static_cast<size_t>(n) + 24;
  *pose_out = Vector3d(pose[0], pose[1], pose[2]);
  return true;
}
