// Drop-in NDTFrame (see ndtframe.h): host-side map building with the reference's semantics
// (lib/ndtpso_slam/ndtframe.cpp:19-66,144-235,240-266), scan matching on the GPU.
#include "ndtpso_slam/ndtframe.h"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "ndtpso_b200.h"
#include "ndtpso_slam/core.h"

using ndtpso_b200::CellWindow;

NDTFrame::NDTFrame(Vector3d trans, unsigned short w, unsigned short h, double side, bool calculate_cells_params, NDTPSOConfig config
#if BUILD_OCCUPANCY_GRID
                   ,
                   double /*occupancy_grid_cell_size*/
#endif
                   )
    : width(w), height(h), built(false), cell_side(side), s_trans(std::move(trans)), s_config(std::move(config)),
      zero_windows_(calculate_cells_params) {
  widthNumOfCells = uint16_t(std::ceil(width / cell_side));    // ndtframe.cpp:27-29
  heightNumOfCells = uint16_t(std::ceil(height / cell_side));
  numOfCells = widthNumOfCells * heightNumOfCells;
  s_x_min = -width / 2.;  // ndtframe.cpp:57-65
  s_x_max = width / 2.;
  s_y_min = -height / 2.;
  s_y_max = height / 2.;
  mean_.assign(2 * (size_t)numOfCells, 0.);
  icov_.assign(4 * (size_t)numOfCells, 0.);
  built_.assign(numOfCells, 0);
  slot_of_.assign(numOfCells, -1);
}

NDTFrame::~NDTFrame() { dropMirror(); }

int NDTFrame::getCellIndex(Vector2d point, int grid_width, double side) {
  if (!((point.x() > s_x_min) && (point.x() < s_x_max) && (point.y() > s_y_min) && (point.y() < s_y_max))) return -1;  // strict
  return static_cast<int>(std::floor((point.x() + (width / 2.)) / side) + grid_width * (std::floor((point.y() + (height / 2.)) / side)));
}

void NDTFrame::addPoint(Vector2d& point) {
  // a point added directly is not mirrored on the device: a mirrored map falls back to uploading its table per align
  if (dev_ || zero_windows_) mirror_ok_ = false;
  if (dev_) dropMirror();
  laser_.valid = false;
  addPointInternal(point);
}

namespace {
// a process-wide counter: a scan frame that is deleted and re-allocated at the same address must not pass for the scan the
// device already holds (the node does exactly that, ndtpso_slam_node.cpp:228-229)
unsigned long next_version() {
  static std::atomic<unsigned long> counter{0};
  return ++counter;
}
}  // namespace

void NDTFrame::addPointInternal(Vector2d& point) {
  version_ = next_version();
  const int idx = getCellIndex(point, widthNumOfCells, cell_side);
  if (idx < 0 || static_cast<unsigned>(idx) >= numOfCells) return;
  int s = slot_of_[idx];
  if (s < 0) {
    s = slot_of_[idx] = static_cast<int>(windows_.size());
    windows_.emplace_back(new CellWindow(zero_windows_));
  }
  windows_[s]->add(point);
  built_[idx] = 0;  // NDTCell::addPoint clears the cell's flag (ndtcell.cpp:33)
  built = false;
  scan_cache_valid_ = false;
}

void NDTFrame::loadLaser(const vector<float>& laser_data, const float& min_angle, const float& angle_increment, const float& max_range) {
  built = false;
  if (dev_ || zero_windows_) mirror_ok_ = false;  // a scan loaded into the map itself is not mirrored
  if (dev_) dropMirror();
  const bool first_fill = windows_.empty();
  const bool shift = !s_trans.isZero(1e-6);
  const unsigned n = static_cast<unsigned>(laser_data.size());
  for (unsigned i = 0; i < n; ++i) {
    const float r = laser_data[i];
    if (!((r > 0.) && (r < max_range) && (r > s_config.laserIgnoreEpsilon))) continue;  // ndtframe.cpp:165
    const float theta = index_to_angle(i, angle_increment, min_angle);
    Vector2d p = laser_to_point(r, theta);
    if (shift) p = transform_point(p, s_trans);
    addPointInternal(p);
  }
  // remembered so that a mirrored map can take this scan as 4 bytes per beam
  laser_.valid = first_fill;
  if (first_fill) {
    laser_.ranges = laser_data;
    laser_.min_angle = min_angle;
    laser_.angle_increment = angle_increment;
    laser_.max_range = max_range;
  }
}

void NDTFrame::update(Vector3d trans, NDTFrame* const new_frame) {
  built = false;
  const bool map_was_empty = windows_.empty();
  laser_.valid = false;
  for (const Vector2d& q : new_frame->scanPoints()) {  // slot 0 of every created cell, cell order (ndtframe.cpp:190-196)
    Vector2d p = transform_point(q, trans);
    addPointInternal(p);
  }
  last_update_h2d_ = 0;
  if (mirror_ok_ && zero_windows_ && !mirrorUpdate(trans, new_frame, map_was_empty)) {
    mirror_ok_ = false;
    dropMirror();
  }
}

void NDTFrame::build() {
  sp_index_.clear();
  sp_mean_.clear();
  sp_icov_.clear();
  for (unsigned i = 0; i < numOfCells; ++i) {  // cell order, like ndtframe.cpp:73-77
    const int s = slot_of_[i];
    if (s < 0) continue;
    if (windows_[s]->build(&mean_[2 * (size_t)i], &icov_[4 * (size_t)i])) built_[i] = 1;
    if (built_[i]) {
      sp_index_.push_back(static_cast<int>(i));
      sp_mean_.insert(sp_mean_.end(), &mean_[2 * (size_t)i], &mean_[2 * (size_t)i] + 2);
      sp_icov_.insert(sp_icov_.end(), &icov_[4 * (size_t)i], &icov_[4 * (size_t)i] + 4);
    }
  }
  built = true;
}

const vector<Vector2d>& NDTFrame::scanPoints() const {
  if (!scan_cache_valid_) {
    scan_cache_.clear();
    for (unsigned i = 0; i < numOfCells; ++i) {
      const int s = slot_of_[i];
      if (s < 0) continue;
      const auto& v = windows_[s]->points[0];
      scan_cache_.insert(scan_cache_.end(), v.begin(), v.end());
    }
    scan_cache_valid_ = true;
  }
  return scan_cache_;
}

size_t NDTFrame::pointCount() const {
  size_t n = 0;
  for (const auto& w : windows_)
    for (const auto& v : w->points) n += v.size();
  return n;
}

static void fill_geometry(const NDTFrame& f, ndtpso_map_view* out) {
  out->w_cells = f.widthNumOfCells;
  out->h_cells = f.heightNumOfCells;
  out->width_m = f.width;
  out->height_m = f.height;
  out->cell_side = f.cell_side;
  out->x_min = f.xMin();
  out->x_max = f.xMax();
  out->y_min = f.yMin();
  out->y_max = f.yMax();
  out->reserved = 0;
}

void NDTFrame::mapView(ndtpso_map_view* out) const {
  if (!built) const_cast<NDTFrame*>(this)->build();  // with the map mirrored on the device the host table is built on demand only
  fill_geometry(*this, out);
  out->mean = mean_.data();
  out->inv_cov = icov_.data();
  out->built = built_.data();
  out->n_sparse = -1;
  out->cell_index = nullptr;
}

void NDTFrame::sparseMapView(ndtpso_map_view* out) const {
  if (!built) const_cast<NDTFrame*>(this)->build();
  fill_geometry(*this, out);
  out->mean = sp_mean_.data();
  out->inv_cov = sp_icov_.data();
  out->built = nullptr;
  out->n_sparse = static_cast<int32_t>(sp_index_.size());
  out->cell_index = sp_index_.data();
}

Vector3d NDTFrame::align(Vector3d initial_guess, const NDTFrame* const new_frame) {
  return align(std::move(initial_guess), new_frame, PSOConfig());  // the reference always runs the defaults (ndtframe.cpp:257)
}

Vector3d NDTFrame::align(Vector3d initial_guess, const NDTFrame* const new_frame, const PSOConfig& conf) {
  // spread of the initial swarm: fixed for the first two calls, then twice the last pose step (ndtframe.cpp:253)
  Vector3d deviation = s_iter < 2 ? Vector3d(.1, .1, 3.1415E-3) : (s_pose_diff * 2.).array().abs();
  ++s_iter;
  Vector3d pose;
  last_align_h2d_ = 0;
  // cost_function builds the map lazily before the first evaluation (core.cpp:27-28).  NDTCell::build is not idempotent (every
  // call re-adds the current slot's statistics), so the host's copy of the table is built here too, exactly as often as the
  // reference's: it stays what the reference would hold, and what the device mirror holds.
  if (!built) build();
  // the mirror keeps its own copy of this bookkeeping and applies the same rule (align_prepare_kernel)
  if (!(dev_ && mirror_ok_ && mirrorAlign(initial_guess, new_frame, conf, &pose))) {
    if (dev_) {  // no mirror for this map (or it could not take this call): from here on the table is uploaded per align
      mirror_ok_ = false;
      dropMirror();
    }
    pose = pso_optimization(std::move(initial_guess), this, new_frame, deviation, conf);
  }
  s_pose_diff = pose - s_prev_pose;
  s_prev_pose = pose;
  return pose;
}

void NDTFrame::addPose(double timestamp, const Vector3d& pose, const Vector3d& odom) {
  s_timestamps.push_back(timestamp);
  s_poses.push_back(pose);
  s_odoms.push_back(odom);
}

void NDTFrame::resetCells() {
  if (dev_ || zero_windows_) mirror_ok_ = false;
  if (dev_) dropMirror();
  version_ = next_version();
  for (auto& w : windows_) w->reset();  // NDTCell::reset (ndtcell.cpp:80-91) leaves `built`, mean and Sigma^-1 as they are
  scan_cache_valid_ = false;            // the cached scan holds the points that were just dropped
}

void NDTFrame::dumpMap(const char* filename, bool save_poses, bool save_points, bool /*save_image*/, short /*density*/
#if BUILD_OCCUPANCY_GRID
                       ,
                       bool /*save_occupancy_grid*/
#endif
) {
  // CSV + gnuplot export in the reference's format (ndtframe.cpp:268-391); images need OpenCV and are not produced
  char name[512];
  FILE *fp_pose = nullptr, *fp_map = nullptr;
  if (save_poses) {
    snprintf(name, sizeof name, "%s.pose.csv", filename);
    fp_pose = fopen(name, "w");
    if (fp_pose) fprintf(fp_pose, "timestamp,xP,yP,thP,xO,yO,thO\n");
  }
  if (save_points) {
    snprintf(name, sizeof name, "%s.map.csv", filename);
    fp_map = fopen(name, "w");
    if (fp_map) fprintf(fp_map, "x,y\n");
  }
  if ((save_poses && !fp_pose) || (save_points && !fp_map)) {
    printf("%s: Cannot open files, cannot save!\n ", __func__);
    if (fp_pose) fclose(fp_pose);
    if (fp_map) fclose(fp_map);
    return;
  }
  if (fp_map) {
    for (unsigned i = 0; i < numOfCells; ++i) {
      if (slot_of_[i] < 0) continue;
      for (const auto& v : windows_[slot_of_[i]]->points)
        for (const auto& p : v) fprintf(fp_map, "%.5f,%.5f\n", p.x(), p.y());
    }
  }
  if (fp_pose && save_points) {  // the reference writes poses only under save_points (ndtframe.cpp:346)
    for (size_t i = 0; i < s_poses.size(); ++i)
      fprintf(fp_pose, "%.6f,%.5f,%.5f,%.5f\n", s_timestamps[i], s_poses[i].x(), s_poses[i].y(), s_poses[i].z());
  }
  if (fp_pose) fclose(fp_pose);
  if (fp_map) fclose(fp_map);
  if (save_poses || save_points) {
    snprintf(name, sizeof name, "%s.gnuplot", filename);
    FILE* g = fopen(name, "w");
    if (!g) return;
    fprintf(g, "set datafile separator ','\nset key autotitle columnhead\nset size ratio -1\nplot ");
    if (save_points) fprintf(g, "'%s.map.csv' title 'Map' with points pointsize 0.2 pointtype 5 linecolor rgb '#555555'", filename);
    if (save_poses)
      fprintf(g, ", \\\n'%s.pose.csv' using 2:3 title 'Pose (LiDAR)' with linespoints linewidth 0.7 pointtype 6 pointsize 0.7 linecolor rgb '#ff0000'",
              filename);
    fprintf(g, "\npause 1000\n");
    fclose(g);
  }
}
