// Drop-in NDTFrame: the class the reference's ROS node drives (src/ndtpso_slam_node.cpp:64-78,
// 186-230) with the same constructor and method signatures as include/ndtpso_slam/ndtframe.h:31-72,
// so the node recompiles against this header unchanged.  align() runs on the GPU through the C ABI
// (include/ndtpso_b200.h).
//
// The map lives in HBM.  A frame that is filled the way the node fills its reference frame — through update() only, from
// the first point on — is mirrored by a device-resident frame (include/ndtpso_dframes.h): update() runs NDTFrame::update +
// NDTCell::addPoint as a kernel, align() runs NDTFrame::build, the table compaction and the swarm there, and what crosses PCIe
// per scan is the scan itself (4 bytes per beam when the scan frame was filled by one loadLaser call; NDTPSO_SHIM_EXACT_SCAN=1
// sends the host-computed points, 16 bytes each, which keeps the device map bit-identical to the host's), the random numbers
// and the pose.  The host keeps the points (dumpMap, pointCount) and its own copy of the table, built as often as the reference
// builds its (NDTCell::build is not idempotent), so mapView() shows what the reference would hold.  A frame that receives points any other way (addPoint / loadLaser on the map itself, resetCells)
// drops the mirror and uploads its table with every align, as before; the swarm runs on the GPU either way.
//
// Differences from the reference, all outside the node's use of the class:
//  * `cells` is not a public vector<NDTCell>: storage is a dense (mean, Sigma^-1, built) table plus
//    lazily allocated window state (ndtcell.h).  Read access: mapView(), cellBuilt(), cellMean(), ...
//  * the occupancy grid and the PNG export of dumpMap are not built (visualisation, out of scope);
//    dumpMap writes the same CSV and gnuplot files.
//  * NDTFrame::transform is omitted (dead code with a use-after-free in the reference, ndtframe.cpp:119-140).
//  * a point whose flat cell index falls past the end of the table is dropped instead of written
//    out of bounds (the reference's behaviour there is undefined).
#ifndef NDTPSO_B200_SHIM_NDTFRAME_H
#define NDTPSO_B200_SHIM_NDTFRAME_H

#include <cstdint>
#include <eigen3/Eigen/Core>
#include <memory>
#include <utility>
#include <vector>

#include "ndtpso_slam/ndtcell.h"

using namespace Eigen;
using std::vector;

struct ndtpso_map_view;

class NDTFrame {
 public:
  uint16_t width, height, widthNumOfCells, heightNumOfCells;
  bool built;
  unsigned int numOfCells;
  double cell_side;

  NDTFrame(Vector3d trans, unsigned short width = 20, unsigned short height = 20, double cell_side = 1.0,
           bool calculate_cells_params = true, NDTPSOConfig config = NDTPSOConfig()
#if BUILD_OCCUPANCY_GRID
                                                                        ,
           double occupancy_grid_cell_size = .0
#endif
  );
  ~NDTFrame();
  NDTFrame(const NDTFrame&) = delete;
  NDTFrame& operator=(const NDTFrame&) = delete;

  void loadLaser(const vector<float>& laser_data, const float& min_angle, const float& angle_increment, const float& max_range);
  void update(Vector3d trans, NDTFrame* new_frame);
  void addPoint(Vector2d& point);
  inline void setTrans(Vector3d trans) { s_trans = std::move(trans); }
  void build();
  int getCellIndex(Vector2d point, int grid_width, double cell_side);
  Vector3d align(Vector3d initial_guess, const NDTFrame* const new_frame);
  // the overload the author evidently intended (the node's PSO parameters never reach the reference's align)
  Vector3d align(Vector3d initial_guess, const NDTFrame* const new_frame, const PSOConfig& conf);
  void dumpMap(const char* filename, bool save_poses = true, bool save_points = true, bool save_image = true, short density = 50
#if BUILD_OCCUPANCY_GRID
               ,
               bool save_occupancy_grid = true
#endif
  );
  void addPose(double timestamp, const Vector3d& pose, const Vector3d& odom = Vector3d::Zero());
  void resetCells();

  // ---- read access to what the scan matcher consumes
  void mapView(ndtpso_map_view* out) const;       // dense view, valid until the next addPoint/build
  void sparseMapView(ndtpso_map_view* out) const;  // built cells only (refreshed by build())
  const vector<Vector2d>& scanPoints() const;      // window slot 0 of every cell in cell order (core.cpp:33-36)
  bool cellBuilt(unsigned i) const { return built_[i] != 0; }
  bool cellCreated(unsigned i) const { return slot_of_[i] >= 0; }
  Vector2d cellMean(unsigned i) const { return Vector2d(mean_[2 * i], mean_[2 * i + 1]); }
  const double* cellInvCov(unsigned i) const { return &icov_[4 * i]; }
  double xMin() const { return s_x_min; }
  double xMax() const { return s_x_max; }
  double yMin() const { return s_y_min; }
  double yMax() const { return s_y_max; }
  size_t pointCount() const;  // all windows of all cells
  int alignCalls() const { return s_iter; }
  // ---- the device mirror
  bool deviceResident() const;  // the map is mirrored in HBM and align() uses it
  // bytes moved by the most recent align() / update() of this frame through the mirror (0 without one)
  size_t lastAlignH2DBytes() const { return last_align_h2d_; }
  size_t lastUpdateH2DBytes() const { return last_update_h2d_; }
  // device copy of the table (synchronises): mean [numOfCells][2], inv_cov [numOfCells][4], built [numOfCells]; false without a mirror
  bool downloadDeviceMap(double* mean, double* inv_cov, uint8_t* built);

 private:
  Vector3d s_trans{Vector3d::Zero()}, s_prev_pose{Vector3d::Zero()}, s_pose_diff{Vector3d::Zero()};
  vector<Vector3d> s_poses, s_odoms;
  vector<double> s_timestamps;
  double s_x_min, s_x_max, s_y_min, s_y_max;
  NDTPSOConfig s_config;
  int s_iter{0};
  bool zero_windows_;

  // dense hot-path table
  vector<double> mean_;    // [numOfCells][2]
  vector<double> icov_;    // [numOfCells][4]
  vector<uint8_t> built_;  // [numOfCells]
  // window state of the cells that ever received a point
  vector<int> slot_of_;  // [numOfCells] -> index into windows_, or -1
  vector<std::unique_ptr<ndtpso_b200::CellWindow>> windows_;
  // sparse copy of the built cells, refreshed by build()
  vector<int> sp_index_;
  vector<double> sp_mean_, sp_icov_;
  mutable vector<Vector2d> scan_cache_;
  mutable bool scan_cache_valid_{false};

  // ---- device mirror (ndtframe_device.cpp)
  struct DeviceMirror;
  DeviceMirror* dev_{nullptr};
  bool mirror_ok_{true};     // false once a point entered the map by another way than update(), or the mirror failed
  unsigned long version_{0}; // bumped by every change of this frame's points: identifies a scan for the mirror
  // how this frame's points came to be, when by exactly one loadLaser into an empty frame: enough to redo it on the
  // device from 4 bytes per beam
  struct LaserInput {
    vector<float> ranges;
    float min_angle{0}, angle_increment{0}, max_range{0};
    bool valid{false};
  } laser_;
  size_t last_align_h2d_{0}, last_update_h2d_{0};
  void addPointInternal(Vector2d& point);
  bool mirrorUpdate(const Vector3d& trans, const NDTFrame* new_frame, bool map_was_empty);
  bool mirrorAlign(const Vector3d& guess, const NDTFrame* new_frame, const PSOConfig& conf, Vector3d* pose);
  bool mirrorSyncScan(const NDTFrame* new_frame, size_t* h2d);
  void dropMirror();
  friend struct NDTFrameDeviceAccess;
};

#endif
