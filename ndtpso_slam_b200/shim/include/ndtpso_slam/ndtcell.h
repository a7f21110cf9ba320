// NDT cell statistics of the drop-in library.
//
// The reference keeps one 7.7 KB NDTCell object per grid cell (100-slot sliding window of partial
// sums, counts, covariances and point lists; include/ndtpso_slam/ndtcell.h:63-81).  Here a cell's
// window state is a CellWindow allocated only for cells that ever received a point, and the three
// quantities the scan matcher reads (mean, inverse covariance, built flag) live in dense arrays
// owned by the frame, so the matcher's view of the map is zero-copy.
// The arithmetic of addPoint/build/inverse follows lib/ndtpso_slam/ndtcell.cpp:21-68,93-111
// operation for operation (tests/test_shim_frames.py checks the tables bit for bit).
#ifndef NDTPSO_B200_SHIM_NDTCELL_H
#define NDTPSO_B200_SHIM_NDTCELL_H

#include <eigen3/Eigen/Core>
#include <vector>

#include "ndtpso_slam/config.h"

namespace ndtpso_b200 {

struct Sym2 {  // 2x2 matrix, row-major
  double m00, m01, m10, m11;
};

struct CellWindow {
  double part_sum[NDT_WINDOW_SIZE][2];
  Sym2 part_cov[NDT_WINDOW_SIZE];
  int part_count[NDT_WINDOW_SIZE];
  double cur_sum[2];
  double glob_sum[2];
  Sym2 glob_cov;
  int cur_count, glob_count;
  unsigned slot;
  std::vector<Eigen::Vector2d> points[NDT_WINDOW_SIZE];
  explicit CellWindow(bool zero_windows);
  void add(const Eigen::Vector2d& p);
  // closes the statistics of the current slot; returns true and fills mean/inv when the cell has > 2 points
  bool build(double* mean2, double* inv4);
  void reset();
};

}  // namespace ndtpso_b200
#endif
