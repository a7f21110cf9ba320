// Window statistics of one NDT cell (see ndtcell.h).  Arithmetic follows the reference's
// NDTCell::addPoint / build / s_calc_covar_inverse (lib/ndtpso_slam/ndtcell.cpp:21-68,93-111) so
// that the (mean, inverse covariance) tables come out bit-identical; compile with
// -ffp-contract=off (the reference is built for baseline x86-64: no fused multiply-add).
#include "ndtpso_slam/ndtcell.h"

#include <cmath>
#include <cstring>

namespace ndtpso_b200 {

CellWindow::CellWindow(bool zero_windows) : cur_count(0), glob_count(0), slot(0) {
  // The reference zeroes the window only when asked (its one-cell scan frames skip it and are
  // never built); zeroing always is harmless and keeps build() defined for every frame.
  (void)zero_windows;
  std::memset(part_sum, 0, sizeof part_sum);
  std::memset(part_cov, 0, sizeof part_cov);
  std::memset(part_count, 0, sizeof part_count);
  cur_sum[0] = cur_sum[1] = 0.;
  glob_sum[0] = glob_sum[1] = 0.;
  glob_cov = Sym2{0., 0., 0., 0.};
}

void CellWindow::add(const Eigen::Vector2d& p) {
  if (cur_count == 0) points[slot].clear();  // first point after the slot was closed (ndtcell.cpp:22-27)
  ++cur_count;
  cur_sum[0] += p.x();
  cur_sum[1] += p.y();
  points[slot].push_back(p);
}

// eigenvalues of a real 2x2 matrix in closed form; the reference asks Eigen's EigenSolver, which is
// only used for the ratio test below
static void eig2(const Sym2& a, double* l0, double* l1) {
  const double half_tr = (a.m00 + a.m11) / 2.;
  const double half_df = (a.m00 - a.m11) / 2.;
  const double disc = half_df * half_df + a.m01 * a.m10;
  const double root = disc > 0. ? std::sqrt(disc) : 0.;
  *l0 = half_tr + root;
  *l1 = half_tr - root;
}

bool CellWindow::build(double* mean2, double* inv4) {
  // sliding window in O(1): global += current - what the slot held before (ndtcell.h:13-15 WINDOW_ADD)
  for (int k = 0; k < 2; ++k) {
    glob_sum[k] = glob_sum[k] + cur_sum[k] - part_sum[slot][k];
    part_sum[slot][k] = cur_sum[k];
  }
  glob_count = glob_count + cur_count - part_count[slot];
  part_count[slot] = cur_count;

  bool is_built = false;
  if (glob_count > 2) {  // ndtcell.cpp:43
    const double n = glob_count;
    const double mx = glob_sum[0] / n, my = glob_sum[1] / n;
    Sym2 cov{0., 0., 0., 0.};
    for (const auto& pt : points[slot]) {  // scatter of the CURRENT slot about the GLOBAL mean (ndtcell.cpp:49-52)
      const double dx = pt.x() - mx, dy = pt.y() - my;
      cov.m00 += dx * dx;
      cov.m01 += dx * dy;
      cov.m10 += dy * dx;
      cov.m11 += dy * dy;
    }
    Sym2& g = glob_cov;
    Sym2& old = part_cov[slot];
    g.m00 = g.m00 + cov.m00 - old.m00;
    g.m01 = g.m01 + cov.m01 - old.m01;
    g.m10 = g.m10 + cov.m10 - old.m10;
    g.m11 = g.m11 + cov.m11 - old.m11;
    old = cov;

    // inverse covariance with the eigenvalue-ratio floor (ndtcell.cpp:93-111)
    const Sym2 c{g.m00 / n, g.m01 / n, g.m10 / n, g.m11 / n};
    double e0, e1;
    eig2(c, &e0, &e1);
    const double large = e0 > e1 ? e0 : e1;
    const double small = e0 < e1 ? e0 : e1;
    double det;
    if (small < .001 * large)
      det = .001 * large * large;  // inflates: the adjugate is kept, only the determinant is replaced
    else
      det = c.m00 * c.m11 - c.m01 * c.m10;
    mean2[0] = mx;
    mean2[1] = my;
    inv4[0] = c.m11 / det;
    inv4[1] = -c.m01 / det;
    inv4[2] = -c.m10 / det;
    inv4[3] = c.m00 / det;
    is_built = true;
  }
  if (cur_count > NDT_MAX_POINTS_PER_CELL) {  // the slot is full: open the next one (ndtcell.cpp:61-65)
    slot = (slot + 1) % NDT_WINDOW_SIZE;
    cur_count = 0;
    cur_sum[0] = cur_sum[1] = 0.;
  }
  return is_built;
}

void CellWindow::reset() {
  cur_sum[0] = cur_sum[1] = 0.;
  glob_sum[0] = glob_sum[1] = 0.;
  cur_count = glob_count = 0;
  glob_cov = Sym2{0., 0., 0., 0.};
  slot = 0;
  for (auto& v : points) v.clear();
}

}  // namespace ndtpso_b200
