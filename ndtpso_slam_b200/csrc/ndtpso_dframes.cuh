// Device-resident reference frames (include/ndtpso_dframes.h): the NDT map building and scan
// ingestion of libndtpso_slam as CUDA kernels, so that a tracked robot's map never leaves HBM.
//
//   load_laser_kernel    NDTFrame::loadLaser   lib/ndtpso_slam/ndtframe.cpp:144-185
//                        index_to_angle / laser_to_point   include/ndtpso_slam/core.h:40-47
//   frame_update_kernel  NDTFrame::update      lib/ndtpso_slam/ndtframe.cpp:187-198
//                        NDTFrame::addPoint    lib/ndtpso_slam/ndtframe.cpp:215-225
//                        NDTCell::addPoint     lib/ndtpso_slam/ndtcell.cpp:21-34
//   frame_build_kernel   NDTFrame::build       lib/ndtpso_slam/ndtframe.cpp:68-117
//                        NDTCell::build        lib/ndtpso_slam/ndtcell.cpp:36-68
//                        s_calc_covar_inverse  lib/ndtpso_slam/ndtcell.cpp:93-111
//   align_prepare_kernel / align_finish_kernel   NDTFrame::align   lib/ndtpso_slam/ndtframe.cpp:251-266
//
// One CTA per frame.  The reference adds points to a cell one at a time in scan order, and a cell's
// running sums depend on that order, so the kernels keep it: a scan's points are tagged with their
// cell, sorted by (cell, scan position) with a shared-memory bitonic sort, and one thread per touched
// cell appends its run sequentially.  Cells are independent of each other, so this is the only
// serial chain.  All arithmetic that feeds the tables is written with __dadd_rn / __dmul_rn /
// __ddiv_rn / __dsqrt_rn: no contraction, the reference's association order (it is built for
// baseline x86-64, which has no fused multiply-add).
//
// Storage per frame: dense (mean, Sigma^-1, built) arrays — the table cost_function reads, and what
// K0 compacts for the PSO kernel — plus a pool of window states for the cells that ever received a
// point (the reference keeps a 7.7 KB NDTCell for every cell of the grid):
//   cur_sum, cur_count, slot          s_current_partial_sum / s_current_count / s_current_window_id
//   glob_sum, glob_cov, glob_count    s_global_sum / s_global_covar_sum / s_global_count
//   part_sum/part_cov/part_count[100] s_partial_sums / s_partial_covars / s_partial_counts
//   ring[R], slot_start/slot_len[100] points_vector[100]: the points of the last slots in insertion
//                                     order; slot s owns ring positions [start, start+len)
#pragma once
#include "ndtpso_kernels.cuh"

namespace ndtpso {

constexpr int kWindow = 100;        // NDT_WINDOW_SIZE           config.h:8
constexpr int kMaxPerCell = 50;     // NDT_MAX_POINTS_PER_CELL   config.h:5
constexpr int kDfThreads = 512;
constexpr int kDfBuildThreads = 512;

enum { DF_CELL_POOL_FULL = 1, DF_WINDOW_TRUNCATED = 2, DF_INDEX_PAST_END = 4, DF_IRREGULAR_SIGMA = 8 };

struct DevFrames {
  int n, gw, gh, ncells, max_beams, max_cells, ring;
  int scan_gw, scan_ncells;
  double hw, hh, cs;                      // width/2., height/2., cell_side
  double x_min, x_max, y_min, y_max;      // ndtframe.cpp:57-65
  double scan_cs;                         // cell side of the scan frame
  float laser_eps;
  DevProblem* probs;    // [n]  pts -> scan_pts row, n_pts, guess, dev, seed, rnd
  double2* scan_pts;    // [n][max_beams]
  double2* tmp_pts;     // [n][max_beams]
  float* ranges;        // [n][max_beams]
  double* mean;         // [n][ncells][2]
  double* icov;         // [n][ncells][4]
  uint8_t* built;       // [n][ncells]
  int* slot_of;         // [n][ncells]  cell -> pool index, -1 = never received a point
  int* n_created;       // [n]
  int* flags;           // [n]
  // pool, [n][max_cells]
  int* cell_of;
  double2* cur_sum;
  double2* glob_sum;
  double* glob_cov;     // [..][4]
  int* cur_count;
  int* glob_count;
  int* slot;
  unsigned* total;      // points ever appended to the cell (ring position of the next one)
  // window, [n][max_cells][kWindow]
  double2* part_sum;
  double* part_cov;     // [..][4]
  int* part_count;
  unsigned* slot_start;
  unsigned* slot_len;
  double2* ringbuf;     // [n][max_cells][ring]
  // align bookkeeping, [n]
  int* s_iter;
  double* prev_pose;    // [n][3] s_prev_pose
  double* pose_diff;    // [n][3] s_pose_diff
  double* node_pose;    // [n][3] the caller's previous_pose_ / current_pose_ (ndtpso_slam_node.cpp:194)
  double* results;      // [n][4] x, y, theta, cost of the last align
};

// ---- block helpers ---------------------------------------------------------------------------
// exclusive prefix count of `flag` over the block in thread order; *total = block sum
__device__ __forceinline__ int block_excl_count(bool flag, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  const int in_warp = __popc(m & ((1u << lane) - 1u));
  __syncthreads();  // s_warp may still be read from a previous call
  if (lane == 0) s_warp[warp] = __popc(m);
  __syncthreads();
  int before = 0, all = 0;
  for (int w = 0; w < nw; ++w) {
    const int c = s_warp[w];
    if (w < warp) before += c;
    all += c;
  }
  *total = all;
  return before + in_warp;
}

// ascending bitonic sort of np2 (power of two) 64-bit keys in shared memory
__device__ __forceinline__ void block_bitonic_sort(unsigned long long* key, int np2) {
  for (int k = 2; k <= np2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < np2; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const unsigned long long a = key[i], b = key[p];
          const bool up = (i & k) == 0;
          if ((a > b) == up) {
            key[i] = b;
            key[p] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

// NDTFrame::getCellIndex (ndtframe.cpp:240-249): strict bounds, the sum of the two floors is formed
// in double and then truncated.  Returns -1 outside the frame.
__device__ __forceinline__ long long cell_index(double x, double y, double x_min, double x_max, double y_min, double y_max, double hw,
                                                double hh, double cs, int gw) {
  if (!((x > x_min) && (x < x_max) && (y > y_min) && (y < y_max))) return -1;
  const double fx = floor(__ddiv_rn(__dadd_rn(x, hw), cs));
  const double fy = floor(__ddiv_rn(__dadd_rn(y, hh), cs));
  return static_cast<long long>(__dadd_rn(fx, __dmul_rn(static_cast<double>(gw), fy)));
}

// transform_point (core.h:28-31) with the pose's cos/sin given: (px*c - py*s) + tx, (px*s + py*c) + ty
__device__ __forceinline__ double2 transform_nofma(double2 p, double tx, double ty, double c, double s) {
  return make_double2(__dadd_rn(__dadd_rn(__dmul_rn(p.x, c), -__dmul_rn(p.y, s)), tx),
                      __dadd_rn(__dadd_rn(__dmul_rn(p.x, s), __dmul_rn(p.y, c)), ty));
}

// ---- loadLaser ---------------------------------------------------------------------------------
// trans_cs: nullptr, or [n][5] = {x, y, cos(theta), sin(theta), apply} of each scan frame's s_trans
// (cos/sin taken on the host).  Dynamic shared memory: np2 * 8 bytes when the scan frame has more
// than one cell (the points are then put in cell-index-major order), else 0.
__global__ void __launch_bounds__(kDfThreads) load_laser_kernel(DevFrames F, int n_beams, float angle_min, float angle_inc, float range_max,
                                                                const double* __restrict__ trans_cs, int np2) {
  extern __shared__ __align__(16) unsigned char df_smem[];
  __shared__ int s_warp[kDfThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* r_in = F.ranges + (size_t)b * F.max_beams;
  const bool sorting = F.scan_ncells > 1;
  double2* dst = (sorting ? F.tmp_pts : F.scan_pts) + (size_t)b * F.max_beams;
  unsigned long long* key = reinterpret_cast<unsigned long long*>(df_smem);
  double tx = 0., ty = 0., tc = 1., ts = 0.;
  bool shift = false;
  if (trans_cs) {
    const double* t = trans_cs + 5 * (size_t)b;
    tx = t[0];
    ty = t[1];
    tc = t[2];
    ts = t[3];
    shift = t[4] != 0.;
  }
  int base = 0, my_flags = 0;
  for (int i0 = 0; i0 < n_beams; i0 += kDfThreads) {
    const int i = i0 + tid;
    bool valid = false;
    double2 p = make_double2(0., 0.);
    long long idx = -1;
    if (i < n_beams) {
      const float r = r_in[i];
      if ((static_cast<double>(r) > 0.) && (r < range_max) && (r > F.laser_eps)) {  // ndtframe.cpp:165
        const float theta = __fadd_rn(__fmul_rn(__uint2float_rn(static_cast<unsigned>(i)), angle_inc), angle_min);  // core.h:40-42
        double s, c;
        sincos(static_cast<double>(theta), &s, &c);
        p = make_double2(__dmul_rn(static_cast<double>(r), c), __dmul_rn(static_cast<double>(r), s));  // core.h:45-47
        if (shift) p = transform_nofma(p, tx, ty, tc, ts);
        idx = cell_index(p.x, p.y, F.x_min, F.x_max, F.y_min, F.y_max, F.hw, F.hh, F.scan_cs, F.scan_gw);
        if (idx >= F.scan_ncells) {  // past the end of the scan frame's cells: undefined in the reference, dropped here
          my_flags |= DF_INDEX_PAST_END;
          idx = -1;
        }
        valid = idx >= 0;
      }
    }
    int total;
    const int pos = base + block_excl_count(valid, s_warp, &total);
    if (valid) {
      dst[pos] = p;
      if (sorting) key[pos] = (static_cast<unsigned long long>(idx) << 32) | static_cast<unsigned>(pos);
    }
    base += total;
  }
  if (my_flags) atomicOr(&F.flags[b], my_flags);
  if (tid == 0) F.probs[b].n_pts = base;
  if (sorting) {
    for (int i = base + tid; i < np2; i += kDfThreads) key[i] = ~0ull;
    __syncthreads();
    block_bitonic_sort(key, np2);
    double2* out = F.scan_pts + (size_t)b * F.max_beams;
    for (int i = tid; i < base; i += kDfThreads) out[i] = dst[static_cast<unsigned>(key[i])];
  }
}

// ---- update --------------------------------------------------------------------------------------
// pose_cs: [n][4] = {x, y, cos(theta), sin(theta)} from the host, or nullptr = node_pose with the
// device's sincos.  Dynamic shared memory: np2 * 8 bytes.
__global__ void __launch_bounds__(kDfThreads) frame_update_kernel(DevFrames F, const double* __restrict__ pose_cs, int np2) {
  extern __shared__ __align__(16) unsigned char df_smem[];
  __shared__ int s_warp[kDfThreads / 32];
  __shared__ int s_created;
  unsigned long long* key = reinterpret_cast<unsigned long long*>(df_smem);
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n_pts = F.probs[b].n_pts;
  const double2* src = F.scan_pts + (size_t)b * F.max_beams;
  double2* q = F.tmp_pts + (size_t)b * F.max_beams;
  double tx, ty, c, s;
  if (pose_cs) {
    tx = pose_cs[4 * (size_t)b];
    ty = pose_cs[4 * (size_t)b + 1];
    c = pose_cs[4 * (size_t)b + 2];
    s = pose_cs[4 * (size_t)b + 3];
  } else {
    tx = F.node_pose[3 * (size_t)b];
    ty = F.node_pose[3 * (size_t)b + 1];
    sincos(F.node_pose[3 * (size_t)b + 2], &s, &c);
  }
  int my_flags = 0;
  for (int i = tid; i < np2; i += kDfThreads) {
    unsigned long long k = ~0ull;
    if (i < n_pts) {
      const double2 p = transform_nofma(src[i], tx, ty, c, s);  // ndtframe.cpp:193
      q[i] = p;
      long long idx = cell_index(p.x, p.y, F.x_min, F.x_max, F.y_min, F.y_max, F.hw, F.hh, F.cs, F.gw);  // ndtframe.cpp:217
      if (idx >= F.ncells) {
        my_flags |= DF_INDEX_PAST_END;
        idx = -1;
      }
      if (idx >= 0) k = (static_cast<unsigned long long>(idx) << 32) | static_cast<unsigned>(i);
    }
    key[i] = k;
  }
  if (tid == 0) s_created = F.n_created[b];
  __syncthreads();
  block_bitonic_sort(key, np2);

  int* slot_of = F.slot_of + (size_t)b * F.ncells;
  const size_t pool0 = (size_t)b * F.max_cells;
  // pass 1: cells that receive their first point get the next pool entries, in ascending cell order
  for (int p0 = 0; p0 < n_pts; p0 += kDfThreads) {
    const int p = p0 + tid;
    bool fresh = false;
    int cell = -1;
    if (p < n_pts && key[p] != ~0ull) {
      cell = static_cast<int>(key[p] >> 32);
      const bool head = (p == 0) || (static_cast<int>(key[p - 1] >> 32) != cell);
      fresh = head && slot_of[cell] < 0;
    }
    int total;
    const int rank = block_excl_count(fresh, s_warp, &total);
    const int base = s_created;
    if (fresh) {
      const int ci = base + rank;
      if (ci < F.max_cells) {
        slot_of[cell] = ci;
        F.cell_of[pool0 + ci] = cell;
      } else {
        my_flags |= DF_CELL_POOL_FULL;
      }
    }
    __syncthreads();
    if (tid == 0) s_created = min(base + total, F.max_cells);
    __syncthreads();
  }
  if (tid == 0) F.n_created[b] = s_created;
  __syncthreads();
  // pass 2: one thread per touched cell appends its run in scan order (NDTCell::addPoint, ndtcell.cpp:21-34)
  uint8_t* built = F.built + (size_t)b * F.ncells;
  for (int p = tid; p < n_pts; p += kDfThreads) {
    if (key[p] == ~0ull) continue;
    const int cell = static_cast<int>(key[p] >> 32);
    if (p > 0 && static_cast<int>(key[p - 1] >> 32) == cell) continue;  // not the head of its run
    const int ci = slot_of[cell];
    if (ci < 0) continue;  // pool full
    const size_t e = pool0 + ci;
    const int sl = F.slot[e];
    int cc = F.cur_count[e];
    double2 sum = F.cur_sum[e];
    unsigned tot = F.total[e];
    unsigned st = F.slot_start[e * kWindow + sl], ln = F.slot_len[e * kWindow + sl];
    double2* ring = F.ringbuf + e * (size_t)F.ring;
    for (int r = p; r < n_pts && key[r] != ~0ull && static_cast<int>(key[r] >> 32) == cell; ++r) {
      const double2 pt = q[static_cast<unsigned>(key[r])];
      if (cc == 0) {  // first point after the slot was closed: its old content goes (ndtcell.cpp:22-27)
        st = tot;
        ln = 0;
      }
      ++cc;
      sum.x = __dadd_rn(sum.x, pt.x);
      sum.y = __dadd_rn(sum.y, pt.y);
      ring[tot & static_cast<unsigned>(F.ring - 1)] = pt;
      ++tot;
      ++ln;
    }
    F.cur_count[e] = cc;
    F.cur_sum[e] = sum;
    F.total[e] = tot;
    F.slot_start[e * kWindow + sl] = st;
    F.slot_len[e * kWindow + sl] = ln;
    built[cell] = 0;  // ndtcell.cpp:33
  }
  if (my_flags) atomicOr(&F.flags[b], my_flags);
}

// ---- build -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDfBuildThreads) frame_build_kernel(DevFrames F) {
  const int b = blockIdx.x;
  const int n_created = F.n_created[b];
  const size_t pool0 = (size_t)b * F.max_cells;
  double* mean = F.mean + (size_t)b * F.ncells * 2;
  double* icov = F.icov + (size_t)b * F.ncells * 4;
  uint8_t* built = F.built + (size_t)b * F.ncells;
  int my_flags = 0;
  // DF_IRREGULAR_SIGMA describes the table as THIS build leaves it (every created cell is looked at below), so that a frame whose
  // one odd cell has been rebuilt into a regular one goes back to the point-sliced kernel; the other bits record events and stay
  if (threadIdx.x == 0) atomicAnd(&F.flags[b], ~DF_IRREGULAR_SIGMA);
  __syncthreads();
  for (int ci = threadIdx.x; ci < n_created; ci += blockDim.x) {
    const size_t e = pool0 + ci;
    const int cell = F.cell_of[e];
    const int sl = F.slot[e];
    const size_t w = e * kWindow + sl;
    // sliding window in O(1): global = global + current - what the slot held before (ndtcell.h:13-15)
    const double2 cur = F.cur_sum[e];
    double2 gs = F.glob_sum[e];
    const double2 old = F.part_sum[w];
    gs.x = __dadd_rn(__dadd_rn(gs.x, cur.x), -old.x);
    gs.y = __dadd_rn(__dadd_rn(gs.y, cur.y), -old.y);
    F.glob_sum[e] = gs;
    F.part_sum[w] = cur;
    const int cc = F.cur_count[e];
    const int gc = F.glob_count[e] + cc - F.part_count[w];
    F.glob_count[e] = gc;
    F.part_count[w] = cc;
    if (gc > 2) {  // ndtcell.cpp:43
      const double n = static_cast<double>(gc);
      const double mx = __ddiv_rn(gs.x, n), my = __ddiv_rn(gs.y, n);
      // scatter of the CURRENT slot's points about the GLOBAL mean (ndtcell.cpp:49-52)
      double c00 = 0., c01 = 0., c10 = 0., c11 = 0.;
      const unsigned st = F.slot_start[w], ln = F.slot_len[w];
      if (ln > 0 && F.total[e] - st > static_cast<unsigned>(F.ring)) my_flags |= DF_WINDOW_TRUNCATED;
      const double2* ring = F.ringbuf + e * (size_t)F.ring;
      for (unsigned k = 0; k < ln; ++k) {
        const double2 pt = ring[(st + k) & static_cast<unsigned>(F.ring - 1)];
        const double dx = __dadd_rn(pt.x, -mx), dy = __dadd_rn(pt.y, -my);
        c00 = __dadd_rn(c00, __dmul_rn(dx, dx));
        c01 = __dadd_rn(c01, __dmul_rn(dx, dy));
        c10 = __dadd_rn(c10, __dmul_rn(dy, dx));
        c11 = __dadd_rn(c11, __dmul_rn(dy, dy));
      }
      double* g = F.glob_cov + 4 * e;
      double* pc = F.part_cov + 4 * w;
      const double g00 = __dadd_rn(__dadd_rn(g[0], c00), -pc[0]);
      const double g01 = __dadd_rn(__dadd_rn(g[1], c01), -pc[1]);
      const double g10 = __dadd_rn(__dadd_rn(g[2], c10), -pc[2]);
      const double g11 = __dadd_rn(__dadd_rn(g[3], c11), -pc[3]);
      g[0] = g00;
      g[1] = g01;
      g[2] = g10;
      g[3] = g11;
      pc[0] = c00;
      pc[1] = c01;
      pc[2] = c10;
      pc[3] = c11;
      // inverse covariance with the eigenvalue-ratio floor (ndtcell.cpp:93-111)
      const double v00 = __ddiv_rn(g00, n), v01 = __ddiv_rn(g01, n), v10 = __ddiv_rn(g10, n), v11 = __ddiv_rn(g11, n);
      const double half_tr = __ddiv_rn(__dadd_rn(v00, v11), 2.);
      const double half_df = __ddiv_rn(__dadd_rn(v00, -v11), 2.);
      const double disc = __dadd_rn(__dmul_rn(half_df, half_df), __dmul_rn(v01, v10));
      const double root = disc > 0. ? __dsqrt_rn(disc) : 0.;
      const double e0 = __dadd_rn(half_tr, root), e1 = __dadd_rn(half_tr, -root);
      const double large = e0 > e1 ? e0 : e1;
      const double small = e0 < e1 ? e0 : e1;
      double det;
      if (small < __dmul_rn(.001, large))
        det = __dmul_rn(__dmul_rn(.001, large), large);  // the adjugate is kept, only the determinant is replaced
      else
        det = __dadd_rn(__dmul_rn(v00, v11), -__dmul_rn(v01, v10));
      const double s00 = __ddiv_rn(v11, det), s01 = __ddiv_rn(-v01, det), s10 = __ddiv_rn(-v10, det), s11 = __ddiv_rn(v00, det);
      mean[2 * (size_t)cell] = mx;
      mean[2 * (size_t)cell + 1] = my;
      icov[4 * (size_t)cell] = s00;
      icov[4 * (size_t)cell + 1] = s01;
      icov[4 * (size_t)cell + 2] = s10;
      icov[4 * (size_t)cell + 3] = s11;
      built[cell] = 1;
      // what the point-sliced PSO kernel requires of a table (the host checks the same on its path)
      const bool regular = (__double_as_longlong(s01) == __double_as_longlong(s10)) && s00 >= 0. && s11 >= 0. &&
                           (s00 * s11 - s01 * s10 >= 0.) && s00 < 1e300 && s11 < 1e300 && isfinite(mx) && isfinite(my);
      if (!regular) my_flags |= DF_IRREGULAR_SIGMA;
    } else if (built[cell]) {  // the cell keeps its table row: the row still counts
      const double mx = mean[2 * (size_t)cell], my = mean[2 * (size_t)cell + 1];
      const double s00 = icov[4 * (size_t)cell], s01 = icov[4 * (size_t)cell + 1], s10 = icov[4 * (size_t)cell + 2], s11 = icov[4 * (size_t)cell + 3];
      const bool regular = (__double_as_longlong(s01) == __double_as_longlong(s10)) && s00 >= 0. && s11 >= 0. &&
                           (s00 * s11 - s01 * s10 >= 0.) && s00 < 1e300 && s11 < 1e300 && isfinite(mx) && isfinite(my);
      if (!regular) my_flags |= DF_IRREGULAR_SIGMA;
    }
    if (cc > kMaxPerCell) {  // the slot is full: open the next one (ndtcell.cpp:61-65)
      F.slot[e] = (sl + 1) % kWindow;
      F.cur_count[e] = 0;
      F.cur_sum[e] = make_double2(0., 0.);
    }
  }
  if (my_flags) atomicOr(&F.flags[b], my_flags);
}

// ---- align bookkeeping ---------------------------------------------------------------------------
// guess: [n][3] device or nullptr (= node_pose); seeds: [n] device or nullptr (= 1, the default
// seed of a never-seeded process); rnd_base + b*rnd_stride = where K1 writes problem b's stream.
// from_host: the streams were uploaded (drop-in mode: drawn from the process-global std::rand()), K1 skips these problems.
__global__ void align_prepare_kernel(DevFrames F, const double* __restrict__ guess, const unsigned* __restrict__ seeds, int* rnd_base,
                                     size_t rnd_stride, int from_host) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= F.n) return;
  DevProblem& pr = F.probs[b];
  const int it = F.s_iter[b];
  for (int k = 0; k < 3; ++k) {
    // ndtframe.cpp:253: fixed spread for the first two calls, then |2 * s_pose_diff|
    pr.dev[k] = it < 2 ? (k == 2 ? 3.1415E-3 : .1) : fabs(__dmul_rn(F.pose_diff[3 * (size_t)b + k], 2.));
    pr.guess[k] = guess ? guess[3 * (size_t)b + k] : F.node_pose[3 * (size_t)b + k];
  }
  F.s_iter[b] = it + 1;  // ndtframe.cpp:255
  pr.seed = seeds ? seeds[b] : 1u;
  pr.rnd = rnd_base + (size_t)b * rnd_stride;
  pr.rnd_from_host = from_host;
}

__global__ void align_finish_kernel(DevFrames F) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= F.n) return;
  for (int k = 0; k < 3; ++k) {
    const double p = F.results[4 * (size_t)b + k];
    F.pose_diff[3 * (size_t)b + k] = __dadd_rn(p, -F.prev_pose[3 * (size_t)b + k]);  // ndtframe.cpp:263
    F.prev_pose[3 * (size_t)b + k] = p;                                              // ndtframe.cpp:264
    F.node_pose[3 * (size_t)b + k] = p;                                              // current_pose_, ndtpso_slam_node.cpp:194
  }
}

__global__ void set_node_pose_kernel(DevFrames F, const double* __restrict__ poses) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * F.n) F.node_pose[i] = poses ? poses[i] : 0.;
}

}  // namespace ndtpso
