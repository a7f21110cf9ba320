// Device side of the B200-native PSO/NDT scan matcher (sm_100a).
//
// Kernels (one launch each per batch):
//   compact_map_kernel  K0  dense/sparse (mu, Sigma^-1, built) table -> u16 row-strip lookup grid
//                           + packed 48-byte records {mu, -Sigma^-1/2}
//   rng_fill_kernel     K1  glibc TYPE_3 rand() stream of srand(seed), 32 values per warp step
//   pso_kernel          K2  the hot path: P+1+P*I NDT cost evaluations and the swarm update with
//                           the reference's SEQUENTIAL gbest order (speculate-and-replay)
//   cost_kernel             cost_function alone, for unit parity
//   fp64_peak_kernel        DFMA throughput probe (roofline denominator)
//
// Reference semantics followed (paths under /root/reference):
//   cost_function                lib/ndtpso_slam/core.cpp:26-48
//   transform_point              include/ndtpso_slam/core.h:28-31
//   NDTFrame::getCellIndex       lib/ndtpso_slam/ndtframe.cpp:240-249
//   NDTCell::normalDistribution  lib/ndtpso_slam/ndtcell.cpp:70-78
//   Particle ctor                lib/ndtpso_slam/core.cpp:13-23
//   pso_optimization             lib/ndtpso_slam/core.cpp:50-116
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fast_exp.h"

namespace ndtpso {

// ------------------------------------------------------------------------------------------
// Device-resident descriptors (built on the host, one H2D with the rest of the batch)
// ------------------------------------------------------------------------------------------
enum { HDR_NREC = 0, HDR_ROW0, HDR_NROWS, HDR_MODE, HDR_WORDS = 4 };
enum { MAP_COMPACT = 0, MAP_DENSE_DIRECT = 1 };  // hdr[HDR_MODE]

struct DevMap {
  double x_min, x_max, y_min, y_max;
  double hw, hh;        // width/2., height/2. (ndtframe.cpp:245-246)
  double cs, inv_cs;    // cell_side and 1/cell_side (exact when cs is a power of two)
  int gw, gh, ncells;
  int fast_geom;        // 1: cs is a power of two and the bounds are symmetric (x_min == -x_max, y_min == -y_max)
  // input table (device pointers)
  const double* mean;      // [rows][2]
  const double* icov;      // [rows][4]
  const uint8_t* built;    // [ncells] (dense form) or nullptr
  const int* cell_index;   // [n_sparse] (sparse form) or nullptr
  int n_sparse;            // < 0: dense
  int _pad;
  // compact table written by K0
  unsigned short* grid;    // [nrows*gw + 1] record id per cell of rows row0..row0+nrows-1; last = null id
  double* rec;             // [(n_rec + 1)][6]: {mx, my, -S00/2, -S01/2, -S10/2, -S11/2}; last = null record
  int* hdr;                // HDR_WORDS ints
};

struct DevProblem {
  const double2* pts;
  const int* rnd;       // rand() outputs, n_draws of them
  double guess[3];
  double dev[3];
  int n_pts;
  int map_id;
  unsigned seed;
  int rnd_from_host;    // 1: `rnd` was uploaded, K1 skips this problem
};

// Result exchange fused into the PSO kernels' epilogue (multi-GPU): instead of a collective after the
// kernel, the CTA that solved problem b stores its 32-byte result straight into the gathered result
// buffer of EVERY rank (its own and, over NVLink, its peers'), and the last CTA of the launch raises
// this rank's arrival flag on every rank.  world == 0: off.
constexpr int kStatsWords = 4;  // per problem: rounds, gbest updates, fp64 cost evaluations, evaluations settled by the fp32 screen
constexpr int kMaxPeers = 8;
struct PeerExchange {
  double* out[kMaxPeers];     // rank r's gathered buffer [world * n_per_rank][4] (this epoch's half)
  unsigned* flag[kMaxPeers];  // rank r's arrival flags [world]
  unsigned* done;             // this rank's counter of finished problems
  int world, my_rank, offset; // offset = first row of this rank's block
  unsigned epoch;
};

struct PsoParams {
  int P, I;
  double w, c1, c2, wd;
  int n_draws;      // 3 + 3P + 6PI (GLIR: 3(P + 2) + 6PI)
  int variant;      // 0 = pso_optimization (core.cpp:50-116), 1 = glir_pso_optimization (core.cpp:118-186; generic kernel only)
  int smem_bytes;   // dynamic shared memory given to pso_kernel
  int hot_chunk;    // point-sliced kernel: speculation window while gbest improves often (0 = always the whole swarm)
  int hot_thresh;   // improvements in an iteration that keep the next one's window small
  // fp32 screening of the point-sliced kernel (ndtpso_pso_sliced.cuh): 0 = off
  int screen;
  float scr_du;     // bound on the fp32 error of a transformed point's cell coordinate, in cell sides (rounded up)
  float scr_beta_c; // 0.5 - beta: a point closer than beta cell sides to a cell edge counts as worst case
  PeerExchange ex;
};

// called by the one thread that wrote problem b's result `o` = {x, y, theta, cost}
__device__ __forceinline__ void publish_result(const PeerExchange& ex, int b, int n_problems, const double* o) {
  if (ex.world <= 0) return;
  const double2 xy = make_double2(o[0], o[1]), tc = make_double2(o[2], o[3]);
  for (int r = 0; r < ex.world; ++r) {
    double2* dst = reinterpret_cast<double2*>(ex.out[r] + 4 * (size_t)(ex.offset + b));
    dst[0] = xy;
    dst[1] = tc;
  }
  __threadfence_system();  // the stores above are visible system-wide before this problem counts as done
  const unsigned old = atomicAdd(ex.done, 1u);
  if (old == static_cast<unsigned>(n_problems - 1)) {  // last problem of this launch
    *ex.done = 0u;          // the next launch on this stream starts after this one has ended
    __threadfence_system();  // orders every other CTA's stores (observed through the counter) before the flags
    for (int r = 0; r < ex.world; ++r)
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ex.flag[r] + ex.my_rank), "r"(ex.epoch) : "memory");
  }
}

// Stream-ordered wait for every rank's flag to reach `epoch` (bounded: sets *err after `timeout_cycles`).
__global__ void exchange_wait_kernel(const unsigned* flags, int world, unsigned epoch, long long timeout_cycles, int* err) {
  const int r = threadIdx.x;
  if (r >= world) return;
  const long long t0 = clock64();
  for (;;) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + r) : "memory");
    if (static_cast<int>(v - epoch) >= 0) break;
    if (clock64() - t0 > timeout_cycles) {
      atomicExch(err, 1 + r);
      break;
    }
    __nanosleep(200);
  }
}

__constant__ double c_exp_table[kExpTableSize] = {
#include "exp_table.inc"
};

// ------------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + 1-D bulk TMA (cp.async.bulk -> SASS UBLKCP)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__host__ __device__ __forceinline__ int round16(int x) { return (x + 15) & ~15; }

// ------------------------------------------------------------------------------------------
// K0: compact the NDT table.  One CTA per map.
//   pass A: count built cells per thread chunk + first/last occupied grid row -> block scan
//   pass B: fill the row strip with the null id, then write record ids and packed records
// Record order = ascending cell index (deterministic).  The grid covers whole rows
// row0..row0+nrows-1 so that the reference's FLAT index ix + gw*iy (including its wrap into the
// next row when (x + W/2)/cs rounds up to gw, ndtframe.cpp:245) addresses it directly.
// ------------------------------------------------------------------------------------------
constexpr int K0_THREADS = 1024;

__global__ void __launch_bounds__(K0_THREADS) compact_map_kernel(const DevMap* __restrict__ maps) {
  const DevMap& m = maps[blockIdx.x];
  const int tid = threadIdx.x;
  const bool sparse = m.n_sparse >= 0;
  const int rows = sparse ? m.n_sparse : m.ncells;
  const int chunk = (rows + K0_THREADS - 1) / K0_THREADS;
  const int lo = min(rows, tid * chunk), hi = min(rows, lo + chunk);

  __shared__ int s_scan[K0_THREADS];
  __shared__ int s_box[2];
  if (tid == 0) {
    s_box[0] = INT_MAX;  // min iy
    s_box[1] = -1;       // max iy
  }
  __syncthreads();

  int cnt = 0, ay = INT_MAX, by = -1;
  for (int r = lo; r < hi; ++r) {
    int cell;
    if (sparse) {
      cell = m.cell_index[r];
    } else {
      if (!m.built[r]) continue;
      cell = r;
    }
    const int iy = cell / m.gw;
    ay = min(ay, iy);
    by = max(by, iy);
    ++cnt;
  }
  if (cnt) {
    atomicMin(&s_box[0], ay);
    atomicMax(&s_box[1], by);
  }
  s_scan[tid] = cnt;
  __syncthreads();
  for (int off = 1; off < K0_THREADS; off <<= 1) {  // inclusive Hillis-Steele scan
    int v = (tid >= off) ? s_scan[tid - off] : 0;
    __syncthreads();
    s_scan[tid] += v;
    __syncthreads();
  }
  const int n_rec = s_scan[K0_THREADS - 1];
  int slot = s_scan[tid] - cnt;  // exclusive prefix

  const int row0 = n_rec ? s_box[0] : 0;
  const int nrows = n_rec ? s_box[1] - s_box[0] + 1 : 0;
  const bool compact_ok = n_rec <= 65534;
  if (tid == 0) {
    m.hdr[HDR_NREC] = n_rec;
    m.hdr[HDR_ROW0] = row0;
    m.hdr[HDR_NROWS] = nrows;
    m.hdr[HDR_MODE] = compact_ok ? MAP_COMPACT : MAP_DENSE_DIRECT;
  }
  if (!compact_ok) return;  // uniform: the hot kernel reads the dense arrays directly

  const int span = nrows * m.gw;
  const unsigned short null_id = static_cast<unsigned short>(n_rec);
  for (int g = tid; g <= span; g += K0_THREADS) m.grid[g] = null_id;
  if (tid < 6) m.rec[6 * (size_t)n_rec + tid] = 0.;  // null record (never contributes; keeps loads in range)
  __syncthreads();

  const int base = row0 * m.gw;
  for (int r = lo; r < hi; ++r) {
    int cell;
    if (sparse) {
      cell = m.cell_index[r];
    } else {
      if (!m.built[r]) continue;
      cell = r;
    }
    m.grid[cell - base] = static_cast<unsigned short>(slot);
    double* rec = m.rec + 6 * (size_t)slot;
    const double2 mu = *reinterpret_cast<const double2*>(m.mean + 2 * (size_t)r);
    const double2 s0 = *reinterpret_cast<const double2*>(m.icov + 4 * (size_t)r);
    const double2 s1 = *reinterpret_cast<const double2*>(m.icov + 4 * (size_t)r + 2);
    // -S/2 is an exact scaling: -(d'Sd)/2 (ndtcell.cpp:74-75) == d'(-S/2)d bit for bit
    *reinterpret_cast<double2*>(rec) = mu;
    *reinterpret_cast<double2*>(rec + 2) = make_double2(-0.5 * s0.x, -0.5 * s0.y);
    *reinterpret_cast<double2*>(rec + 4) = make_double2(-0.5 * s1.x, -0.5 * s1.y);
    ++slot;
  }
}

// ------------------------------------------------------------------------------------------
// K1: glibc rand() after srand(seed)  (stdlib/random_r.c TYPE_3: degree 31, separation 3)
//
// Written as one linear sequence z[n] = z[n-31] + z[n-3] (mod 2^32) with
// z[0..27] = state[3..30], z[28..30] = state[0..2], state[] = the 16807 Lehmer fill of srandom_r.
// rand() number k (after srandom_r's 310 discarded outputs) is z[341 + k] >> 1.
// Applying the recurrence 11 times gives  z[n] = sum_k C(11,k) z[n - 33 - 28k],  all lags >= 33,
// so a warp produces 32 consecutive values per step from its shared-memory history.
// ------------------------------------------------------------------------------------------
constexpr int K1_WARPS = 4;
constexpr int K1_RING = 512;  // >= 341 + 32, power of two
// Persistent generator state of one stream (NDTPSO_RNG_CONTINUE: the reference never seeds, so a
// process' rand() stream continues from one align() to the next): the ring, where it stands, and
// how many generated values have not been handed out yet.
enum { K1_ST_NMOD = K1_RING, K1_ST_AHEAD, K1_ST_INIT, K1_ST_PAD, K1_STATE_WORDS };

// `state` == nullptr: every problem's stream is that of srand(probs[b].seed).
// `state` != nullptr: problem b continues the stream kept in state[b] (seeded with probs[b].seed on first use).
template <bool PERSIST>
__global__ void __launch_bounds__(K1_WARPS * 32) rng_fill_kernel(const DevProblem* __restrict__ probs, int n_problems, int n_draws,
                                                                 uint32_t* __restrict__ state) {
  __shared__ uint32_t ring_all[K1_WARPS][K1_RING];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * K1_WARPS + warp;
  if (b >= n_problems) return;
  if (probs[b].rnd_from_host) return;
  uint32_t* z = ring_all[warp];
  int* out = const_cast<int*>(probs[b].rnd);
  uint32_t* st = PERSIST ? state + (size_t)b * K1_STATE_WORDS : nullptr;

  // next index to generate, in ring coordinates: the same constant for a fresh and for a resumed stream (a saved ring is
  // stored oldest value first and reloaded rotated), so every ring index below is a compile-time constant plus the lane
  constexpr int n_loc = K1_RING + 373;
  int ahead;  // values generated but not handed out yet
  if (PERSIST && st[K1_ST_INIT]) {
    for (int i = lane; i < K1_RING; i += 32) z[(n_loc + i) & (K1_RING - 1)] = st[i];
    ahead = static_cast<int>(st[K1_ST_AHEAD]);
    __syncwarp();
  } else {
    if (lane == 0) {
      uint32_t seed = probs[b].seed;
      if (seed == 0) seed = 1;
      // state[i], i = 0..30 ; z index of state[i] is (i + 28) % 31
      int word = static_cast<int>(seed);  // glibc keeps `word` in an int32_t: seeds >= 2^31 go negative
      z[28] = seed;
      for (int i = 1; i < 31; ++i) {
        const long long hi = word / 127773, lo = word % 127773;
        long long t = 16807 * lo - 2836 * hi;
        if (t < 0) t += 2147483647;
        word = static_cast<int>(t);
        z[(i + 28) % 31] = static_cast<uint32_t>(word);
      }
    }
    __syncwarp();
    // z[31 .. 372]: three values per step (lag 3 is the shortest)
    for (int n = 31; n < 373; n += 3) {
      if (lane < 3 && n + lane < 373) z[n + lane] = z[n + lane - 31] + z[n + lane - 3];
      __syncwarp();
    }
    ahead = 32;  // rand k = z[341 + k] >> 1: the first 32 are already in the history
  }
  // values that are already in the history
  for (int j = lane; j < min(ahead, n_draws); j += 32) out[j] = static_cast<int>(z[(n_loc - ahead + j) & (K1_RING - 1)] >> 1);
  // then 32 per step
  int t0 = 0;
  for (; ahead + t0 < n_draws; t0 += 32) {
    const int n = n_loc + t0 + lane;
    uint32_t v = z[(n - 33) & (K1_RING - 1)] + z[(n - 341) & (K1_RING - 1)];
    v += 11u * (z[(n - 61) & (K1_RING - 1)] + z[(n - 313) & (K1_RING - 1)]);
    v += 55u * (z[(n - 89) & (K1_RING - 1)] + z[(n - 285) & (K1_RING - 1)]);
    v += 165u * (z[(n - 117) & (K1_RING - 1)] + z[(n - 257) & (K1_RING - 1)]);
    v += 330u * (z[(n - 145) & (K1_RING - 1)] + z[(n - 229) & (K1_RING - 1)]);
    v += 462u * (z[(n - 173) & (K1_RING - 1)] + z[(n - 201) & (K1_RING - 1)]);
    __syncwarp();
    z[n & (K1_RING - 1)] = v;
    const int j = ahead + t0 + lane;
    if (j < n_draws) out[j] = static_cast<int>(v >> 1);
    __syncwarp();
  }
  if (PERSIST) {
    for (int i = lane; i < K1_RING; i += 32) st[i] = z[(n_loc + t0 + i) & (K1_RING - 1)];  // oldest first
    if (lane == 0) {
      st[K1_ST_NMOD] = 0u;
      st[K1_ST_AHEAD] = static_cast<uint32_t>(ahead + t0 - n_draws);
      st[K1_ST_INIT] = 1u;
    }
  }
}

// ------------------------------------------------------------------------------------------
// NDT cost of one candidate pose, evaluated by one warp (lanes stride over the points).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;  // identical in every lane; fixed butterfly order => deterministic
}

// ---- fast path: table staged in shared memory, branch-free, custom exp -------------------
// FAST_GEOM = cell side a power of two and symmetric bounds (every frame the reference builds,
// ndtframe.cpp:57-65, with cell sides 0.25/0.5/1/2): bounds test |x| < x_max, cell coordinate
// fma(x, 1/cs, (W/2)/cs) (bit-identical to (x + W/2)/cs because scaling by 2^k is exact).
// Otherwise: four strict compares and a true IEEE division.
struct FastCost {
  const double2* pts;          // shared
  const unsigned short* grid;  // shared, [span + 1]
  const double* rec;           // shared, [(n_rec + 1) * 6]
  const double* etab;          // shared, [64]
  double x_min, x_max, y_min, y_max;
  double hw, hh, cs, inv_cs, hw_s, hh_s;  // hw_s = hw * inv_cs
  int n, gw, base, span, null_id;

  template <bool FAST_GEOM>
  __device__ __forceinline__ double point(int i, double c, double s, double tx, double ty) const {
    const double2 p = pts[i];
    const double x = fma(p.x, c, fma(-p.y, s, tx));  // transform_point, core.h:29-30
    const double y = fma(p.x, s, fma(p.y, c, ty));
    bool inb;
    double u, v;
    if (FAST_GEOM) {
      inb = (fabs(x) < x_max) && (fabs(y) < y_max);  // strict, ndtframe.cpp:242
      u = fma(x, inv_cs, hw_s);
      v = fma(y, inv_cs, hh_s);
    } else {
      inb = (x > x_min) && (x < x_max) && (y > y_min) && (y < y_max);
      u = __ddiv_rn(x + hw, cs);
      v = __ddiv_rn(y + hh, cs);
    }
    const int ix = __double2int_rd(u), iy = __double2int_rd(v);  // floor, ndtframe.cpp:245-246
    const unsigned g = static_cast<unsigned>(ix + gw * iy - base);
    const bool in_strip = inb && (g < static_cast<unsigned>(span));
    const unsigned r = grid[in_strip ? g : static_cast<unsigned>(span)];
    const double* q = rec + 6 * r;
    const double2 mu = *reinterpret_cast<const double2*>(q);
    const double2 h0 = *reinterpret_cast<const double2*>(q + 2);
    const double2 h1 = *reinterpret_cast<const double2*>(q + 4);
    const double d0 = x - mu.x, d1 = y - mu.y;  // normalDistribution, ndtcell.cpp:72-75
    const double r0 = fma(d1, h1.x, d0 * h0.x);
    const double r1 = fma(d1, h1.y, d0 * h0.y);
    const double a = fma(r1, d1, r0 * d0);       // = -(d' S d)/2
    const double e = fast_exp(a, etab);
    return (in_strip && r != static_cast<unsigned>(null_id)) ? e : 0.0;
  }

  template <bool FAST_GEOM>
  __device__ __forceinline__ double eval(double tx, double ty, double th, int lane) const {
    double s, c;
    sincos(th, &s, &c);
    double acc0 = 0., acc1 = 0.;
    int i = lane;
    for (; i + 32 < n; i += 64) {
      const double e0 = point<FAST_GEOM>(i, c, s, tx, ty);
      const double e1 = point<FAST_GEOM>(i + 32, c, s, tx, ty);
      acc0 -= e0;
      acc1 -= e1;
    }
    if (i < n) acc0 -= point<FAST_GEOM>(i, c, s, tx, ty);
    return warp_sum(acc0 + acc1);
  }
};

// ---- fallback: table (and possibly points) in global memory; library exp --------------------
// Used when the compact table does not fit the CTA's shared memory or has more than 65534 built
// cells (then it reads the caller's dense arrays directly).
struct SlowCost {
  const double2* pts;  // generic
  double x_min, x_max, y_min, y_max, hw, hh, cs;
  int n, gw, ncells, base, span, null_id;
  bool dense;
  const unsigned short* grid;
  const double* rec;
  const double* mean;
  const double* icov;
  const uint8_t* built;

  __device__ __forceinline__ double eval(double tx, double ty, double th, int lane) const {
    double s, c;
    sincos(th, &s, &c);
    double acc = 0.;
    for (int i = lane; i < n; i += 32) {
      const double2 p = pts[i];
      const double x = fma(p.x, c, fma(-p.y, s, tx));
      const double y = fma(p.x, s, fma(p.y, c, ty));
      if (!((x > x_min) && (x < x_max) && (y > y_min) && (y < y_max))) continue;
      const int ix = __double2int_rd(__ddiv_rn(x + hw, cs)), iy = __double2int_rd(__ddiv_rn(y + hh, cs));
      const int idx = ix + gw * iy;
      double mx, my, h00, h01, h10, h11;
      if (dense) {
        if (idx < 0 || idx >= ncells || !built[idx]) continue;
        mx = mean[2 * idx];
        my = mean[2 * idx + 1];
        h00 = -0.5 * icov[4 * idx];
        h01 = -0.5 * icov[4 * idx + 1];
        h10 = -0.5 * icov[4 * idx + 2];
        h11 = -0.5 * icov[4 * idx + 3];
      } else {
        const unsigned g = static_cast<unsigned>(idx - base);
        if (g >= static_cast<unsigned>(span)) continue;
        const unsigned r = grid[g];
        if (r == static_cast<unsigned>(null_id)) continue;
        const double* q = rec + 6 * r;
        mx = q[0];
        my = q[1];
        h00 = q[2];
        h01 = q[3];
        h10 = q[4];
        h11 = q[5];
      }
      const double d0 = x - mx, d1 = y - my;
      const double r0 = fma(d1, h10, d0 * h00);
      const double r1 = fma(d1, h11, d0 * h01);
      acc -= exp(fma(r1, d1, r0 * d0));
    }
    return warp_sum(acc);
  }
};

enum { COST_FAST_GEOM = 0, COST_FAST_ANY = 1, COST_SLOW = 2 };

template <int MODE>
struct CostOf;
template <>
struct CostOf<COST_FAST_GEOM> {
  const FastCost& f;
  __device__ __forceinline__ double operator()(double tx, double ty, double th, int lane) const { return f.eval<true>(tx, ty, th, lane); }
};
template <>
struct CostOf<COST_FAST_ANY> {
  const FastCost& f;
  __device__ __forceinline__ double operator()(double tx, double ty, double th, int lane) const { return f.eval<false>(tx, ty, th, lane); }
};
template <>
struct CostOf<COST_SLOW> {
  const SlowCost& f;
  __device__ __forceinline__ double operator()(double tx, double ty, double th, int lane) const { return f.eval(tx, ty, th, lane); }
};

// Eigen Random(): x + (y-x)*Scalar(rand())/Scalar(RAND_MAX) with x=-1, y=1; no fusion.
__device__ __forceinline__ double unit_random(int r) {
  return __dadd_rn(-1.0, __ddiv_rn(__dmul_rn(2.0, static_cast<double>(r)), 2147483647.0));
}

struct __align__(16) Cand {
  double c, x, y, th;
};

// Shared-memory carve-up of pso_kernel (all offsets multiples of 16 bytes)
struct PsoSmem {
  uint64_t* bar;
  double* etab;   // [64]
  Cand* cand;     // [2][P+1]
  double* x;      // [P][3]
  double* v;      // [P][3]
  double* vnew;   // [P][3]
  double* pb;     // [P][3]
  double* pbc;    // [P]
  double* pavg;   // [P]  Particle::pbest_average (GLIR only)
  unsigned char* dyn;  // start of the staged problem data
  int fixed_bytes;
};

__host__ __device__ inline int pso_fixed_smem_bytes(int P) {
  int b = 16;                                       // mbarrier
  b += kExpTableSize * (int)sizeof(double);         // exp table
  b += 2 * (P + 1) * (int)sizeof(Cand);             // candidates, double buffered
  b += (P > 0 ? P : 1) * 14 * (int)sizeof(double);  // x, v, vnew, pb (3 each) + pbc + pavg
  return (b + 15) & ~15;
}

__device__ __forceinline__ PsoSmem carve_smem(unsigned char* base, int P) {
  PsoSmem s;
  const int Pn = P > 0 ? P : 1;
  s.bar = reinterpret_cast<uint64_t*>(base);
  s.etab = reinterpret_cast<double*>(base + 16);
  s.cand = reinterpret_cast<Cand*>(base + 16 + kExpTableSize * sizeof(double));
  double* d = reinterpret_cast<double*>(base + 16 + kExpTableSize * sizeof(double) + 2 * (P + 1) * sizeof(Cand));
  s.x = d;
  s.v = d + 3 * Pn;
  s.vnew = d + 6 * Pn;
  s.pb = d + 9 * Pn;
  s.pbc = d + 12 * Pn;
  s.pavg = d + 13 * Pn;
  s.fixed_bytes = pso_fixed_smem_bytes(P);
  s.dyn = base + s.fixed_bytes;
  return s;
}

// ------------------------------------------------------------------------------------------
// K2 body: the swarm.  NW warps of one CTA share one problem; particle j belongs to warp j % NW
// for the whole run, so particle state is private to its warp and only candidates are shared.
//
// Sequential-gbest semantics (core.cpp:82-106 in single-thread order) via speculate-and-replay:
// every pending particle computes its candidate from the CURRENT gbest; after one barrier every
// warp finds j* = the first pending particle whose candidate beats gbest; particles <= j* commit,
// gbest becomes j*'s candidate and the particles after j* are replayed against it.  Candidates are
// double-buffered so one barrier per round suffices.
// ------------------------------------------------------------------------------------------
template <class Cost, int NW>
__device__ __forceinline__ void pso_body(const Cost& cost, const DevProblem& pr, const PsoParams& prm, const PsoSmem& sm,
                                         double* __restrict__ out, int* __restrict__ stats) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = prm.P, I = prm.I;
  const int* __restrict__ rnd = pr.rnd;
  Cand* cand0 = sm.cand;
  Cand* cand1 = sm.cand + (P + 1);

  // ---- initial swarm: task 0 = the seed particle (core.cpp:53,58), task 1+j = particle j (core.cpp:60-61)
  for (int t = warp; t < P + 1; t += NW) {
    const bool seed_particle = (t == 0);
    double pos[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double dv = seed_particle ? (k == 2 ? 1E-5 : 1E-4) : pr.dev[k];
      pos[k] = __dadd_rn(pr.guess[k], __dmul_rn(unit_random(rnd[3 * t + k]), dv));
    }
    const double c = cost(pos[0], pos[1], pos[2], lane);
    if (lane == 0) cand0[t] = Cand{c, pos[0], pos[1], pos[2]};
  }
  __syncthreads();
  // every warp derives the initial gbest the way core.cpp:58-69 does (strict <, index order)
  double gbc = cand0[0].c, gb0 = cand0[0].x, gb1 = cand0[0].y, gb2 = cand0[0].th;
  for (int j = 0; j < P; ++j) {
    const Cand cd = cand0[1 + j];
    if (cd.c < gbc) {
      gbc = cd.c;
      gb0 = cd.x;
      gb1 = cd.y;
      gb2 = cd.th;
    }
  }
  // owners initialise their particles' private state
  for (int j = warp + lane * NW; j < P; j += 32 * NW) {
    const Cand cd = cand0[1 + j];
    sm.x[3 * j] = cd.x;
    sm.x[3 * j + 1] = cd.y;
    sm.x[3 * j + 2] = cd.th;
    sm.pb[3 * j] = cd.x;
    sm.pb[3 * j + 1] = cd.y;
    sm.pb[3 * j + 2] = cd.th;
    sm.v[3 * j] = sm.v[3 * j + 1] = sm.v[3 * j + 2] = 0.;
    sm.pbc[j] = cd.c;
  }
  __syncwarp();

  // ---- iterations
  int it = 0, start = 0, par = 1, rounds = 0, n_gb = 0;
  double w = prm.w;
  while (it < I) {
    Cand* cand = par ? cand1 : cand0;
    // first pending particle owned by this warp
    const int j0 = start + ((warp - start) % NW + NW) % NW;
    for (int j = j0; j < P; j += NW) {
      const int base = 3 + 3 * P + 6 * P * it + 6 * j;
      double u = 0.;
      if (lane < 6) u = fabs(unit_random(rnd[base + lane]));  // Array2d::Random().abs(), core.cpp:84
      double nx[3], nv[3];
      const double gb[3] = {gb0, gb1, gb2};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double rx = __shfl_sync(0xffffffffu, u, 2 * k);
        const double ry = __shfl_sync(0xffffffffu, u, 2 * k + 1);
        const double xk = sm.x[3 * j + k], vk = sm.v[3 * j + k], pbk = sm.pb[3 * j + k];
        // core.cpp:85-87: ((w*v) + ((c1*rx)*(pb-x))) + ((c2*ry)*(gb-x)), no contraction
        const double t1 = __dmul_rn(w, vk);
        const double t2 = __dmul_rn(__dmul_rn(prm.c1, rx), __dadd_rn(pbk, -xk));
        const double t3 = __dmul_rn(__dmul_rn(prm.c2, ry), __dadd_rn(gb[k], -xk));
        nv[k] = __dadd_rn(__dadd_rn(t1, t2), t3);
        nx[k] = __dadd_rn(xk, nv[k]);  // core.cpp:89
      }
      const double c = cost(nx[0], nx[1], nx[2], lane);
      if (lane == 0) {
        cand[j] = Cand{c, nx[0], nx[1], nx[2]};
        sm.vnew[3 * j] = nv[0];
        sm.vnew[3 * j + 1] = nv[1];
        sm.vnew[3 * j + 2] = nv[2];
      }
    }
    __syncthreads();
    // j* = first pending particle that improves gbest (core.cpp:98)
    int jstar = -1;
    for (int base = start; base < P && jstar < 0; base += 32) {
      const int j = base + lane;
      const bool imp = (j < P) && (cand[j].c < gbc);
      const unsigned mask = __ballot_sync(0xffffffffu, imp);
      if (mask) jstar = base + __ffs(mask) - 1;
    }
    const int end = (jstar >= 0) ? jstar + 1 : P;
    // commit own particles in [start, end)  (core.cpp:89-96)
    for (int j = j0 + lane * NW; j < end; j += 32 * NW) {
      const Cand cd = cand[j];
      sm.x[3 * j] = cd.x;
      sm.x[3 * j + 1] = cd.y;
      sm.x[3 * j + 2] = cd.th;
      sm.v[3 * j] = sm.vnew[3 * j];
      sm.v[3 * j + 1] = sm.vnew[3 * j + 1];
      sm.v[3 * j + 2] = sm.vnew[3 * j + 2];
      if (cd.c < sm.pbc[j]) {
        sm.pbc[j] = cd.c;
        sm.pb[3 * j] = cd.x;
        sm.pb[3 * j + 1] = cd.y;
        sm.pb[3 * j + 2] = cd.th;
      }
    }
    __syncwarp();
    if (jstar >= 0) {  // core.cpp:102-103
      const Cand cd = cand[jstar];
      gbc = cd.c;
      gb0 = cd.x;
      gb1 = cd.y;
      gb2 = cd.th;
      ++n_gb;
    }
    start = end;
    if (start >= P) {
      start = 0;
      ++it;
      w = __dmul_rn(w, prm.wd);  // core.cpp:108
    }
    par ^= 1;
    ++rounds;
  }

  if (threadIdx.x == 0) {
    out[0] = gb0;
    out[1] = gb1;
    out[2] = gb2;
    out[3] = gbc;
    if (stats) {
      stats[0] = rounds;
      stats[1] = n_gb;
      stats[2] = 0;
      stats[3] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// GLIR variant (glir_pso_optimization, core.cpp:118-186; "UNTESTED" in the reference, which never calls it).  Same
// warp-per-particle, speculate-and-replay structure as pso_body: a pending particle's candidate depends on the current
// gbest (position AND cost: omega, c1 = c2 and best_ratio, core.cpp:146-150) and on its own committed state.
//   * global_best is a particle of its own drawn with the caller's deviation (:125; zero_devi is never used): task 0.
//   * P + 1 particles are constructed (:132-135), 3(P + 2) draws before the first iteration; particles[P] is never
//     compared (:137) nor iterated (:145): its draws are skipped over and its cost is not evaluated.
//   * every operation is a separate IEEE add/multiply/divide in the reference's association order (no contraction).
// Costs enter the arithmetic here (not only comparisons), so poses follow the reference to rounding-level differences
// of the cost (~1e-13 relative), not bit for bit.
// ------------------------------------------------------------------------------------------
template <class Cost, int NW>
__device__ __forceinline__ void glir_body(const Cost& cost, const DevProblem& pr, const PsoParams& prm, const PsoSmem& sm,
                                          double* __restrict__ out, int* __restrict__ stats) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int P = prm.P, I = prm.I;
  const int* __restrict__ rnd = pr.rnd;
  Cand* cand0 = sm.cand;
  Cand* cand1 = sm.cand + (P + 1);

  for (int t = warp; t < P + 1; t += NW) {  // Particle ctor, core.cpp:13-23
    double pos[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) pos[k] = __dadd_rn(pr.guess[k], __dmul_rn(unit_random(rnd[3 * t + k]), pr.dev[k]));
    const double c = cost(pos[0], pos[1], pos[2], lane);
    if (lane == 0) cand0[t] = Cand{c, pos[0], pos[1], pos[2]};
  }
  __syncthreads();
  double gbc = cand0[0].c, gb0 = cand0[0].x, gb1 = cand0[0].y, gb2 = cand0[0].th;
  for (int j = 0; j < P; ++j) {  // core.cpp:137-140, particles[0 .. P-1] in order
    const Cand cd = cand0[1 + j];
    if (cd.c < gbc) {
      gbc = cd.c;
      gb0 = cd.x;
      gb1 = cd.y;
      gb2 = cd.th;
    }
  }
  for (int j = warp + lane * NW; j < P; j += 32 * NW) {
    const Cand cd = cand0[1 + j];
    sm.x[3 * j] = cd.x;
    sm.x[3 * j + 1] = cd.y;
    sm.x[3 * j + 2] = cd.th;
    sm.pb[3 * j] = cd.x;
    sm.pb[3 * j + 1] = cd.y;
    sm.pb[3 * j + 2] = cd.th;
    sm.v[3 * j] = sm.v[3 * j + 1] = sm.v[3 * j + 2] = 0.;
    sm.pbc[j] = cd.c;
    sm.pavg[j] = cd.c;  // core.cpp:22
  }
  __syncwarp();

  int it = 0, start = 0, par = 1, rounds = 0, n_gb = 0;
  while (it < I) {
    Cand* cand = par ? cand1 : cand0;
    const int j0 = start + ((warp - start) % NW + NW) % NW;
    for (int j = j0; j < P; j += NW) {
      const int base = 3 * (P + 2) + 6 * P * it + 6 * j;
      double u = 0.;
      if (lane < 6) u = fabs(unit_random(rnd[base + lane]));  // Array2d::Random().abs(), core.cpp:149
      const double omega = __dadd_rn(1.1, -__ddiv_rn(gbc, __ddiv_rn(sm.pavg[j], static_cast<double>(j + 1))));  // core.cpp:146
      const double c12 = __dadd_rn(1.0, __ddiv_rn(gbc, sm.pbc[j]));                                             // core.cpp:147
      double nx[3], nv[3];
      const double gb[3] = {gb0, gb1, gb2};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double rx = __shfl_sync(0xffffffffu, u, 2 * k);
        const double ry = __shfl_sync(0xffffffffu, u, 2 * k + 1);
        const double xk = sm.x[3 * j + k], vk = sm.v[3 * j + k], pbk = sm.pb[3 * j + k];
        const double ratio = __ddiv_rn(pbk, gb[k]);  // core.cpp:150
        // core.cpp:151-153: ((omega*v) + ((c1*rx)*(ratio*pb - x))) + ((c2*ry)*((1./ratio)*gb - x))
        const double t1 = __dmul_rn(omega, vk);
        const double t2 = __dmul_rn(__dmul_rn(c12, rx), __dadd_rn(__dmul_rn(ratio, pbk), -xk));
        const double t3 = __dmul_rn(__dmul_rn(c12, ry), __dadd_rn(__dmul_rn(__ddiv_rn(1., ratio), gb[k]), -xk));
        nv[k] = __dadd_rn(__dadd_rn(t1, t2), t3);
        nx[k] = __dadd_rn(xk, nv[k]);  // core.cpp:155
      }
      const double c = cost(nx[0], nx[1], nx[2], lane);
      if (lane == 0) {
        cand[j] = Cand{c, nx[0], nx[1], nx[2]};
        sm.vnew[3 * j] = nv[0];
        sm.vnew[3 * j + 1] = nv[1];
        sm.vnew[3 * j + 2] = nv[2];
      }
    }
    __syncthreads();
    int jstar = -1;  // first pending particle that improves gbest (core.cpp:168)
    for (int base = start; base < P && jstar < 0; base += 32) {
      const int j = base + lane;
      const bool imp = (j < P) && (cand[j].c < gbc);
      const unsigned mask = __ballot_sync(0xffffffffu, imp);
      if (mask) jstar = base + __ffs(mask) - 1;
    }
    const int end = (jstar >= 0) ? jstar + 1 : P;
    for (int j = j0 + lane * NW; j < end; j += 32 * NW) {  // commit, core.cpp:155-166
      const Cand cd = cand[j];
      sm.x[3 * j] = cd.x;
      sm.x[3 * j + 1] = cd.y;
      sm.x[3 * j + 2] = cd.th;
      sm.v[3 * j] = sm.vnew[3 * j];
      sm.v[3 * j + 1] = sm.vnew[3 * j + 1];
      sm.v[3 * j + 2] = sm.vnew[3 * j + 2];
      double pbcj = sm.pbc[j];
      if (cd.c < pbcj) {
        pbcj = cd.c;
        sm.pbc[j] = cd.c;
        sm.pb[3 * j] = cd.x;
        sm.pb[3 * j + 1] = cd.y;
        sm.pb[3 * j + 2] = cd.th;
      }
      sm.pavg[j] = __dadd_rn(sm.pavg[j], pbcj);  // core.cpp:166
    }
    __syncwarp();
    if (jstar >= 0) {  // core.cpp:172-173: the particle's best fields, which the commit above just set to this candidate
      const Cand cd = cand[jstar];  // (cost < gbest_cost <= pbest_cost_j for every j < P, an invariant of :137-140,:161-173)
      gbc = cd.c;
      gb0 = cd.x;
      gb1 = cd.y;
      gb2 = cd.th;
      ++n_gb;
    }
    start = end;
    if (start >= P) {
      start = 0;
      ++it;
    }
    par ^= 1;
    ++rounds;
  }

  if (threadIdx.x == 0) {
    out[0] = gb0;
    out[1] = gb1;
    out[2] = gb2;
    out[3] = gbc;
    if (stats) {
      stats[0] = rounds;
      stats[1] = n_gb;
      stats[2] = 0;
      stats[3] = 0;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Staging: points + compact table -> shared memory with bulk TMA; builds the cost evaluators.
// ------------------------------------------------------------------------------------------
struct Staged {
  FastCost fast;
  SlowCost slow;
  int mode;  // COST_*
};

// dyn: 16-byte aligned shared memory of dyn_bytes; etab: 64 doubles of shared memory
__device__ __forceinline__ Staged stage_problem(const DevProblem& pr, const DevMap& mp, unsigned char* dyn, int dyn_bytes, double* etab,
                                                uint64_t* bar) {
  Staged st;
  const int n_rec = mp.hdr[HDR_NREC];
  const int row0 = mp.hdr[HDR_ROW0], nrows = mp.hdr[HDR_NROWS];
  const int mode = mp.hdr[HDR_MODE];
  const int span = nrows * mp.gw, base = row0 * mp.gw;

  const int pts_bytes = pr.n_pts * 16;
  const int rec_bytes = (n_rec + 1) * 48;
  const int grid_bytes = round16((span + 1) * 2);
  const bool pts_fit = pts_bytes <= dyn_bytes;
  const bool table_fit = (mode == MAP_COMPACT) && (pts_bytes + rec_bytes + grid_bytes <= dyn_bytes);

  if (threadIdx.x < kExpTableSize) etab[threadIdx.x] = c_exp_table[threadIdx.x];
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t tx = 0;
  if (pts_fit) tx += pts_bytes;
  if (table_fit) tx += rec_bytes + grid_bytes;
  if (tx > 0) {
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, tx);
      if (pts_fit && pts_bytes) tma_load_1d(dyn, pr.pts, pts_bytes, bar);
      if (table_fit) tma_load_1d(dyn + pts_bytes, mp.rec, rec_bytes, bar);
      if (table_fit) tma_load_1d(dyn + pts_bytes + rec_bytes, mp.grid, grid_bytes, bar);
    }
    mbar_wait(bar, 0);
  }

  FastCost& f = st.fast;
  f.pts = reinterpret_cast<const double2*>(dyn);
  f.rec = reinterpret_cast<const double*>(dyn + pts_bytes);
  f.grid = reinterpret_cast<const unsigned short*>(dyn + pts_bytes + rec_bytes);
  f.etab = etab;
  f.x_min = mp.x_min;
  f.x_max = mp.x_max;
  f.y_min = mp.y_min;
  f.y_max = mp.y_max;
  f.hw = mp.hw;
  f.hh = mp.hh;
  f.cs = mp.cs;
  f.inv_cs = mp.inv_cs;
  f.hw_s = mp.hw * mp.inv_cs;
  f.hh_s = mp.hh * mp.inv_cs;
  f.n = pr.n_pts;
  f.gw = mp.gw;
  f.base = base;
  f.span = span;
  f.null_id = n_rec;

  SlowCost& s = st.slow;
  s.pts = pts_fit ? reinterpret_cast<const double2*>(dyn) : pr.pts;
  s.x_min = mp.x_min;
  s.x_max = mp.x_max;
  s.y_min = mp.y_min;
  s.y_max = mp.y_max;
  s.hw = mp.hw;
  s.hh = mp.hh;
  s.cs = mp.cs;
  s.n = pr.n_pts;
  s.gw = mp.gw;
  s.ncells = mp.ncells;
  s.base = base;
  s.span = span;
  s.null_id = n_rec;
  s.dense = (mode != MAP_COMPACT);
  s.grid = mp.grid;
  s.rec = mp.rec;
  s.mean = mp.mean;
  s.icov = mp.icov;
  s.built = mp.built;

  st.mode = table_fit ? (mp.fast_geom ? COST_FAST_GEOM : COST_FAST_ANY) : COST_SLOW;
  return st;
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) pso_kernel(const DevProblem* __restrict__ probs, const DevMap* __restrict__ maps, PsoParams prm,
                                                      double* __restrict__ out, int* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const DevProblem& pr = probs[b];
  const DevMap& mp = maps[pr.map_id];
  PsoSmem sm = carve_smem(smem_raw, prm.P);
  const Staged st = stage_problem(pr, mp, sm.dyn, prm.smem_bytes - sm.fixed_bytes, sm.etab, sm.bar);
  double* o = out + 4 * (size_t)b;
  int* s = stats ? stats + kStatsWords * (size_t)b : nullptr;
  if (prm.variant == 1) {
    if (st.mode == COST_FAST_GEOM) {
      glir_body<CostOf<COST_FAST_GEOM>, NW>(CostOf<COST_FAST_GEOM>{st.fast}, pr, prm, sm, o, s);
    } else if (st.mode == COST_FAST_ANY) {
      glir_body<CostOf<COST_FAST_ANY>, NW>(CostOf<COST_FAST_ANY>{st.fast}, pr, prm, sm, o, s);
    } else {
      glir_body<CostOf<COST_SLOW>, NW>(CostOf<COST_SLOW>{st.slow}, pr, prm, sm, o, s);
    }
  } else if (st.mode == COST_FAST_GEOM) {
    pso_body<CostOf<COST_FAST_GEOM>, NW>(CostOf<COST_FAST_GEOM>{st.fast}, pr, prm, sm, o, s);
  } else if (st.mode == COST_FAST_ANY) {
    pso_body<CostOf<COST_FAST_ANY>, NW>(CostOf<COST_FAST_ANY>{st.fast}, pr, prm, sm, o, s);
  } else {
    pso_body<CostOf<COST_SLOW>, NW>(CostOf<COST_SLOW>{st.slow}, pr, prm, sm, o, s);
  }
  if (threadIdx.x == 0) publish_result(prm.ex, b, gridDim.x, o);
}

constexpr int kCostFixedSmem = 16 + kExpTableSize * (int)sizeof(double);

// cost_function alone: CTA per problem, warps stride over the candidate poses.
template <int NW>
__global__ void __launch_bounds__(NW * 32) cost_kernel(const DevProblem* __restrict__ probs, const DevMap* __restrict__ maps, int n_poses,
                                                       const double* __restrict__ poses, double* __restrict__ out, int smem_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const DevProblem& pr = probs[b];
  const DevMap& mp = maps[pr.map_id];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* etab = reinterpret_cast<double*>(smem_raw + 16);
  const Staged st = stage_problem(pr, mp, smem_raw + kCostFixedSmem, smem_bytes - kCostFixedSmem, etab, bar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = warp; q < n_poses; q += NW) {
    const double* ps = poses + 3 * ((size_t)b * n_poses + q);
    double c;
    if (st.mode == COST_FAST_GEOM)
      c = st.fast.eval<true>(ps[0], ps[1], ps[2], lane);
    else if (st.mode == COST_FAST_ANY)
      c = st.fast.eval<false>(ps[0], ps[1], ps[2], lane);
    else
      c = st.slow.eval(ps[0], ps[1], ps[2], lane);
    if (lane == 0) out[(size_t)b * n_poses + q] = c;
  }
}

// DFMA throughput probe: 8 independent chains per thread.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double bconst) {
  double r0 = threadIdx.x, r1 = r0 + 1, r2 = r0 + 2, r3 = r0 + 3, r4 = r0 + 4, r5 = r0 + 5, r6 = r0 + 6, r7 = r0 + 7;
  for (int i = 0; i < iters; ++i) {
    r0 = fma(r0, a, bconst);
    r1 = fma(r1, a, bconst);
    r2 = fma(r2, a, bconst);
    r3 = fma(r3, a, bconst);
    r4 = fma(r4, a, bconst);
    r5 = fma(r5, a, bconst);
    r6 = fma(r6, a, bconst);
    r7 = fma(r7, a, bconst);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
}

}  // namespace ndtpso
