// Device side of the B200-native PSO/NDT scan matcher (sm_100a).
//
// Kernels (one launch each per batch):
//   compact_map_kernel  K0  dense/sparse (mu, Sigma^-1, built) table -> u16 lookup grid over the
//                           bounding box of the built cells + packed 48-byte records
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

namespace ndtpso {

// ------------------------------------------------------------------------------------------
// Device-resident descriptors (built on the host, one H2D with the rest of the batch)
// ------------------------------------------------------------------------------------------
enum { HDR_NREC = 0, HDR_BX0, HDR_BY0, HDR_BW, HDR_BH, HDR_MODE, HDR_WORDS = 8 };
enum { MAP_COMPACT = 0, MAP_DENSE_DIRECT = 1 };  // hdr[HDR_MODE]

struct DevMap {
  double x_min, x_max, y_min, y_max;
  double hw, hh;        // width/2., height/2. (ndtframe.cpp:245-246)
  double cs, inv_cs;    // cell_side and, when cs is a power of two, its exact reciprocal
  int gw, gh, ncells, cs_pow2;
  // input table (device pointers)
  const double* mean;      // [rows][2]
  const double* icov;      // [rows][4]
  const uint8_t* built;    // [ncells] (dense form) or nullptr
  const int* cell_index;   // [n_sparse] (sparse form) or nullptr
  int n_sparse;            // < 0: dense
  int _pad;
  // compact table written by K0
  unsigned short* grid;    // capacity ncells (+ padding to 16 B)
  double* rec;             // capacity rows*6: {mx, my, S00, S01, S10, S11}
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

struct PsoParams {
  int P, I;
  double w, c1, c2, wd;
  int n_draws;      // 3 + 3P + 6PI
  int smem_bytes;   // dynamic shared memory given to pso_kernel
};

// ------------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + 1-D bulk TMA (cp.async.bulk -> SASS UBLKCP)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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

__device__ __forceinline__ int round16(int x) { return (x + 15) & ~15; }

// ------------------------------------------------------------------------------------------
// K0: compact the NDT table.  One CTA per map.
//   pass A: count built cells per thread chunk + bounding box -> block scan
//   pass B: fill the bbox grid with 0xFFFF, then write record slots and packed records
// Record order = ascending cell index (deterministic).
// ------------------------------------------------------------------------------------------
constexpr int K0_THREADS = 256;

__global__ void __launch_bounds__(K0_THREADS) compact_map_kernel(const DevMap* __restrict__ maps) {
  const DevMap& m = maps[blockIdx.x];
  const int tid = threadIdx.x;
  const bool sparse = m.n_sparse >= 0;
  const int rows = sparse ? m.n_sparse : m.ncells;
  const int chunk = (rows + K0_THREADS - 1) / K0_THREADS;
  const int lo = min(rows, tid * chunk), hi = min(rows, lo + chunk);

  __shared__ int s_scan[K0_THREADS];
  __shared__ int s_box[4];
  if (tid == 0) {
    s_box[0] = INT_MAX;  // min ix
    s_box[1] = INT_MAX;  // min iy
    s_box[2] = -1;       // max ix
    s_box[3] = -1;       // max iy
  }
  __syncthreads();

  int cnt = 0, ax = INT_MAX, ay = INT_MAX, bx = -1, by = -1;
  for (int r = lo; r < hi; ++r) {
    int cell;
    if (sparse) {
      cell = m.cell_index[r];
    } else {
      if (!m.built[r]) continue;
      cell = r;
    }
    const int ix = cell % m.gw, iy = cell / m.gw;
    ax = min(ax, ix);
    ay = min(ay, iy);
    bx = max(bx, ix);
    by = max(by, iy);
    ++cnt;
  }
  if (cnt) {
    atomicMin(&s_box[0], ax);
    atomicMin(&s_box[1], ay);
    atomicMax(&s_box[2], bx);
    atomicMax(&s_box[3], by);
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

  const int bx0 = n_rec ? s_box[0] : 0, by0 = n_rec ? s_box[1] : 0;
  const int bw = n_rec ? s_box[2] - s_box[0] + 1 : 0, bh = n_rec ? s_box[3] - s_box[1] + 1 : 0;
  const bool compact_ok = n_rec <= 65534;
  if (tid == 0) {
    m.hdr[HDR_NREC] = n_rec;
    m.hdr[HDR_BX0] = bx0;
    m.hdr[HDR_BY0] = by0;
    m.hdr[HDR_BW] = bw;
    m.hdr[HDR_BH] = bh;
    m.hdr[HDR_MODE] = compact_ok ? MAP_COMPACT : MAP_DENSE_DIRECT;
  }
  if (!compact_ok) return;  // uniform: the hot kernel reads the dense arrays directly

  for (int g = tid; g < bw * bh; g += K0_THREADS) m.grid[g] = 0xFFFFu;
  __syncthreads();

  for (int r = lo; r < hi; ++r) {
    int cell;
    if (sparse) {
      cell = m.cell_index[r];
    } else {
      if (!m.built[r]) continue;
      cell = r;
    }
    const int ix = cell % m.gw, iy = cell / m.gw;
    m.grid[(iy - by0) * bw + (ix - bx0)] = static_cast<unsigned short>(slot);
    double* rec = m.rec + 6 * (size_t)slot;
    const double2 mu = *reinterpret_cast<const double2*>(m.mean + 2 * (size_t)r);
    const double2 s0 = *reinterpret_cast<const double2*>(m.icov + 4 * (size_t)r);
    const double2 s1 = *reinterpret_cast<const double2*>(m.icov + 4 * (size_t)r + 2);
    *reinterpret_cast<double2*>(rec) = mu;
    *reinterpret_cast<double2*>(rec + 2) = s0;
    *reinterpret_cast<double2*>(rec + 4) = s1;
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

__global__ void __launch_bounds__(K1_WARPS * 32) rng_fill_kernel(const DevProblem* __restrict__ probs, int n_problems, int n_draws) {
  __shared__ uint32_t ring_all[K1_WARPS][K1_RING];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * K1_WARPS + warp;
  if (b >= n_problems) return;
  if (probs[b].rnd_from_host) return;
  uint32_t* z = ring_all[warp];
  int* out = const_cast<int*>(probs[b].rnd);

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
  // outputs that are already in the history: rand k = z[341 + k] >> 1, k = 0..31
  if (lane < n_draws) out[lane] = static_cast<int>(z[341 + lane] >> 1);
  // then 32 per step: n = 373 + 32 s + lane  <->  k = n - 341
  for (int n0 = 373; n0 - 341 < n_draws; n0 += 32) {
    const int n = n0 + lane;
    uint32_t v = z[(n - 33) & (K1_RING - 1)] + z[(n - 341) & (K1_RING - 1)];
    v += 11u * (z[(n - 61) & (K1_RING - 1)] + z[(n - 313) & (K1_RING - 1)]);
    v += 55u * (z[(n - 89) & (K1_RING - 1)] + z[(n - 285) & (K1_RING - 1)]);
    v += 165u * (z[(n - 117) & (K1_RING - 1)] + z[(n - 257) & (K1_RING - 1)]);
    v += 330u * (z[(n - 145) & (K1_RING - 1)] + z[(n - 229) & (K1_RING - 1)]);
    v += 462u * (z[(n - 173) & (K1_RING - 1)] + z[(n - 201) & (K1_RING - 1)]);
    __syncwarp();  // all reads of the slots about to be overwritten are done
    z[n & (K1_RING - 1)] = v;
    const int k = n - 341;
    if (k < n_draws) out[k] = static_cast<int>(v >> 1);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------
// NDT cost of one candidate pose, evaluated by one warp (lanes stride over the points).
// ------------------------------------------------------------------------------------------
enum { TABLE_SMEM = 0, TABLE_GLOBAL = 1, TABLE_DENSE = 2 };

struct MapCtx {
  double x_min, x_max, y_min, y_max, hw, hh, cs, inv_cs;
  int gw, ncells;
  int bx0, by0, bw, bh;
  const unsigned short* grid;  // TABLE_SMEM / TABLE_GLOBAL
  const double* rec;           // TABLE_SMEM / TABLE_GLOBAL
  const double* mean;          // TABLE_DENSE
  const double* icov;
  const uint8_t* built;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;  // identical in every lane; fixed butterfly order => deterministic
}

template <int TABLE, bool POW2>
__device__ __forceinline__ double warp_cost(const MapCtx& m, const double2* __restrict__ pts, int n, double tx, double ty, double th,
                                            int lane) {
  double s, c;
  sincos(th, &s, &c);
  double acc = 0.;
#pragma unroll 2
  for (int i = lane; i < n; i += 32) {
    const double2 p = pts[i];
    const double x = p.x * c - p.y * s + tx;  // transform_point, core.h:29-30
    const double y = p.x * s + p.y * c + ty;
    if ((x > m.x_min) && (x < m.x_max) && (y > m.y_min) && (y < m.y_max)) {  // strict, ndtframe.cpp:242
      const double fx = floor(POW2 ? (x + m.hw) * m.inv_cs : (x + m.hw) / m.cs);
      const double fy = floor(POW2 ? (y + m.hh) * m.inv_cs : (y + m.hh) / m.cs);
      int ix = static_cast<int>(fx), iy = static_cast<int>(fy);
      if (ix >= m.gw) {  // (x + W/2)/cs rounded up to gw: the reference's flat index wraps into the next row
        ix -= m.gw;
        iy += 1;
      }
      const double* rec = nullptr;
      double mx, my, s00, s01, s10, s11;
      bool hit = false;
      if (TABLE == TABLE_DENSE) {
        const int idx = ix + m.gw * iy;
        if (idx < m.ncells && m.built[idx]) {
          hit = true;
          mx = m.mean[2 * idx];
          my = m.mean[2 * idx + 1];
          s00 = m.icov[4 * idx];
          s01 = m.icov[4 * idx + 1];
          s10 = m.icov[4 * idx + 2];
          s11 = m.icov[4 * idx + 3];
        }
      } else {
        const unsigned gx = static_cast<unsigned>(ix - m.bx0), gy = static_cast<unsigned>(iy - m.by0);
        if (gx < static_cast<unsigned>(m.bw) && gy < static_cast<unsigned>(m.bh)) {
          const unsigned r = m.grid[gy * m.bw + gx];
          if (r != 0xFFFFu) {
            hit = true;
            rec = m.rec + 6 * r;
            const double2 a = *reinterpret_cast<const double2*>(rec);
            const double2 b = *reinterpret_cast<const double2*>(rec + 2);
            const double2 d = *reinterpret_cast<const double2*>(rec + 4);
            mx = a.x;
            my = a.y;
            s00 = b.x;
            s01 = b.y;
            s10 = d.x;
            s11 = d.y;
          }
        }
      }
      if (hit) {  // normalDistribution, ndtcell.cpp:72-75
        const double d0 = x - mx, d1 = y - my;
        const double r0 = d0 * s00 + d1 * s10;
        const double r1 = d0 * s01 + d1 * s11;
        acc -= exp(-(r0 * d0 + r1 * d1) / 2.);
      }
    }
  }
  return warp_sum(acc);
}

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
  Cand* cand;     // [2][P+1]
  double* x;      // [P][3]
  double* v;      // [P][3]
  double* vnew;   // [P][3]
  double* pb;     // [P][3]
  double* pbc;    // [P]
  unsigned char* dyn;  // start of the staged problem data
  int fixed_bytes;
};

__host__ __device__ inline int pso_fixed_smem_bytes(int P) {
  int b = 16;                                  // mbarrier
  b += 2 * (P + 1) * (int)sizeof(Cand);        // candidates, double buffered
  b += (P > 0 ? P : 1) * 13 * (int)sizeof(double);  // x, v, vnew, pb (3 each) + pbc
  return (b + 15) & ~15;
}

__device__ __forceinline__ PsoSmem carve_smem(unsigned char* base, int P) {
  PsoSmem s;
  const int Pn = P > 0 ? P : 1;
  s.bar = reinterpret_cast<uint64_t*>(base);
  s.cand = reinterpret_cast<Cand*>(base + 16);
  double* d = reinterpret_cast<double*>(base + 16 + 2 * (P + 1) * sizeof(Cand));
  s.x = d;
  s.v = d + 3 * Pn;
  s.vnew = d + 6 * Pn;
  s.pb = d + 9 * Pn;
  s.pbc = d + 12 * Pn;
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
template <int TABLE, bool POW2, int NW>
__device__ __forceinline__ void pso_body(const MapCtx& m, const double2* __restrict__ pts, int n_pts, const DevProblem& pr,
                                         const PsoParams& prm, const PsoSmem& sm, double* __restrict__ out, int* __restrict__ stats) {
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
    const double c = warp_cost<TABLE, POW2>(m, pts, n_pts, pos[0], pos[1], pos[2], lane);
    if (lane == 0) {
      cand0[t] = Cand{c, pos[0], pos[1], pos[2]};
    }
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
    int j0 = start + ((warp - start) % NW + NW) % NW;
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
      const double c = warp_cost<TABLE, POW2>(m, pts, n_pts, nx[0], nx[1], nx[2], lane);
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
    }
  }
}

// Stage one problem's points and compact table into shared memory with bulk TMA and pick the
// table mode.  Returns the MapCtx and the points pointer the cost loop should use.
struct Staged {
  MapCtx m;
  const double2* pts;  // global points (used when they do not fit in shared memory)
  int table;           // TABLE_*
  int pts_bytes, rec_bytes;
  bool pts_fit;
};

__device__ __forceinline__ Staged stage_problem(const DevProblem& pr, const DevMap& mp, unsigned char* dyn, int dyn_bytes, uint64_t* bar) {
  Staged st;
  MapCtx& m = st.m;
  m.x_min = mp.x_min;
  m.x_max = mp.x_max;
  m.y_min = mp.y_min;
  m.y_max = mp.y_max;
  m.hw = mp.hw;
  m.hh = mp.hh;
  m.cs = mp.cs;
  m.inv_cs = mp.inv_cs;
  m.gw = mp.gw;
  m.ncells = mp.ncells;
  const int n_rec = mp.hdr[HDR_NREC];
  m.bx0 = mp.hdr[HDR_BX0];
  m.by0 = mp.hdr[HDR_BY0];
  m.bw = mp.hdr[HDR_BW];
  m.bh = mp.hdr[HDR_BH];
  const int mode = mp.hdr[HDR_MODE];
  m.grid = mp.grid;
  m.rec = mp.rec;
  m.mean = mp.mean;
  m.icov = mp.icov;
  m.built = mp.built;

  const int pts_bytes = pr.n_pts * 16;
  const int rec_bytes = n_rec * 48;
  const int grid_bytes = round16(m.bw * m.bh * 2);
  const bool pts_fit = pts_bytes <= dyn_bytes;
  const bool table_fit = (mode == MAP_COMPACT) && (pts_bytes + rec_bytes + grid_bytes <= dyn_bytes);
  st.table = (mode == MAP_COMPACT) ? (table_fit ? TABLE_SMEM : TABLE_GLOBAL) : TABLE_DENSE;
  st.pts = pr.pts;
  st.pts_bytes = pts_bytes;
  st.rec_bytes = rec_bytes;
  st.pts_fit = pts_fit;

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
      if (table_fit && rec_bytes) tma_load_1d(dyn + pts_bytes, mp.rec, rec_bytes, bar);
      if (table_fit && grid_bytes) tma_load_1d(dyn + pts_bytes + rec_bytes, mp.grid, grid_bytes, bar);
    }
    mbar_wait(bar, 0);
  }
  return st;
}

// Pointers into the staged copy, derived from the shared-memory base so that the compiler emits
// LDS (not generic loads) in the TABLE_SMEM instantiation.
__device__ __forceinline__ MapCtx smem_table(const Staged& st, const unsigned char* dyn) {
  MapCtx m = st.m;
  m.rec = reinterpret_cast<const double*>(dyn + st.pts_bytes);
  m.grid = reinterpret_cast<const unsigned short*>(dyn + st.pts_bytes + st.rec_bytes);
  return m;
}

template <int NW>
__global__ void __launch_bounds__(NW * 32) pso_kernel(const DevProblem* __restrict__ probs, const DevMap* __restrict__ maps, PsoParams prm,
                                                      double* __restrict__ out, int* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const DevProblem& pr = probs[b];
  const DevMap& mp = maps[pr.map_id];
  PsoSmem sm = carve_smem(smem_raw, prm.P);
  const Staged st = stage_problem(pr, mp, sm.dyn, prm.smem_bytes - sm.fixed_bytes, sm.bar);
  double* o = out + 4 * (size_t)b;
  int* s = stats ? stats + 2 * (size_t)b : nullptr;
  const bool pow2 = mp.cs_pow2 != 0;
  if (st.table == TABLE_SMEM) {
    const MapCtx m = smem_table(st, sm.dyn);
    const double2* spts = reinterpret_cast<const double2*>(sm.dyn);
    if (pow2)
      pso_body<TABLE_SMEM, true, NW>(m, spts, pr.n_pts, pr, prm, sm, o, s);
    else
      pso_body<TABLE_SMEM, false, NW>(m, spts, pr.n_pts, pr, prm, sm, o, s);
  } else {
    // generic pointer: shared when the points fit, else global
    const double2* gpts = st.pts_fit ? reinterpret_cast<const double2*>(sm.dyn) : st.pts;
    if (st.table == TABLE_GLOBAL) {
      if (pow2)
        pso_body<TABLE_GLOBAL, true, NW>(st.m, gpts, pr.n_pts, pr, prm, sm, o, s);
      else
        pso_body<TABLE_GLOBAL, false, NW>(st.m, gpts, pr.n_pts, pr, prm, sm, o, s);
    } else {
      if (pow2)
        pso_body<TABLE_DENSE, true, NW>(st.m, gpts, pr.n_pts, pr, prm, sm, o, s);
      else
        pso_body<TABLE_DENSE, false, NW>(st.m, gpts, pr.n_pts, pr, prm, sm, o, s);
    }
  }
}

// cost_function alone: CTA per problem, warps stride over the candidate poses.
template <int NW>
__global__ void __launch_bounds__(NW * 32) cost_kernel(const DevProblem* __restrict__ probs, const DevMap* __restrict__ maps, int n_poses,
                                                       const double* __restrict__ poses, double* __restrict__ out, int smem_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const DevProblem& pr = probs[b];
  const DevMap& mp = maps[pr.map_id];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  unsigned char* dyn = smem_raw + 16;
  const Staged st = stage_problem(pr, mp, dyn, smem_bytes - 16, bar);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool pow2 = mp.cs_pow2 != 0;
  const MapCtx ms = smem_table(st, dyn);
  const double2* spts = reinterpret_cast<const double2*>(dyn);
  const double2* gpts = st.pts_fit ? spts : st.pts;
  for (int q = warp; q < n_poses; q += NW) {
    const double* ps = poses + 3 * ((size_t)b * n_poses + q);
    double c;
    if (st.table == TABLE_SMEM)
      c = pow2 ? warp_cost<TABLE_SMEM, true>(ms, spts, pr.n_pts, ps[0], ps[1], ps[2], lane)
               : warp_cost<TABLE_SMEM, false>(ms, spts, pr.n_pts, ps[0], ps[1], ps[2], lane);
    else if (st.table == TABLE_GLOBAL)
      c = pow2 ? warp_cost<TABLE_GLOBAL, true>(st.m, gpts, pr.n_pts, ps[0], ps[1], ps[2], lane)
               : warp_cost<TABLE_GLOBAL, false>(st.m, gpts, pr.n_pts, ps[0], ps[1], ps[2], lane);
    else
      c = pow2 ? warp_cost<TABLE_DENSE, true>(st.m, gpts, pr.n_pts, ps[0], ps[1], ps[2], lane)
               : warp_cost<TABLE_DENSE, false>(st.m, gpts, pr.n_pts, ps[0], ps[1], ps[2], lane);
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
