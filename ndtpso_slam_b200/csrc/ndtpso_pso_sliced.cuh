// K2, point-sliced form: the production PSO kernel.
//
// One CTA per scan-match problem, T = 32*NW threads.  Thread t keeps scan points
// t, t+T, t+2T, ... (NPT of them) in REGISTERS for the whole run; the compact NDT table
// ({mu, -Sigma^-1/2} records + u16 row-strip grid) is staged once into shared memory by bulk TMA.
// Every round has three phases:
//   A  thread j owns particle j: velocity/position update from its private state and the current
//      gbest (core.cpp:83-90, no FMA contraction), sincos of the candidate heading -> pose[j]
//   -- barrier --
//   B  every warp evaluates EVERY pending candidate on its own slice of the scan, JB candidates at
//      a time (branch-free NDT score, fast_exp, JB*NPT independent evaluations in flight per lane),
//      packed warp-shuffle reduction -> partial[j][warp].  With the fp32 screen (one CTA per problem,
//      see "fp32 screen" below) B is B1, a rigorous fp32 lower bound of every pending candidate's
//      cost, then B2, the fp64 evaluation of only those candidates the bound does not rule out.
//   -- barrier --
//   C  every warp sums the partials of the pending candidates (fixed order), finds j* = the first
//      candidate that beats gbest (ballot), owners commit particles <= j* (core.cpp:89-105);
//      particles after j* are replayed in the next round against the new gbest.
// pose[] and partial[] are double buffered, particle state is private to its owner thread, so two
// barriers per round suffice.  Work per warp in phase B is identical for every warp whatever the
// number of pending candidates: no load imbalance, and replay rounds cost only what they recompute.
#pragma once
#include "ndtpso_kernels.cuh"

namespace ndtpso {

constexpr int kSlicedMaxNPT = 6;

// Code-generation variants of the point evaluation (bit mask), selectable for measurement:
enum {
  VAR_FLOOR_ON_FP64 = 1,  // floor() by a round-down magic add on the fp64 pipe instead of F2I.F64.FLOOR (conversion pipe)
  VAR_KF_ON_FP64 = 2,     // k as a double by subtracting the magic constant (fp64 pipe) instead of I2F.F64
  VAR_PIN_CONSTS = 4,     // keep 16/ln2 and 1/7! in vector registers (loaded through a lane-dependent address) instead of
                          // re-materialising them from uniform registers with two moves per use: -4 instructions per evaluation
};
#ifndef NDTPSO_SCREEN_JB8
#define NDTPSO_SCREEN_JB8 1  // the screen takes candidates eight at a time first (more loads in flight, fewer reductions): +2 %
#endif
#ifndef NDTPSO_PROD_VARIANT
#define NDTPSO_PROD_VARIANT 4
#endif
constexpr int kProdVariant = NDTPSO_PROD_VARIANT;

// Phase timing (tools/score_bench.cu builds with -DNDTPSO_PHASE_TIMING): thread 0 of every CTA adds the
// cycles it spends in {prologue+init, phase A, phase B, phase C} (barrier waits included) to g_phase_cycles.
#ifdef NDTPSO_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[8];
#define NDTPSO_PHASE_DECL long long ph_t0 = clock64();
#define NDTPSO_PHASE_MARK(i)                                                         \
  if (threadIdx.x == 0) {                                                            \
    const long long ph_t1 = clock64();                                               \
    atomicAdd(&g_phase_cycles[i], static_cast<unsigned long long>(ph_t1 - ph_t0));   \
    ph_t0 = ph_t1;                                                                   \
  }
#else
#define NDTPSO_PHASE_DECL
#define NDTPSO_PHASE_MARK(i)
#endif

struct __align__(16) Pose {
  double x, y, c, s, th, pad;  // {x, y} and {cos, sin} are the two 16-byte loads of phase B
};

// Shared memory: [mbarrier 16][exp table 128][exp constants 64][records][grid] at FIXED offsets
// (so the hot loop's table addresses are a constant and one register), then the swarm arrays.
#if (NDTPSO_PROD_VARIANT & 4)
constexpr int kSlicedTableOffset = 16 + (kExpTableSize + 8 + 64) * (int)sizeof(double);  // + per-lane copies of two constants
#else
constexpr int kSlicedTableOffset = 16 + (kExpTableSize + 8) * (int)sizeof(double);
#endif

struct SlicedSmem {
  uint64_t* bar;
  double* etab;     // [16]
  double* cst;      // [8] fast_exp constants
  unsigned char* table;  // records, then grid
  Pose* pose;       // [2][P+1]
  double* partial;  // [2][(P+1)*PW]
  double* wpart;    // [(P+1)*NW] warp partials of the cluster form (unused when CL == 1)
  double* ubuf;     // [2][6P] |Random()| coefficients of the current / next iteration (core.cpp:84)
  double* cost0;    // [P+1] initial costs
  double* x;        // [P][3]  owner-private particle state
  double* v;        // [P][3]
  double* vnew;     // [P][3]
  double* pb;       // [P][3]
  double* pbc;      // [P]
  // fp32 screening (CL == 1 only; null when off)
  float4* pose32;   // [P+1][2] {x, y, cos, sin}, {-sin, cos, 0, 0} of the current round's candidates (the register pairs the packed transform takes)
  float* lbpart;    // [(P+1)*NW] per-warp partial sums of the upper bounds
  int* surv;        // [P+2] candidates that need the fp64 evaluation; surv[P+1] = their number
  float* rec32;     // [(n_rec+1)][8] {h00, h01, h11, hs, mx, my, -, -}
};

// PW = partials per candidate summed in phase C; WP = warp partials per candidate of the cluster form (0 when CL == 1)
__host__ __device__ inline int sliced_swarm_smem_bytes(int P, int PW, int WP) {
  const int Pn = P > 0 ? P : 1;
  int b = 2 * (P + 1) * (int)sizeof(Pose);
  b += 2 * (P + 1) * PW * (int)sizeof(double);
  b += (P + 1) * WP * (int)sizeof(double);
  b += 2 * 6 * Pn * (int)sizeof(double);
  b += (P + 1) * (int)sizeof(double);
  b += Pn * 13 * (int)sizeof(double);
  return (b + 15) & ~15;
}
// shared memory of the fp32 screen without its record table: pose32, lbpart, surv
__host__ __device__ inline int sliced_screen_fixed_bytes(int P, int NW) {
  return (P + 1) * 32 + round16((P + 1) * NW * 4) + round16((P + 2) * 4);
}
// total dynamic shared memory: table_bytes = (n_rec + 1) * 48 + round16((span + 1) * 2).  With the screen (screen_recs > 0:
// the largest record count of the batch, null record included) the fp32 records take screen_recs * 32 bytes more.
__host__ __device__ inline int sliced_smem_bytes(int P, int PW, int WP, int table_bytes, int screen_recs = 0) {
  return kSlicedTableOffset + table_bytes + sliced_swarm_smem_bytes(P, PW, WP) +
         (screen_recs > 0 ? sliced_screen_fixed_bytes(P, PW) + round16(screen_recs * 32) : 0);
}

__device__ __forceinline__ SlicedSmem carve_sliced(unsigned char* base, int P, int PW, int WP, int table_bytes, int screen = 0) {
  SlicedSmem s;
  const int Pn = P > 0 ? P : 1;
  s.bar = reinterpret_cast<uint64_t*>(base);
  s.etab = reinterpret_cast<double*>(base + 16);
  s.cst = s.etab + kExpTableSize;
  s.table = base + kSlicedTableOffset;
  unsigned char* p = s.table + table_bytes;
  s.pose = reinterpret_cast<Pose*>(p);
  p += 2 * (P + 1) * sizeof(Pose);
  s.partial = reinterpret_cast<double*>(p);
  p += 2 * (size_t)(P + 1) * PW * sizeof(double);
  s.wpart = reinterpret_cast<double*>(p);
  p += (size_t)(P + 1) * WP * sizeof(double);
  s.ubuf = reinterpret_cast<double*>(p);
  p += 2 * 6 * (size_t)Pn * sizeof(double);
  s.cost0 = reinterpret_cast<double*>(p);
  p += (P + 1) * sizeof(double);
  double* d = reinterpret_cast<double*>(p);
  s.x = d;
  s.v = d + 3 * Pn;
  s.vnew = d + 6 * Pn;
  s.pb = d + 9 * Pn;
  s.pbc = d + 12 * Pn;
  s.pose32 = nullptr;
  s.lbpart = nullptr;
  s.surv = nullptr;
  s.rec32 = nullptr;
  if (screen) {
    p = base + kSlicedTableOffset + table_bytes + sliced_swarm_smem_bytes(P, PW, WP);
    s.pose32 = reinterpret_cast<float4*>(p);
    p += (P + 1) * 32;
    s.lbpart = reinterpret_cast<float*>(p);
    p += round16((P + 1) * PW * 4);
    s.surv = reinterpret_cast<int*>(p);
    p += round16((P + 2) * 4);
    s.rec32 = reinterpret_cast<float*>(p);
  }
  return s;
}

// Loop-invariant operands of the point evaluation, held in registers.
struct SliceCtx {
  const unsigned short* grid;  // shared
  const double* rec;           // shared
  const double* etab;          // shared
  double x_min, x_max, y_min, y_max, hw, hh, cs, inv_cs, hw_s, hh_s;
  double l2e, ln2hi, ln2lo, c7, c6, c5, c4, c3;  // fast_exp constants kept out of the immediate field
  int gw, base, span, null_id;
};

// One scan point against one candidate pose: subtracts exp(-(d' S d)/2) from acc iff the point is
// inside the frame (strict), its cell is built, and the value is a normal double (>= 2.2e-308;
// smaller ones are flushed to zero, see fast_exp.h).  Padding points of the last slice are stored
// as (1e200, 0): whatever the pose, |x'| or |y'| is then ~1e200, i.e. out of bounds, so they need
// no validity flag.  The host only selects this kernel for tables whose every Sigma^-1 is finite,
// symmetric and positive semi-definite (what NDTCell::build produces), so the exponent is <= 0 up
// to rounding and can never overflow; anything else takes the generic kernel (library exp).
template <bool FAST_GEOM, int VAR>
__device__ __forceinline__ void slice_point(const SliceCtx& m, const double2 p, double tx, double ty, double c, double s, double& acc) {
  const double x = fma(p.x, c, fma(-p.y, s, tx));  // transform_point, core.h:29-30
  const double y = fma(p.x, s, fma(p.y, c, ty));
  bool inb;
  double u, v;
  if (FAST_GEOM) {
    inb = (fabs(x) < m.x_max) && (fabs(y) < m.y_max);  // strict, ndtframe.cpp:242
    u = fma(x, m.inv_cs, m.hw_s);                      // == (x + W/2)/cs exactly (cs = 2^k)
    v = fma(y, m.inv_cs, m.hh_s);
  } else {
    inb = (x > m.x_min) && (x < m.x_max) && (y > m.y_min) && (y < m.y_max);
    u = __ddiv_rn(x + m.hw, m.cs);
    v = __ddiv_rn(y + m.hh, m.cs);
  }
  int ix, iy;  // floor, ndtframe.cpp:245-246
  if (VAR & VAR_FLOOR_ON_FP64) {
    ix = __double2loint(__dadd_rd(u, kExpMagic));  // valid for |u| < 2^31; out-of-range u only occurs out of bounds
    iy = __double2loint(__dadd_rd(v, kExpMagic));
  } else {
    ix = __double2int_rd(u);
    iy = __double2int_rd(v);
  }
  const unsigned g = static_cast<unsigned>(ix + m.gw * iy - m.base);
  const bool in_strip = inb && (g < static_cast<unsigned>(m.span));
  const unsigned r = m.grid[in_strip ? g : static_cast<unsigned>(m.span)];
  const double* q = m.rec + 6 * r;
  // the sliced kernel only runs on symmetric tables (S01 == S10 bit for bit, which is what
  // NDTCell::s_calc_covar_inverse produces, ndtcell.cpp:109-110): 40 bytes per record instead of 48
  const double2 mu = *reinterpret_cast<const double2*>(q);
  const double2 h0 = *reinterpret_cast<const double2*>(q + 2);  // {-S00/2, -S01/2}
  const double h11 = q[5];
  const double d0 = x - mu.x, d1 = y - mu.y;  // normalDistribution, ndtcell.cpp:72-75
  const double r0 = fma(d1, h0.y, d0 * h0.x);  // d0*S00 + d1*S10
  const double r1 = fma(d1, h11, d0 * h0.y);   // d0*S01 + d1*S11
  const double a = fma(r1, d1, r0 * d0);       // = -(d' S d)/2
  // fast_exp (fast_exp.h), with its constants in registers
  const double l2e = m.l2e, c7 = m.c7;
  const double kd = fma(a, l2e, kExpMagic);
  const int k = __double2loint(kd);
  const double kf = (VAR & VAR_KF_ON_FP64) ? (kd - kExpMagic) : static_cast<double>(k);
  double rr = fma(kf, m.ln2hi, a);
  rr = fma(kf, m.ln2lo, rr);
  double pl = fma(rr, c7, m.c6);
  pl = fma(pl, rr, m.c5);
  pl = fma(pl, rr, m.c4);
  pl = fma(pl, rr, m.c3);
  pl = fma(pl, rr, 0.5);
  pl = fma(pl, rr, 1.0);
  const double em1 = pl * rr;
  const double t = m.etab[k & (kExpTableSize - 1)];
  const double val = fma(t, em1, t);
  const int ahi = __double2hiint(a);
  // the result is a normal double for -708 <= a <= 709 (hi words of a: 0xC0862000 / 0x40862800)
  const bool use = in_strip && (r != static_cast<unsigned>(m.null_id)) && (static_cast<unsigned>(ahi) <= 0xC0862000u);
  const int ehi = use ? __double2hiint(val) + ((k >> kExpTableShift) << 20) : 0;
  const int elo = use ? __double2loint(val) : 0;
  acc -= __hiloint2double(ehi, elo);
}

// Packed warp reduction of JB per-lane accumulators (one per candidate): after it, the total of
// candidate packed_slot<JB>(lane) sits in every lane of its group.  JB candidates share the five
// shuffle levels, so the cost is 5 + (JB - 1) exchanges instead of 5*JB, and the JB chains hide
// each other's latency.  The tree is fixed => deterministic.
template <int JB>
__device__ __forceinline__ double packed_warp_sum(const double (&a)[JB], int lane);
template <>
__device__ __forceinline__ double packed_warp_sum<1>(const double (&a)[1], int) {
  return warp_sum(a[0]);
}
template <>
__device__ __forceinline__ double packed_warp_sum<2>(const double (&a)[2], int lane) {
  const bool hi16 = (lane & 16) != 0;
  double k = hi16 ? a[1] : a[0];
  k += __shfl_xor_sync(0xffffffffu, hi16 ? a[0] : a[1], 16);  // lanes < 16: candidate 0, >= 16: candidate 1
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) k += __shfl_xor_sync(0xffffffffu, k, off);
  return k;
}
template <>
__device__ __forceinline__ double packed_warp_sum<4>(const double (&a)[4], int lane) {
  const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
  double k01 = hi16 ? a[1] : a[0];
  k01 += __shfl_xor_sync(0xffffffffu, hi16 ? a[0] : a[1], 16);
  double k23 = hi16 ? a[3] : a[2];
  k23 += __shfl_xor_sync(0xffffffffu, hi16 ? a[2] : a[3], 16);
  double k = hi8 ? k23 : k01;
  k += __shfl_xor_sync(0xffffffffu, hi8 ? k01 : k23, 8);  // bit 3 clear: candidates {0,1}; set: {2,3}
#pragma unroll
  for (int off = 4; off > 0; off >>= 1) k += __shfl_xor_sync(0xffffffffu, k, off);
  return k;
}
template <>
__device__ __forceinline__ double packed_warp_sum<8>(const double (&a)[8], int lane) {
  const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0, hi4 = (lane & 4) != 0;
  double k01 = hi16 ? a[1] : a[0];
  k01 += __shfl_xor_sync(0xffffffffu, hi16 ? a[0] : a[1], 16);
  double k23 = hi16 ? a[3] : a[2];
  k23 += __shfl_xor_sync(0xffffffffu, hi16 ? a[2] : a[3], 16);
  double k45 = hi16 ? a[5] : a[4];
  k45 += __shfl_xor_sync(0xffffffffu, hi16 ? a[4] : a[5], 16);
  double k67 = hi16 ? a[7] : a[6];
  k67 += __shfl_xor_sync(0xffffffffu, hi16 ? a[6] : a[7], 16);
  double ka = hi8 ? k23 : k01;
  ka += __shfl_xor_sync(0xffffffffu, hi8 ? k01 : k23, 8);
  double kb = hi8 ? k67 : k45;
  kb += __shfl_xor_sync(0xffffffffu, hi8 ? k45 : k67, 8);
  double k = hi4 ? kb : ka;
  k += __shfl_xor_sync(0xffffffffu, hi4 ? ka : kb, 4);  // bit 2 clear: candidates 0..3; set: 4..7
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;
}
template <int JB>
__device__ __forceinline__ int packed_slot(int lane) {
  if (JB == 1) return 0;
  if (JB == 2) return (lane >> 4) & 1;
  if (JB == 4) return ((lane >> 4) & 1) + ((lane >> 3) & 1) * 2;
  return ((lane >> 4) & 1) + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1) * 4;
}
template <int JB>
__device__ __forceinline__ bool packed_writer(int lane) {
  return (lane & (32 / JB - 1)) == 0;
}

// ---- fp32 screen ------------------------------------------------------------------------------------
// 9 of 10 candidate poses a swarm evaluates cannot improve their particle's best (median cost ratio 0.23), and the PSO only
// ever asks "is this cost below pbest?" of them (core.cpp:94): their exact value is never used.  So every pending candidate
// is first bounded in fp32 — a RIGOROUS lower bound L <= cost — and only those with L < pbest get the fp64 evaluation.
// Results are bit-identical with the screen on or off: a candidate is dropped only when its fp64 cost provably fails the
// comparison, and survivors are evaluated exactly as before.
//
// The bound, per scan point.  Let H = Sigma^-1/2 (positive semi-definite; the host only selects this kernel for such
// tables), A(d) = d'Hd, so the point's term is -exp(-A(d*)) with d* its exact offset from the cell mean.
//   * Position.  The fp32 transform is within delta_d of the fp64 one in each coordinate (delta_d from the magnitudes of
//     the scan, the frame and fp32 rounding; PsoParams::scr_dd2 = delta_d^2).
//   * Cell.  If the fp32 point lies within beta cell sides of a cell edge (which includes the frame's border: frames are
//     whole cells), fp32 and fp64 may disagree on the cell: the point counts as the worst case, exp(.) = 1.  Otherwise both
//     pick the same cell c, and |d_k| <= dmax_k(c), the distance from the mean to the far side of the cell.
//     A point outside the frame has the term 0 in fp64, so whatever e >= 0 the screen computes for it bounds it: there is no
//     bounds test (screen_point), and the error terms below, derived for a point inside its cell, need not hold for it.
//   * Exponent.  With d = d* + eps, |eps_k| <= delta_d, and Cauchy-Schwarz + AM-GM on the cross term,
//         A(d*) >= (1 - t) A(d) - A(eps)/t,          A(eps) <= hs delta_d^2,   hs = H00 + 2|H01| + H11.
//     A(d) = z0^2 + z1^2 with z = L'd (Cholesky factor L of H).  In fp32 (factor rounded, two FMAs) each z_k is off by at
//     most ez_k = 3 * 2^-24 * (sum of |L_kj| dmax_j), and (|z| - ez)^2 >= (1 - t) z^2 - ez^2/t, and the sum of the two
//     squares carries 2^-22 relative rounding.  Together, with t chosen per record (sqrt of the absolute terms, clamped to
//     [2^-10, 2^-3]: it balances the relative loosening t A against the absolute one for A of order one):
//         -A(d*) <= -(1 - t)^2 (1 - 2^-22) Atilde + kappa,     kappa = (ez0^2 + ez1^2)/t + hs delta_d^2 / t
//     The record stores L scaled by sqrt((1 - t)^2 (1 - 2^-22) log2(e)) and kappa log2(e): the screen evaluates
//         e = ex2(min(kappa2 - z0~^2 - z1~^2, 0)) >= exp(-A(d*)).
//     The null record (unbuilt cell, outside the strip) has kappa2 = -1e30: e = 0 without a test.
//   * ex2.approx (2 ulp), the fp32 product/sums of the accumulation: a 2^-14 slack on the total, and 1e-6 absolute for
//     results flushed to zero:  L = -(sum (1 + 2^-14)) - 1e-6.
// cost >= L because each fp64 term is >= -(upper bound of its exponential).
struct ScreenCtx {
  const float* rec32;          // shared: [n_rec + 1][8] = {l00, l11, l10, kappa2, -mx, -my, -, -}
  const unsigned short* grid;  // shared
  float2 k2, off2;  // (1/cs, 1/cs) and ((W/2)/cs - 0.5, (H/2)/cs - 0.5): the cell coordinates minus one half
  float beta_c;     // 0.5 - beta
  int gw, span;
  unsigned base;
};

constexpr float kScreenMagic = 12582912.0f;      // 1.5 * 2^23: adding it rounds to an integer
constexpr int kScreenMagicBits = 0x4B400000;     // its bit pattern

// One scan point against one candidate, with Blackwell's packed fp32 arithmetic (FFMA2 / FADD2 / FMUL2: two IEEE results per
// instruction) wherever x and y go through the same operation — the loop is bound by instruction issue, not by the fp32 pipe.
//   px2 = (px, px), py2 = (py, py);  cs = (cos, sin), sc = (-sin, cos), txy = (tx, ty) of the candidate
__device__ __forceinline__ void screen_point(const ScreenCtx& m, const float2 px2, const float2 py2, const float2 txy, const float2 cs,
                                             const float2 sc, float& acc) {
  const float2 xy = __ffma2_rn(px2, cs, __ffma2_rn(py2, sc, txy));  // transform_point
  // No bounds test: a point outside the frame contributes 0 to the fp64 cost, so ANY e >= 0 bounds its term.  Outside in y
  // its index leaves the strip and clamps to the null record; outside in x it aliases to a cell of a neighbouring row, a
  // frame width away, whose Gaussian is 0 there to fp32 (d is taken from the true position).  Points that fp32 and fp64
  // place on different sides of the frame's border lie within beta of a cell edge (frames are whole cells): `unc` below.
  const float2 uv = __ffma2_rn(xy, m.k2, m.off2);                   // cell coordinates - 0.5
  const float2 t2 = __fadd2_rn(uv, make_float2(kScreenMagic, kScreenMagic));  // round to nearest of (u - 0.5) = floor(u), |u| < 2^22
  const float2 fl = __fadd2_rn(t2, make_float2(-kScreenMagic, -kScreenMagic));
  const float2 df = __ffma2_rn(fl, make_float2(-1.f, -1.f), uv);    // fractional part - 0.5
  const bool unc = fmaxf(fabsf(df.x), fabsf(df.y)) > m.beta_c;      // within beta of a cell edge
  // ix + gw*iy - base with ix = bits(t) - magic bits: the constants are folded into `base`
  const unsigned g = __float_as_uint(t2.x) + static_cast<unsigned>(m.gw) * __float_as_uint(t2.y) - m.base;
  const unsigned gs = min(g, static_cast<unsigned>(m.span));
  const unsigned r = m.grid[gs];
  const float* q = m.rec32 + 8 * r;
  const float4 l = *reinterpret_cast<const float4*>(q);        // l00, l11, l10, kappa2
  const float2 nmu = *reinterpret_cast<const float2*>(q + 4);  // -mx, -my
  const float2 d = __fadd2_rn(xy, nmu);
  const float2 zz = __fmul2_rn(make_float2(l.x, l.y), d);      // l00 d0, l11 d1
  const float z0 = fmaf(l.z, d.y, zz.x);
  const float xe = fmaf(-zz.y, zz.y, fmaf(-z0, z0, l.w));
  float e;  // ex2.approx: 2 ulp, results below 2^-126 flushed to zero; both covered by the total's slack
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fminf(xe, 0.f)));
  acc += unc ? 1.0f : e;
}

// warp sums of JB per-lane accumulators, packed like packed_warp_sum (same slot assignment)
template <int JB>
__device__ __forceinline__ float packed_warp_sum_f(const float (&a)[JB], int lane);
template <>
__device__ __forceinline__ float packed_warp_sum_f<2>(const float (&a)[2], int lane) {
  const bool hi16 = (lane & 16) != 0;
  float k = hi16 ? a[1] : a[0];
  k += __shfl_xor_sync(0xffffffffu, hi16 ? a[0] : a[1], 16);
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) k += __shfl_xor_sync(0xffffffffu, k, off);
  return k;
}
template <>
__device__ __forceinline__ float packed_warp_sum_f<4>(const float (&a)[4], int lane) {
  const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
  float k01 = hi16 ? a[1] : a[0];
  k01 += __shfl_xor_sync(0xffffffffu, hi16 ? a[0] : a[1], 16);
  float k23 = hi16 ? a[3] : a[2];
  k23 += __shfl_xor_sync(0xffffffffu, hi16 ? a[2] : a[3], 16);
  float k = hi8 ? k23 : k01;
  k += __shfl_xor_sync(0xffffffffu, hi8 ? k01 : k23, 8);
#pragma unroll
  for (int off = 4; off > 0; off >>= 1) k += __shfl_xor_sync(0xffffffffu, k, off);
  return k;
}

template <>
__device__ __forceinline__ float packed_warp_sum_f<8>(const float (&a)[8], int lane) {
  const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0, hi4 = (lane & 4) != 0;
  float k01 = hi16 ? a[1] : a[0];
  k01 += __shfl_xor_sync(0xffffffffu, hi16 ? a[0] : a[1], 16);
  float k23 = hi16 ? a[3] : a[2];
  k23 += __shfl_xor_sync(0xffffffffu, hi16 ? a[2] : a[3], 16);
  float k45 = hi16 ? a[5] : a[4];
  k45 += __shfl_xor_sync(0xffffffffu, hi16 ? a[4] : a[5], 16);
  float k67 = hi16 ? a[7] : a[6];
  k67 += __shfl_xor_sync(0xffffffffu, hi16 ? a[6] : a[7], 16);
  float ka = hi8 ? k23 : k01;
  ka += __shfl_xor_sync(0xffffffffu, hi8 ? k01 : k23, 8);
  float kb = hi8 ? k67 : k45;
  kb += __shfl_xor_sync(0xffffffffu, hi8 ? k45 : k67, 8);
  float k = hi4 ? kb : ka;
  k += __shfl_xor_sync(0xffffffffu, hi4 ? ka : kb, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;
}

// screen of candidates j .. j+JB-1 (clamped to hi-1) on this warp's slice: lbpart[j*NW + warp] = sum of the upper bounds
template <int NPT, int JB>
__device__ __forceinline__ void screen_batch(const ScreenCtx& m, const float2 (&px2)[NPT], const float2 (&py2)[NPT], const float4* pose32,
                                             float* lbpart, int j, int hi, int NW, int warp, int lane) {
  float acc[JB];
  float4 ps[JB];
  float2 sc2[JB];
#pragma unroll
  for (int b = 0; b < JB; ++b) {
    const float4* q = pose32 + 2 * min(j + b, hi - 1);
    ps[b] = q[0];
    sc2[b] = *reinterpret_cast<const float2*>(q + 1);
    acc[b] = 0.f;
  }
#pragma unroll
  for (int b = 0; b < JB; ++b) {
    const float2 txy = make_float2(ps[b].x, ps[b].y), cs = make_float2(ps[b].z, ps[b].w), sc = sc2[b];
#pragma unroll
    for (int k = 0; k < NPT; ++k) screen_point(m, px2[k], py2[k], txy, cs, sc, acc[b]);
  }
  const float tot = packed_warp_sum_f<JB>(acc, lane);
  const int jj = j + packed_slot<JB>(lane);
  if (packed_writer<JB>(lane) && jj < hi) lbpart[jj * NW + warp] = tot;
}

// fp64 evaluation of the candidates listed in surv[i .. i+JB-1] (clamped to the last entry)
template <int NPT, int JB, bool FAST_GEOM, int VAR>
__device__ __forceinline__ void score_batch_listed(const SliceCtx& m, const double2 (&pt)[NPT], const Pose* pose, double* wpart, const int* surv,
                                                   int i, int ns, int NW, int warp, int lane) {
  double acc[JB];
  double2 txy[JB], cs[JB];
#pragma unroll
  for (int b = 0; b < JB; ++b) {
    const Pose* ps = pose + surv[min(i + b, ns - 1)];
    txy[b] = *reinterpret_cast<const double2*>(&ps->x);
    cs[b] = *reinterpret_cast<const double2*>(&ps->c);
    acc[b] = 0.;
  }
#pragma unroll
  for (int b = 0; b < JB; ++b) {
#pragma unroll
    for (int k = 0; k < NPT; ++k) slice_point<FAST_GEOM, VAR>(m, pt[k], txy[b].x, txy[b].y, cs[b].x, cs[b].y, acc[b]);
  }
  const double tot = packed_warp_sum<JB>(acc, lane);
  const int ii = i + packed_slot<JB>(lane);
  if (packed_writer<JB>(lane) && ii < ns) wpart[surv[ii] * NW + warp] = tot;
}

// ---- thread-block cluster support -------------------------------------------------------------
// A problem may be solved by a cluster of CL CTAs (one per SM) when the batch is too small to fill
// the GPU: the CTAs are arranged as G candidate groups x S point slices (CL = G*S).  CTA (g, s)
// keeps slice s of the scan in registers and scores the candidate batches assigned to group g.
// Its warps' partial scores are first summed inside the CTA, then the CTA pushes one value per
// candidate into the shared memory of ALL CTAs of the cluster (DSMEM stores); after one cluster
// barrier every CTA holds every (candidate, slice) partial and runs phases A and C redundantly on
// its own copy of the swarm (deterministic => identical in every CTA).
struct Topo {
  int CL, S, G;      // cluster size, point slices, candidate groups
  int rank, s, g;    // this CTA
  int NW, PW;        // warps per CTA; partials per candidate that phase C sums (NW when CL == 1, else S)
};

__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store v at the same shared-memory offset as `local` in CTA `rank` of this cluster
__device__ __forceinline__ void st_cluster_f64(const double* local, unsigned rank, double v) {
  unsigned raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local)), "r"(rank));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(raddr), "d"(v) : "memory");
}

// phase B: this warp scores the candidates [lo, hi) of `pose` that belong to its CTA's group on its
// slice of the scan, JB candidates at a time.  wpart[j*NW + warp] receives this warp's partial.
// One batch: candidates j .. j+JB-1 (clamped to hi-1; results of clamped slots are not stored).
template <int NPT, int JB, bool FAST_GEOM, int VAR>
__device__ __forceinline__ void score_batch(const SliceCtx& m, const double2 (&pt)[NPT], const Pose* pose, double* wpart, int j, int hi,
                                            int NW, int warp, int lane) {
  double acc[JB];
  double2 txy[JB], cs[JB];
#pragma unroll
  for (int b = 0; b < JB; ++b) {
    const Pose* ps = pose + min(j + b, hi - 1);
    txy[b] = *reinterpret_cast<const double2*>(&ps->x);
    cs[b] = *reinterpret_cast<const double2*>(&ps->c);
    acc[b] = 0.;
  }
#pragma unroll
  for (int b = 0; b < JB; ++b) {
#pragma unroll
    for (int k = 0; k < NPT; ++k) slice_point<FAST_GEOM, VAR>(m, pt[k], txy[b].x, txy[b].y, cs[b].x, cs[b].y, acc[b]);
  }
  const double tot = packed_warp_sum<JB>(acc, lane);
  const int jj = j + packed_slot<JB>(lane);
  if (packed_writer<JB>(lane) && jj < hi) wpart[jj * NW + warp] = tot;
}

template <int NPT, int JB, int CL, bool FAST_GEOM, int VAR>
__device__ __forceinline__ void score_candidates(const SliceCtx& m, const double2 (&pt)[NPT], const Pose* pose, double* wpart, int lo,
                                                 int hi, const Topo& tp, int warp, int lane) {
  if (CL == 1 && JB == 4) {
    // whole batches of 4, then the 1..3 left over in batches of 2: a swarm of 70 costs 70 evaluations, not 72
    int j = lo;
    for (; j + 4 <= hi; j += 4) score_batch<NPT, 4, FAST_GEOM, VAR>(m, pt, pose, wpart, j, hi, tp.NW, warp, lane);
    for (; j < hi; j += 2) score_batch<NPT, 2, FAST_GEOM, VAR>(m, pt, pose, wpart, j, hi, tp.NW, warp, lane);
  } else {
    for (int j = lo + tp.g * JB; j < hi; j += JB * tp.G) score_batch<NPT, JB, FAST_GEOM, VAR>(m, pt, pose, wpart, j, hi, tp.NW, warp, lane);
  }
}

// Cluster form, after phase B: sum the NW warp partials of every candidate this CTA scored and
// store the sum into cpart[j*S + s] of every CTA of the cluster; one (candidate, destination) pair
// per thread.  Ends with the cluster barrier that makes all partials visible everywhere.
template <int JB, int CL>
__device__ __forceinline__ void exchange_partials(const double* wpart, double* cpart, int lo, int hi, const Topo& tp) {
  __syncthreads();
  const int span = JB * tp.G;
  const int mine = ((hi - lo + span - 1) / span) * JB;  // candidates of my group, padding included
  for (int idx = threadIdx.x; idx < mine * CL; idx += blockDim.x) {
    const int jl = idx / CL, dest = idx - jl * CL;
    const int j = lo + tp.g * JB + (jl / JB) * span + (jl % JB);
    if (j < hi) {
      const double* p = wpart + j * tp.NW;
      double c = 0.;
      for (int w = 0; w < tp.NW; ++w) c += p[w];
      st_cluster_f64(cpart + j * tp.S + tp.s, dest, c);
    }
  }
  cluster_barrier();
}

// cost of candidate j = sum of its PW partials (slice-major, warp-minor) in a fixed 4-way interleaved
// order: identical in every thread and CTA that asks, and short enough a dependency chain for phase C
__device__ __forceinline__ double candidate_total(const double* part, int j, int PW) {
  const double* p = part + j * PW;
  double c0 = 0., c1 = 0., c2 = 0., c3 = 0.;
  int w = 0;
  for (; w + 4 <= PW; w += 4) {
    c0 += p[w];
    c1 += p[w + 1];
    c2 += p[w + 2];
    c3 += p[w + 3];
  }
  for (; w < PW; ++w) c0 += p[w];
  return (c0 + c1) + (c2 + c3);
}

// candidate_total, or the +1e300 the screen left in the first partial of a candidate it ruled out (the other partials of
// such a candidate are stale)
__device__ __forceinline__ double candidate_cost(const double* part, int j, int PW) {
  const double p0 = part[j * PW];
  return p0 > 1e299 ? p0 : candidate_total(part, j, PW);
}

template <int NPT, int JB, int CL, bool FAST_GEOM, int VAR>
__device__ __forceinline__ void sliced_body(const SliceCtx& m, const ScreenCtx& sc, const bool scr, const double2 (&pt)[NPT],
                                            const DevProblem& pr, const PsoParams& prm, const SlicedSmem& sm, const Topo& tp,
                                            double* __restrict__ out, int* __restrict__ stats) {
  const int tid = threadIdx.x, T = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int P = prm.P, I = prm.I, PW = tp.PW;
  const int* __restrict__ rnd = pr.rnd;
  NDTPSO_PHASE_DECL
  Pose* pose0 = sm.pose;
  Pose* pose1 = sm.pose + (P + 1);
  // CL == 1: phase C sums the NW warp partials directly (double buffered).
  // CL > 1 : warps write wpart (single buffer), exchange_partials() fills the double-buffered cpart.
  double* part0 = sm.partial;
  double* part1 = sm.partial + (size_t)(P + 1) * PW;
  double* wpart = sm.wpart;
  // the 6P velocity coefficients of iteration `iter` -> ubuf[iter & 1]; issued right before a scoring
  // phase so that the global-memory latency and the conversions hide behind it
  auto prefetch_draws = [&](int iter) {
    if (iter >= I) return;
    double* dst = sm.ubuf + (iter & 1) * 6 * P;
    const int* src = rnd + 3 + 3 * P + 6 * P * iter;
    for (int i = tid; i < 6 * P; i += T) dst[i] = fabs(unit_random(src[i]));  // Array2d::Random().abs(), core.cpp:84
  };

  // ---- initial swarm: task 0 = the seed particle (core.cpp:53,58), task 1+j = particle j (core.cpp:60-61)
  for (int t = tid; t < P + 1; t += T) {
    double pos[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double dv = (t == 0) ? (k == 2 ? 1E-5 : 1E-4) : pr.dev[k];
      pos[k] = __dadd_rn(pr.guess[k], __dmul_rn(unit_random(rnd[3 * t + k]), dv));
    }
    double s, c;
    sincos(pos[2], &s, &c);
    pose0[t] = Pose{pos[0], pos[1], c, s, pos[2], 0.};
  }
  __syncthreads();
  prefetch_draws(0);
  score_candidates<NPT, JB, CL, FAST_GEOM, VAR>(m, pt, pose0, CL == 1 ? part0 : wpart, 0, P + 1, tp, warp, lane);
  if (CL == 1)
    __syncthreads();
  else
    exchange_partials<JB, CL>(wpart, part0, 0, P + 1, tp);
  for (int t = tid; t < P + 1; t += T) sm.cost0[t] = candidate_total(part0, t, PW);
  __syncthreads();
  // every thread derives the initial gbest the way core.cpp:58-69 does (strict <, index order)
  double gbc = sm.cost0[0], gb0 = pose0[0].x, gb1 = pose0[0].y, gb2 = pose0[0].th;
  for (int j = 0; j < P; ++j) {
    const double cj = sm.cost0[1 + j];
    if (cj < gbc) {
      gbc = cj;
      gb0 = pose0[1 + j].x;
      gb1 = pose0[1 + j].y;
      gb2 = pose0[1 + j].th;
    }
  }
  for (int j = tid; j < P; j += T) {  // owner-private state
    const Pose ps = pose0[1 + j];
    sm.x[3 * j] = ps.x;
    sm.x[3 * j + 1] = ps.y;
    sm.x[3 * j + 2] = ps.th;
    sm.pb[3 * j] = ps.x;
    sm.pb[3 * j + 1] = ps.y;
    sm.pb[3 * j + 2] = ps.th;
    sm.v[3 * j] = sm.v[3 * j + 1] = sm.v[3 * j + 2] = 0.;
    sm.pbc[j] = sm.cost0[1 + j];
  }

  // ---- iterations
  int it = 0, start = 0, par = 1, rounds = 0, n_gb = 0, n_f64 = P + 1, n_scr = 0;
  // Speculation window.  While gbest is improving often (the first iterations: every particle jumps towards gbest),
  // a round speculates only on the next `win` particles, so an improvement discards at most a window's worth of
  // evaluations instead of the rest of the swarm.  An iteration starts with the whole swarm as its window unless the
  // previous one saw at least `hot_thresh` improvements.  Any window gives the reference's sequential order.
  // After an improvement the window is `hot_chunk`; every round without one doubles it.
  int hot = prm.hot_chunk > 0 ? 1 : 0, imp_it = 0, win = prm.hot_chunk;
  double w = prm.w;
  NDTPSO_PHASE_MARK(0)
  while (it < I) {
    Pose* pose = par ? pose1 : pose0;
    double* part = par ? part1 : part0;
    const int lim = win > 0 ? min(P, start + win) : P;  // this round covers particles [start, lim)
    // phase A: owners of the pending particles [start, P)
    const int ja = start + ((tid - start) % T + T) % T;  // first pending particle owned by this thread
    const double* ucoef = sm.ubuf + (it & 1) * 6 * P;
    for (int j = ja; j < lim; j += T) {
      double nx[3], nv[3];
      const double gb[3] = {gb0, gb1, gb2};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double rx = ucoef[6 * j + 2 * k];  // draw 3+3P + 6P*it + 6j + 2k (and +1), prefetched
        const double ry = ucoef[6 * j + 2 * k + 1];
        const double xk = sm.x[3 * j + k], vk = sm.v[3 * j + k], pbk = sm.pb[3 * j + k];
        // core.cpp:85-87: ((w*v) + ((c1*rx)*(pb-x))) + ((c2*ry)*(gb-x)), no contraction
        const double t1 = __dmul_rn(w, vk);
        const double t2 = __dmul_rn(__dmul_rn(prm.c1, rx), __dadd_rn(pbk, -xk));
        const double t3 = __dmul_rn(__dmul_rn(prm.c2, ry), __dadd_rn(gb[k], -xk));
        nv[k] = __dadd_rn(__dadd_rn(t1, t2), t3);
        nx[k] = __dadd_rn(xk, nv[k]);  // core.cpp:89
      }
      double s, c;
      sincos(nx[2], &s, &c);
      pose[j] = Pose{nx[0], nx[1], c, s, nx[2], 0.};
      if (CL == 1 && scr) {
        const float cf = static_cast<float>(c), sf = static_cast<float>(s);
        sm.pose32[2 * j] = make_float4(static_cast<float>(nx[0]), static_cast<float>(nx[1]), cf, sf);
        sm.pose32[2 * j + 1] = make_float4(-sf, cf, 0.f, 0.f);
      }
      sm.vnew[3 * j] = nv[0];
      sm.vnew[3 * j + 1] = nv[1];
      sm.vnew[3 * j + 2] = nv[2];
    }
    __syncthreads();
    NDTPSO_PHASE_MARK(1)
    if (start == 0) prefetch_draws(it + 1);  // first round of an iteration
    if (CL == 1 && FAST_GEOM && scr) {
      // phase B1: fp32 lower bound of every pending candidate's cost on this warp's slice
      float2 px2[NPT], py2[NPT];
#pragma unroll
      for (int k = 0; k < NPT; ++k) {  // padding points (1e200, 0) become (1e30, 0): still outside every frame, but finite in fp32
        const float fx = static_cast<float>(fmin(pt[k].x, 1e30)), fy = static_cast<float>(pt[k].y);
        px2[k] = make_float2(fx, fx);
        py2[k] = make_float2(fy, fy);
      }
      {
        int j = start;
#if NDTPSO_SCREEN_JB8
        for (; j + 8 <= lim; j += 8) screen_batch<NPT, 8>(sc, px2, py2, sm.pose32, sm.lbpart, j, lim, tp.NW, warp, lane);
#endif
        for (; j + 4 <= lim; j += 4) screen_batch<NPT, 4>(sc, px2, py2, sm.pose32, sm.lbpart, j, lim, tp.NW, warp, lane);
        for (; j < lim; j += 2) screen_batch<NPT, 2>(sc, px2, py2, sm.pose32, sm.lbpart, j, lim, tp.NW, warp, lane);
      }
      __syncthreads();
      NDTPSO_PHASE_MARK(5)
      // warp 0 lists the candidates whose bound does not already rule out an improvement of their particle's best
      // (core.cpp:94); the others get a cost of +1e300, which phase C treats like any cost that improves nothing
      if (warp == 0) {
        int ns = 0;
        for (int base = start; base < lim; base += 32) {
          const int j = base + lane;
          bool alive = false;
          if (j < lim) {
            double u = 0.;
            if ((tp.NW & 3) == 0) {  // rows of 16-byte multiples: vector loads, four independent sums
              const float4* row = reinterpret_cast<const float4*>(sm.lbpart + j * tp.NW);
              double u0 = 0., u1 = 0., u2 = 0., u3 = 0.;
              for (int w2 = 0; w2 < tp.NW / 4; ++w2) {
                const float4 q = row[w2];
                u0 += static_cast<double>(q.x);
                u1 += static_cast<double>(q.y);
                u2 += static_cast<double>(q.z);
                u3 += static_cast<double>(q.w);
              }
              u = (u0 + u1) + (u2 + u3);
            } else {
              for (int w2 = 0; w2 < tp.NW; ++w2) u += static_cast<double>(sm.lbpart[j * tp.NW + w2]);
            }
            const double lower = -(u * (1. + 6.103515625e-5)) - 1e-6;  // slack: ex2.approx, fp32 products and sums (2^-14), flushed denormals
            alive = !(lower >= sm.pbc[j]);
            if (!alive) part[j * PW] = 1e300;  // candidate_cost() looks at this entry first
          }
          const unsigned mask = __ballot_sync(0xffffffffu, alive);
          if (alive) sm.surv[ns + __popc(mask & ((1u << lane) - 1u))] = j;
          ns += __popc(mask);
        }
        if (lane == 0) sm.surv[P + 1] = ns;
      }
      __syncthreads();
      NDTPSO_PHASE_MARK(6)
      // phase B2: the fp64 evaluation of the survivors
      const int ns = sm.surv[P + 1];
      n_f64 += ns;
      n_scr += (lim - start) - ns;
      {
        int i = 0;
        for (; i + 4 <= ns; i += 4) score_batch_listed<NPT, 4, FAST_GEOM, VAR>(m, pt, pose, part, sm.surv, i, ns, tp.NW, warp, lane);
        for (; i + 2 <= ns; i += 2) score_batch_listed<NPT, 2, FAST_GEOM, VAR>(m, pt, pose, part, sm.surv, i, ns, tp.NW, warp, lane);
        if (i < ns) score_batch_listed<NPT, 1, FAST_GEOM, VAR>(m, pt, pose, part, sm.surv, i, ns, tp.NW, warp, lane);  // survivors are few: no padding
      }
    } else {
      n_f64 += lim - start;
      score_candidates<NPT, JB, CL, FAST_GEOM, VAR>(m, pt, pose, CL == 1 ? part : wpart, start, lim, tp, warp, lane);  // phase B
    }
    NDTPSO_PHASE_MARK(2)
    if (CL == 1)
      __syncthreads();
    else
      exchange_partials<JB, CL>(wpart, part, start, lim, tp);
    NDTPSO_PHASE_MARK(3)
    // phase C: j* = first pending particle that improves gbest (core.cpp:98)
    int jstar = -1;
    double cstar = 0.;
    for (int base = start; base < lim && jstar < 0; base += 32) {
      const int j = base + lane;
      const double cj = (j < lim) ? candidate_cost(part, j, PW) : 0.;
      const bool imp = (j < lim) && (cj < gbc);
      const unsigned mask = __ballot_sync(0xffffffffu, imp);
      if (mask) {
        const int src = __ffs(mask) - 1;
        jstar = base + src;
        cstar = __shfl_sync(0xffffffffu, cj, src);
      }
    }
    const int end = (jstar >= 0) ? jstar + 1 : lim;
    for (int j = ja; j < end; j += T) {  // commit own particles in [start, end)  (core.cpp:89-96)
      const Pose ps = pose[j];
      const double cj = candidate_cost(part, j, PW);
      sm.x[3 * j] = ps.x;
      sm.x[3 * j + 1] = ps.y;
      sm.x[3 * j + 2] = ps.th;
      sm.v[3 * j] = sm.vnew[3 * j];
      sm.v[3 * j + 1] = sm.vnew[3 * j + 1];
      sm.v[3 * j + 2] = sm.vnew[3 * j + 2];
      if (cj < sm.pbc[j]) {
        sm.pbc[j] = cj;
        sm.pb[3 * j] = ps.x;
        sm.pb[3 * j + 1] = ps.y;
        sm.pb[3 * j + 2] = ps.th;
      }
    }
    if (jstar >= 0) {  // core.cpp:102-103
      gbc = cstar;
      gb0 = pose[jstar].x;
      gb1 = pose[jstar].y;
      gb2 = pose[jstar].th;
      ++n_gb;
      ++imp_it;
      win = prm.hot_chunk;
    } else if (win > 0) {
      win = min(2 * win, 1 << 20);
    }
    start = end;
    if (start >= P) {
      start = 0;
      ++it;
      w = __dmul_rn(w, prm.wd);  // core.cpp:108
      hot = (prm.hot_chunk > 0 && imp_it >= prm.hot_thresh) ? 1 : 0;
      win = hot ? prm.hot_chunk : 0;
      imp_it = 0;
    }
    par ^= 1;
    ++rounds;
    NDTPSO_PHASE_MARK(4)
  }

  if (tid == 0 && tp.rank == 0) {
    out[0] = gb0;
    out[1] = gb1;
    out[2] = gb2;
    out[3] = gbc;
    if (stats) {
      stats[0] = rounds;
      stats[1] = n_gb;
      stats[2] = n_f64;
      stats[3] = n_scr;
    }
  }
  if (CL > 1) cluster_barrier();  // no CTA may exit while peers can still store into its shared memory
}

// Prologue shared by the production kernel and the phase-B microbenchmark: stages the compact
// table with two bulk TMA copies, loads this thread's scan points into registers meanwhile
// (coalesced 16-byte loads), and fills the loop-invariant context.
template <int NPT>
__device__ __forceinline__ SlicedSmem sliced_prologue(unsigned char* smem_raw, const DevProblem& pr, const DevMap& mp, int P, const Topo& tp,
                                                      SliceCtx& m, double2 (&pt)[NPT], int screen = 0, ScreenCtx* sc = nullptr,
                                                      const PsoParams* prm = nullptr) {
  const int tid = threadIdx.x, T = blockDim.x;
  const int n_rec = mp.hdr[HDR_NREC];
  const int row0 = mp.hdr[HDR_ROW0], nrows = mp.hdr[HDR_NROWS];
  const int span = nrows * mp.gw;
  const int rec_bytes = (n_rec + 1) * 48;
  const int grid_bytes = round16((span + 1) * 2);
  const SlicedSmem sm = carve_sliced(smem_raw, P, tp.PW, tp.CL == 1 ? 0 : tp.NW, rec_bytes + grid_bytes, screen);

  if (tid < kExpTableSize) sm.etab[tid] = c_exp_table[tid];
  // constants go through volatile shared memory so the compiler keeps them in registers instead of
  // re-materialising 64-bit immediates inside the loop
  volatile double* cst = reinterpret_cast<volatile double*>(sm.cst);
#if (NDTPSO_PROD_VARIANT & 4)
  if (tid < 32) {
    const ExpConsts ec = exp_consts();
    sm.cst[8 + tid] = ec.l2e;
    sm.cst[40 + tid] = ec.c7;
  }
#endif
  if (tid == 0) {
    const ExpConsts ec = exp_consts();
    cst[0] = ec.l2e;
    cst[1] = -ec.hi;
    cst[2] = -ec.lo;
    cst[3] = ec.c7;
    cst[4] = ec.c6;
    cst[5] = ec.c5;
    cst[6] = ec.c4;
    cst[7] = ec.c3;
    mbar_init(sm.bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(sm.bar, rec_bytes + grid_bytes);
    tma_load_1d(sm.table, mp.rec, rec_bytes, sm.bar);
    tma_load_1d(sm.table + rec_bytes, mp.grid, grid_bytes, sm.bar);
  }
  // slice s of the scan: points k*(S*T) + s*T + tid
#pragma unroll
  for (int k = 0; k < NPT; ++k) {
    const int i = (k * tp.S + tp.s) * T + tid;
    pt[k] = (i < pr.n_pts) ? pr.pts[i] : make_double2(1e200, 0.);  // padding: out of bounds for every pose
  }
  m.rec = reinterpret_cast<const double*>(sm.table);
  m.grid = reinterpret_cast<const unsigned short*>(sm.table + rec_bytes);
  m.etab = sm.etab;
  m.x_min = mp.x_min;
  m.x_max = mp.x_max;
  m.y_min = mp.y_min;
  m.y_max = mp.y_max;
  m.hw = mp.hw;
  m.hh = mp.hh;
  m.cs = mp.cs;
  m.inv_cs = mp.inv_cs;
  m.hw_s = mp.hw * mp.inv_cs;
  m.hh_s = mp.hh * mp.inv_cs;
#if (NDTPSO_PROD_VARIANT & 4)
  {  // per-lane copies: a lane-dependent address is not uniform, so the two constants stay in vector registers
    volatile double* lanes = reinterpret_cast<volatile double*>(sm.cst + 8);
    m.l2e = lanes[tid & 31];
    m.c7 = lanes[32 + (tid & 31)];
  }
#else
  m.l2e = cst[0];
#endif
  m.ln2hi = cst[1];
  m.ln2lo = cst[2];
#if !(NDTPSO_PROD_VARIANT & 4)
  m.c7 = cst[3];
#endif
  m.c6 = cst[4];
  m.c5 = cst[5];
  m.c4 = cst[6];
  m.c3 = cst[7];
  m.gw = mp.gw;
  m.base = row0 * mp.gw;
  m.span = span;
  m.null_id = n_rec;
  mbar_wait(sm.bar, 0);
  if (screen) {
    // fp32 records of the screen (see above): one per built cell, found through the grid so that the cell's extent is
    // known.  Read after the next __syncthreads (the body has one before its first use).
    const double* rec = reinterpret_cast<const double*>(sm.table);
    const double dd2 = static_cast<double>(prm->scr_dd2);
    const double dd = sqrt(dd2);
    const double u24 = 5.9604644775390625e-08;
    for (int g = tid; g <= span; g += T) {
      const int r = (g < span) ? m.grid[g] : n_rec;
      if (g < span && r == n_rec) continue;  // unbuilt cell
      float* o = sm.rec32 + 8 * r;
      if (r == n_rec) {  // the null record
        o[0] = o[1] = o[2] = 0.f;
        o[3] = -1e30f;
        o[4] = o[5] = o[6] = o[7] = 0.f;
        continue;
      }
      const double* q = rec + 6 * r;
      const double mx = q[0], my = q[1], H00 = -q[2], H01 = -q[3], H11 = -q[5];
      const int cell = g + row0 * mp.gw;
      const double cx = (cell % mp.gw) * mp.cs - mp.hw, cy = (cell / mp.gw) * mp.cs - mp.hh;  // the cell's low corner
      const double dm0 = fmax(fabs(cx - mx), fabs(cx + mp.cs - mx)) + dd, dm1 = fmax(fabs(cy - my), fabs(cy + mp.cs - my)) + dd;
      const double l00 = H00 > 0. ? sqrt(H00) : 0.;
      const double l10 = l00 > 0. ? H01 / l00 : 0.;
      const double l11 = sqrt(fmax(H11 - l10 * l10, 0.));
      // a Cholesky factor that rounding made too large would overstate A: shrink it by what sqrt/div/rounding can add
      const double sh = 1. - 1e-12;
      const double ez0 = 3. * u24 * (l00 * dm0 + fabs(l10) * dm1), ez1 = 3. * u24 * l11 * dm1;
      const double hs = H00 + 2. * fabs(H01) + H11;
      // t trades the relative loosening t*A against the absolute one kappa0/t: balanced for A of order one
      const double kappa0 = ez0 * ez0 + ez1 * ez1 + hs * dd2;
      const double t = fmin(fmax(sqrt(kappa0), 0x1p-10), 0x1p-3);
      const double scale = sqrt((1. - t) * (1. - t) * (1. - 2.384185791015625e-07) * 1.4426950408889634);
      const double kappa = kappa0 / t * 1.000001;  // and the rounding of the sum it enters
      o[0] = static_cast<float>(l00 * scale * sh);
      o[1] = static_cast<float>(l11 * scale * sh);
      o[2] = static_cast<float>(l10 * scale * sh);
      o[3] = __double2float_ru(kappa * 1.4426950408889634);
      o[4] = -static_cast<float>(mx);
      o[5] = -static_cast<float>(my);
      o[6] = o[7] = 0.f;
    }
    sc->rec32 = sm.rec32;
    sc->grid = m.grid;
    sc->k2 = make_float2(static_cast<float>(mp.inv_cs), static_cast<float>(mp.inv_cs));
    sc->off2 = make_float2(static_cast<float>(mp.hw * mp.inv_cs - 0.5), static_cast<float>(mp.hh * mp.inv_cs - 0.5));
    sc->beta_c = prm->scr_beta_c;
    sc->gw = mp.gw;
    sc->base = static_cast<unsigned>(m.base) + static_cast<unsigned>(kScreenMagicBits) * (1u + static_cast<unsigned>(mp.gw));  // folds the magic bits of both coordinates
    sc->span = m.span;
  }
  return sm;
}

template <int CL>
__device__ __forceinline__ Topo make_topo(int groups) {
  Topo tp;
  tp.CL = CL;
  tp.G = (CL == 1) ? 1 : groups;
  tp.S = CL / tp.G;
  tp.rank = (CL == 1) ? 0 : static_cast<int>(cluster_ctarank());
  tp.s = tp.rank % tp.S;
  tp.g = tp.rank / tp.S;
  tp.NW = blockDim.x >> 5;
  tp.PW = (CL == 1) ? tp.NW : tp.S;
  return tp;
}

// Host guarantees: every table is compact and symmetric and fits the dynamic shared memory;
// n_pts <= NPT * S * blockDim.x; the grid is n_problems * CL CTAs launched as clusters of CL.
template <int NPT, int JB, int CL, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) pso_sliced_kernel(const DevProblem* __restrict__ probs, const DevMap* __restrict__ maps,
                                                              PsoParams prm, int groups, double* __restrict__ out, int* __restrict__ stats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x / CL;
  const DevProblem& pr = probs[b];
  const DevMap& mp = maps[pr.map_id];
  const Topo tp = make_topo<CL>(groups);
  SliceCtx m;
  ScreenCtx sc;
  double2 pt[NPT];
  const int screen = (CL == 1) ? prm.screen : 0;  // shared memory is laid out for it whenever the launch asks for it
  const SlicedSmem sm = sliced_prologue<NPT>(smem_raw, pr, mp, prm.P, tp, m, pt, screen, &sc, &prm);
  if (CL > 1) cluster_barrier();  // every CTA's shared memory is carved before anyone stores into it
  double* o = out + 4 * (size_t)b;
  int* s = stats ? stats + kStatsWords * (size_t)b : nullptr;
  if (mp.fast_geom)
    sliced_body<NPT, JB, CL, true, kProdVariant>(m, sc, screen != 0, pt, pr, prm, sm, tp, o, s);
  else
    sliced_body<NPT, JB, CL, false, kProdVariant>(m, sc, false, pt, pr, prm, sm, tp, o, s);  // the screen's geometry is the fast one
  if (threadIdx.x == 0 && tp.rank == 0) publish_result(prm.ex, b, gridDim.x / CL, o);  // the thread that wrote o
}

// The screen's lower bound for given poses (ndtpso_screen_bounds): the very code phase B1 runs, so that tests can check
// "bound <= fp64 cost" pose by pose instead of only through the decisions it leads to.  One CTA per problem, T threads holding
// NPT points each; poses [n][m][3], out [n][m].  prm.P = m - 1 sizes the shared-memory arrays.
template <int NPT, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) screen_bound_kernel(const DevProblem* __restrict__ probs, const DevMap* __restrict__ maps, PsoParams prm,
                                                              int m_poses, const double* __restrict__ poses, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int b = blockIdx.x;
  const DevProblem& pr = probs[b];
  const DevMap& mp = maps[pr.map_id];
  const Topo tp = make_topo<1>(1);
  SliceCtx m;
  ScreenCtx sc;
  double2 pt[NPT];
  const SlicedSmem sm = sliced_prologue<NPT>(smem_raw, pr, mp, prm.P, tp, m, pt, 1, &sc, &prm);
  const int tid = threadIdx.x, T = blockDim.x, warp = tid >> 5, lane = tid & 31;
  for (int j = tid; j < m_poses; j += T) {
    const double* p = poses + 3 * ((size_t)b * m_poses + j);
    double s, c;
    sincos(p[2], &s, &c);
    const float cf = static_cast<float>(c), sf = static_cast<float>(s);
    sm.pose32[2 * j] = make_float4(static_cast<float>(p[0]), static_cast<float>(p[1]), cf, sf);
    sm.pose32[2 * j + 1] = make_float4(-sf, cf, 0.f, 0.f);
  }
  __syncthreads();
  float2 px2[NPT], py2[NPT];
#pragma unroll
  for (int k = 0; k < NPT; ++k) {
    const float fx = static_cast<float>(fmin(pt[k].x, 1e30)), fy = static_cast<float>(pt[k].y);
    px2[k] = make_float2(fx, fx);
    py2[k] = make_float2(fy, fy);
  }
  {
    int j = 0;
    for (; j + 8 <= m_poses; j += 8) screen_batch<NPT, 8>(sc, px2, py2, sm.pose32, sm.lbpart, j, m_poses, tp.NW, warp, lane);
    for (; j + 4 <= m_poses; j += 4) screen_batch<NPT, 4>(sc, px2, py2, sm.pose32, sm.lbpart, j, m_poses, tp.NW, warp, lane);
    for (; j < m_poses; j += 2) screen_batch<NPT, 2>(sc, px2, py2, sm.pose32, sm.lbpart, j, m_poses, tp.NW, warp, lane);
  }
  __syncthreads();
  for (int j = tid; j < m_poses; j += T) {
    double u = 0.;
    for (int w2 = 0; w2 < tp.NW; ++w2) u += static_cast<double>(sm.lbpart[j * tp.NW + w2]);
    out[(size_t)b * m_poses + j] = mp.fast_geom ? -(u * (1. + 6.103515625e-5)) - 1e-6 : -1e300;  // the same slack as phase B1
  }
}

// Phase-B microbenchmark: every CTA stages problem blockIdx.x % n_problems and scores `ncand`
// synthetic candidates around its guess `reps` times.  out[blockIdx.x] = a checksum.
template <int NPT, int JB, int VAR, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) score_bench_kernel(const DevProblem* __restrict__ probs, const DevMap* __restrict__ maps,
                                                               int n_problems, int ncand, int reps, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const DevProblem& pr = probs[blockIdx.x % n_problems];
  const DevMap& mp = maps[pr.map_id];
  const Topo tp = make_topo<1>(1);
  SliceCtx m;
  double2 pt[NPT];
  const SlicedSmem sm = sliced_prologue<NPT>(smem_raw, pr, mp, ncand - 1, tp, m, pt);
  const int tid = threadIdx.x, T = blockDim.x, NW = T >> 5, warp = tid >> 5, lane = tid & 31;
  for (int j = tid; j < ncand; j += T) {
    const double th = pr.guess[2] + 1e-3 * (j % 17 - 8);
    double s, c;
    sincos(th, &s, &c);
    sm.pose[j] = Pose{pr.guess[0] + 0.01 * (j % 13 - 6), pr.guess[1] + 0.01 * (j % 11 - 5), c, s, th, 0.};
  }
  __syncthreads();
  for (int r = 0; r < reps; ++r) {
    if (mp.fast_geom)
      score_candidates<NPT, JB, 1, true, VAR>(m, pt, sm.pose, sm.partial, 0, ncand, tp, warp, lane);
    else
      score_candidates<NPT, JB, 1, false, VAR>(m, pt, sm.pose, sm.partial, 0, ncand, tp, warp, lane);
    __syncthreads();
  }
  if (tid == 0) {
    double c = 0.;
    for (int j = 0; j < ncand; ++j)
      for (int w = 0; w < NW; ++w) c += sm.partial[j * NW + w];
    out[blockIdx.x] = c;
  }
}

}  // namespace ndtpso
