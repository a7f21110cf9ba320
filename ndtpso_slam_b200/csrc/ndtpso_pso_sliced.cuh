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
#ifndef NDTPSO_SCREEN_JB
#define NDTPSO_SCREEN_JB 8  // candidates the screen takes at a time first (more loads in flight, fewer reductions), then 4, then pairs
#endif
#ifndef NDTPSO_SCREEN_DEFER
#define NDTPSO_SCREEN_DEFER 1  // the screen reduces a batch's sums while the next batch is being evaluated (2.178 -> 2.171 ms per 256 matches)
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

// Shared memory: [mbarrier 16][exp table 128][exp constants 64][per-lane copies of two of them 512][frame geometry 64], then
// the staged table at a FIXED offset (the screen's records come first, so the hot loop addresses them as
// constant + the byte offset it reads from the grid), then the swarm arrays.
constexpr int kSlicedGeomDoubles = 8;  // x_max, y_max, 1/cs, (W/2)/cs, (H/2)/cs (what the fp64 evaluation re-reads every round)
constexpr int kSlicedTableOffset = 16 + (kExpTableSize + 8 + 64 + kSlicedGeomDoubles) * (int)sizeof(double);

struct SlicedSmem {
  uint64_t* bar;
  double* etab;     // [16]
  double* cst;      // [8] fast_exp constants
  unsigned char* table;  // fp32 screen records (screened launches), then the grid, then the fp64 records
  Pose* pose;       // [2][P+1]
  double* partial;  // [NB][(P+1)*PW], NB = 2 in the cluster form, 1 otherwise (see sliced_swarm_smem_bytes)
  double* wpart;    // [(P+1)*NW] warp partials of the cluster form (unused when CL == 1)
  double* ubuf;     // [2][6P] |Random()| coefficients of the current / next iteration (core.cpp:84)
  double* cost0;    // [P+1] initial costs: aliases the second half of ubuf, which is first written after the barrier that ends the first phase A
  double* x;        // [P][3]  owner-private particle state
  double* v;        // [P][3]
  double* vnew;     // [P][3]
  double* pb;       // [P][3]
  double* pbc;      // [P]
  // fp32 screening (CL == 1 only; null when off)
  float4* pose32;   // [P+1] {tu, tv, c/cs, s/cs} of the current round's candidates (cell units)
  float* lbpart;    // [(P+1)*NW] per-warp partial sums of the upper bounds
  unsigned short* wsurv;  // [NW][P+2] per warp: the candidates that need the fp64 evaluation, ascending
};

// PW = partials per candidate summed in phase C; WP = warp partials per candidate of the cluster form (0 when CL == 1).
// The partials are written by phase B and read by phase C, and a barrier (the one that ends phase A) separates a round's
// phase C from the next round's phase B, so one buffer suffices; the cluster form fills them through remote stores that a
// fast CTA may issue while a slow one still reads the previous round's, hence its two buffers.
__host__ __device__ inline int sliced_swarm_smem_bytes(int P, int PW, int WP) {
  const int Pn = P > 0 ? P : 1;
  int b = 2 * (P + 1) * (int)sizeof(Pose);
  b += (WP > 0 ? 2 : 1) * (P + 1) * PW * (int)sizeof(double);
  b += (P + 1) * WP * (int)sizeof(double);
  b += 2 * 6 * Pn * (int)sizeof(double);
  b += 13 * Pn * (int)sizeof(double);
  return (b + 15) & ~15;
}
// The screen's records, n of them (null record included), in blocks of eight: [8 x {l00, l11, c0, c1}][8 x {l10, -, -, -}].
// Record r's first part sits at 256 (r / 8) + 16 (r % 8), its second part 128 bytes further: consecutive records — the
// neighbouring cells a warp's points fall into — lie in different shared-memory banks (tools/microbench/lds_wavefronts.cu:
// with whole records 32 bytes apart, records r and r + 4 collide).  20 bytes are loaded per evaluation (LDS.128 + LDS.32): the
// loop is bound by shared-memory wavefronts, which a warp-wide load costs in proportion to its width (3.1 / 1.7 / 1.0 for
// 16 / 8 / 4 bytes with the ~6 records a warp's 32 beams hit), so the additive term of the exponent is ONE value for the
// whole table (ScreenCtx::kappa2) instead of a sixth field of every record.
__host__ __device__ inline int screen_rec_bytes(int n) { return ((n + 7) / 8) * 256; }
__host__ __device__ inline unsigned screen_rec_offset(unsigned r) { return 256u * (r >> 3) + 16u * (r & 7u); }
// shared memory of the fp32 screen without its record table: pose32, lbpart, wsurv
__host__ __device__ inline int sliced_screen_fixed_bytes(int P, int NW) {
  return (P + 1) * 16 + round16((P + 1) * NW * 4) + round16(NW * (P + 2) * 2);
}
// total dynamic shared memory: table_bytes = (n_rec + 1) * 48 + round16((span + 1) * 2).  With the screen (screen_recs > 0:
// the largest record count of the batch, null record included) the fp32 records take screen_recs * 32 bytes more.
__host__ __device__ inline int sliced_smem_bytes(int P, int PW, int WP, int table_bytes, int screen_recs = 0) {
  return kSlicedTableOffset + table_bytes + sliced_swarm_smem_bytes(P, PW, WP) +
         (screen_recs > 0 ? sliced_screen_fixed_bytes(P, PW) + screen_rec_bytes(screen_recs) : 0);
}

// table_bytes = this CTA's staged table: [fp32 records (screen only)][grid][fp64 records].  Every pointer is `base` plus an
// integer offset (no pointer/integer round trips), so the compiler keeps them in the shared address space (LDS, not LD).
__device__ __forceinline__ SlicedSmem carve_sliced(unsigned char* base, int P, int PW, int WP, int table_bytes, int screen = 0) {
  SlicedSmem s;
  const int Pn = P > 0 ? P : 1;
  s.bar = reinterpret_cast<uint64_t*>(base);
  s.etab = reinterpret_cast<double*>(base + 16);
  s.cst = s.etab + kExpTableSize;
  s.table = base + kSlicedTableOffset;
  int o = kSlicedTableOffset + table_bytes;
  s.pose = reinterpret_cast<Pose*>(base + o);
  o += 2 * (P + 1) * (int)sizeof(Pose);
  s.partial = reinterpret_cast<double*>(base + o);
  o += (WP > 0 ? 2 : 1) * (P + 1) * PW * (int)sizeof(double);
  s.wpart = reinterpret_cast<double*>(base + o);
  o += (P + 1) * WP * (int)sizeof(double);
  s.ubuf = reinterpret_cast<double*>(base + o);
  s.cost0 = s.ubuf + 6 * Pn;  // P + 1 <= 6 Pn
  o += 2 * 6 * Pn * (int)sizeof(double);
  double* d = reinterpret_cast<double*>(base + o);
  s.x = d;
  s.v = d + 3 * Pn;
  s.vnew = d + 6 * Pn;
  s.pb = d + 9 * Pn;
  s.pbc = d + 12 * Pn;
  o = kSlicedTableOffset + table_bytes + sliced_swarm_smem_bytes(P, PW, WP);  // rounded to 16
  s.pose32 = nullptr;
  s.lbpart = nullptr;
  s.wsurv = nullptr;
  if (screen) {
    s.pose32 = reinterpret_cast<float4*>(base + o);
    o += (P + 1) * 16;
    s.lbpart = reinterpret_cast<float*>(base + o);
    o += round16((P + 1) * PW * 4);
    s.wsurv = reinterpret_cast<unsigned short*>(base + o);
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
  unsigned goff;  // screened launches: the grid holds goff + screen_rec_offset(record id), the shared-memory address of the screen's record
};

// One scan point against one candidate pose: subtracts exp(-(d' S d)/2) from acc iff the point is
// inside the frame (strict), its cell is built, and the value is a normal double (>= 2.2e-308;
// smaller ones are flushed to zero, see fast_exp.h).  Padding points of the last slice are stored
// as (1e200, 0): whatever the pose, |x'| or |y'| is then ~1e200, i.e. out of bounds, so they need
// no validity flag.  The host only selects this kernel for tables whose every Sigma^-1 is finite,
// symmetric and positive semi-definite (what NDTCell::build produces), so the exponent is <= 0 up
// to rounding and can never overflow; anything else takes the generic kernel (library exp).
template <bool FAST_GEOM, int VAR, int GSHIFT = 0>
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
  const unsigned ge = m.grid[in_strip ? g : static_cast<unsigned>(m.span)];
  const unsigned ga = ge - m.goff;
  // screened launches keep the shared address of the screen's record in the grid: goff + screen_rec_offset(id)
  const unsigned r = GSHIFT ? ((ga >> 8) << 3) | ((ga >> 4) & 7u) : ge;
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
// Everything is done in CELL units (frames the screen accepts are square, whole cells of a power-of-two side cs, so 1/cs is
// exact).  Notation: u = 2^-24 (every fp32 operation returns x(1 + eta), |eta| <= u); for a scan point p and a candidate
// (tx, ty, c, s):   U* = (p_x c - p_y s + tx + W/2)/cs - 1/2   (and V* alike): the point's cell coordinate minus one half
// as the fp64 evaluation sees it (its own rounding, ~2^-50 relative, is folded into the constants below).
//
//   (1) Cell coordinate.  Phase A stores, per candidate, ck = fl(c/cs), sk = fl(s/cs), tu = fl(tx/cs + (W/2)/cs - 1/2)
//       (computed in fp64, rounded once); the point is held as px~ = fl(p_x), py~ = fl(p_y).  The screen computes
//           Uf = fl(px~ ck + fl(py~ (-sk) + tu))                                              (two FMAs)
//       Expanding,  Uf - U* = p_x c/cs [(1+e1)(1+e2)(1+e7) - 1] - p_y s/cs [(1+e3)(1+e4)(1+e6)(1+e7) - 1]
//                           + tau [(1+e5)(1+e6)(1+e7) - 1],   tau = tx/cs + (W/2)/cs - 1/2,
//       and |(1+e)^3 - 1| <= 3u + 3u^2 + u^3, |(1+e)^4 - 1| <= 4u + 6u^2 + 4u^3 + u^4 — ALL orders, not only the first:
//           |Uf - U*| <= (3u + 4u^2)(|p_x|/cs + |tau|) + (4u + 7u^2)|p_y|/cs.
//       A point that the fp64 evaluation places INSIDE the frame has |tx| <= W/2 + sqrt2 pmax (pmax = largest |coordinate| of
//       the scan), hence |tau| <= (W + 1.4143 pmax)/cs + 1/2, and the first-order part is u (11.25 pmax + 3 W)/cs + 1.5u.
//       The host (screen_params, ndtpso_capi.cu) sets
//           du := 2^-24 (12 pmax' + 3.01 W)/cs + 2^-23,   pmax' = max(pmax, 1)            (PsoParams::scr_du)
//       whose margin over the first-order part, u (0.75 pmax' + 0.01 W)/cs + 0.5u, exceeds the higher-order terms
//       (<= 11 u^2 (pmax + W)/cs + 2u^2) and the fp64 side's own rounding (<= 2^-48 (pmax + W)/cs) by orders of magnitude
//       for every batch the host admits ((pmax + W)/cs <= 2^22).  Hence |Uf - U*| <= du for every in-frame point.
//       A point OUTSIDE the frame contributes 0 to the fp64 cost, so whatever e >= 0 the screen computes for it bounds its
//       term: nothing has to hold for it (its index is clamped into the staged strip, so every load is in range).
//   (2) Cell.  n = fl(Uf + 1.5 2^23) - 1.5 2^23 is Uf rounded to the nearest integer (exact for |Uf| < 2^22), i.e. the
//       floor of the cell coordinate, and df = Uf - n is exact, |df| <= 1/2.  If |df| <= 1/2 - beta with beta >= du, then
//       U* lies in the same unit interval: fp32 and fp64 agree on the cell (this also covers the frame's border: frames are
//       whole cells, so a point within du of the border is within du of a cell edge).  Otherwise the point counts as the
//       worst case, exp(.) = 1.
//   (3) Offset from the cell's mean, in cell units: o = (mu + W/2)/cs - (n + 1/2), in [-1/2, 1/2] for a mean inside its cell (any
//       o is handled).  The fp64 evaluation's offset is d* = (U* - n) - o; the screen works with d^ = df - o (never formed: o is
//       folded into the constants c0, c1 of (4)), d^ = d* + eps with |eps| <= du <= delta := du + u(2|o| + 0.51) (the larger
//       delta of the form that subtracted a rounded o explicitly is kept), and |d*_k| <= dmax_k := max(|-1/2 - o_k|, |1/2 - o_k|)
//       (the point is in the cell), |d^_k| <= dmax_k + delta.
//   (4) Exponent.  H = (Sigma^-1/2) cs^2 (positive semi-definite: the host only selects this kernel for such tables),
//       A(d) = d'Hd = |L'd|^2 with H = LL' (Cholesky), so the point's term is -exp(-A(d*)).  For every t in (0, 1) and reals
//       a, b:  (a - b)^2 >= (1 - t) a^2 - (1/t - 1) b^2   (2ab <= t a^2 + b^2/t); applied to the vectors L'd~ and L'eps,
//           A(d*) >= (1 - t) A(d~) - (1/t - 1) A(eps),          A(eps) <= hs delta^2,   hs = H00 + 2|H01| + H11.
//       The screen evaluates z~0 = fl(l~10 df1 + fl(l~00 df0 + c~0)), z~1 = fl(l~11 df1 + c~1) with l~ = fl(l sqrt(S)),
//       c~0 = fl(-(l00 o0 + l10 o1) sqrt(S)), c~1 = fl(-l11 o1 sqrt(S)) (fp64, rounded once), S below; in exact arithmetic these
//       are sqrt(S) z_k(d^).  Each term (l~00 df0, c~0, l~10 df1; l~11 df1, c~1) carries at most three roundings,
//       (1 + u)^3 <= 1 + 4u, and |df_k| <= 1/2, |c0| <= sqrt(S)(l00 |o0| + |l10| |o1|), |c1| = sqrt(S) l11 |o1|, so
//       |z~_k - sqrt(S) z_k(d^)| <= sqrt(S) ez_k with
//           ez_0 = 4u (l00 (dmax_0 + delta) + |l10| (dmax_1 + delta)),   ez_1 = 4u l11 (dmax_1 + delta)      (dmax_k = 1/2 + |o_k|),
//       and again z_k^2 >= (1 - t) z~_k^2/S - (1/t - 1) ez_k^2 (also when |z~_k| < sqrt(S) ez_k: the right side is then <= 0).
//       xe = fl(-z~0^2 + fl(-z~1^2 + kappa2)) >= kappa2 (1 - 2u') - (z~0^2 + z~1^2)(1 + 2u'), u' = u(1 + u) (the squares
//       are exact inside the FMAs).  With S = (1 - t)^2 (1 - 2^-22) log2(e) and
//           kappa2 (1 - 2^-22) >= 1.000001 log2(e) [(ez_0^2 + ez_1^2) + hs delta^2]/t
//       the chain gives  xe >= -A(d*) log2(e), i.e.  e = ex2(xe) >= exp(-A(d*))  (xe may exceed 0 by at most kappa2: e is then
//       a little above 1, still an upper bound).  kappa2 is ONE fp32 value for the whole table (kept in a register; a record
//       is 20 bytes instead of 24) and t is what each record needs to get by with it:
//           kappa0_r = (ez_0^2 + ez_1^2) + hs delta^2,    t_r = 1.000001 log2(e) kappa0_r / (kappa2 (1 - 2^-22)),
//           kappa2 = log2(e) sqrt(mean_r kappa0_r):
//       the additive loosening 0.69 kappa2 of every term against the relative loosening t_r A of the exponents, balanced for
//       A = 1, the mean exponent of a two-dimensional Gaussian's mass (screen_kappa2 below; two passes over the records in
//       the prologue, fixed-order reductions).  A record that would need t_r > 1/2 (a needle far sharper than the rest of its
//       table) gets the trivial bound instead: a zero factor, z = 0, e = 2^kappa2 >= 1 for every point of its cell — so one
//       degenerate cell costs its own points' terms, not the tightness of the whole table.
//       The Cholesky factor is computed in fp64 and shrunk by what its own rounding could add (1e-12 relative on l00, l10;
//       4e-15 H11 absolute on l11^2 before the root, which also covers cancellation in H11 - l10^2).
//       The null record (unbuilt cell, outside the strip) is {l00 = 0, l11 = 0, c0 = 1e18, c1 = 0, l10 = 0}: z0 = 1e18,
//       xe = kappa2 - 1e36, e = 0 without a test.
//   (5) Sum.  ex2.approx is within 2 ulp (2^-22); the per-lane accumulation (NPT multiply-adds with weights 0 or 1), the warp tree (5) and the sum over
//       the warps (NW - 1) are fp32 additions of non-negative terms: relative error <= (NPT + NW + 4) u < 2^-19 for every
//       shape launched.  The fp64 evaluation's own deviation from exact arithmetic (FMA roundings in the exponent, ~1e-12
//       relative in exp; fast_exp 1 ulp; tree sum) is below 2^-36.  Total:  L = -(sum (1 + 2^-14)) - 1e-6, the absolute term
//       for results flushed to zero (ex2.approx.ftz below 2^-126; the fp64 side flushes too, which only raises its cost).
// cost >= L because each fp64 term is >= -(upper bound of its exponential).
struct ScreenCtx {
  // shared: records {l00, l11, c0, c1}, {l10, -, -, -} in blocks of eight (screen_rec_bytes) from the 32-bit shared address rec32
  unsigned rec32;
  const unsigned short* grid;  // shared: the 32-bit shared ADDRESS of the cell's record (rec32 + screen_rec_offset(id); below 2^16)
  float beta_c;                // 0.5 - beta
  float kappa2;                // the exponent's additive term, one value for the table (derivation (4))
  int gw;
  unsigned span, nbase;        // nbase = -(first cell of the strip + the magic bits of both coordinates)
};

// shared-memory loads by 32-bit address: the record's address comes straight out of the grid, no pointer arithmetic
__device__ __forceinline__ float4 lds_f4(unsigned a) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds_f1_128(unsigned a) {
  float v;
  asm("ld.shared.f32 %0, [%1+128];" : "=f"(v) : "r"(a));
  return v;
}

constexpr float kScreenMagic = 12582912.0f;      // 1.5 * 2^23: adding it rounds to an integer
constexpr int kScreenMagicBits = 0x4B400000;     // its bit pattern

// what phase A leaves for the screen: the candidate in cell units, as the register pairs the packed transform takes
__device__ __forceinline__ void store_pose32(float4* pose32, int j, double x, double y, double c, double s, double inv_cs, double off_u,
                                             double off_v) {
  const float ck = static_cast<float>(c * inv_cs), sk = static_cast<float>(s * inv_cs);  // c/cs is exact in fp64: one rounding
  pose32[j] = make_float4(static_cast<float>(fma(x, inv_cs, off_u)), static_cast<float>(fma(y, inv_cs, off_v)), ck, sk);
}

// This lane's scan points for the screen: the (x, y), (-y, x) pairs of the packed transform, and per slot a weight (1 for a
// scan point, 0 for padding).  A padding slot holds a COPY of one of the lane's own scan points (of the scan's first point in
// a lane that has none), so it goes through the arithmetic like any point, cannot add a new "near a cell edge" case, and is
// dropped from the sum by its weight.
template <int NPT>
struct ScreenPts {
  float2 px2[NPT], py2[NPT];
  float w[NPT];
  float wsum;  // number of scan points this lane holds
};

// One scan point against one candidate, with Blackwell's packed fp32 arithmetic (FFMA2 / FADD2 / FMUL2: two IEEE results per
// instruction) wherever the two coordinates go through the same operation: the upper bound of the point's exp(.) for a point
// that is not within beta of a cell edge; `edge` collects max(|df.x|, |df.y|) over the points of a lane so that the caller
// can test all of them at once (a 3-input FMNMX per point instead of compare + select).
//   px2 = (px, py), py2 = (-py, px);  tuv = (tu, tv), cs = (ck, ck), sc = (sk, sk) of the candidate (scalar operands of the
//   packed instructions): (u, v) = ck (px, py) + sk (-py, px) + (tu, tv).
// Three shared-memory loads: grid entry (2 bytes), record (16 + 4 bytes).
__device__ __forceinline__ float screen_point(const ScreenCtx& m, const float2 px2, const float2 py2, const float2 tuv, const float2 cs,
                                              const float2 sc, float& edge) {
  const float2 uv = __ffma2_rn(px2, cs, __ffma2_rn(py2, sc, tuv));             // cell coordinates - 0.5
  const float2 t2 = __fadd2_rn(uv, make_float2(kScreenMagic, kScreenMagic));   // round to nearest = floor of the cell coordinate, |uv| < 2^22
  const float2 fl = __fadd2_rn(t2, make_float2(-kScreenMagic, -kScreenMagic));
  const float2 df = __fadd2_rn(uv, make_float2(-fl.x, -fl.y));                 // exact: position in the cell - 0.5
  // ix + gw*iy - (first cell of the strip) with ix = bits(t) - magic bits: the constants are folded into `nbase`;
  // min(g + nbase, span) is one VIADDMNMX
  const unsigned g = __float_as_uint(t2.x) + static_cast<unsigned>(m.gw) * __float_as_uint(t2.y);
  const unsigned ra = m.grid[__viaddmin_u32(g, m.nbase, m.span)];
  const float4 l = lds_f4(ra);        // l00, l11, c0, c1
  const float l10 = lds_f1_128(ra);
  const float2 zz = __ffma2_rn(make_float2(l.x, l.y), df, make_float2(l.z, l.w));  // l00 df0 + c0, l11 df1 + c1 (= z1)
  const float z0 = fmaf(l10, df.y, zz.x);
  const float xe = fmaf(-z0, z0, fmaf(-zz.y, zz.y, m.kappa2));  // z1's square first: it does not wait for l10
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(xe));                       // 2 ulp, results below 2^-126 flushed: covered by the total's slack
  edge = fmaxf(fmaxf(edge, fabsf(df.x)), fabsf(df.y));
  return e;
}

// warp sums of JB per-lane accumulators, packed like packed_warp_sum (same slot assignment)
template <int JB>
__device__ __forceinline__ float packed_warp_sum_f(const float (&a)[JB], int lane);
template <>
__device__ __forceinline__ float packed_warp_sum_f<1>(const float (&a)[1], int) {
  float k = a[0];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) k += __shfl_xor_sync(0xffffffffu, k, off);
  return k;
}
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
// (FULL: all JB candidates exist, j + JB <= hi: no clamping of the pose index, no test before the store)
template <int NPT, int JB, bool FULL>
__device__ __forceinline__ void screen_batch(const ScreenCtx& m, const ScreenPts<NPT>& p, const float4* pose32, float* lbpart, int j, int hi,
                                             int NW, int warp, int lane) {
  float acc[JB];
#pragma unroll
  for (int b = 0; b < JB; ++b) {
    const float4 ps = pose32[FULL ? j + b : min(j + b, hi - 1)];
    const float2 tuv = make_float2(ps.x, ps.y), cs = make_float2(ps.z, ps.z), sc = make_float2(ps.w, ps.w);
    float edge = 0.f;
    acc[b] = screen_point(m, p.px2[0], p.py2[0], tuv, cs, sc, edge) * p.w[0];  // the weight drops padding slots (copies of a scan point)
#pragma unroll
    for (int k = 1; k < NPT; ++k) acc[b] = fmaf(screen_point(m, p.px2[k], p.py2[k], tuv, cs, sc, edge), p.w[k], acc[b]);
    // a point within beta of a cell edge counts as the worst case, exp(.) = 1: if any of this lane's points is, all of them do
    // (each term is <= 1 and the terms already added are >= 0, so adding the lane's point count keeps an upper bound)
    acc[b] = fmaf(edge > m.beta_c ? 1.f : 0.f, p.wsum, acc[b]);  // FSET.BF + FFMA
  }
  const float tot = packed_warp_sum_f<JB>(acc, lane);
  const int jj = j + packed_slot<JB>(lane);
  if (packed_writer<JB>(lane) && (FULL || jj < hi)) lbpart[jj * NW + warp] = tot;
}

// the per-lane sums of a full batch of JB candidates (no reduction)
template <int NPT, int JB>
__device__ __forceinline__ void screen_batch_acc(const ScreenCtx& m, const ScreenPts<NPT>& p, const float4* pose32, int j, float (&acc)[JB]) {
#pragma unroll
  for (int b = 0; b < JB; ++b) {
    const float4 ps = pose32[j + b];
    const float2 tuv = make_float2(ps.x, ps.y), cs = make_float2(ps.z, ps.z), sc = make_float2(ps.w, ps.w);
    float edge = 0.f;
    acc[b] = screen_point(m, p.px2[0], p.py2[0], tuv, cs, sc, edge) * p.w[0];
#pragma unroll
    for (int k = 1; k < NPT; ++k) acc[b] = fmaf(screen_point(m, p.px2[k], p.py2[k], tuv, cs, sc, edge), p.w[k], acc[b]);
    acc[b] = fmaf(edge > m.beta_c ? 1.f : 0.f, p.wsum, acc[b]);
  }
}

// the screen over candidates [lo, hi): NDTPSO_SCREEN_JB at a time, then the rest in pairs
template <int NPT>
__device__ __forceinline__ void screen_candidates(const ScreenCtx& m, const ScreenPts<NPT>& p, const float4* pose32, float* lbpart, int lo, int hi,
                                                  int NW, int warp, int lane) {
  int j = lo;
#if NDTPSO_SCREEN_DEFER
  // the warp reduction of a batch (a chain of shuffles with no load in flight) is deferred by one batch, so that it overlaps
  // the next batch's evaluations
  constexpr int JB = NDTPSO_SCREEN_JB;
  if (j + JB <= hi) {
    float prev[JB];
    screen_batch_acc<NPT, JB>(m, p, pose32, j, prev);
    int jp = j;
    for (j += JB; j + JB <= hi; j += JB) {
      float acc[JB];
      screen_batch_acc<NPT, JB>(m, p, pose32, j, acc);
      const float tot = packed_warp_sum_f<JB>(prev, lane);
      if (packed_writer<JB>(lane)) lbpart[(jp + packed_slot<JB>(lane)) * NW + warp] = tot;
#pragma unroll
      for (int b = 0; b < JB; ++b) prev[b] = acc[b];
      jp = j;
    }
    const float tot = packed_warp_sum_f<JB>(prev, lane);
    if (packed_writer<JB>(lane)) lbpart[(jp + packed_slot<JB>(lane)) * NW + warp] = tot;
  }
#else
  for (; j + NDTPSO_SCREEN_JB <= hi; j += NDTPSO_SCREEN_JB) screen_batch<NPT, NDTPSO_SCREEN_JB, true>(m, p, pose32, lbpart, j, hi, NW, warp, lane);
#endif
#if NDTPSO_SCREEN_JB > 4
  for (; j + 4 <= hi; j += 4) screen_batch<NPT, 4, true>(m, p, pose32, lbpart, j, hi, NW, warp, lane);
#endif
  for (; j < hi; j += 2) screen_batch<NPT, 2, false>(m, p, pose32, lbpart, j, hi, NW, warp, lane);
}

// this lane's fp32 copies of its scan points (pt[k] = point k*T + tid of the scan, padding = (1e200, 0)); `first` = the scan's
// first point, which stands in for the points of a lane that has none
template <int NPT>
__device__ __forceinline__ void screen_points(const double2 (&pt)[NPT], const double2 first, ScreenPts<NPT>& p) {
  const bool none = pt[0].x > 1e199;
  const double2 own = none ? first : pt[0];
#pragma unroll
  for (int k = 0; k < NPT; ++k) {
    const bool pad = pt[k].x > 1e199;
    const double2 q = pad ? own : pt[k];
    const float fx = static_cast<float>(q.x), fy = static_cast<float>(q.y);
    p.px2[k] = make_float2(fx, fy);
    p.py2[k] = make_float2(-fy, fx);
    p.w[k] = pad ? 0.f : 1.f;
  }
  p.wsum = 0.f;
#pragma unroll
  for (int k = 0; k < NPT; ++k) p.wsum += p.w[k];
}

// fp64 evaluation of the candidates listed in surv[i .. i+JB-1] (clamped to the last entry)
template <int NPT, int JB, bool FAST_GEOM, int VAR, int GSHIFT>
__device__ __forceinline__ void score_batch_listed(const SliceCtx& m, const double2 (&pt)[NPT], const Pose* pose, double* wpart,
                                                   const unsigned short* surv, int i, int ns, int NW, int warp, int lane) {
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
    for (int k = 0; k < NPT; ++k) slice_point<FAST_GEOM, VAR, GSHIFT>(m, pt[k], txy[b].x, txy[b].y, cs[b].x, cs[b].y, acc[b]);
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

// two batches of JB candidates at once (cluster form: one CTA per SM, registers to spare, few warps to hide the fp64 latency):
// 2 JB NPT independent evaluations in flight per lane, the two packed reductions interleave
template <int NPT, int JB, bool FAST_GEOM, int VAR>
__device__ __forceinline__ void score_batch_pair(const SliceCtx& m, const double2 (&pt)[NPT], const Pose* pose, double* wpart, int j, int NW,
                                                 int warp, int lane) {
  double acc0[JB], acc1[JB];
#pragma unroll
  for (int b = 0; b < JB; ++b) {
    const Pose* p0 = pose + j + b;
    const Pose* p1 = pose + j + JB + b;
    const double2 t0 = *reinterpret_cast<const double2*>(&p0->x), c0 = *reinterpret_cast<const double2*>(&p0->c);
    const double2 t1 = *reinterpret_cast<const double2*>(&p1->x), c1 = *reinterpret_cast<const double2*>(&p1->c);
    acc0[b] = 0.;
    acc1[b] = 0.;
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
      slice_point<FAST_GEOM, VAR>(m, pt[k], t0.x, t0.y, c0.x, c0.y, acc0[b]);
      slice_point<FAST_GEOM, VAR>(m, pt[k], t1.x, t1.y, c1.x, c1.y, acc1[b]);
    }
  }
  const double tot0 = packed_warp_sum<JB>(acc0, lane);
  const double tot1 = packed_warp_sum<JB>(acc1, lane);
  if (packed_writer<JB>(lane)) {
    const int jj = j + packed_slot<JB>(lane);
    wpart[jj * NW + warp] = tot0;
    wpart[(jj + JB) * NW + warp] = tot1;
  }
}

template <int NPT, int JB, int CL, bool FAST_GEOM, int VAR>
__device__ __forceinline__ void score_candidates(const SliceCtx& m, const double2 (&pt)[NPT], const Pose* pose, double* wpart, int lo,
                                                 int hi, const Topo& tp, int warp, int lane) {
  if (CL == 1 && JB == 4) {
    // whole batches of 4, then the 1..3 left over in batches of 2: a swarm of 70 costs 70 evaluations, not 72
    int j = lo;
    for (; j + 4 <= hi; j += 4) score_batch<NPT, 4, FAST_GEOM, VAR>(m, pt, pose, wpart, j, hi, tp.NW, warp, lane);
    for (; j < hi; j += 2) score_batch<NPT, 2, FAST_GEOM, VAR>(m, pt, pose, wpart, j, hi, tp.NW, warp, lane);
  } else if (CL > 1) {
    // cluster form: group g scores an equal, contiguous share of the pending candidates (70 on 4 groups = 18 + 18 + 18 + 16), whole
    // batches of JB first and the rest in batches of 4 and 2: every group's critical path is the same two-and-a-bit batches
    // (dealing whole batches of 8 round-robin gave one group three batches and the others two, and all waited for the one)
    const int q = (hi - lo + tp.G - 1) / tp.G;
    int j = lo + tp.g * q;
    const int ghi = min(j + q, hi);
    for (; j + 2 * JB <= ghi; j += 2 * JB) score_batch_pair<NPT, JB, FAST_GEOM, VAR>(m, pt, pose, wpart, j, tp.NW, warp, lane);
    for (; j + JB <= ghi; j += JB) score_batch<NPT, JB, FAST_GEOM, VAR>(m, pt, pose, wpart, j, ghi, tp.NW, warp, lane);
    if (JB > 4)
      for (; j + 4 <= ghi; j += 4) score_batch<NPT, 4, FAST_GEOM, VAR>(m, pt, pose, wpart, j, ghi, tp.NW, warp, lane);
    if (JB > 2)
      for (; j < ghi; j += 2) score_batch<NPT, 2, FAST_GEOM, VAR>(m, pt, pose, wpart, j, ghi, tp.NW, warp, lane);
    else
      for (; j < ghi; j += JB) score_batch<NPT, JB, FAST_GEOM, VAR>(m, pt, pose, wpart, j, ghi, tp.NW, warp, lane);
  } else {
    for (int j = lo; j < hi; j += JB) score_batch<NPT, JB, FAST_GEOM, VAR>(m, pt, pose, wpart, j, hi, tp.NW, warp, lane);
  }
}

// Cluster form, after phase B: sum the NW warp partials of every candidate this CTA scored and
// store the sum into cpart[j*S + s] of every CTA of the cluster; one (candidate, destination) pair
// per thread.  Ends with the cluster barrier that makes all partials visible everywhere.
template <int JB, int CL>
__device__ __forceinline__ void exchange_partials(const double* wpart, double* cpart, int lo, int hi, const Topo& tp) {
  __syncthreads();
  const int q = (hi - lo + tp.G - 1) / tp.G;  // my group's share: see score_candidates
  const int glo = lo + tp.g * q;
  const int mine = max(min(glo + q, hi) - glo, 0);
  for (int idx = threadIdx.x; idx < mine * CL; idx += blockDim.x) {
    const int jl = idx / CL, dest = idx - jl * CL;
    const int j = glo + jl;
    const double* p = wpart + j * tp.NW;
    double c = 0.;
    for (int w = 0; w < tp.NW; ++w) c += p[w];
    st_cluster_f64(cpart + j * tp.S + tp.s, dest, c);
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

// ---- pieces shared by the two bodies ----------------------------------------------------------------
// initial swarm: task 0 = the seed particle (core.cpp:53,58), task 1+j = particle j (core.cpp:60-61)
__device__ __forceinline__ void init_candidates(const DevProblem& pr, const int* __restrict__ rnd, int P, Pose* pose0) {
  for (int t = threadIdx.x; t < P + 1; t += blockDim.x) {
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
}

// owner-private state of particle j from its initial candidate
__device__ __forceinline__ void init_particles(const SlicedSmem& sm, int P, const Pose* pose0) {
  for (int j = threadIdx.x; j < P; j += blockDim.x) {
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
}

// velocity/position update of particle j against gbest (core.cpp:83-90, no contraction); candidate -> pose[j], velocity -> vnew
__device__ __forceinline__ Pose update_particle(const SlicedSmem& sm, const PsoParams& prm, const double* ucoef, int j, double w, double gb0,
                                                double gb1, double gb2) {
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
  sm.vnew[3 * j] = nv[0];
  sm.vnew[3 * j + 1] = nv[1];
  sm.vnew[3 * j + 2] = nv[2];
  return Pose{nx[0], nx[1], c, s, nx[2], 0.};
}

// particle j takes its candidate (core.cpp:89-96); cj = the candidate's cost, or anything >= pbest when it is known not to improve
__device__ __forceinline__ void commit_particle(const SlicedSmem& sm, int j, const Pose& ps, double cj) {
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

// Generic body: one CTA or a cluster per problem, every pending candidate evaluated in fp64.
template <int NPT, int JB, int CL, bool FAST_GEOM, int VAR>
__device__ __forceinline__ void sliced_body(const SliceCtx& m, const double2 (&pt)[NPT], const DevProblem& pr, const PsoParams& prm,
                                            const SlicedSmem& sm, const Topo& tp, double* __restrict__ out, int* __restrict__ stats) {
  const int tid = threadIdx.x, T = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int P = prm.P, I = prm.I, PW = tp.PW;
  const int* __restrict__ rnd = pr.rnd;
  NDTPSO_PHASE_DECL
  Pose* pose0 = sm.pose;
  Pose* pose1 = sm.pose + (P + 1);
  // CL == 1: phase C sums the NW warp partials directly (one buffer: see sliced_swarm_smem_bytes).
  // CL > 1 : warps write wpart (single buffer), exchange_partials() fills the double-buffered cpart.
  double* part0 = sm.partial;
  double* part1 = CL == 1 ? sm.partial : sm.partial + (size_t)(P + 1) * PW;
  double* wpart = sm.wpart;
  // the 6P velocity coefficients of iteration `iter` -> ubuf[iter & 1]; issued right before a scoring
  // phase so that the global-memory latency and the conversions hide behind it
  auto prefetch_draws = [&](int iter) {
    if (iter >= I) return;
    double* dst = sm.ubuf + (iter & 1) * 6 * P;
    const int* src = rnd + 3 + 3 * P + 6 * P * iter;
    for (int i = tid; i < 6 * P; i += T) dst[i] = fabs(unit_random(src[i]));  // Array2d::Random().abs(), core.cpp:84
  };

  init_candidates(pr, rnd, P, pose0);
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
  init_particles(sm, P, pose0);

  // ---- iterations
  int it = 0, start = 0, par = 1, rounds = 0, n_gb = 0, n_f64 = P + 1;
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
    for (int j = ja; j < lim; j += T) pose[j] = update_particle(sm, prm, ucoef, j, w, gb0, gb1, gb2);
    __syncthreads();
    NDTPSO_PHASE_MARK(1)
    if (start == 0) prefetch_draws(it + 1);  // first round of an iteration
    n_f64 += lim - start;
    score_candidates<NPT, JB, CL, FAST_GEOM, VAR>(m, pt, pose, CL == 1 ? part : wpart, start, lim, tp, warp, lane);  // phase B
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
      const double cj = (j < lim) ? candidate_total(part, j, PW) : 0.;
      const bool imp = (j < lim) && (cj < gbc);
      const unsigned mask = __ballot_sync(0xffffffffu, imp);
      if (mask) {
        const int src = __ffs(mask) - 1;
        jstar = base + src;
        cstar = __shfl_sync(0xffffffffu, cj, src);
      }
    }
    const int end = (jstar >= 0) ? jstar + 1 : lim;
    for (int j = ja; j < end; j += T) commit_particle(sm, j, pose[j], candidate_total(part, j, PW));  // own particles in [start, end)
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
      stats[3] = 0;
    }
  }
  if (CL > 1) cluster_barrier();  // no CTA may exit while peers can still store into its shared memory
}

// the loop-invariant operands of the fp64 point evaluation, re-read from shared memory (see sliced_prologue): the screened
// body fetches them at the start of every fp64 phase instead of carrying ~30 registers through the screen's loop
__device__ __forceinline__ void load_slice_ctx(const SlicedSmem& sm, const ScreenCtx& sc, const double* rec, int base, int n_rec, SliceCtx& m) {
  const volatile double* cst = reinterpret_cast<const volatile double*>(sm.cst);
  const volatile double* geo = cst + 8 + 64;
  m.grid = sc.grid;
  m.rec = rec;
  m.base = base;
  m.etab = sm.etab;
  m.x_max = geo[0];
  m.y_max = geo[1];
  m.x_min = -m.x_max;
  m.y_min = -m.y_max;
  m.inv_cs = geo[2];
  m.hw_s = geo[3];
  m.hh_s = geo[4];
  m.hw = m.hh = m.cs = 0.;  // the fast geometry does not use them
  m.l2e = cst[8 + (threadIdx.x & 31)];  // per-lane copies: a lane-dependent address keeps them in vector registers
  m.c7 = cst[40 + (threadIdx.x & 31)];
  m.ln2hi = cst[1];
  m.ln2lo = cst[2];
  m.c6 = cst[4];
  m.c5 = cst[5];
  m.c4 = cst[6];
  m.c3 = cst[7];
  m.gw = sc.gw;
  m.span = static_cast<int>(sc.span);
  m.null_id = n_rec;
  m.goff = sc.rec32;
}

#ifndef NDTPSO_B2_RELOAD
#define NDTPSO_B2_RELOAD 1  // the fp64 phase re-reads its points (L2) and constants (shared memory) every round instead of holding them in registers through the screen
#endif

// Screened body (one CTA per problem, fast geometry): every pending candidate is first bounded in fp32 (phase B1); each warp
// then lists, for itself, the candidates the bound does not rule out and evaluates them in fp64 on its slice (phase B2);
// phase C looks at the survivors only.  Three barriers per round: after A, after B1, after B2.
template <int NPT, int VAR>
__device__ __forceinline__ void sliced_body_screened(const ScreenCtx& sc, const SliceCtx& m0, const double2 (&pt0)[NPT], const DevProblem& pr,
                                                     const PsoParams& prm, const SlicedSmem& sm, const Topo& tp, int n_rec,
                                                     double* __restrict__ out, int* __restrict__ stats) {
  const int tid = threadIdx.x, T = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int P = prm.P, I = prm.I, NW = tp.NW;
  const int* __restrict__ rnd = pr.rnd;
  NDTPSO_PHASE_DECL
  Pose* pose0 = sm.pose;
  Pose* pose1 = sm.pose + (P + 1);
  double* part = sm.partial;
  unsigned short* mysurv = sm.wsurv + warp * (P + 2);
  const double inv_cs = m0.inv_cs, off_u = m0.hw_s - 0.5, off_v = m0.hh_s - 0.5;
  auto prefetch_draws = [&](int iter) {
    if (iter >= I) return;
    double* dst = sm.ubuf + (iter & 1) * 6 * P;
    const int* src = rnd + 3 + 3 * P + 6 * P * iter;
    for (int i = tid; i < 6 * P; i += T) dst[i] = fabs(unit_random(src[i]));  // Array2d::Random().abs(), core.cpp:84
  };
  // fp64 evaluation of the candidates in mysurv[0 .. ns) on this warp's slice -> part[j*NW + warp]
  auto score_listed = [&](const SliceCtx& m, const double2 (&pt)[NPT], const Pose* pose, int ns) {
    int i = 0;
    for (; i + 4 <= ns; i += 4) score_batch_listed<NPT, 4, true, VAR, 5>(m, pt, pose, part, mysurv, i, ns, NW, warp, lane);
    for (; i + 2 <= ns; i += 2) score_batch_listed<NPT, 2, true, VAR, 5>(m, pt, pose, part, mysurv, i, ns, NW, warp, lane);
    if (i < ns) score_batch_listed<NPT, 1, true, VAR, 5>(m, pt, pose, part, mysurv, i, ns, NW, warp, lane);  // survivors are few: no padding
  };

  init_candidates(pr, rnd, P, pose0);
  for (int t = lane; t < P + 1; t += 32) mysurv[t] = static_cast<unsigned short>(t);  // the initial swarm is evaluated in full
  __syncthreads();
  prefetch_draws(0);
  score_listed(m0, pt0, pose0, P + 1);
  __syncthreads();
  for (int t = tid; t < P + 1; t += T) sm.cost0[t] = candidate_total(part, t, NW);
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
  init_particles(sm, P, pose0);
  ScreenPts<NPT> sp;
  screen_points<NPT>(pt0, pr.n_pts > 0 ? pr.pts[0] : make_double2(0., 0.), sp);

  int it = 0, start = 0, par = 1, rounds = 0, n_gb = 0, n_f64 = P + 1, n_scr = 0;
  int hot = prm.hot_chunk > 0 ? 1 : 0, imp_it = 0, win = prm.hot_chunk;  // speculation window: see sliced_body
  double w = prm.w;
  NDTPSO_PHASE_MARK(0)
  while (it < I) {
    Pose* pose = par ? pose1 : pose0;
    const int lim = win > 0 ? min(P, start + win) : P;  // this round covers particles [start, lim)
    // phase A: owners of the pending particles [start, lim)
    const int ja = start + ((tid - start) % T + T) % T;  // first pending particle owned by this thread
    const double* ucoef = sm.ubuf + (it & 1) * 6 * P;
    for (int j = ja; j < lim; j += T) {
      const Pose ps = update_particle(sm, prm, ucoef, j, w, gb0, gb1, gb2);
      pose[j] = ps;
      store_pose32(sm.pose32, j, ps.x, ps.y, ps.c, ps.s, inv_cs, off_u, off_v);
    }
    __syncthreads();
    NDTPSO_PHASE_MARK(1)
    if (start == 0) prefetch_draws(it + 1);  // first round of an iteration
    // phase B1: fp32 lower bound of every pending candidate's cost on this warp's slice
    screen_candidates<NPT>(sc, sp, sm.pose32, sm.lbpart, start, lim, NW, warp, lane);
#if NDTPSO_B2_RELOAD
    // the fp64 copies of this thread's points for phase B2: requested now, they arrive while the barrier and the list pass
    double2 pt[NPT];
#pragma unroll
    for (int k = 0; k < NPT; ++k) {
      const int i = k * T + tid;
      pt[k] = make_double2(1e200, 0.);  // padding: out of bounds for every pose
      if (i < pr.n_pts) asm volatile("ld.global.v2.f64 {%0, %1}, [%2];" : "=d"(pt[k].x), "=d"(pt[k].y) : "l"(pr.pts + i));
    }
#endif
    __syncthreads();
    NDTPSO_PHASE_MARK(5)
    // every warp lists, for itself, the candidates whose bound does not already rule out an improvement of their particle's
    // best (core.cpp:94); no barrier needed before phase B2.  The others are marked with a cost of +1e300 in their first
    // partial (by warp 0), which phase C's commit treats like any cost that improves nothing.
    int ns = 0;
    for (int base = start; base < lim; base += 32) {
      const int j = base + lane;
      bool alive = false;
      if (j < lim) {
        const float* row = sm.lbpart + j * NW;
        float u0 = 0.f, u1 = 0.f, u2 = 0.f, u3 = 0.f;
        int w2 = 0;
        if ((NW & 3) == 0) {  // rows of 16-byte multiples: vector loads
          for (; w2 < NW; w2 += 4) {
            const float4 q = *reinterpret_cast<const float4*>(row + w2);
            u0 += q.x;
            u1 += q.y;
            u2 += q.z;
            u3 += q.w;
          }
        }
        for (; w2 < NW; ++w2) u0 += row[w2];
        const double u = static_cast<double>((u0 + u1) + (u2 + u3));
        const double lower = -(u * (1. + 6.103515625e-5)) - 1e-6;  // slack: ex2.approx, fp32 sums (2^-14 in all), flushed denormals
        alive = !(lower >= sm.pbc[j]);
        if (!alive && warp == 0) part[j * NW] = 1e300;
      }
      const unsigned mask = __ballot_sync(0xffffffffu, alive);
      if (alive) mysurv[ns + __popc(mask & ((1u << lane) - 1u))] = static_cast<unsigned short>(j);
      ns += __popc(mask);
    }
    __syncwarp();
    NDTPSO_PHASE_MARK(6)
    // phase B2: the fp64 evaluation of the survivors
    n_f64 += ns;
    n_scr += (lim - start) - ns;
    if (ns > 0) {
#if NDTPSO_B2_RELOAD
      SliceCtx m;
      load_slice_ctx(sm, sc, m0.rec, m0.base, n_rec, m);
      score_listed(m, pt, pose, ns);
#else
      score_listed(m0, pt0, pose, ns);
#endif
    }
    NDTPSO_PHASE_MARK(2)
    __syncthreads();
    NDTPSO_PHASE_MARK(3)
    // phase C: j* = first pending particle that improves gbest (core.cpp:98); only survivors can (gbest <= every pbest)
    int jstar = -1;
    double cstar = 0.;
    for (int base = 0; base < ns && jstar < 0; base += 32) {
      const int i = base + lane;
      const int j = (i < ns) ? mysurv[i] : 0;
      const double cj = (i < ns) ? candidate_total(part, j, NW) : 0.;
      const bool imp = (i < ns) && (cj < gbc);
      const unsigned mask = __ballot_sync(0xffffffffu, imp);
      if (mask) {
        const int src = __ffs(mask) - 1;  // the list is ascending: the first survivor that improves is the first particle that does
        jstar = __shfl_sync(0xffffffffu, j, src);
        cstar = __shfl_sync(0xffffffffu, cj, src);
      }
    }
    const int end = (jstar >= 0) ? jstar + 1 : lim;
    for (int j = ja; j < end; j += T) commit_particle(sm, j, pose[j], candidate_cost(part, j, NW));  // own particles in [start, end)
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

  if (tid == 0) {
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
}

// What the screen's record of one built cell is made of (derivation (3), (4) above "struct ScreenCtx"), in fp64:
// q = the cell's fp64 record {mu_x, mu_y, -S00/2, -S01/2, -S10/2, -S11/2}, cell = its flat index, du = the coordinate error bound.
struct ScreenRec {
  double l00, l10, l11;  // Cholesky factor of H = (Sigma^-1 / 2) cs^2, shrunk by its own rounding
  double ox, oy;         // the mean's offset from the cell's centre, in cell sides
  double kappa0;         // (ez_0^2 + ez_1^2) + hs delta^2: the absolute terms of the exponent's bound, before the division by t
};
__device__ __forceinline__ ScreenRec screen_record(const double* q, int cell, const DevMap& mp, double du) {
  ScreenRec r;
  const double u24 = 5.9604644775390625e-08;
  const double cs2 = mp.cs * mp.cs;
  const double H00 = -q[2] * cs2, H01 = -q[3] * cs2, H11 = -q[5] * cs2;  // (Sigma^-1/2) in cell units
  r.ox = (q[0] + mp.hw) * mp.inv_cs - ((cell % mp.gw) + 0.5);
  r.oy = (q[1] + mp.hh) * mp.inv_cs - ((cell / mp.gw) + 0.5);
  const double dl0 = du + u24 * (2. * fabs(r.ox) + 0.51), dl1 = du + u24 * (2. * fabs(r.oy) + 0.51), dl = fmax(dl0, dl1);
  const double dm0 = fmax(fabs(-0.5 - r.ox), fabs(0.5 - r.ox)) + dl0, dm1 = fmax(fabs(-0.5 - r.oy), fabs(0.5 - r.oy)) + dl1;
  const double sh = 1. - 1e-12;
  r.l00 = H00 > 0. ? sqrt(H00) * sh : 0.;
  r.l10 = r.l00 > 0. ? H01 / r.l00 * sh : 0.;
  r.l11 = sqrt(fmax(H11 - r.l10 * r.l10 - 4e-15 * H11, 0.)) * sh;
  const double ez0 = 4. * u24 * (r.l00 * dm0 + fabs(r.l10) * dm1), ez1 = 4. * u24 * r.l11 * dm1;
  const double hs = H00 + 2. * fabs(H01) + H11;
  r.kappa0 = ez0 * ez0 + ez1 * ez1 + hs * dl * dl;
  return r;
}
// The table's kappa2 from the sum of its records' kappa0 (derivation (4)), rounded up to fp32; any positive finite value is
// valid (the records adapt their t to it), this one balances the two loosenings.
__device__ __forceinline__ float screen_kappa2(double k0sum, int n_rec) {
  const double mean = n_rec > 0 ? k0sum / n_rec : 0.;
  double k = 1.4426950408889634 * sqrt(mean);
  if (!(k >= 9.5367431640625e-07)) k = 9.5367431640625e-07;  // never below 2^-20 (and not NaN)
  if (!(k <= 0.25)) k = 0.25;                                  // a table of needles: most records then take the trivial bound
  return __double2float_ru(k * 1.000001);
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
  const int rec32_bytes = screen ? screen_rec_bytes(n_rec + 1) : 0;
  const SlicedSmem sm = carve_sliced(smem_raw, P, tp.PW, tp.CL == 1 ? 0 : tp.NW, rec32_bytes + grid_bytes + rec_bytes, screen);
  unsigned char* s_rec32 = sm.table;
  unsigned short* s_grid = reinterpret_cast<unsigned short*>(sm.table + rec32_bytes);
  unsigned char* s_rec = sm.table + rec32_bytes + grid_bytes;

  if (tid < kExpTableSize) sm.etab[tid] = c_exp_table[tid];
  // constants go through volatile shared memory so the compiler keeps them in registers instead of
  // re-materialising 64-bit immediates inside the loop
  volatile double* cst = reinterpret_cast<volatile double*>(sm.cst);
  if (tid < 32) {
    const ExpConsts ec = exp_consts();
    sm.cst[8 + tid] = ec.l2e;
    sm.cst[40 + tid] = ec.c7;
  }
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
    volatile double* geo = cst + 8 + 64;  // what load_slice_ctx re-reads
    geo[0] = mp.x_max;
    geo[1] = mp.y_max;
    geo[2] = mp.inv_cs;
    geo[3] = mp.hw * mp.inv_cs;
    geo[4] = mp.hh * mp.inv_cs;
    mbar_init(sm.bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(sm.bar, rec_bytes + grid_bytes);
    tma_load_1d(s_rec, mp.rec, rec_bytes, sm.bar);
    tma_load_1d(s_grid, mp.grid, grid_bytes, sm.bar);
  }
  // slice s of the scan: points k*(S*T) + s*T + tid
#pragma unroll
  for (int k = 0; k < NPT; ++k) {
    const int i = (k * tp.S + tp.s) * T + tid;
    pt[k] = (i < pr.n_pts) ? pr.pts[i] : make_double2(1e200, 0.);  // padding: out of bounds for every pose
  }
  m.rec = reinterpret_cast<const double*>(s_rec);
  m.grid = s_grid;
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
  {  // per-lane copies: a lane-dependent address is not uniform, so the two constants stay in vector registers
    volatile double* lanes = reinterpret_cast<volatile double*>(sm.cst + 8);
    m.l2e = lanes[tid & 31];
    m.c7 = lanes[32 + (tid & 31)];
  }
  m.ln2hi = cst[1];
  m.ln2lo = cst[2];
  m.c6 = cst[4];
  m.c5 = cst[5];
  m.c4 = cst[6];
  m.c3 = cst[7];
  m.gw = mp.gw;
  m.base = row0 * mp.gw;
  m.span = span;
  m.null_id = n_rec;
  m.goff = 0u;
  mbar_wait(sm.bar, 0);
  if (screen) {
    // fp32 records of the screen (derivation above "struct ScreenCtx"): one per built cell, found through the grid so that the
    // cell is known; the grid entry is then replaced by the record's shared address.  Every thread owns whole grid entries, and
    // the body has a __syncthreads before anyone else reads them.
    const double* rec = reinterpret_cast<const double*>(s_rec);
    const unsigned rec32_addr = smem_u32(s_rec32);  // sliced_screen_fits() made sure every record's address is below 2^16
    m.goff = rec32_addr;
    const double du = static_cast<double>(prm->scr_du);
    const int cell0 = row0 * mp.gw;
    // pass 1: what every record needs of the exponent's additive term (kappa0_r), its maximum and its mean -> kappa2
    double k0sum = 0.;
    for (int g = tid; g < span; g += T) {
      const int r = s_grid[g];
      if (r == n_rec) continue;
      const double k0 = screen_record(rec + 6 * r, g + cell0, mp, du).kappa0;
      k0sum += k0 < 1e300 ? k0 : 1e300;  // a non-finite kappa0 (its record takes the trivial bound) must not poison the mean
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) k0sum += __shfl_xor_sync(0xffffffffu, k0sum, off);  // fixed trees: the same kappa2 in every run
    double* red = sm.partial;  // NW doubles: the swarm arrays are free until the body starts
    if ((tid & 31) == 0) red[tid >> 5] = k0sum;
    __syncthreads();
    k0sum = 0.;
    for (int w = 0; w < tp.NW; ++w) k0sum += red[w];
    __syncthreads();  // `red` is read before anything else may write sm.partial
    const float kappa2 = screen_kappa2(k0sum, n_rec);
    const double kd = static_cast<double>(kappa2) * (1. - 2.384185791015625e-07);
    // pass 2: the records
    for (int g = tid; g <= span; g += T) {
      const int r = (g < span) ? s_grid[g] : n_rec;
      s_grid[g] = static_cast<unsigned short>(rec32_addr + screen_rec_offset(r));
      if (g < span && r == n_rec) continue;  // unbuilt cell: points at the null record
      float* o = reinterpret_cast<float*>(s_rec32 + screen_rec_offset(r));  // o[0..3] = first part, o[32] = second part (128 bytes on)
      if (r == n_rec) {  // the null record: z0 = 1e18 whatever the point
        o[0] = 0.f;
        o[1] = 0.f;
        o[2] = 1e18f;
        o[3] = 0.f;
        o[32] = 0.f;
        continue;
      }
      const ScreenRec sr = screen_record(rec + 6 * r, g + cell0, mp, du);
      const double t = 1.000001 * 1.4426950408889634 * sr.kappa0 / kd;
      // !(t <= 1/2): this record needs more than the table's kappa2 gives (or is not finite): the trivial bound, e = 2^kappa2
      const double scale = (t <= 0.5) ? sqrt((1. - t) * (1. - t) * (1. - 2.384185791015625e-07) * 1.4426950408889634) * (1. - 1e-12) : 0.;
      // the factor's entries are rounded to fp32 here; that rounding is the first of the three ez counts per product
      // the offset is folded into the constants of the two rows: z0 = l00 df0 + l10 df1 + c0, z1 = l11 df1 + c1 (computed in fp64,
      // rounded once: the first of at most three roundings the term carries)
      o[0] = scale > 0. ? static_cast<float>(sr.l00 * scale) : 0.f;
      o[1] = scale > 0. ? static_cast<float>(sr.l11 * scale) : 0.f;
      o[2] = scale > 0. ? static_cast<float>(-(sr.l00 * sr.ox + sr.l10 * sr.oy) * scale) : 0.f;
      o[3] = scale > 0. ? static_cast<float>(-(sr.l11 * sr.oy) * scale) : 0.f;
      o[32] = scale > 0. ? static_cast<float>(sr.l10 * scale) : 0.f;
    }
    sc->kappa2 = kappa2;
    sc->rec32 = rec32_addr;
    sc->grid = s_grid;
    sc->beta_c = prm->scr_beta_c;
    sc->gw = mp.gw;
    sc->nbase = 0u - (static_cast<unsigned>(m.base) + static_cast<unsigned>(kScreenMagicBits) * (1u + static_cast<unsigned>(mp.gw)));  // folds the magic bits of both coordinates
    sc->span = static_cast<unsigned>(m.span);
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

// the screen's records start at a fixed offset of the dynamic shared memory and are addressed by 16-bit shared addresses
constexpr int kScreenMaxRecords = 1976;  // null record included: 1024 (room for the window's base) + kSlicedTableOffset + screen_rec_bytes(1976) < 2^16
__device__ __forceinline__ bool sliced_screen_fits(const unsigned char* smem_raw, int n_rec) {
  return smem_u32(smem_raw) + kSlicedTableOffset + static_cast<unsigned>(screen_rec_bytes(n_rec + 1)) <= 65536u;
}

// Host guarantees: every table is compact and symmetric and fits the dynamic shared memory;
// n_pts <= NPT * S * blockDim.x; the grid is n_problems * CL CTAs launched as clusters of CL; with prm.screen every table
// has fewer than kScreenMaxRecords built cells and the fast geometry.
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
  // shared memory is laid out for the screen whenever the launch asks for it; its records must sit below 2^16 in the shared window
  const int screen = (CL == 1 && mp.fast_geom && sliced_screen_fits(smem_raw, mp.hdr[HDR_NREC])) ? prm.screen : 0;
  const SlicedSmem sm = sliced_prologue<NPT>(smem_raw, pr, mp, prm.P, tp, m, pt, screen, &sc, &prm);
  if (CL > 1) cluster_barrier();  // every CTA's shared memory is carved before anyone stores into it
  double* o = out + 4 * (size_t)b;
  int* s = stats ? stats + kStatsWords * (size_t)b : nullptr;
  if (CL == 1 && screen)
    sliced_body_screened<NPT, kProdVariant>(sc, m, pt, pr, prm, sm, tp, mp.hdr[HDR_NREC], o, s);
  else if (mp.fast_geom)
    sliced_body<NPT, JB, CL, true, kProdVariant>(m, pt, pr, prm, sm, tp, o, s);
  else
    sliced_body<NPT, JB, CL, false, kProdVariant>(m, pt, pr, prm, sm, tp, o, s);
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
  const int tid = threadIdx.x, T = blockDim.x, warp = tid >> 5, lane = tid & 31;
  if (!mp.fast_geom || !sliced_screen_fits(smem_raw, mp.hdr[HDR_NREC])) {  // no screen for this table: the trivial bound
    for (int j = tid; j < m_poses; j += T) out[(size_t)b * m_poses + j] = -1e300;
    return;
  }
  const SlicedSmem sm = sliced_prologue<NPT>(smem_raw, pr, mp, prm.P, tp, m, pt, 1, &sc, &prm);
  for (int j = tid; j < m_poses; j += T) {
    const double* p = poses + 3 * ((size_t)b * m_poses + j);
    double s, c;
    sincos(p[2], &s, &c);
    store_pose32(sm.pose32, j, p[0], p[1], c, s, m.inv_cs, m.hw_s - 0.5, m.hh_s - 0.5);
  }
  __syncthreads();
  ScreenPts<NPT> sp;
  screen_points<NPT>(pt, pr.n_pts > 0 ? pr.pts[0] : make_double2(0., 0.), sp);
  screen_candidates<NPT>(sc, sp, sm.pose32, sm.lbpart, 0, m_poses, tp.NW, warp, lane);
  __syncthreads();
  for (int j = tid; j < m_poses; j += T) {
    float u = 0.f;
    for (int w2 = 0; w2 < tp.NW; ++w2) u += sm.lbpart[j * tp.NW + w2];
    out[(size_t)b * m_poses + j] = -(static_cast<double>(u) * (1. + 6.103515625e-5)) - 1e-6;  // the same slack as phase B1
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
