// Host side of libndtpso_b200.so: the C ABI of include/ndtpso_b200.h.
//
// A batch is laid out in ONE device allocation ("arena"):
//   [DevProblem n][DevMap M][points][per map: mean, inv_cov, built | cell_index][host rand streams]   <- uploaded, one H2D
//   [results n x 4][stats n x 2][per map: hdr, grid, records][device rand streams]                    <- device-only
// and mirrored (upload part only) in one pinned staging buffer.  Buffers are recycled through a
// small pool in the context, so a steady stream of align calls performs no cudaMalloc.
#include "../../include/ndtpso_b200.h"

#include <sched.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "ndtpso_kernels.cuh"
#include "ndtpso_pso_sliced.cuh"
#include "ndtpso_dframes.cuh"
#include "../../include/ndtpso_dframes.h"

using namespace ndtpso;

namespace {

constexpr int kMaxChunks = 8;  // pipeline chunks of one ndtpso_align_batch call
struct PoolBuf {
  void* ptr;
  size_t bytes;
};

constexpr size_t kAlign = 256;
constexpr int kDefaultHotChunk = 16;  // speculation window of the sliced kernel while gbest improves often (tools/chunk_sweep.py)
inline size_t align_up(size_t v, size_t a = kAlign) { return (v + a - 1) / a * a; }

}  // namespace

namespace {
class HostPool;
}

struct ndtpso_ctx {
  int device = 0;
  HostPool* pool = nullptr;  // this context's staging workers (created on first use)
  int opt_host_threads = 0;  // staging threads incl. the caller: 0 = auto (the cores this process may use, at most 8)
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  cudaStream_t chunk_stream[kMaxChunks] = {};  // pipelined align_batch
  cudaStream_t copy_stream = nullptr;  // uploads of batch k+1 overlap the kernels of batch k
  cudaStream_t pipe_stream[2] = {nullptr, nullptr};  // ndtpso_align_submit alternates between them (see there)
  unsigned pipe_next = 0;
  std::string err;
  int opt_warps = 0;
  int64_t opt_smem = 0;
  int opt_cluster = 0;
  int opt_kernel = 0;  // 0 auto, 1 warp-per-particle (generic), 2 point-sliced
  int opt_npt = 0;     // points per thread of the sliced kernel, 0 auto
  int opt_chunks = 0;      // pipelined align_batch: number of chunks (1 = off, 0 = auto: 3 from 128 problems on)
  int opt_screen = -1;     // fp32 screening of the sliced kernel: -1 / 1 on whenever the batch qualifies and its tables fit shared memory, 0 off
  int opt_hot_chunk = -1;  // speculation window of the sliced kernel while gbest improves often: -1 auto, 0 off
  int opt_cand_batch = 0;  // candidates scored together by the sliced kernel: 0 auto (largest), 1, 2, 4
  int64_t launches = 0;
  int64_t last_h2d = 0, last_d2h = 0;  // bytes moved by the most recent upload / results call
  int sm_count = 0;
  int clock_khz = 2000000;
  int64_t opt_exchange_timeout_ms = 10000;  // bounded wait for the peers' results
  int max_smem_optin = 0;
  std::vector<PoolBuf> dev_pool, pin_pool;
  int smem_attr_set[33] = {0};
  std::vector<const void*> sliced_attr_set;  // point-sliced kernel instances whose function attributes are set on this context's device
};

struct ndtpso_exchange {
  ndtpso_ctx* ctx = nullptr;
  int world = 1, rank = 0, n_per_rank = 0;
  unsigned char* base = nullptr;  // [2][world*n][4] fp64 | flags[world] u32 | done u32 | err i32
  size_t half_bytes = 0, o_flags = 0, o_done = 0, o_err = 0, bytes = 0;
  unsigned char* peer_base[NDTPSO_MAX_RANKS] = {};
  bool opened[NDTPSO_MAX_RANKS] = {};
  bool connected = false;
  unsigned epoch = 0;
  double* pin = nullptr;  // [world*n][4] read-back staging
};

struct ndtpso_batch {
  ndtpso_ctx* ctx = nullptr;
  int n = 0, n_maps = 0;
  PsoParams prm{};
  bool has_pso = false;
  bool any_device_rng = false;
  PoolBuf dev{nullptr, 0}, pin{nullptr, 0};
  size_t upload_bytes = 0;
  DevProblem* d_probs = nullptr;
  DevMap* d_maps = nullptr;
  double* d_out = nullptr;
  int* d_stats = nullptr;
  uint32_t* d_rng_state = nullptr;  // per-problem persistent rand() streams (device-resident frames), or nullptr
  ndtpso_exchange* ex = nullptr;    // results are also published to every rank's gathered buffer (multi-GPU)
  // what the fp32 screen's error bound needs (ndtpso_pso_sliced.cuh): largest |coordinate| of any scan point, largest frame
  // half-extent, largest 1/cell_side and grid width; scr_ok: every frame is whole cells of a power-of-two side, symmetric
  double scr_pmax = 0., scr_ext = 0., scr_inv_cs = 0.;
  int scr_gw = 0;
  bool scr_ok = true;
  int need_dyn_smem = 0;  // points + records + grid of the largest problem
  int max_pts = 0;        // largest scan
  int max_table_smem = 0; // records + grid of the largest table
  int max_n_rec = 0;      // built cells of the largest table
  bool all_compact = true;  // every table has <= 65534 built cells
  bool all_symmetric = true;  // every built cell: S01 == S10 bit for bit, finite, positive semi-definite (NDTCell::build always does)
  bool solved = false;
  bool results_enqueued = false;  // align_submit already queued the D2H of the results behind the kernels
  cudaEvent_t ev_done = nullptr;
  cudaEvent_t ev_up = nullptr;  // upload finished (copy stream)
  bool no_cluster = false;  // chunk of a pipelined align_batch: the chunks together fill the GPU
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // around K0, K1, K2 of the last solve
};

namespace {

int fail(ndtpso_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

#define CUDA_TRY(ctx, call)                                                                                         \
  do {                                                                                                              \
    cudaError_t e__ = (call);                                                                                       \
    if (e__ != cudaSuccess) {                                                                                       \
      return fail((ctx), e__ == cudaErrorMemoryAllocation ? NDTPSO_ERR_NOMEM : NDTPSO_ERR_CUDA,                     \
                  std::string(#call) + ": " + cudaGetErrorString(e__));                                             \
    }                                                                                                               \
  } while (0)

int pool_take(ndtpso_ctx* ctx, std::vector<PoolBuf>& pool, size_t bytes, bool pinned, PoolBuf* out) {
  int best = -1;
  for (int i = 0; i < (int)pool.size(); ++i)
    if (pool[i].bytes >= bytes && (best < 0 || pool[i].bytes < pool[best].bytes)) best = i;
  if (best >= 0) {
    *out = pool[best];
    pool.erase(pool.begin() + best);
    return NDTPSO_OK;
  }
  void* p = nullptr;
  const size_t cap = align_up(bytes + bytes / 4, 1 << 16);  // headroom so slightly larger batches still fit
  if (pinned) {
    CUDA_TRY(ctx, cudaMallocHost(&p, cap));
  } else {
    CUDA_TRY(ctx, cudaMalloc(&p, cap));
  }
  *out = PoolBuf{p, cap};
  return NDTPSO_OK;
}

void pool_give(std::vector<PoolBuf>& pool, PoolBuf b, bool pinned) {
  if (!b.ptr) return;
  if (pool.size() >= 8) {  // keep the pool small: drop the smallest
    int small = 0;
    for (int i = 1; i < (int)pool.size(); ++i)
      if (pool[i].bytes < pool[small].bytes) small = i;
    if (pool[small].bytes < b.bytes) std::swap(pool[small], b);
    if (pinned)
      cudaFreeHost(b.ptr);
    else
      cudaFree(b.ptr);
    return;
  }
  pool.push_back(b);
}

bool is_pow2_double(double v) {
  if (!(v > 0.) || !std::isfinite(v)) return false;
  int e;
  return std::frexp(v, &e) == 0.5;
}

struct MapKey {
  const void *mean, *icov, *built, *cidx;
  int32_t w, h, ns;
  bool operator<(const MapKey& o) const {
    return std::tie(mean, icov, built, cidx, w, h, ns) < std::tie(o.mean, o.icov, o.built, o.cidx, o.w, o.h, o.ns);
  }
};

int validate_problem(ndtpso_ctx* ctx, const ndtpso_problem& p, int b) {
  const ndtpso_map_view& m = p.map;
  char buf[160];
  if (p.n_points < 0 || (p.n_points > 0 && !p.points_xy)) {
    snprintf(buf, sizeof buf, "problem %d: bad points (n_points=%d)", b, p.n_points);
    return fail(ctx, NDTPSO_ERR_ARG, buf);
  }
  if (m.w_cells <= 0 || m.h_cells <= 0 || (int64_t)m.w_cells * m.h_cells > (int64_t)INT_MAX / 8 || !(m.cell_side > 0.)) {
    snprintf(buf, sizeof buf, "problem %d: bad grid %d x %d, cell_side %g", b, m.w_cells, m.h_cells, m.cell_side);
    return fail(ctx, NDTPSO_ERR_ARG, buf);
  }
  if (m.n_sparse < 0) {
    if (!m.mean || !m.inv_cov || !m.built) {
      snprintf(buf, sizeof buf, "problem %d: dense map with null table pointer", b);
      return fail(ctx, NDTPSO_ERR_ARG, buf);
    }
  } else {
    if (m.n_sparse > 0 && (!m.mean || !m.inv_cov || !m.cell_index)) {
      snprintf(buf, sizeof buf, "problem %d: sparse map with null table pointer", b);
      return fail(ctx, NDTPSO_ERR_ARG, buf);
    }
    if (m.n_sparse > 65534) {
      snprintf(buf, sizeof buf, "problem %d: sparse map with %d rows (max 65534)", b, m.n_sparse);
      return fail(ctx, NDTPSO_ERR_LIMIT, buf);
    }
  }
  return NDTPSO_OK;
}

// Persistent host workers for the staging loops (table scans, gathers, memcpy): spawning threads per batch cost more than
// the loops themselves.  One pool per CONTEXT, created on first use: contexts of one process (one per GPU, ndtpso_multi) stage
// their shards concurrently, each on its own workers.  run() hands out indices in small chunks and the calling thread works too.
class HostPool {
 public:
  explicit HostPool(int threads) {
    // Threads incl. the caller: the cores this process may use, at most 8; NDTPSO_HOST_THREADS or NDTPSO_OPT_HOST_THREADS
    // override.  Not divided by the ranks sharing the host: on a 32-core host with 8 ranks, 4 threads per rank staged
    // slower than 8 oversubscribed ones (e2e 704 k vs 807 k scan-matches/s on 8 GPUs).
    int budget = threads;
    if (budget <= 0) {
      int hw = (int)std::thread::hardware_concurrency();
      cpu_set_t set;
      if (sched_getaffinity(0, sizeof set, &set) == 0 && CPU_COUNT(&set) > 0) hw = CPU_COUNT(&set);
      budget = std::min(8, std::max(1, hw));
      if (const char* e = std::getenv("NDTPSO_HOST_THREADS")) budget = std::max(1, std::min(64, std::atoi(e)));
    }
    const int nw = std::max(0, std::min(63, budget - 1));
    for (int t = 0; t < nw; ++t) workers_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }
  int threads() const { return (int)workers_.size() + 1; }
  template <class F>
  void run(int n, F f) {
    const int want = std::min<int>((int)workers_.size() + 1, n / 8);  // not worth a thread for fewer than 8 items
    if (want <= 1) {
      for (int i = 0; i < n; ++i) f(i);
      return;
    }
    std::lock_guard<std::mutex> serial(run_mutex_);
    std::function<void(int)> fn = f;
    {
      std::lock_guard<std::mutex> lk(m_);
      job_ = &fn;
      n_ = n;
      next_.store(0);
      active_ = want - 1;   // workers that should join in
      running_ = want - 1;  // workers that have not finished yet
      ++epoch_;
    }
    cv_.notify_all();
    work(fn, n);
    std::unique_lock<std::mutex> lk(m_);
    done_cv_.wait(lk, [&] { return running_ == 0; });
    job_ = nullptr;
  }

 private:
  void work(std::function<void(int)>& fn, int n) {
    for (;;) {
      const int i0 = next_.fetch_add(4);
      if (i0 >= n) break;
      for (int i = i0; i < std::min(n, i0 + 4); ++i) fn(i);
    }
  }
  void loop() {
    unsigned seen = 0;
    for (;;) {
      std::function<void(int)>* fn = nullptr;
      int n = 0;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || (epoch_ != seen && active_ > 0); });
        if (stop_) return;
        seen = epoch_;
        --active_;
        fn = job_;
        n = n_;
      }
      work(*fn, n);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--running_ == 0) done_cv_.notify_one();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_, run_mutex_;
  std::condition_variable cv_, done_cv_;
  std::function<void(int)>* job_ = nullptr;
  std::atomic<int> next_{0};
  int n_ = 0, active_ = 0, running_ = 0;
  unsigned epoch_ = 0;
  bool stop_ = false;
};

// Runs f(i) for i in [0, n) on the context's staging workers (staging is memcpy / gather bound).
template <class F>
void parallel_for(ndtpso_ctx* ctx, int n, F f) {
  if (!ctx->pool) ctx->pool = new HostPool(ctx->opt_host_threads);
  ctx->pool->run(n, f);
}

// What the host learns from the first pass over a table — the `built` flags only (sequential, 1 byte per cell): which cells
// are built and the grid rows they span (=> shared-memory need of the compact form).  Whether every Sigma^-1 is symmetric,
// finite and positive semi-definite is checked while the rows are copied (check_row), the one pass that touches them.
struct MapScan {
  std::vector<int> cells;  // built cell indices, ascending (dense input); empty for sparse input
  int n_rec = 0, row0 = 0, nrows = 0;
  bool symmetric = true, ok = true;  // symmetric: also finite and positive semi-definite
};

// S01 == S10 bit for bit, finite, positive semi-definite (NaN/inf fail the comparisons): what NDTCell::build produces, and
// what the point-sliced kernel assumes (its exponent is then <= 0)
inline bool check_row(const double* S, const double* mu) {
  return memcmp(&S[1], &S[2], sizeof(double)) == 0 && S[0] >= 0. && S[3] >= 0. && S[0] * S[3] - S[1] * S[2] >= 0. && S[0] < 1e300 && S[3] < 1e300 &&
         std::isfinite(mu[0]) && std::isfinite(mu[1]);
}

void scan_map(const ndtpso_map_view& m, bool want_cells, MapScan* out) {
  const int ncells = m.w_cells * m.h_cells;
  int lo = INT_MAX, hi = -1, n_rec = 0;
  if (m.n_sparse >= 0) {
    for (int r = 0; r < m.n_sparse; ++r) {
      const int cell = m.cell_index[r];
      if (cell < 0 || cell >= ncells || (r > 0 && cell <= m.cell_index[r - 1])) {
        out->ok = false;
        return;
      }
    }
    n_rec = m.n_sparse;
    if (n_rec) {
      lo = m.cell_index[0];
      hi = m.cell_index[n_rec - 1];
    }
  } else {
    if (want_cells) out->cells.reserve(1024);
    int c = 0;
    for (; c + 8 <= ncells; c += 8) {  // tables are sparse: test 8 flags at a time
      uint64_t w;
      memcpy(&w, m.built + c, 8);
      if (!w) continue;
      for (int k = 0; k < 8; ++k)
        if (m.built[c + k]) {
          lo = std::min(lo, c + k);
          hi = c + k;
          ++n_rec;
          if (want_cells) out->cells.push_back(c + k);
        }
    }
    for (; c < ncells; ++c)
      if (m.built[c]) {
        lo = std::min(lo, c);
        hi = c;
        ++n_rec;
        if (want_cells) out->cells.push_back(c);
      }
  }
  out->n_rec = n_rec;
  out->row0 = n_rec ? lo / m.w_cells : 0;  // cells ascend, so the first and last built cell give the row span
  out->nrows = n_rec ? hi / m.w_cells - out->row0 + 1 : 0;
}

// Builds the arena for `n` problems and fills the pinned staging copy.
//   conf == nullptr     cost-only batch (no rand streams)
//   compact_on_host     dense tables are staged in the sparse form (built cells only): the host
//                       scans the `built` flags anyway, and the H2D copy shrinks from 49 bytes per
//                       cell to 52 bytes per BUILT cell.  Used by the host-buffer entry points;
//                       ndtpso_batch_create keeps dense tables dense (K0 compacts them on the device).
int batch_build(ndtpso_ctx* ctx, int32_t n, const ndtpso_problem* problems, const ndtpso_pso_config* conf, bool compact_on_host,
                size_t extra_upload_bytes, size_t extra_device_bytes, ndtpso_batch** out, size_t* extra_upload_off,
                size_t* extra_device_off) {
  if (!ctx || !out || n < 0 || (n > 0 && !problems)) return fail(ctx, NDTPSO_ERR_ARG, "batch: null argument or negative count");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  PsoParams prm{};
  if (conf) {
    if (conf->iterations < 0 || conf->population < 0) return fail(ctx, NDTPSO_ERR_ARG, "PSO config: negative iterations/population");
    if (conf->variant != NDTPSO_VARIANT_PSO && conf->variant != NDTPSO_VARIANT_GLIR)
      return fail(ctx, NDTPSO_ERR_ARG, "PSO config: unknown variant");
    prm.variant = conf->variant;
    const int64_t draws = ndtpso_rand_draws(conf);
    if (draws > INT_MAX / 2) return fail(ctx, NDTPSO_ERR_LIMIT, "PSO config: 3+3P+6PI exceeds the supported stream length");
    prm.P = conf->population;
    prm.I = conf->iterations;
    prm.w = conf->w;
    prm.c1 = conf->c1;
    prm.c2 = conf->c2;
    prm.wd = conf->w_dumping;
    prm.n_draws = (int)draws;
    if (pso_fixed_smem_bytes(prm.P) > ctx->max_smem_optin - 1024)
      return fail(ctx, NDTPSO_ERR_LIMIT, "PSO config: population too large for one CTA's shared memory");
  }
  for (int b = 0; b < n; ++b) {
    int rc = validate_problem(ctx, problems[b], b);
    if (rc) return rc;
    if (conf && problems[b].rand_stream && problems[b].rand_count < prm.n_draws)
      return fail(ctx, NDTPSO_ERR_ARG, "problem: rand_stream shorter than 3+3P+6PI");
  }

  // ---- maps: dedupe by identity of the table pointers
  std::map<MapKey, int> map_ids;
  std::vector<const ndtpso_map_view*> maps;
  std::vector<int> map_of(n);
  for (int b = 0; b < n; ++b) {
    const ndtpso_map_view& m = problems[b].map;
    MapKey k{m.mean, m.inv_cov, m.built, m.cell_index, m.w_cells, m.h_cells, m.n_sparse};
    auto it = map_ids.find(k);
    if (it == map_ids.end()) {
      it = map_ids.emplace(k, (int)maps.size()).first;
      maps.push_back(&m);
    }
    map_of[b] = it->second;
  }
  const int M = (int)maps.size();

  // ---- one pass over every table
  std::vector<MapScan> scans(M);
  parallel_for(ctx, M, [&](int i) { scan_map(*maps[i], compact_on_host, &scans[i]); });
  for (int i = 0; i < M; ++i)
    if (!scans[i].ok) return fail(ctx, NDTPSO_ERR_ARG, "sparse map: cell_index must be strictly ascending and inside the grid");

  // ---- arena layout
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes);
    return o;
  };
  const size_t o_probs = take(sizeof(DevProblem) * (size_t)std::max(n, 1));
  const size_t o_maps = take(sizeof(DevMap) * (size_t)std::max(M, 1));
  std::vector<size_t> o_pts(n), o_rnd(n, 0);
  for (int b = 0; b < n; ++b) o_pts[b] = take((size_t)problems[b].n_points * 16);
  std::vector<size_t> o_mean(M), o_icov(M), o_built(M), o_cidx(M);
  std::vector<int> rows(M);
  std::vector<char> staged_sparse(M);
  for (int i = 0; i < M; ++i) {
    const ndtpso_map_view& m = *maps[i];
    const int ncells = m.w_cells * m.h_cells;
    staged_sparse[i] = (m.n_sparse >= 0) || (compact_on_host && scans[i].n_rec <= 65534);
    rows[i] = m.n_sparse >= 0 ? m.n_sparse : (staged_sparse[i] ? scans[i].n_rec : ncells);
    o_mean[i] = take((size_t)rows[i] * 16);
    o_icov[i] = take((size_t)rows[i] * 32);
    o_built[i] = staged_sparse[i] ? 0 : take((size_t)ncells);
    o_cidx[i] = staged_sparse[i] ? take((size_t)rows[i] * 4) : 0;
  }
  bool any_device_rng = false;
  if (conf)
    for (int b = 0; b < n; ++b) {
      if (problems[b].rand_stream)
        o_rnd[b] = take((size_t)prm.n_draws * 4);
      else
        any_device_rng = true;
    }
  const size_t o_extra_up = take(extra_upload_bytes);
  const size_t upload_bytes = off;
  const size_t o_out = take(sizeof(double) * 4 * (size_t)std::max(n, 1));
  const size_t o_stats = take(sizeof(int) * kStatsWords * (size_t)std::max(n, 1));
  std::vector<size_t> o_hdr(M), o_grid(M), o_rec(M);
  for (int i = 0; i < M; ++i) {
    const ndtpso_map_view& m = *maps[i];
    o_hdr[i] = take(sizeof(int) * HDR_WORDS);
    o_grid[i] = take(((size_t)scans[i].nrows * m.w_cells + 1) * 2 + 16);  // row strip + null slot
    o_rec[i] = take(((size_t)scans[i].n_rec + 1) * 48);                   // records + null record
  }
  if (conf)
    for (int b = 0; b < n; ++b)
      if (!problems[b].rand_stream) o_rnd[b] = take((size_t)prm.n_draws * 4);
  const size_t o_extra_dev = take(extra_device_bytes);
  const size_t total_bytes = off;

  ndtpso_batch* bt = new (std::nothrow) ndtpso_batch();
  if (!bt) return fail(ctx, NDTPSO_ERR_NOMEM, "batch: host allocation failed");
  bt->ctx = ctx;
  bt->n = n;
  bt->n_maps = M;
  bt->prm = prm;
  bt->has_pso = conf != nullptr;
  bt->any_device_rng = any_device_rng;
  bt->upload_bytes = upload_bytes;
  int rc = pool_take(ctx, ctx->dev_pool, total_bytes, false, &bt->dev);
  if (rc == NDTPSO_OK) rc = pool_take(ctx, ctx->pin_pool, upload_bytes, true, &bt->pin);
  if (rc != NDTPSO_OK) {
    ndtpso_batch_destroy(bt);
    return rc;
  }
  unsigned char* h = static_cast<unsigned char*>(bt->pin.ptr);
  unsigned char* d = static_cast<unsigned char*>(bt->dev.ptr);
  bt->d_probs = reinterpret_cast<DevProblem*>(d + o_probs);
  bt->d_maps = reinterpret_cast<DevMap*>(d + o_maps);
  bt->d_out = reinterpret_cast<double*>(d + o_out);
  bt->d_stats = reinterpret_cast<int*>(d + o_stats);

  // ---- fill the staging buffer (tables and scans in parallel: pure memcpy / gather)
  DevMap* hm = reinterpret_cast<DevMap*>(h + o_maps);
  std::vector<int> map_dyn(M, 0);
  parallel_for(ctx, M, [&](int i) {
    const ndtpso_map_view& m = *maps[i];
    MapScan& sc = scans[i];
    const int ncells = m.w_cells * m.h_cells;
    DevMap dm{};
    dm.x_min = m.x_min;
    dm.x_max = m.x_max;
    dm.y_min = m.y_min;
    dm.y_max = m.y_max;
    dm.hw = m.width_m / 2.;
    dm.hh = m.height_m / 2.;
    dm.cs = m.cell_side;
    dm.inv_cs = 1.0 / m.cell_side;
    dm.fast_geom = (is_pow2_double(m.cell_side) && m.x_min == -m.x_max && m.y_min == -m.y_max) ? 1 : 0;
    dm.gw = m.w_cells;
    dm.gh = m.h_cells;
    dm.ncells = ncells;
    dm.n_sparse = staged_sparse[i] ? rows[i] : -1;
    dm.mean = reinterpret_cast<const double*>(d + o_mean[i]);
    dm.icov = reinterpret_cast<const double*>(d + o_icov[i]);
    dm.built = staged_sparse[i] ? nullptr : d + o_built[i];
    dm.cell_index = staged_sparse[i] ? reinterpret_cast<const int*>(d + o_cidx[i]) : nullptr;
    dm.grid = reinterpret_cast<unsigned short*>(d + o_grid[i]);
    dm.rec = reinterpret_cast<double*>(d + o_rec[i]);
    dm.hdr = reinterpret_cast<int*>(d + o_hdr[i]);
    hm[i] = dm;
    bool regular = true;
    if (m.n_sparse >= 0 || !staged_sparse[i]) {  // as given
      if (rows[i]) {
        memcpy(h + o_mean[i], m.mean, (size_t)rows[i] * 16);
        memcpy(h + o_icov[i], m.inv_cov, (size_t)rows[i] * 32);
      }
      if (m.n_sparse >= 0) {
        if (rows[i]) memcpy(h + o_cidx[i], m.cell_index, (size_t)rows[i] * 4);
        for (int r = 0; r < rows[i]; ++r) regular = regular && check_row(m.inv_cov + 4 * (size_t)r, m.mean + 2 * (size_t)r);
      } else {
        memcpy(h + o_built[i], m.built, (size_t)ncells);
        for (int c = 0; c < ncells; ++c)
          if (m.built[c]) regular = regular && check_row(m.inv_cov + 4 * (size_t)c, m.mean + 2 * (size_t)c);
      }
    } else {  // dense -> sparse on the host: the one pass over the built rows (scattered in the caller's table, so the rows of
              // the cells a few steps ahead are prefetched while the current one is copied and checked)
      double* hmean = reinterpret_cast<double*>(h + o_mean[i]);
      double* hicov = reinterpret_cast<double*>(h + o_icov[i]);
      int* hcidx = reinterpret_cast<int*>(h + o_cidx[i]);
      constexpr int kAhead = 12;
      for (int r = 0; r < std::min(kAhead, sc.n_rec); ++r) {
        __builtin_prefetch(m.mean + 2 * (size_t)sc.cells[r]);
        __builtin_prefetch(m.inv_cov + 4 * (size_t)sc.cells[r]);
      }
      for (int r = 0; r < sc.n_rec; ++r) {
        if (r + kAhead < sc.n_rec) {
          __builtin_prefetch(m.mean + 2 * (size_t)sc.cells[r + kAhead]);
          __builtin_prefetch(m.inv_cov + 4 * (size_t)sc.cells[r + kAhead]);
        }
        const size_t c = (size_t)sc.cells[r];
        hcidx[r] = (int)c;
        memcpy(hmean + 2 * (size_t)r, m.mean + 2 * c, 16);
        memcpy(hicov + 4 * (size_t)r, m.inv_cov + 4 * c, 32);
        regular = regular && check_row(hicov + 4 * (size_t)r, hmean + 2 * (size_t)r);
      }
    }
    sc.symmetric = regular;
    map_dyn[i] = (sc.n_rec + 1) * 48 + round16((sc.nrows * m.w_cells + 1) * 2);
  });
  for (int i = 0; i < M; ++i) {
    const ndtpso_map_view& mv = *maps[i];
    const double wc = mv.width_m / mv.cell_side, hc = mv.height_m / mv.cell_side;
    if (!(is_pow2_double(mv.cell_side) && mv.x_min == -mv.x_max && mv.y_min == -mv.y_max && wc == std::floor(wc) && hc == std::floor(hc) &&
          mv.x_max == mv.width_m / 2. && mv.y_max == mv.height_m / 2. && mv.x_max == mv.y_max))  // whole cells, square
      bt->scr_ok = false;
    bt->scr_ext = std::max(bt->scr_ext, std::max(std::fabs(mv.x_max), std::fabs(mv.y_max)));
    bt->scr_inv_cs = std::max(bt->scr_inv_cs, 1.0 / mv.cell_side);
    bt->scr_gw = std::max(bt->scr_gw, std::max(mv.w_cells, mv.h_cells));
    bt->max_table_smem = std::max(bt->max_table_smem, map_dyn[i]);
    bt->max_n_rec = std::max(bt->max_n_rec, scans[i].n_rec);
    if (scans[i].n_rec > 65534) bt->all_compact = false;
    if (!scans[i].symmetric) bt->all_symmetric = false;
  }
  DevProblem* hp = reinterpret_cast<DevProblem*>(h + o_probs);
  std::vector<double> pmax_of(std::max(n, 1), 0.);
  parallel_for(ctx, n, [&](int b) {
    const ndtpso_problem& p = problems[b];
    DevProblem dp{};
    dp.pts = reinterpret_cast<const double2*>(d + o_pts[b]);
    dp.n_pts = p.n_points;
    dp.map_id = map_of[b];
    for (int k = 0; k < 3; ++k) {
      dp.guess[k] = p.guess[k];
      dp.dev[k] = p.deviation[k];
    }
    dp.seed = p.seed;
    dp.rnd = conf ? reinterpret_cast<const int*>(d + o_rnd[b]) : nullptr;
    dp.rnd_from_host = (conf && p.rand_stream) ? 1 : 0;
    hp[b] = dp;
    if (p.n_points) memcpy(h + o_pts[b], p.points_xy, (size_t)p.n_points * 16);
    if (conf && p.rand_stream) memcpy(h + o_rnd[b], p.rand_stream, (size_t)prm.n_draws * 4);
    double pm = 0.;
    bool finite = true;
    for (int i = 0; i < 2 * p.n_points; ++i) {
      const double v = std::fabs(p.points_xy[i]);
      pm = v > pm ? v : pm;
      finite = finite && (v <= 1e30);  // false for NaN too
    }
    pmax_of[b] = finite ? pm : INFINITY;
  });
  for (int b = 0; b < n; ++b) bt->scr_pmax = std::max(bt->scr_pmax, pmax_of[b]);
  int need = 0;
  for (int b = 0; b < n; ++b) {
    need = std::max(need, problems[b].n_points * 16 + map_dyn[map_of[b]]);
    bt->max_pts = std::max(bt->max_pts, problems[b].n_points);
  }
  bt->need_dyn_smem = need;
  if (extra_upload_off) *extra_upload_off = o_extra_up;
  if (extra_device_off) *extra_device_off = o_extra_dev;
  *out = bt;
  return NDTPSO_OK;
}

int batch_upload(ndtpso_batch* bt) {
  ndtpso_ctx* ctx = bt->ctx;
  if (ctx->stream == ctx->own_stream) {
    // own stream: upload on the copy stream so that it overlaps kernels of earlier batches; the
    // compute stream waits for it (the arena of an in-flight batch is never reused, see the pool)
    if (!ctx->copy_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    if (!bt->ev_up) CUDA_TRY(ctx, cudaEventCreateWithFlags(&bt->ev_up, cudaEventDisableTiming));
    CUDA_TRY(ctx, cudaMemcpyAsync(bt->dev.ptr, bt->pin.ptr, bt->upload_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
    CUDA_TRY(ctx, cudaEventRecord(bt->ev_up, ctx->copy_stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, bt->ev_up, 0));
  } else {
    CUDA_TRY(ctx, cudaMemcpyAsync(bt->dev.ptr, bt->pin.ptr, bt->upload_bytes, cudaMemcpyHostToDevice, ctx->stream));
  }
  ctx->last_h2d = (int64_t)bt->upload_bytes;
  return NDTPSO_OK;
}

int pick_warps(const ndtpso_ctx* ctx) {
  int w = ctx->opt_warps;
  if (w == 4 || w == 8 || w == 16 || w == 32) return w;
  return 8;
}

int pick_smem(const ndtpso_ctx* ctx, int fixed_bytes, int need_dyn) {
  int64_t s = ctx->opt_smem > 0 ? ctx->opt_smem : (int64_t)fixed_bytes + need_dyn;
  s = std::max<int64_t>(s, fixed_bytes);
  s = std::min<int64_t>(s, ctx->max_smem_optin);
  return (int)((s + 15) & ~15);
}

template <int NW>
int launch_pso(ndtpso_batch* bt, int smem) {
  ndtpso_ctx* ctx = bt->ctx;
  if (ctx->smem_attr_set[NW] < smem) {
    CUDA_TRY(ctx, cudaFuncSetAttribute(pso_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin));
    ctx->smem_attr_set[NW] = ctx->max_smem_optin;
  }
  PsoParams prm = bt->prm;
  prm.smem_bytes = smem;
  pso_kernel<NW><<<bt->n, NW * 32, smem, ctx->stream>>>(bt->d_probs, bt->d_maps, prm, bt->d_out, bt->d_stats);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return NDTPSO_OK;
}

// Point-sliced kernel: T = 32*NW threads hold NPT scan points each and score JB candidates at a
// time; a problem may be spread over a cluster of CL CTAs (G candidate groups x S point slices).
// launch_sliced returns 1 when the batch does not qualify (table too large for shared memory,
// scan too long, > 65534 built cells, asymmetric Sigma^-1).
// Constants of the fp32 screen's error bound for this batch (derivation: "fp32 screen" in ndtpso_pso_sliced.cuh); false = the
// batch does not qualify.  Everything is in cell sides.
//   du:   bound on |Uf - U*|, the fp32 error of a transformed point's cell coordinate, for every point the fp64 evaluation
//         places inside its frame: first-order part 2^-24 (11.25 pmax + 3 W)/cs + 1.5 2^-24; taken as
//         2^-24 (12 pmax' + 3.01 W)/cs + 2^-23 with pmax' = max(pmax, 1), which covers every higher-order term and the fp64
//         side's own rounding as long as (pmax + W)/cs <= 2^22 (checked here).  pmax, W, 1/cs: the largest of the batch.
//   beta: band around cell edges inside which fp32 and fp64 may pick different cells; it only has to be >= du, and is taken
//         twice that.  (A lane of the screen that has ANY point in the band counts all its points as worst case, so the band's
//         width costs tightness: at 2 du about 0.2 % of a typical cost.)
bool screen_params(const ndtpso_batch* bt, PsoParams* prm, bool ignore_option = false) {
  prm->screen = 0;
  const ndtpso_ctx* ctx = bt->ctx;
  if ((ctx->opt_screen == 0 && !ignore_option) || !bt->scr_ok || !(bt->scr_pmax <= 1e6) || !(bt->scr_ext > 0.) || bt->scr_gw >= (1 << 20)) return false;
  if (bt->max_n_rec + 1 > kScreenMaxRecords) return false;  // the staged grid holds 16-bit shared addresses of the screen's records
  const double u24 = 5.9604644775390625e-08;  // 2^-24
  const double pmax = std::max(bt->scr_pmax, 1.0), W = 2. * bt->scr_ext;
  if ((pmax + W) * bt->scr_inv_cs > 4194304.) return false;
  const double du = u24 * (12. * pmax + 3.01 * W) * bt->scr_inv_cs + 2. * u24;
  const double beta = 2. * du;
  if (beta > 0.05) return false;
  prm->screen = 1;
  prm->scr_du = (float)(du * 1.000001);  // rounded to fp32: 1e-6 relative covers the rounding
  prm->scr_beta_c = (float)(0.5 - beta);
  return true;
}

template <int NPT, int JB, int CL, int MAXT, int MINB>
int launch_sliced_cfg(ndtpso_batch* bt, int nw, int groups, int smem, bool screen = false) {
  ndtpso_ctx* ctx = bt->ctx;
  auto kern = pso_sliced_kernel<NPT, JB, CL, MAXT, MINB>;
  // function attributes are per device; remembered per context (a context never changes its device)
  const void* kern_id = reinterpret_cast<const void*>(kern);
  if (std::find(ctx->sliced_attr_set.begin(), ctx->sliced_attr_set.end(), kern_id) == ctx->sliced_attr_set.end()) {
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin));
    if (CL > 8) CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    ctx->sliced_attr_set.push_back(kern_id);
  }
  PsoParams prm = bt->prm;
  prm.smem_bytes = smem;
  // one CTA per problem: rounds are cheap (two barriers), so a small window pays; a cluster round costs a cluster barrier
  prm.hot_chunk = ctx->opt_hot_chunk >= 0 ? (ctx->opt_hot_chunk & 0xffff) : (CL == 1 ? kDefaultHotChunk : 0);
  prm.hot_thresh = ctx->opt_hot_chunk >= 0x10000 ? (ctx->opt_hot_chunk >> 16) : 1;
  if (!(CL == 1 && screen && screen_params(bt, &prm))) prm.screen = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)bt->n * CL, 1, 1);
  cfg.blockDim = dim3((unsigned)nw * 32, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, kern, (const DevProblem*)bt->d_probs, (const DevMap*)bt->d_maps, prm, groups, bt->d_out, bt->d_stats));
  ctx->launches++;
  return NDTPSO_OK;
}

// maximum warps per CTA of each points-per-thread variant (its __launch_bounds__)
constexpr int kSlicedMaxWarps[kSlicedMaxNPT + 1] = {0, 20, 20, 12, 10, 8, 8};

// Cluster form of the sliced kernel for one cluster size: S = min(CL, 4) point slices, G = CL / S
// candidate groups, 4 candidates per batch.  Returns 1 when the shape does not fit (scan too long
// for 12 warps x 2 points per thread, table too large) or when fewer than n clusters of this size
// can be resident at once (cudaOccupancyMaxActiveClusters; e.g. only ~14 clusters of 8 fit the
// GPCs of a B200, so 16 problems are faster on clusters of 4) unless the size was forced.
template <int CL>
int try_cluster(ndtpso_batch* bt, bool forced) {
#ifdef NDTPSO_DEV_FAST  // experiment builds (tools/variant_time.py): only the production shape is instantiated
  return 1;
#else
  ndtpso_ctx* ctx = bt->ctx;
  const int n = std::max(bt->max_pts, 1);
  const int S = std::min(CL, 4), G = CL / S;
  int npt = 1, nw = (n + 32 * S - 1) / (32 * S);
  if (nw > 12) {
    npt = 2;
    nw = (n + 64 * S - 1) / (64 * S);
  }
  if (nw > 12) return 1;
  nw = std::max(nw, 2);
  const int smem = round16(sliced_smem_bytes(bt->prm.P, S, nw, bt->max_table_smem));
  if (smem > ctx->max_smem_optin) return 1;
  if (!forced) {
    const void* kern = npt == 1 ? (const void*)pso_sliced_kernel<1, 8, CL, 384, 1> : (const void*)pso_sliced_kernel<2, 4, CL, 384, 1>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin) != cudaSuccess ||
        (CL > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess)) {
      cudaGetLastError();
      return 1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)bt->n * CL, 1, 1);
    cfg.blockDim = dim3((unsigned)nw * 32, 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess) {
      cudaGetLastError();
      return 1;
    }
    if (bt->n > max_clusters) return 1;
  }
  return npt == 1 ? launch_sliced_cfg<1, 8, CL, 384, 1>(bt, nw, G, smem) : launch_sliced_cfg<2, 4, CL, 384, 1>(bt, nw, G, smem);
#endif
}

#ifndef NDTPSO_NPT4_T
#define NDTPSO_NPT4_T 320  // __launch_bounds__ of the 4-points-per-thread shape: threads, CTAs per SM
#endif
#ifndef NDTPSO_NPT4_MINB
#define NDTPSO_NPT4_MINB 2
#endif
#ifndef NDTPSO_MINB3
#define NDTPSO_MINB3 2  // CTAs per SM the 3-points-per-thread shape is compiled for (2 => 80 registers); tools/variant_time.py
#endif
int launch_sliced(ndtpso_batch* bt) {
  ndtpso_ctx* ctx = bt->ctx;
  if (!bt->all_compact || !bt->all_symmetric) return 1;
  const int n = std::max(bt->max_pts, 1);
  // Spread a problem over a cluster of SMs only while the batch alone leaves SMs idle: the
  // largest cluster size whose clusters can all be resident together.
  if (ctx->opt_cluster > 1) {
    int rc = 1;
    switch (ctx->opt_cluster) {
      case 2: rc = try_cluster<2>(bt, true); break;
      case 4: rc = try_cluster<4>(bt, true); break;
      case 8: rc = try_cluster<8>(bt, true); break;
      default: rc = try_cluster<16>(bt, true); break;
    }
    if (rc != 1) return rc;
  } else if (ctx->opt_cluster == 0 && !bt->no_cluster && bt->n * 2 <= ctx->sm_count) {
    int rc = 1;
    if (bt->n * 16 <= ctx->sm_count) rc = try_cluster<16>(bt, false);
    if (rc == 1 && bt->n * 8 <= ctx->sm_count) rc = try_cluster<8>(bt, false);
    if (rc == 1 && bt->n * 4 <= ctx->sm_count) rc = try_cluster<4>(bt, false);
    // a cluster of two (unscreened) loses to one screened CTA per problem (profiles/r02c_cluster_sizes.txt: 64 problems 1.69 vs 1.47 ms):
    // only for batches the screen does not take
    PsoParams probe{};
    if (rc == 1 && !(ctx->opt_screen != 0 && screen_params(bt, &probe, true))) rc = try_cluster<2>(bt, false);
    if (rc != 1) return rc;
  }
  int npt = 0, nw = 0;
  if (ctx->opt_npt > 0) {
    npt = std::min(ctx->opt_npt, kSlicedMaxNPT);
    nw = (n + 32 * npt - 1) / (32 * npt);
  } else if (ctx->opt_warps > 0) {
    npt = (n + 32 * ctx->opt_warps - 1) / (32 * ctx->opt_warps);
    if (npt > kSlicedMaxNPT) return 1;
    nw = (n + 32 * npt - 1) / (32 * npt);
  } else {
    // default: about 12 warps per CTA (two CTAs per SM at <= 80 registers measured fastest, see
    // profiles/r01_phaseB_microbench.md): the fewest points per thread that needs at most 12 warps
    for (npt = 1; npt <= kSlicedMaxNPT; ++npt) {
      nw = (n + 32 * npt - 1) / (32 * npt);
      if (nw <= 12) break;
    }
    if (npt > kSlicedMaxNPT) return 1;
  }
  if (npt < 1 || npt > kSlicedMaxNPT || nw > kSlicedMaxWarps[npt]) return 1;
  nw = std::max(nw, 4);
  PsoParams probe{};
  // the launch shape is chosen as if the screen were on whenever the batch qualifies for it, whatever NDTPSO_OPT_SCREEN says:
  // a cost is a tree sum over the CTA's warps, so only the same shape gives bit-identical costs with the screen on and off
  bool scr = screen_params(bt, &probe, true);
  auto smem_of = [&](int warps, bool screen) { return round16(sliced_smem_bytes(bt->prm.P, warps, 0, bt->max_table_smem, screen ? bt->max_n_rec + 1 : 0)); };
  int smem = smem_of(nw, scr);
  if (scr && smem > ctx->max_smem_optin) {  // the screen's tables do not fit: fp64 only
    scr = false;
    smem = smem_of(nw, false);
  }
  if (ctx->opt_npt <= 0 && ctx->opt_warps <= 0 && smem > ctx->max_smem_optin / 2) {
    // one CTA per SM anyway (large swarm or table: BASELINE configs[4], 200 particles): registers are plentiful then, so take
    // the shape with the most warps that fits (tools/cfg5_time.py: 2 points x 17 warps 7.5 ms, 3 x 12 warps 8.0 ms per 148 matches)
    for (int p2 = 1; p2 < npt; ++p2) {
      const int w2 = std::max((n + 32 * p2 - 1) / (32 * p2), 4);
      if (w2 > kSlicedMaxWarps[p2] || smem_of(w2, scr) > ctx->max_smem_optin) continue;
      npt = p2;
      nw = w2;
      smem = smem_of(nw, scr);
      break;
    }
  }
  if (scr && ctx->opt_screen == 0) {  // switched off by the caller: same shape, without the screen's tables
    scr = false;
    smem = smem_of(nw, false);
  }
  if (smem > ctx->max_smem_optin) return 1;
  const int jb = ctx->opt_cand_batch;
#ifdef NDTPSO_DEV_FAST
  return npt == 3 ? launch_sliced_cfg<3, 4, 1, 384, NDTPSO_MINB3>(bt, nw, 1, smem, scr) : npt == 4 ? launch_sliced_cfg<4, 4, 1, NDTPSO_NPT4_T, NDTPSO_NPT4_MINB>(bt, nw, 1, smem, scr) : 1;
#else
  switch (npt) {
    case 1: return jb == 1 ? launch_sliced_cfg<1, 1, 1, 640, 1>(bt, nw, 1, smem, scr) : jb == 2 ? launch_sliced_cfg<1, 2, 1, 640, 1>(bt, nw, 1, smem, scr)
                                                                                        : launch_sliced_cfg<1, 4, 1, 640, 1>(bt, nw, 1, smem, scr);
    case 2: return jb == 1 ? launch_sliced_cfg<2, 1, 1, 640, 1>(bt, nw, 1, smem, scr) : jb == 2 ? launch_sliced_cfg<2, 2, 1, 640, 1>(bt, nw, 1, smem, scr)
                                                                                        : launch_sliced_cfg<2, 4, 1, 640, 1>(bt, nw, 1, smem, scr);
    case 3: return jb == 1 ? launch_sliced_cfg<3, 1, 1, 384, NDTPSO_MINB3>(bt, nw, 1, smem, scr) : jb == 2 ? launch_sliced_cfg<3, 2, 1, 384, NDTPSO_MINB3>(bt, nw, 1, smem, scr)
                                                                                        : launch_sliced_cfg<3, 4, 1, 384, NDTPSO_MINB3>(bt, nw, 1, smem, scr);
    case 4: return jb == 1 ? launch_sliced_cfg<4, 1, 1, NDTPSO_NPT4_T, NDTPSO_NPT4_MINB>(bt, nw, 1, smem, scr) : launch_sliced_cfg<4, 2, 1, NDTPSO_NPT4_T, NDTPSO_NPT4_MINB>(bt, nw, 1, smem, scr);
    case 5: return jb == 1 ? launch_sliced_cfg<5, 1, 1, 256, 2>(bt, nw, 1, smem, scr) : launch_sliced_cfg<5, 2, 1, 256, 2>(bt, nw, 1, smem, scr);
    default: return jb == 1 ? launch_sliced_cfg<6, 1, 1, 256, 2>(bt, nw, 1, smem, scr) : launch_sliced_cfg<6, 2, 1, 256, 2>(bt, nw, 1, smem, scr);
  }
#endif
}

// Points the next PSO launch of `bt` at the gathered buffers of all ranks (next epoch), or switches publishing off.
void exchange_arm(ndtpso_batch* bt) {
  PeerExchange& px = bt->prm.ex;
  ndtpso_exchange* ex = bt->ex;
  if (!ex || !ex->connected) {
    px.world = 0;
    return;
  }
  ++ex->epoch;
  const size_t half = (ex->epoch & 1u) * ex->half_bytes;
  for (int r = 0; r < ex->world; ++r) {
    px.out[r] = reinterpret_cast<double*>(ex->peer_base[r] + half);
    px.flag[r] = reinterpret_cast<unsigned*>(ex->peer_base[r] + ex->o_flags);
  }
  px.done = reinterpret_cast<unsigned*>(ex->base + ex->o_done);
  px.world = ex->world;
  px.my_rank = ex->rank;
  px.offset = ex->rank * ex->n_per_rank;
  px.epoch = ex->epoch;
}

// K2: the point-sliced kernel when the batch qualifies, else the generic warp-per-particle kernel
// (any scan length, any table size)
int launch_pso_any(ndtpso_batch* bt) {
  ndtpso_ctx* ctx = bt->ctx;
  const bool glir = bt->prm.variant == NDTPSO_VARIANT_GLIR;  // the point-sliced kernel implements pso_optimization only
  int rc = (ctx->opt_kernel == 1 || glir) ? 1 : launch_sliced(bt);
  if (rc == 1) {
    if (ctx->opt_kernel == 2 && !glir) return fail(ctx, NDTPSO_ERR_LIMIT, "batch does not qualify for the point-sliced kernel");
    const int fixed = pso_fixed_smem_bytes(bt->prm.P);
    const int smem = pick_smem(ctx, fixed, bt->need_dyn_smem);
    switch (pick_warps(ctx)) {
      case 4: rc = launch_pso<4>(bt, smem); break;
      case 16: rc = launch_pso<16>(bt, smem); break;
      case 32: rc = launch_pso<32>(bt, smem); break;
      default: rc = launch_pso<8>(bt, smem); break;
    }
  }
  return rc;
}

// K1: the rand() streams of the problems that did not bring their own
int launch_rng(ndtpso_batch* bt) {
  ndtpso_ctx* ctx = bt->ctx;
  const int grid = (bt->n + K1_WARPS - 1) / K1_WARPS;
  if (bt->d_rng_state)
    rng_fill_kernel<true><<<grid, K1_WARPS * 32, 0, ctx->stream>>>(bt->d_probs, bt->n, bt->prm.n_draws, bt->d_rng_state);
  else
    rng_fill_kernel<false><<<grid, K1_WARPS * 32, 0, ctx->stream>>>(bt->d_probs, bt->n, bt->prm.n_draws, nullptr);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return NDTPSO_OK;
}

template <int NPT>
int launch_screen_bound(ndtpso_batch* bt, int nw, int smem, const PsoParams& prm, int m, const double* d_poses, double* d_out) {
  ndtpso_ctx* ctx = bt->ctx;
  auto kern = screen_bound_kernel<NPT, 640>;
  CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin));
  kern<<<bt->n, nw * 32, smem, ctx->stream>>>(bt->d_probs, bt->d_maps, prm, m, d_poses, d_out);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return NDTPSO_OK;
}

int launch_compact(ndtpso_batch* bt) {
  ndtpso_ctx* ctx = bt->ctx;
  if (bt->n_maps == 0) return NDTPSO_OK;
  compact_map_kernel<<<bt->n_maps, K0_THREADS, 0, ctx->stream>>>(bt->d_maps);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return NDTPSO_OK;
}

}  // namespace

extern "C" {

int ndtpso_abi_version(void) { return NDTPSO_ABI_VERSION; }

void ndtpso_pso_config_default(ndtpso_pso_config* conf) {
  if (!conf) return;
  conf->iterations = 50;   // PSO_ITERATIONS        config.h:20
  conf->population = 30;   // PSO_POPULATION_SIZE   config.h:21
  conf->num_threads = -1;  //                       config.h:30
  conf->variant = NDTPSO_VARIANT_PSO;
  conf->w = .8;            // PSO_W                 config.h:23
  conf->c1 = 2.;           // PSO_C1                config.h:24
  conf->c2 = 2.;           // PSO_C2                config.h:25
  conf->w_dumping = 1.;    // PSO_W_DUMPING_COEF    config.h:22
}

int64_t ndtpso_rand_draws(const ndtpso_pso_config* conf) {
  if (!conf) return 0;
  const int64_t P = conf->population, I = conf->iterations;
  if (conf->variant == NDTPSO_VARIANT_GLIR) return 3 * (P + 2) + 6 * P * I;  // core.cpp:125,132,135,149
  return 3 + 3 * P + 6 * P * I;
}

int ndtpso_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int ndtpso_ctx_create(int device, ndtpso_ctx** out) {
  if (!out) return NDTPSO_ERR_ARG;
  *out = nullptr;
  const int n = ndtpso_device_count();
  if (n <= 0 || device < 0 || device >= n) return NDTPSO_ERR_NODEVICE;
  ndtpso_ctx* ctx = new (std::nothrow) ndtpso_ctx();
  if (!ctx) return NDTPSO_ERR_NOMEM;
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess ||
      cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    cudaGetLastError();
    delete ctx;
    return NDTPSO_ERR_CUDA;
  }
  if (prop.major < 10) {  // sm_100a code only
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return NDTPSO_ERR_NODEVICE;
  }
  ctx->stream = ctx->own_stream;
  ctx->sm_count = prop.multiProcessorCount;
  if (prop.clockRate > 0) ctx->clock_khz = prop.clockRate;
  ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  *out = ctx;
  return NDTPSO_OK;
}

void ndtpso_ctx_destroy(ndtpso_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& b : ctx->dev_pool) cudaFree(b.ptr);
  for (auto& b : ctx->pin_pool) cudaFreeHost(b.ptr);
  for (auto& cs : ctx->chunk_stream)
    if (cs) cudaStreamDestroy(cs);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  for (auto& ps : ctx->pipe_stream)
    if (ps) cudaStreamDestroy(ps);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx->pool;
  delete ctx;
}

int ndtpso_ctx_set_stream(ndtpso_ctx* ctx, void* cuda_stream) {
  if (!ctx) return NDTPSO_ERR_ARG;
  ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  return NDTPSO_OK;
}

const char* ndtpso_last_error(const ndtpso_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int ndtpso_ctx_set_option(ndtpso_ctx* ctx, int option, int64_t value) {
  if (!ctx) return NDTPSO_ERR_ARG;
  switch (option) {
    case NDTPSO_OPT_WARPS_PER_CTA:
      if (value < 0 || value > 32) return fail(ctx, NDTPSO_ERR_ARG, "warps per CTA must be in 0..32");
      ctx->opt_warps = (int)value;
      return NDTPSO_OK;
    case NDTPSO_OPT_KERNEL:
      if (value < 0 || value > 2) return fail(ctx, NDTPSO_ERR_ARG, "kernel must be 0 (auto), 1 (warp-per-particle) or 2 (point-sliced)");
      ctx->opt_kernel = (int)value;
      return NDTPSO_OK;
    case NDTPSO_OPT_PIPELINE_CHUNKS:
      if (value < 0 || value > kMaxChunks) return fail(ctx, NDTPSO_ERR_ARG, "pipeline chunks must be in 0..8 (0 = auto)");
      ctx->opt_chunks = (int)value;
      return NDTPSO_OK;
    case NDTPSO_OPT_CANDIDATE_BATCH:
      if (value != 0 && value != 1 && value != 2 && value != 4) return fail(ctx, NDTPSO_ERR_ARG, "candidate batch must be 0, 1, 2 or 4");
      ctx->opt_cand_batch = (int)value;
      return NDTPSO_OK;
    case NDTPSO_OPT_POINTS_PER_THREAD:
      if (value < 0 || value > kSlicedMaxNPT) return fail(ctx, NDTPSO_ERR_ARG, "points per thread out of range");
      ctx->opt_npt = (int)value;
      return NDTPSO_OK;
    case NDTPSO_OPT_SCREEN:
      if (value < -1 || value > 1) return fail(ctx, NDTPSO_ERR_ARG, "screen must be -1 (auto), 0 (off) or 1 (on)");
      ctx->opt_screen = (int)value;
      return NDTPSO_OK;
    case NDTPSO_OPT_HOT_CHUNK:
      if (value < -1 || (value & 0xffff) > 4096 || value > 0xffffff) return fail(ctx, NDTPSO_ERR_ARG, "speculation window must be -1 (auto), 0 (off) or a particle count");
      ctx->opt_hot_chunk = (int)value;
      return NDTPSO_OK;
    case NDTPSO_OPT_EXCHANGE_TIMEOUT_MS:
      if (value < 1 || value > 600000) return fail(ctx, NDTPSO_ERR_ARG, "exchange timeout must be in 1..600000 ms");
      ctx->opt_exchange_timeout_ms = value;
      return NDTPSO_OK;
    case NDTPSO_OPT_HOST_THREADS:
      if (value < 0 || value > 64) return fail(ctx, NDTPSO_ERR_ARG, "host threads must be in 0..64 (0 = auto)");
      if (ctx->pool && (int)value != ctx->opt_host_threads) {  // takes effect with the next batch
        delete ctx->pool;
        ctx->pool = nullptr;
      }
      ctx->opt_host_threads = (int)value;
      return NDTPSO_OK;
    case NDTPSO_OPT_SMEM_BYTES:
      if (value < 0 || value > ctx->max_smem_optin) return fail(ctx, NDTPSO_ERR_ARG, "shared memory bytes out of range");
      ctx->opt_smem = value;
      return NDTPSO_OK;
    case NDTPSO_OPT_CLUSTER:
      if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8 && value != 16)
        return fail(ctx, NDTPSO_ERR_ARG, "cluster size must be 0 (auto), 1, 2, 4, 8 or 16");
      ctx->opt_cluster = (int)value;
      return NDTPSO_OK;
    default:
      return fail(ctx, NDTPSO_ERR_ARG, "unknown option");
  }
}

int64_t ndtpso_ctx_launch_count(const ndtpso_ctx* ctx) { return ctx ? ctx->launches : 0; }

int ndtpso_ctx_last_transfer_bytes(const ndtpso_ctx* ctx, int64_t* h2d, int64_t* d2h) {
  if (!ctx) return NDTPSO_ERR_ARG;
  if (h2d) *h2d = ctx->last_h2d;
  if (d2h) *d2h = ctx->last_d2h;
  return NDTPSO_OK;
}

int ndtpso_ctx_synchronize(ndtpso_ctx* ctx) {
  if (!ctx) return NDTPSO_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return NDTPSO_OK;
}

static int batch_create_impl(ndtpso_ctx* ctx, int32_t n, const ndtpso_problem* problems, const ndtpso_pso_config* conf, bool compact_on_host,
                             ndtpso_batch** out) {
  if (!conf) return fail(ctx, NDTPSO_ERR_ARG, "batch_create: null config");
  if (out) *out = nullptr;
  ndtpso_batch* bt = nullptr;
  int rc = batch_build(ctx, n, problems, conf, compact_on_host, 0, 0, &bt, nullptr, nullptr);
  if (rc) return rc;
  rc = batch_upload(bt);
  if (rc) {
    ndtpso_batch_destroy(bt);
    return rc;
  }
  *out = bt;
  return NDTPSO_OK;
}

int ndtpso_batch_create(ndtpso_ctx* ctx, int32_t n, const ndtpso_problem* problems, const ndtpso_pso_config* conf, ndtpso_batch** out) {
  // tables stay in the form the caller gave: dense ones are compacted by K0 on the device at every solve
  return batch_create_impl(ctx, n, problems, conf, false, out);
}

int ndtpso_batch_solve(ndtpso_batch* bt) {
  if (!bt || !bt->has_pso) return NDTPSO_ERR_ARG;
  ndtpso_ctx* ctx = bt->ctx;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  if (bt->n == 0) {
    bt->solved = true;
    return NDTPSO_OK;
  }
  if (!bt->ev[0])
    for (auto& e : bt->ev) CUDA_TRY(ctx, cudaEventCreate(&e));
  // the upload ran on the copy stream: whatever stream is current NOW (ndtpso_ctx_set_stream may have changed it since
  // ndtpso_batch_create) must not start before it has finished
  if (bt->ev_up) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, bt->ev_up, 0));
  CUDA_TRY(ctx, cudaEventRecord(bt->ev[0], ctx->stream));
  int rc = launch_compact(bt);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaEventRecord(bt->ev[1], ctx->stream));
  if (bt->any_device_rng) {
    rc = launch_rng(bt);
    if (rc) return rc;
  }
  CUDA_TRY(ctx, cudaEventRecord(bt->ev[2], ctx->stream));
  exchange_arm(bt);
  rc = launch_pso_any(bt);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaEventRecord(bt->ev[3], ctx->stream));
  bt->solved = true;
  return NDTPSO_OK;
}

int ndtpso_batch_kernel_times(ndtpso_batch* bt, double* out_ms) {
  if (!bt || !out_ms) return NDTPSO_ERR_ARG;
  ndtpso_ctx* ctx = bt->ctx;
  if (!bt->solved || !bt->ev[0]) return fail(ctx, NDTPSO_ERR_ARG, "batch_kernel_times before batch_solve");
  CUDA_TRY(ctx, cudaEventSynchronize(bt->ev[3]));
  for (int i = 0; i < 3; ++i) {
    float ms = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, bt->ev[i], bt->ev[i + 1]));
    out_ms[i] = ms;
  }
  return NDTPSO_OK;
}

void* ndtpso_batch_device_results(ndtpso_batch* bt) { return bt ? bt->d_out : nullptr; }

int ndtpso_batch_results(ndtpso_batch* bt, double* out_pose, double* out_cost) {
  if (!bt) return NDTPSO_ERR_ARG;
  ndtpso_ctx* ctx = bt->ctx;
  if (!bt->solved) return fail(ctx, NDTPSO_ERR_ARG, "batch_results before batch_solve");
  if (bt->n == 0) return NDTPSO_OK;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));  // a process with one context per GPU calls in from any thread
  // results come back through the (already consumed) head of the pinned staging buffer
  double* h = static_cast<double*>(bt->pin.ptr);
  if (bt->results_enqueued) {
    CUDA_TRY(ctx, cudaEventSynchronize(bt->ev_done));  // waits for THIS batch only, not for batches queued behind it
  } else {
    CUDA_TRY(ctx, cudaMemcpyAsync(h, bt->d_out, sizeof(double) * 4 * (size_t)bt->n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  ctx->last_d2h = (int64_t)sizeof(double) * 4 * bt->n;
  for (int b = 0; b < bt->n; ++b) {
    if (out_pose) {
      out_pose[3 * b] = h[4 * b];
      out_pose[3 * b + 1] = h[4 * b + 1];
      out_pose[3 * b + 2] = h[4 * b + 2];
    }
    if (out_cost) out_cost[b] = h[4 * b + 3];
  }
  return NDTPSO_OK;
}

int ndtpso_batch_stats(ndtpso_batch* bt, int32_t* out) {
  if (!bt || !out) return NDTPSO_ERR_ARG;
  ndtpso_ctx* ctx = bt->ctx;
  if (!bt->solved) return fail(ctx, NDTPSO_ERR_ARG, "batch_stats before batch_solve");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  std::vector<int32_t> all((size_t)kStatsWords * bt->n);
  CUDA_TRY(ctx, cudaMemcpy(all.data(), bt->d_stats, sizeof(int) * kStatsWords * (size_t)bt->n, cudaMemcpyDeviceToHost));
  for (int b = 0; b < bt->n; ++b) {
    out[2 * b] = all[(size_t)kStatsWords * b];
    out[2 * b + 1] = all[(size_t)kStatsWords * b + 1];
  }
  return NDTPSO_OK;
}

int ndtpso_batch_stats_ex(ndtpso_batch* bt, int32_t* out) {
  if (!bt || !out) return NDTPSO_ERR_ARG;
  ndtpso_ctx* ctx = bt->ctx;
  if (!bt->solved) return fail(ctx, NDTPSO_ERR_ARG, "batch_stats before batch_solve");
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  CUDA_TRY(ctx, cudaMemcpy(out, bt->d_stats, sizeof(int) * kStatsWords * (size_t)bt->n, cudaMemcpyDeviceToHost));
  return NDTPSO_OK;
}

void ndtpso_batch_destroy(ndtpso_batch* bt) {
  if (!bt) return;
  ndtpso_ctx* ctx = bt->ctx;
  if (ctx) {
    cudaSetDevice(ctx->device);
    if (bt->results_enqueued && bt->solved)
      cudaEventSynchronize(bt->ev_done);  // this batch is finished; batches queued behind it keep running
    else
      cudaStreamSynchronize(ctx->stream);  // nothing in flight may still read the buffers
    if (bt->ev_done) cudaEventDestroy(bt->ev_done);
    if (bt->ev_up) cudaEventDestroy(bt->ev_up);
    pool_give(ctx->dev_pool, bt->dev, false);
    pool_give(ctx->pin_pool, bt->pin, true);
    for (auto& e : bt->ev)
      if (e) cudaEventDestroy(e);
  }
  delete bt;
}

int ndtpso_align_batch(ndtpso_ctx* ctx, int32_t n, const ndtpso_problem* problems, const ndtpso_pso_config* conf, double* out_pose,
                       double* out_cost) {
  if (!ctx || !conf || (n > 0 && !out_pose)) return fail(ctx, NDTPSO_ERR_ARG, "align_batch: null argument");
  // Large batches on the context's own stream are pipelined: the batch is cut into up to 4 chunks,
  // each staged (host scan + pack of the built cells), uploaded and launched on its own stream, so
  // the host stages chunk k+1 while the GPU already works on chunk k.  Results are identical to the
  // one-shot path (a problem's result does not depend on the batch it is in).
  // auto: three chunks from 128 problems on (tools/onecall_chunks.py, one B200, cfg2: 128 problems 2.71 -> 2.36 ms per call,
  // 256: 3.91 -> 3.21, 512: 6.84 -> 5.71; no difference at 64)
  const int want = ctx->opt_chunks > 0 ? ctx->opt_chunks : (n >= 128 ? 3 : 1);
  const int chunks = (ctx->stream == ctx->own_stream && n >= 64) ? std::min(std::min(want, kMaxChunks), n / 32) : 1;
  if (chunks <= 1) {
    ndtpso_batch* bt = nullptr;
    int rc = batch_create_impl(ctx, n, problems, conf, true, &bt);  // built cells only cross PCIe
    if (rc) return rc;
    rc = ndtpso_batch_solve(bt);
    if (rc == NDTPSO_OK) rc = ndtpso_batch_results(bt, out_pose, out_cost);
    ndtpso_batch_destroy(bt);
    return rc;
  }
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ndtpso_batch* bts[kMaxChunks] = {};
  int lo[kMaxChunks + 1];
  for (int k = 0; k <= chunks; ++k) lo[k] = (int)((int64_t)n * k / chunks);
  int rc = NDTPSO_OK;
  int64_t h2d = 0;
  cudaStream_t saved = ctx->stream;
  for (int k = 0; k < chunks && rc == NDTPSO_OK; ++k) {
    if (!ctx->chunk_stream[k] && cudaStreamCreateWithFlags(&ctx->chunk_stream[k], cudaStreamNonBlocking) != cudaSuccess) {
      rc = fail(ctx, NDTPSO_ERR_CUDA, "cudaStreamCreateWithFlags failed");
      break;
    }
    ctx->stream = ctx->chunk_stream[k];
    rc = batch_create_impl(ctx, lo[k + 1] - lo[k], problems + lo[k], conf, true, &bts[k]);
    if (rc == NDTPSO_OK) {
      h2d += ctx->last_h2d;
      bts[k]->no_cluster = true;
      rc = ndtpso_batch_solve(bts[k]);
    }
  }
  for (int k = 0; k < chunks; ++k) {
    if (!bts[k]) continue;
    ctx->stream = ctx->chunk_stream[k];
    if (rc == NDTPSO_OK) rc = ndtpso_batch_results(bts[k], out_pose + 3 * (size_t)lo[k], out_cost ? out_cost + lo[k] : nullptr);
    ndtpso_batch_destroy(bts[k]);
  }
  ctx->stream = saved;
  ctx->last_h2d = h2d;
  ctx->last_d2h = (int64_t)sizeof(double) * 4 * n;
  return rc;
}

int ndtpso_align_submit(ndtpso_ctx* ctx, int32_t n, const ndtpso_problem* problems, const ndtpso_pso_config* conf, ndtpso_batch** out) {
  if (!ctx || !conf || !out) return fail(ctx, NDTPSO_ERR_ARG, "align_submit: null argument");
  *out = nullptr;
  // Consecutive submissions on the context's own stream go to two alternating streams.  A batch of n CTAs rarely fills a
  // whole number of waves (256 problems on 148 SMs x 2 CTAs leave 40 slots empty, and the SMs with one CTA finish early):
  // with the next batch on another stream its CTAs take those slots at once instead of waiting for the kernel to drain.
  cudaStream_t saved = ctx->stream;
  const bool piped = ctx->stream == ctx->own_stream;
  if (piped) {
    cudaStream_t& ps = ctx->pipe_stream[ctx->pipe_next & 1u];
    if (!ps) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
    ctx->stream = ps;
    ++ctx->pipe_next;
  }
  struct Restore {
    ndtpso_ctx* c;
    cudaStream_t s;
    ~Restore() { c->stream = s; }
  } restore{ctx, saved};
  ndtpso_batch* bt = nullptr;
  int rc = batch_create_impl(ctx, n, problems, conf, true, &bt);  // stage (built cells only) + asynchronous H2D
  if (rc) return rc;
  bt->no_cluster = piped && n >= 64;  // batches overlap: together they fill the GPU
  rc = ndtpso_batch_solve(bt);  // asynchronous launches
  if (rc == NDTPSO_OK && bt->n > 0) {
    // queue the read-back right behind the kernels (the head of the staging buffer is free again:
    // the upload it held completed before the kernels started) and mark the batch's completion
    cudaError_t e = cudaEventCreateWithFlags(&bt->ev_done, cudaEventDisableTiming);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(bt->pin.ptr, bt->d_out, sizeof(double) * 4 * (size_t)bt->n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaEventRecord(bt->ev_done, ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, NDTPSO_ERR_CUDA, cudaGetErrorString(e));
    bt->results_enqueued = (rc == NDTPSO_OK);
  }
  if (rc) {
    ndtpso_batch_destroy(bt);
    return rc;
  }
  *out = bt;
  return NDTPSO_OK;
}

int ndtpso_align_collect(ndtpso_batch* batch, double* out_pose, double* out_cost) {
  if (!batch) return NDTPSO_ERR_ARG;
  const int rc = ndtpso_batch_results(batch, out_pose, out_cost);
  ndtpso_batch_destroy(batch);
  return rc;
}

int ndtpso_cost_batch(ndtpso_ctx* ctx, int32_t n, const ndtpso_problem* problems, int32_t n_poses, const double* poses, double* out_cost) {
  if (!ctx || n_poses < 0 || (n > 0 && n_poses > 0 && (!poses || !out_cost))) return fail(ctx, NDTPSO_ERR_ARG, "cost_batch: null argument");
  if (n == 0 || n_poses == 0) return NDTPSO_OK;
  const size_t pose_bytes = sizeof(double) * 3 * (size_t)n * n_poses;
  const size_t cost_bytes = sizeof(double) * (size_t)n * n_poses;
  ndtpso_batch* bt = nullptr;
  size_t o_up = 0, o_dev = 0;
  int rc = batch_build(ctx, n, problems, nullptr, true, pose_bytes, cost_bytes, &bt, &o_up, &o_dev);
  if (rc) return rc;
  auto cleanup = [&](int code) {
    ndtpso_batch_destroy(bt);
    return code;
  };
  memcpy(static_cast<unsigned char*>(bt->pin.ptr) + o_up, poses, pose_bytes);
  rc = batch_upload(bt);
  if (rc) return cleanup(rc);
  rc = launch_compact(bt);
  if (rc) return cleanup(rc);
  constexpr int NW = 8;
  const int smem = pick_smem(ctx, kCostFixedSmem, bt->need_dyn_smem);
  cudaError_t e = cudaFuncSetAttribute(cost_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin);
  if (e != cudaSuccess) return cleanup(fail(ctx, NDTPSO_ERR_CUDA, cudaGetErrorString(e)));
  unsigned char* d = static_cast<unsigned char*>(bt->dev.ptr);
  cost_kernel<NW><<<n, NW * 32, smem, ctx->stream>>>(bt->d_probs, bt->d_maps, n_poses, reinterpret_cast<const double*>(d + o_up),
                                                     reinterpret_cast<double*>(d + o_dev), smem);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cleanup(fail(ctx, NDTPSO_ERR_CUDA, cudaGetErrorString(e)));
  ctx->launches++;
  e = cudaMemcpyAsync(out_cost, d + o_dev, cost_bytes, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cleanup(fail(ctx, NDTPSO_ERR_CUDA, cudaGetErrorString(e)));
  return cleanup(NDTPSO_OK);
}

int ndtpso_screen_bounds(ndtpso_ctx* ctx, int32_t n, const ndtpso_problem* problems, int32_t n_poses, const double* poses, double* out_lower) {
  if (!ctx || n_poses < 0 || n_poses > 1024 || (n > 0 && n_poses > 0 && (!poses || !out_lower)))
    return fail(ctx, NDTPSO_ERR_ARG, "screen_bounds: null argument or more than 1024 poses");
  if (n == 0 || n_poses == 0) return NDTPSO_OK;
  const size_t pose_bytes = sizeof(double) * 3 * (size_t)n * n_poses;
  const size_t out_bytes = sizeof(double) * (size_t)n * n_poses;
  ndtpso_batch* bt = nullptr;
  size_t o_up = 0, o_dev = 0;
  int rc = batch_build(ctx, n, problems, nullptr, true, pose_bytes, out_bytes, &bt, &o_up, &o_dev);
  if (rc) return rc;
  auto cleanup = [&](int code) {
    ndtpso_batch_destroy(bt);
    return code;
  };
  PsoParams prm{};
  prm.P = n_poses - 1;
  const int saved = ctx->opt_screen;
  ctx->opt_screen = 1;
  const bool ok = bt->all_compact && bt->all_symmetric && screen_params(bt, &prm);
  ctx->opt_screen = saved;
  if (!ok) return cleanup(fail(ctx, NDTPSO_ERR_LIMIT, "screen_bounds: the batch does not qualify for the fp32 screen"));
  const int npts = std::max(bt->max_pts, 1);
  int npt = 1;
  while (npt <= kSlicedMaxNPT && (npts + 32 * npt - 1) / (32 * npt) > 20) ++npt;
  if (npt > kSlicedMaxNPT) return cleanup(fail(ctx, NDTPSO_ERR_LIMIT, "screen_bounds: scan too long"));
  const int nw = std::max((npts + 32 * npt - 1) / (32 * npt), 4);
  const int smem = round16(sliced_smem_bytes(prm.P, nw, 0, bt->max_table_smem, bt->max_n_rec + 1));
  if (smem > ctx->max_smem_optin) return cleanup(fail(ctx, NDTPSO_ERR_LIMIT, "screen_bounds: tables too large for shared memory"));
  memcpy(static_cast<unsigned char*>(bt->pin.ptr) + o_up, poses, pose_bytes);
  rc = batch_upload(bt);
  if (rc) return cleanup(rc);
  rc = launch_compact(bt);
  if (rc) return cleanup(rc);
  unsigned char* d = static_cast<unsigned char*>(bt->dev.ptr);
  const double* d_poses = reinterpret_cast<const double*>(d + o_up);
  double* d_out = reinterpret_cast<double*>(d + o_dev);
#ifdef NDTPSO_DEV_FAST
  npt = 3;
#endif
  switch (npt) {
#ifndef NDTPSO_DEV_FAST
    case 1: rc = launch_screen_bound<1>(bt, nw, smem, prm, n_poses, d_poses, d_out); break;
    case 2: rc = launch_screen_bound<2>(bt, nw, smem, prm, n_poses, d_poses, d_out); break;
    case 4: rc = launch_screen_bound<4>(bt, nw, smem, prm, n_poses, d_poses, d_out); break;
    case 5: rc = launch_screen_bound<5>(bt, nw, smem, prm, n_poses, d_poses, d_out); break;
    case 6: rc = launch_screen_bound<6>(bt, nw, smem, prm, n_poses, d_poses, d_out); break;
#endif
    default: rc = launch_screen_bound<3>(bt, nw, smem, prm, n_poses, d_poses, d_out); break;
  }
  if (rc) return cleanup(rc);
  cudaError_t e = cudaMemcpyAsync(out_lower, d_out, out_bytes, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return cleanup(fail(ctx, NDTPSO_ERR_CUDA, cudaGetErrorString(e)));
  return cleanup(NDTPSO_OK);
}

int ndtpso_exchange_create(ndtpso_ctx* ctx, int32_t world, int32_t rank, int32_t n_per_rank, ndtpso_exchange** out, void* out_ipc_handle) {
  if (!ctx || !out || world < 1 || world > NDTPSO_MAX_RANKS || rank < 0 || rank >= world || n_per_rank < 1)
    return fail(ctx, NDTPSO_ERR_ARG, "exchange_create: bad argument");
  *out = nullptr;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  ndtpso_exchange* ex = new (std::nothrow) ndtpso_exchange();
  if (!ex) return fail(ctx, NDTPSO_ERR_NOMEM, "exchange_create: host allocation failed");
  ex->ctx = ctx;
  ex->world = world;
  ex->rank = rank;
  ex->n_per_rank = n_per_rank;
  ex->half_bytes = align_up(sizeof(double) * 4 * (size_t)world * n_per_rank);
  ex->o_flags = 2 * ex->half_bytes;
  ex->o_done = ex->o_flags + align_up(sizeof(unsigned) * NDTPSO_MAX_RANKS);
  ex->o_err = ex->o_done + kAlign;
  ex->bytes = ex->o_err + kAlign;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ex->base), ex->bytes);  // its own allocation: the IPC handle maps exactly this
  if (e == cudaSuccess) e = cudaMemset(ex->base, 0, ex->bytes);
  if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void**>(&ex->pin), sizeof(double) * 4 * (size_t)world * n_per_rank);
  if (e == cudaSuccess && out_ipc_handle) {
    cudaIpcMemHandle_t h;
    static_assert(sizeof(cudaIpcMemHandle_t) == NDTPSO_IPC_HANDLE_BYTES, "IPC handle size");
    e = cudaIpcGetMemHandle(&h, ex->base);
    if (e == cudaSuccess) memcpy(out_ipc_handle, &h, sizeof h);
  }
  if (e != cudaSuccess) {
    const std::string msg = cudaGetErrorString(e);
    cudaGetLastError();
    ndtpso_exchange_destroy(ex);
    return fail(ctx, NDTPSO_ERR_CUDA, "exchange_create: " + msg);
  }
  ex->peer_base[rank] = ex->base;
  ex->connected = (world == 1);
  *out = ex;
  return NDTPSO_OK;
}

int ndtpso_exchange_connect(ndtpso_exchange* ex, const void* all_handles) {
  if (!ex || !all_handles) return NDTPSO_ERR_ARG;
  ndtpso_ctx* ctx = ex->ctx;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  for (int r = 0; r < ex->world; ++r) {
    if (r == ex->rank || ex->opened[r]) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const unsigned char*>(all_handles) + (size_t)r * NDTPSO_IPC_HANDLE_BYTES, sizeof h);
    void* p = nullptr;
    CUDA_TRY(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ex->peer_base[r] = static_cast<unsigned char*>(p);
    ex->opened[r] = true;
  }
  ex->connected = true;
  return NDTPSO_OK;
}

int ndtpso_exchange_connect_local(ndtpso_exchange* ex, ndtpso_exchange* const* peers) {
  if (!ex || !peers) return NDTPSO_ERR_ARG;
  ndtpso_ctx* ctx = ex->ctx;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  for (int r = 0; r < ex->world; ++r) {
    if (r == ex->rank) continue;
    if (!peers[r] || peers[r]->world != ex->world || peers[r]->n_per_rank != ex->n_per_rank || peers[r]->rank != r)
      return fail(ctx, NDTPSO_ERR_ARG, "exchange_connect_local: peers do not match");
    const int pd = peers[r]->ctx->device;
    if (pd != ctx->device) {
      int can = 0;
      CUDA_TRY(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, pd));
      if (!can) return fail(ctx, NDTPSO_ERR_CUDA, "exchange_connect_local: no peer access between the devices");
      cudaError_t e = cudaDeviceEnablePeerAccess(pd, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, NDTPSO_ERR_CUDA, cudaGetErrorString(e));
      cudaGetLastError();
    }
    ex->peer_base[r] = peers[r]->base;
  }
  ex->connected = true;
  return NDTPSO_OK;
}

int ndtpso_batch_attach_exchange(ndtpso_batch* bt, ndtpso_exchange* ex) {
  if (!bt) return NDTPSO_ERR_ARG;
  if (ex && (ex->ctx != bt->ctx || ex->n_per_rank != bt->n || !ex->connected))
    return fail(bt->ctx, NDTPSO_ERR_ARG, "batch_attach_exchange: exchange not connected, or of another context / batch size");
  bt->ex = ex;
  return NDTPSO_OK;
}

int ndtpso_exchange_wait(ndtpso_exchange* ex) {
  if (!ex || !ex->connected) return NDTPSO_ERR_ARG;
  ndtpso_ctx* ctx = ex->ctx;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const long long timeout_cycles = (long long)ctx->opt_exchange_timeout_ms * ctx->clock_khz;  // a peer that never arrives is an error, not a hang
  exchange_wait_kernel<<<1, 32, 0, ctx->stream>>>(reinterpret_cast<const unsigned*>(ex->base + ex->o_flags), ex->world, ex->epoch,
                                                   timeout_cycles, reinterpret_cast<int*>(ex->base + ex->o_err));
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  return NDTPSO_OK;
}

void* ndtpso_exchange_device_results(ndtpso_exchange* ex) { return ex ? ex->base + (ex->epoch & 1u) * ex->half_bytes : nullptr; }

int ndtpso_exchange_results(ndtpso_exchange* ex, double* out_pose, double* out_cost) {
  if (!ex) return NDTPSO_ERR_ARG;
  ndtpso_ctx* ctx = ex->ctx;
  int rc = ndtpso_exchange_wait(ex);
  if (rc) return rc;
  const size_t rows = (size_t)ex->world * ex->n_per_rank;
  int err = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(ex->pin, ndtpso_exchange_device_results(ex), sizeof(double) * 4 * rows, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaMemcpyAsync(&err, ex->base + ex->o_err, sizeof err, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (err) {
    // report the timeout once: a peer that catches up later must not fail every following call
    CUDA_TRY(ctx, cudaMemsetAsync(ex->base + ex->o_err, 0, sizeof(int), ctx->stream));
    return fail(ctx, NDTPSO_ERR_CUDA, "exchange: rank " + std::to_string(err - 1) + " did not arrive within the time limit");
  }
  for (size_t i = 0; i < rows; ++i) {
    if (out_pose) {
      out_pose[3 * i] = ex->pin[4 * i];
      out_pose[3 * i + 1] = ex->pin[4 * i + 1];
      out_pose[3 * i + 2] = ex->pin[4 * i + 2];
    }
    if (out_cost) out_cost[i] = ex->pin[4 * i + 3];
  }
  return NDTPSO_OK;
}

void ndtpso_exchange_destroy(ndtpso_exchange* ex) {
  if (!ex) return;
  if (ex->ctx) {
    cudaSetDevice(ex->ctx->device);
    cudaStreamSynchronize(ex->ctx->stream);
  }
  for (int r = 0; r < NDTPSO_MAX_RANKS; ++r)
    if (ex->opened[r]) cudaIpcCloseMemHandle(ex->peer_base[r]);
  if (ex->base) cudaFree(ex->base);
  if (ex->pin) cudaFreeHost(ex->pin);
  delete ex;
}

int ndtpso_measure_fp64_peak(ndtpso_ctx* ctx, double* out_tflops) {
  if (!ctx || !out_tflops) return NDTPSO_ERR_ARG;
  CUDA_TRY(ctx, cudaSetDevice(ctx->device));
  const int blocks = ctx->sm_count * 8, threads = 256, iters = 1 << 14;
  double* d = nullptr;
  CUDA_TRY(ctx, cudaMalloc(&d, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best_ms = 1e30;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, ctx->stream);
    fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0) best_ms = std::min(best_ms, (double)ms);
    ctx->launches++;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaError_t e = cudaGetLastError();
  cudaFree(d);
  if (e != cudaSuccess) return fail(ctx, NDTPSO_ERR_CUDA, cudaGetErrorString(e));
  const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
  *out_tflops = flops / (best_ms * 1e-3) / 1e12;
  return NDTPSO_OK;
}

}  // extern "C"

#include "ndtpso_multi_host.inc"
#include "ndtpso_dframes_host.inc"
