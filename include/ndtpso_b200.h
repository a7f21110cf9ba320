/* ndtpso_b200.h — C ABI of the B200-native PSO/NDT scan-matching hot path.
 *
 * This is the drop-in boundary for libndtpso_slam's scan matcher.  The reference has
 * no plugin mechanism; the seam is link-time: its ROS node calls
 *     current_pose_ = ref_frame_->align(previous_pose_, current_frame_);
 *                                   (src/ndtpso_slam_node.cpp:194)
 * and NDTFrame::align (lib/ndtpso_slam/ndtframe.cpp:251-266) calls
 *     pso_optimization(guess, this, new_frame, deviation)         (ndtframe.cpp:257)
 * declared at include/ndtpso_slam/core.h:16-19, which evaluates
 *     cost_function(trans, ref_frame, new_frame)                   (core.h:49-50)
 * P+1+P*I times.  The entry points below are what a C/C++/ctypes binding for that path
 * binds: plain pointers and sizes, no Eigen, STL or torch types.  INTEGRATION.md shows
 * the few lines a maintainer adds to lib/ndtpso_slam/core.cpp to route through them.
 *
 * All floating point is IEEE fp64, as in the reference.  Every function returns an int
 * status (0 = NDTPSO_OK, negative = error) and never throws or aborts; the reference
 * path itself cannot fail, so a caller may treat any non-zero
 * status as fatal.  There is NO CPU fallback behind this ABI: without a CUDA device
 * ndtpso_ctx_create fails with NDTPSO_ERR_NODEVICE.
 *
 * Threading: one in-flight call per context; contexts are independent (one per GPU).
 */
#ifndef NDTPSO_B200_H
#define NDTPSO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDTPSO_ABI_VERSION 1

enum {
  NDTPSO_OK = 0,
  NDTPSO_ERR_ARG = -1,      /* null pointer, negative size, unsupported configuration */
  NDTPSO_ERR_CUDA = -2,     /* a CUDA runtime call or kernel failed; see ndtpso_last_error */
  NDTPSO_ERR_NOMEM = -3,    /* host or device allocation failed */
  NDTPSO_ERR_NODEVICE = -4, /* no usable CUDA device (the product has no CPU path) */
  NDTPSO_ERR_LIMIT = -5     /* population too large for one CTA's shared memory */
};

typedef struct ndtpso_ctx ndtpso_ctx;     /* one GPU, one stream, cached device/pinned arenas */
typedef struct ndtpso_batch ndtpso_batch; /* a batch of problems resident in HBM */

/* Mirror of `struct PSOConfig` (include/ndtpso_slam/config.h:27-38), field for field.
 * num_threads is accepted and ignored: the device path is deterministic and equals the
 * reference run with num_threads = 1 (its OpenMP mode is racy). */
typedef struct ndtpso_pso_config {
  int32_t iterations;   /* PSOConfig::iterations      default 50  (PSO_ITERATIONS) */
  int32_t population;   /* PSOConfig::populationSize  default 30  (PSO_POPULATION_SIZE) */
  int32_t num_threads;  /* PSOConfig::num_threads     default -1 */
  int32_t variant;      /* NDTPSO_VARIANT_*: which of the reference's two optimisers runs (0 = pso_optimization);
                           occupies the padding of the reference's PSOConfig, so the layout is unchanged */
  double w;             /* coeff.w          default .8 */
  double c1;            /* coeff.c1         default 2. */
  double c2;            /* coeff.c2         default 2. */
  double w_dumping;     /* coeff.w_dumping  default 1. */
} ndtpso_pso_config;

/* ndtpso_pso_config::variant */
enum {
  NDTPSO_VARIANT_PSO = 0,  /* pso_optimization       lib/ndtpso_slam/core.cpp:50-116: what NDTFrame::align() runs */
  NDTPSO_VARIANT_GLIR = 1  /* glir_pso_optimization  lib/ndtpso_slam/core.cpp:118-186 (decl core.h:21-22): the GLIR-PSO variant the
                              reference ships as "UNTESTED" and never calls.  omega, c1 = c2 and a best-position ratio are
                              recomputed per particle from the best costs; w, c1, c2, w_dumping are ignored; the reference
                              fixes its population to PSO_POPULATION_SIZE (30), here `population` is honoured.  Runs on the
                              generic warp-per-particle kernel.  It draws 3(P + 2) + 6PI random numbers. */
};

/* Fills *conf with the reference defaults (what 2-argument NDTFrame::align() runs). */
void ndtpso_pso_config_default(ndtpso_pso_config* conf);

/* Read-only view of the reference frame's NDT table: exactly the fields cost_function
 * reads through NDTFrame::getCellIndex (ndtframe.cpp:240-249) and
 * NDTCell::normalDistribution (ndtcell.cpp:70-78).  Cell index = ix + w_cells*iy.
 *
 *   dense  form (n_sparse <  0): mean/inv_cov/built have w_cells*h_cells rows, row i = cell i;
 *   sparse form (n_sparse >= 0): mean/inv_cov have n_sparse rows, row r belongs to cell
 *                                cell_index[r] (strictly ascending), every listed cell is
 *                                built, every other cell is not; `built` is ignored.
 */
typedef struct ndtpso_map_view {
  int32_t w_cells;       /* NDTFrame::widthNumOfCells  (ndtframe.cpp:27) */
  int32_t h_cells;       /* NDTFrame::heightNumOfCells (ndtframe.cpp:28) */
  double width_m;        /* NDTFrame::width  (uint16 metres, ndtframe.h:32) */
  double height_m;       /* NDTFrame::height */
  double cell_side;      /* NDTFrame::cell_side */
  double x_min, x_max;   /* NDTFrame::s_x_min, s_x_max (ndtframe.cpp:57-58) */
  double y_min, y_max;   /* NDTFrame::s_y_min, s_y_max (ndtframe.cpp:64-65) */
  const double* mean;    /* [rows][2]  NDTCell::mean        (ndtcell.h:73) */
  const double* inv_cov; /* [rows][4]  NDTCell::s_inv_covar (ndtcell.h:66), row-major 00,01,10,11 */
  const uint8_t* built;  /* [w_cells*h_cells] NDTCell::built (ndtcell.h:76); dense form only */
  int32_t n_sparse;      /* < 0: dense form */
  int32_t reserved;      /* must be 0 */
  const int32_t* cell_index; /* [n_sparse], sparse form only */
} ndtpso_map_view;

/* One scan-match problem = one pso_optimization call (core.cpp:50-116). */
typedef struct ndtpso_problem {
  ndtpso_map_view map;      /* `ref_frame`; problems may share a table (same pointers) */
  const double* points_xy;  /* [n_points][2]: new_frame->cells[*].points_vector[0] in the
                               iteration order of core.cpp:33-36 */
  int32_t n_points;
  uint32_t seed;            /* used when rand_stream == NULL: the random numbers are those of
                               glibc srand(seed); rand(); rand(); ... */
  double guess[3];          /* initial_guess (x, y, theta) */
  double deviation[3];      /* deviation (core.h:18) */
  const int32_t* rand_stream; /* NULL, or rand_count raw std::rand() outputs drawn by the host
                               in call order (drop-in mode: keeps the process-global stream
                               in lock-step with the CPU build) */
  int64_t rand_count;       /* >= ndtpso_rand_draws(conf) when rand_stream != NULL */
} ndtpso_problem;

/* ---- context ------------------------------------------------------------------- */
int ndtpso_abi_version(void);
int ndtpso_device_count(void);
int ndtpso_ctx_create(int device, ndtpso_ctx** out);
void ndtpso_ctx_destroy(ndtpso_ctx* ctx);
/* Launch on `cuda_stream` (a cudaStream_t) instead of the context's own stream. */
int ndtpso_ctx_set_stream(ndtpso_ctx* ctx, void* cuda_stream);
const char* ndtpso_last_error(const ndtpso_ctx* ctx);

enum {
  NDTPSO_OPT_WARPS_PER_CTA = 1, /* 0 = auto */
  NDTPSO_OPT_SMEM_BYTES = 2,    /* dynamic shared memory per CTA; 0 = auto */
  NDTPSO_OPT_CLUSTER = 3,       /* CTAs cooperating on one problem; 0 = auto */
  NDTPSO_OPT_KERNEL = 4,        /* 0 = auto, 1 = warp-per-particle (generic), 2 = point-sliced */
  NDTPSO_OPT_POINTS_PER_THREAD = 5, /* point-sliced kernel: scan points held per thread; 0 = auto */
  NDTPSO_OPT_CANDIDATE_BATCH = 6,  /* point-sliced kernel: candidates scored together (1, 2, 4); 0 = auto */
  NDTPSO_OPT_PIPELINE_CHUNKS = 7,  /* ndtpso_align_batch: chunks staged/uploaded/solved on separate streams (1..4; default 0 = auto: 3 from 128 problems on) */
  NDTPSO_OPT_EXCHANGE_TIMEOUT_MS = 8, /* ndtpso_exchange_wait: give up after this long (default 10000) */
  NDTPSO_OPT_HOT_CHUNK = 9, /* point-sliced kernel: particles speculated per round while gbest improves often; -1 auto, 0 = whole swarm */
  NDTPSO_OPT_HOST_THREADS = 11, /* host threads (caller included) that stage this context's batches; 0 = auto: the cores the process may use, at most 8 */
  NDTPSO_OPT_SCREEN = 10    /* point-sliced kernel: fp32 lower-bound screen before the fp64 cost (results identical either way); -1 / 1 on
                               whenever the batch qualifies and its tables fit shared memory, 0 off */
};
int ndtpso_ctx_set_option(ndtpso_ctx* ctx, int option, int64_t value);

/* Number of std::rand() calls one pso_optimization makes: 3 + 3P + 6PI (core.cpp:14,58-69,84). */
int64_t ndtpso_rand_draws(const ndtpso_pso_config* conf);

/* ---- the hot path --------------------------------------------------------------- */

/* Replaces pso_optimization (core.h:16-19) for n independent problems: host buffers in,
 * host buffers out.  out_pose[b] = global_best.best_position, out_cost[b] =
 * global_best.best_cost (the reference only prints it, core.cpp:111-114). out_cost may be NULL. */
int ndtpso_align_batch(ndtpso_ctx* ctx, int32_t n, const ndtpso_problem* problems, const ndtpso_pso_config* conf,
                       double* out_pose /* [n][3] */, double* out_cost /* [n] */);

/* The same call split in two so that a throughput-oriented caller can overlap the host-side staging
 * of batch k+1 with the GPU work of batch k:  submit = stage + H2D + kernel launches (asynchronous),
 * collect = D2H of the poses + synchronise + release.  Consecutive submissions run on two alternating
 * streams, so the kernels of batch k+1 fill the SMs batch k leaves idle while it drains; collect waits
 * for its own batch only.  Results do not depend on what else is in flight.  Any number of batches may be in flight;
 * with THREE (submit k+2 before collecting k) the GPU always has two batches queued and the end-to-end rate equals that
 * of batches resident in HBM (tools/e2e_depth.py: 1 -> 65.6 k, 2 -> 117.5 k, 3 -> 126.0 k scan-matches/s on one B200). */
int ndtpso_align_submit(ndtpso_ctx* ctx, int32_t n, const ndtpso_problem* problems, const ndtpso_pso_config* conf, ndtpso_batch** out);
int ndtpso_align_collect(ndtpso_batch* batch, double* out_pose /* [n][3] */, double* out_cost /* [n] */);

/* Replaces cost_function (core.h:49-50): cost of `n_poses` candidate poses per problem. */
int ndtpso_cost_batch(ndtpso_ctx* ctx, int32_t n, const ndtpso_problem* problems, int32_t n_poses,
                      const double* poses /* [n][n_poses][3] */, double* out_cost /* [n][n_poses] */);

/* The rigorous fp32 lower bound the point-sliced PSO kernel screens its candidates with (NDTPSO_OPT_SCREEN), for
 * `n_poses` (<= 1024) candidate poses per problem: out_lower[b][k] <= cost_function(poses[b][k]) always.  The PSO never
 * returns these; the entry point exists so that the bound can be checked pose by pose.  NDTPSO_ERR_LIMIT when the batch
 * does not qualify for the screen (frames that are not square, whole, power-of-two cells; irregular Sigma^-1). */
int ndtpso_screen_bounds(ndtpso_ctx* ctx, int32_t n, const ndtpso_problem* problems, int32_t n_poses,
                         const double* poses /* [n][n_poses][3] */, double* out_lower /* [n][n_poses] */);

/* ---- the same path with the batch kept resident in HBM -------------------------- */
/* upload (H2D, asynchronous on the context's stream) */
int ndtpso_batch_create(ndtpso_ctx* ctx, int32_t n, const ndtpso_problem* problems, const ndtpso_pso_config* conf,
                        ndtpso_batch** out);
/* enqueue the kernels; asynchronous; may be called repeatedly (each call re-solves from scratch) */
int ndtpso_batch_solve(ndtpso_batch* batch);
/* device pointer to [n][4] fp64 = (x, y, theta, cost) per problem, valid after solve completes */
void* ndtpso_batch_device_results(ndtpso_batch* batch);
/* D2H + synchronise; either output may be NULL */
int ndtpso_batch_results(ndtpso_batch* batch, double* out_pose /* [n][3] */, double* out_cost /* [n] */);
/* per-problem counters after a solve: out[b] = {rounds, gbest_updates} */
int ndtpso_batch_stats(ndtpso_batch* batch, int32_t* out /* [n][2] */);
/* out[b] = {rounds, gbest_updates, fp64 cost evaluations, evaluations settled by the fp32 screen alone} */
int ndtpso_batch_stats_ex(ndtpso_batch* batch, int32_t* out /* [n][4] */);
/* device time of the last solve's kernels in ms, measured with CUDA events on the launch stream:
 * out_ms = {table compaction (K0), rand() stream (K1), PSO (K2)}; synchronises on the solve */
int ndtpso_batch_kernel_times(ndtpso_batch* batch, double* out_ms /* [3] */);
void ndtpso_batch_destroy(ndtpso_batch* batch);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
int64_t ndtpso_ctx_launch_count(const ndtpso_ctx* ctx);
/* bytes copied host->device by the most recent upload and device->host by the most recent results
 * read of this context (what bench.py reports as h2d/d2h bytes per step) */
int ndtpso_ctx_last_transfer_bytes(const ndtpso_ctx* ctx, int64_t* h2d, int64_t* d2h);
int ndtpso_ctx_synchronize(ndtpso_ctx* ctx);

/* ---- multi-GPU: the solved poses of every rank on every rank, without a collective call ------ */
/* The path shards over independent problems, one shard per GPU (one process per GPU); its only exchange
 * is the gather of the solved poses at the end.  An exchange replaces that collective by peer stores
 * fused into the PSO kernel's epilogue: the CTA that solved a problem writes its (x, y, theta, cost)
 * into the gathered buffer of every rank over NVLink (CUDA IPC mappings), and the last CTA of a launch
 * raises its rank's arrival flag everywhere.  ndtpso_exchange_wait then enqueues a small kernel that
 * returns once all ranks' flags of the current epoch have arrived (bounded spin, never a hang).
 * Buffers are double-buffered by epoch, so a rank may be one solve ahead of its peers. */
typedef struct ndtpso_exchange ndtpso_exchange;
#define NDTPSO_IPC_HANDLE_BYTES 64
#define NDTPSO_MAX_RANKS 8
/* Allocates this rank's gathered buffer ([2][world * n_per_rank][4] fp64) and writes the 64-byte handle
 * the other ranks need to map it. */
int ndtpso_exchange_create(ndtpso_ctx* ctx, int32_t world, int32_t rank, int32_t n_per_rank, ndtpso_exchange** out,
                           void* out_ipc_handle /* [NDTPSO_IPC_HANDLE_BYTES] */);
/* all_handles: [world][NDTPSO_IPC_HANDLE_BYTES], every rank's handle in rank order (gathered by the
 * caller, e.g. with torch.distributed.all_gather).  Maps the peers' buffers. */
int ndtpso_exchange_connect(ndtpso_exchange* ex, const void* all_handles);
/* The same for ranks that live in ONE process (one context per GPU): peers[r] = rank r's exchange. */
int ndtpso_exchange_connect_local(ndtpso_exchange* ex, ndtpso_exchange* const* peers);
/* Every later ndtpso_batch_solve of `batch` (n == n_per_rank problems) also publishes its results through `ex`;
 * ex = NULL detaches. */
int ndtpso_batch_attach_exchange(ndtpso_batch* batch, ndtpso_exchange* ex);
/* Enqueue the wait for the epoch of the most recent solve on the context's stream (asynchronous). */
int ndtpso_exchange_wait(ndtpso_exchange* ex);
/* device pointer to the gathered results of that epoch, [world * n_per_rank][4] fp64, valid after the wait */
void* ndtpso_exchange_device_results(ndtpso_exchange* ex);
/* wait + D2H + synchronise: out_pose [world * n_per_rank][3], out_cost [world * n_per_rank]; either may be NULL */
int ndtpso_exchange_results(ndtpso_exchange* ex, double* out_pose, double* out_cost);
void ndtpso_exchange_destroy(ndtpso_exchange* ex);

/* ---- several GPUs behind one call: single process, one context per device ------------------------ */
/* The path shards over independent problems: device g of G gets the contiguous block
 * [g*ceil(n/G), min(n, (g+1)*ceil(n/G))).  A C or C++ caller — the node that calls NDTFrame::align
 * (src/ndtpso_slam_node.cpp:194), or a batch replay tool — hands over all n problems in one call; every device
 * stages, uploads and solves its shard concurrently (one host thread and one staging pool per device), and the
 * results come back in problem order.  Results are bit-identical to the single-GPU entry points. */
typedef struct ndtpso_multi ndtpso_multi;
typedef struct ndtpso_multi_batch ndtpso_multi_batch;
/* devices: n_devices CUDA ordinals, or NULL for 0 .. n_devices-1; n_devices <= NDTPSO_MAX_RANKS */
int ndtpso_multi_create(const int32_t* devices, int32_t n_devices, ndtpso_multi** out);
void ndtpso_multi_destroy(ndtpso_multi* m);
int32_t ndtpso_multi_size(const ndtpso_multi* m);
ndtpso_ctx* ndtpso_multi_ctx(ndtpso_multi* m, int32_t i); /* device i's context (options, launch counts) */
const char* ndtpso_multi_last_error(const ndtpso_multi* m);
/* ndtpso_align_batch over all devices: host buffers in, host buffers out */
int ndtpso_align_batch_multi(ndtpso_multi* m, int32_t n, const ndtpso_problem* problems, const ndtpso_pso_config* conf,
                             double* out_pose /* [n][3] */, double* out_cost /* [n] */);
/* the throughput form (ndtpso_align_submit / ndtpso_align_collect): keep three batches in flight */
int ndtpso_align_submit_multi(ndtpso_multi* m, int32_t n, const ndtpso_problem* problems, const ndtpso_pso_config* conf, ndtpso_multi_batch** out);
int ndtpso_align_collect_multi(ndtpso_multi_batch* batch, double* out_pose /* [n][3] */, double* out_cost /* [n] */);
/* the shards kept resident in HBM (ndtpso_batch_create / _solve).  When n is a multiple of the device count the shards'
 * results are also exchanged on the device side, fused into the PSO kernel's epilogue (ndtpso_exchange_*, connected with
 * ndtpso_exchange_connect_local): after a solve EVERY device holds all n results. */
int ndtpso_multi_batch_create(ndtpso_multi* m, int32_t n, const ndtpso_problem* problems, const ndtpso_pso_config* conf, ndtpso_multi_batch** out);
int ndtpso_multi_batch_solve(ndtpso_multi_batch* batch);   /* asynchronous launches on every device */
int ndtpso_multi_batch_results(ndtpso_multi_batch* batch, double* out_pose /* [n][3] */, double* out_cost /* [n] */); /* synchronises */
/* device pointer, on device i, to the gathered [n][4] fp64 results of the last solve (NULL unless the shards are exchanged);
 * valid after ndtpso_multi_batch_results or after synchronising device i's context */
void* ndtpso_multi_batch_device_results(ndtpso_multi_batch* batch, int32_t i);
void ndtpso_multi_batch_destroy(ndtpso_multi_batch* batch);

/* ---- device self-measurement (roofline denominators MEASURED_PEAKS.json lacks) --- */
/* Sustained fp64 FMA throughput of this GPU in TFLOP/s (2 flop per DFMA). */
int ndtpso_measure_fp64_peak(ndtpso_ctx* ctx, double* out_tflops);

#ifdef __cplusplus
}
#endif
#endif /* NDTPSO_B200_H */
