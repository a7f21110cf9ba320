/* ndtpso_dframes.h — C ABI of the DEVICE-RESIDENT reference frames.
 *
 * the two steps either side of the scan matcher in the reference's
 * per-scan callback (src/ndtpso_slam_node.cpp:177-244),
 *
 *     current_frame_->loadLaser(ranges, angle_min, angle_increment, range_max)    :186
 *     current_pose_ = ref_frame_->align(previous_pose_, current_frame_)           :194
 *     ref_frame_->update(current_pose_, current_frame_)                           :198
 *
 * run here as CUDA kernels on maps that never leave HBM.  One `ndtpso_dframes` object holds n
 * independent reference frames (one per tracked robot / replayed trajectory), each with its NDT
 * cells' sliding-window statistics, its dense (mean, Sigma^-1, built) table, the compact table the
 * PSO kernel stages, the current scan's points and the align() bookkeeping (s_iter, s_prev_pose,
 * s_pose_diff, ndtframe.cpp:251-266).  Per scan only the ranges cross PCIe host->device
 * (4 bytes per beam) and the solved poses device->host (32 bytes per frame).
 *
 * Reference functions replaced (file:line under /root/reference):
 *   NDTFrame::NDTFrame            lib/ndtpso_slam/ndtframe.cpp:19-66    ndtpso_dframes_create
 *   NDTFrame::loadLaser           lib/ndtpso_slam/ndtframe.cpp:144-185  ndtpso_dframes_load_laser
 *     index_to_angle, laser_to_point   include/ndtpso_slam/core.h:40-47
 *   NDTFrame::update / addPoint   lib/ndtpso_slam/ndtframe.cpp:187-198,215-235   ndtpso_dframes_update
 *     NDTCell::addPoint           lib/ndtpso_slam/ndtcell.cpp:21-34
 *   NDTFrame::build               lib/ndtpso_slam/ndtframe.cpp:68-117   ndtpso_dframes_build
 *     NDTCell::build              lib/ndtpso_slam/ndtcell.cpp:36-68
 *     NDTCell::s_calc_covar_inverse    lib/ndtpso_slam/ndtcell.cpp:93-111
 *   NDTFrame::align               lib/ndtpso_slam/ndtframe.cpp:251-266  ndtpso_dframes_align
 *   NDTPSONode::scan_matcher_     src/ndtpso_slam_node.cpp:177-244      ndtpso_dframes_track_step
 *
 * Numerics.  update/build use only IEEE add, multiply, divide and square root without
 * contraction, in the reference's order of operations: given the same scan points and the same
 * (x, y, cos, sin) of the pose, the tables are bit-identical to the reference's.  loadLaser and a
 * device-resident pose need cos/sin of a double evaluated on the GPU, which may differ from glibc's
 * in the last place: scan points then agree to <= 2 ulp.
 *
 * All calls are asynchronous on the context's stream unless they return data to the host.
 * Status codes are those of ndtpso_b200.h.  There is no CPU path behind this ABI.
 */
#ifndef NDTPSO_DFRAMES_H
#define NDTPSO_DFRAMES_H

#include "ndtpso_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ndtpso_dframes ndtpso_dframes;

typedef struct ndtpso_dframes_config {
  int32_t n_frames;     /* independent reference frames */
  int32_t width_m;      /* NDTFrame::width  (unsigned short metres, ndtframe.h:32) */
  int32_t height_m;     /* NDTFrame::height */
  int32_t max_beams;    /* longest scan (beams per LaserScan message) */
  double cell_side;     /* NDTFrame::cell_side */
  double scan_cell_side;/* cell side of the scan frame the points are binned in before matching:
                           <= 0 = one cell of the frame's size, what the node uses after its first
                           scan (ndtpso_slam_node.cpp:229-230); > 0 orders the scan points
                           cell-index-major like a frame of that cell side would (core.cpp:33-36) */
  int32_t max_cells;    /* cells per frame that may ever receive a point (pool capacity); 0 = auto */
  int32_t window_points;/* points kept per cell for the sliding window (ring, power of two); the
                           statistics are exact while the last NDT_WINDOW_SIZE slots fit; 0 = auto */
  float laser_ignore_epsilon; /* NDTPSOConfig::laserIgnoreEpsilon (config.h:6,44), default 0.1f */
  int32_t flags;        /* NDTPSO_DF_* creation flags, 0 by default */
} ndtpso_dframes_config;

/* creation flags */
enum {
  NDTPSO_DF_NO_CLUSTER = 1 /* never spread one frame's match over a thread-block cluster (the low-latency form for few frames):
                              for callers that keep several groups of frames in flight on their own streams */
};

/* per-frame status bits (ndtpso_dframes_status) */
enum {
  NDTPSO_DF_CELL_POOL_FULL = 1,   /* more than max_cells cells received points: later cells were dropped */
  NDTPSO_DF_WINDOW_TRUNCATED = 2, /* a re-opened window slot's old points had already left the ring */
  NDTPSO_DF_INDEX_PAST_END = 4,   /* a point's flat cell index fell past the table (undefined in the reference): dropped */
  NDTPSO_DF_IRREGULAR_SIGMA = 8   /* some Sigma^-1 of the table as the last build left it is not finite / symmetric / PSD: the generic PSO kernel is used (recomputed by every build; the other bits are sticky) */
};

/* how ndtpso_dframes_align draws its random numbers */
enum {
  NDTPSO_RNG_SEEDED = 0,    /* srand(seeds[b]) before every align: the batch mode of ndtpso_b200.h */
  NDTPSO_RNG_CONTINUE = 1,  /* the reference as shipped: never seeded, every frame owns one glibc rand()
                               stream (default seed 1) that continues across align calls */
  NDTPSO_RNG_HOST = 2       /* the numbers are the caller's: ndtpso_dframes_align_streams */
};

void ndtpso_dframes_config_default(ndtpso_dframes_config* cfg);
int ndtpso_dframes_create(ndtpso_ctx* ctx, const ndtpso_dframes_config* cfg, ndtpso_dframes** out);
void ndtpso_dframes_destroy(ndtpso_dframes* df);
/* bytes of HBM held by the object */
int64_t ndtpso_dframes_device_bytes(const ndtpso_dframes* df);

/* NDTFrame::loadLaser for the scan frame of every reference frame.
 * ranges: [n_frames][n_beams] float, host (pinned memory is copied asynchronously);
 * scan_trans: NULL or [n_frames][3], the scan frame's s_trans (the node passes its initial pose);
 * a frame whose trans isZero(1e-6) is not transformed (ndtframe.cpp:151-153). */
int ndtpso_dframes_load_laser(ndtpso_dframes* df, const float* ranges, int32_t n_beams, float angle_min, float angle_increment,
                              float range_max, const double* scan_trans);
/* The same for a scan frame of cell side `scan_cell_side` (<= 0: one cell of the frame's size), whatever the object's
 * scan_cell_side: the drop-in NDTFrame passes the cell side of the frame the caller loaded the scan into. */
int ndtpso_dframes_load_laser_binned(ndtpso_dframes* df, const float* ranges, int32_t n_beams, float angle_min, float angle_increment,
                                     float range_max, const double* scan_trans, double scan_cell_side);
/* The same state set directly: points_xy [n_frames][stride_points][2] host, n_points [n_frames]. */
int ndtpso_dframes_set_scan_points(ndtpso_dframes* df, const double* points_xy, const int32_t* n_points, int32_t stride_points);

/* NDTFrame::update(pose, scan frame).  poses: [n_frames][3] host, or NULL = the poses the last
 * ndtpso_dframes_align left on the device. */
int ndtpso_dframes_update(ndtpso_dframes* df, const double* poses);
/* NDTFrame::build() + compaction of the table for the PSO kernel.  Idempotent in the reference's
 * sense: it re-runs NDTCell::build for every created cell (ndtframe.cpp:73-77). */
int ndtpso_dframes_build(ndtpso_dframes* df);

/* NDTFrame::align(guess, scan frame) for every frame, with an explicit PSO configuration
 * (conf = NULL: the reference's defaults, what its 2-argument align() always runs).
 * guess: [n_frames][3] host, or NULL = each frame's previous pose (the node's previous_pose_).
 * Builds the map first when points were added since the last build (core.cpp:27-28).
 * out_pose / out_cost: host, may be NULL (results stay on the device; no synchronisation). */
int ndtpso_dframes_align(ndtpso_dframes* df, const double* guess, const ndtpso_pso_config* conf, int32_t rng_mode,
                         const uint32_t* seeds /* [n_frames], NDTPSO_RNG_SEEDED only */, double* out_pose /* [n][3] */,
                         double* out_cost /* [n] */);

/* The same with the random numbers drawn by the caller: rand_streams [n_frames][ndtpso_rand_draws(conf)] raw std::rand()
 * outputs in call order.  This is what the drop-in NDTFrame::align uses: it draws from the process-global std::rand(), so a
 * sequence of align() calls stays in lock-step with a CPU build of the reference whatever else the process does with rand(). */
int ndtpso_dframes_align_streams(ndtpso_dframes* df, const double* guess, const ndtpso_pso_config* conf, const int32_t* rand_streams,
                                 double* out_pose /* [n][3] */, double* out_cost /* [n] */);

/* One pass of NDTPSONode::scan_matcher_ for every frame: loadLaser; on the first call the pose is
 * `initial_poses` (host [n][3], may be NULL = zero) and no matching is done (:188-192), afterwards
 * align(previous pose); then update(pose).  out_pose: host [n][3]. */
int ndtpso_dframes_track_step(ndtpso_dframes* df, const float* ranges, int32_t n_beams, float angle_min, float angle_increment,
                              float range_max, const double* initial_poses, const ndtpso_pso_config* conf, int32_t rng_mode,
                              const uint32_t* seeds, double* out_pose, double* out_cost);

/* Multi-GPU: robots are split over ranks (one process and one ndtpso_dframes per GPU).  With an exchange attached
 * (ndtpso_exchange_*, ndtpso_b200.h; n_per_rank == n_frames) every later align / track_step also stores its poses into
 * the gathered buffer of every rank from the PSO kernel's epilogue; ndtpso_exchange_wait / _results then give every
 * rank all robots' poses.  ex = NULL detaches. */
int ndtpso_dframes_attach_exchange(ndtpso_dframes* df, ndtpso_exchange* ex);

/* ---- read-back (synchronises) ---------------------------------------------------- */
/* dense table of one frame: mean [C][2], inv_cov [C][4], built [C], C = w_cells*h_cells */
int ndtpso_dframes_download_map(ndtpso_dframes* df, int32_t frame, double* mean, double* inv_cov, uint8_t* built);
/* the scan points of one frame in matching order; *n_points in: capacity, out: count */
int ndtpso_dframes_download_scan(ndtpso_dframes* df, int32_t frame, double* points_xy, int32_t* n_points);
/* out_i = {w_cells, h_cells, n_cells, created cells, built cells, scan points, align calls, status bits} */
int ndtpso_dframes_info(ndtpso_dframes* df, int32_t frame, int32_t* out_i /* [8] */);
int ndtpso_dframes_status(ndtpso_dframes* df, int32_t* flags /* [n_frames] */);
/* counters of the last align, per frame: {rounds, gbest updates, fp64 cost evaluations, evaluations settled by the fp32 screen} */
int ndtpso_dframes_pso_stats(ndtpso_dframes* df, int32_t* out /* [n_frames][4] */);
/* device time in ms of the last call's kernels: {loadLaser, update, build (+compaction), rand() stream, PSO} */
int ndtpso_dframes_kernel_times(ndtpso_dframes* df, double* out_ms /* [5] */);

#ifdef __cplusplus
}
#endif
#endif /* NDTPSO_DFRAMES_H */
