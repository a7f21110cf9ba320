/* ndtpso_frames.h — C ABI over the drop-in NDTFrame (ndtpso_slam_b200/shim), for callers that
 * are not C++: Python (ctypes), C, or another FFI.  It exposes what the reference's ROS node does
 * with its frames (src/ndtpso_slam_node.cpp:64-78,186-230): construct, loadLaser, align, update —
 * plus read access to the table the scan matcher consumes.
 * Map building runs on the host; ndtpso_frame_align runs on the GPU (ndtpso_b200.h) and fails
 * with a negative status when no CUDA device is present.
 */
#ifndef NDTPSO_FRAMES_H
#define NDTPSO_FRAMES_H

#include "ndtpso_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ndtpso_frame ndtpso_frame; /* an NDTFrame */

/* NDTFrame::NDTFrame(trans, width, height, cell_side, calculate_cells_params)  (ndtframe.cpp:19) */
ndtpso_frame* ndtpso_frame_new(const double* trans /* [3] */, int width_m, int height_m, double cell_side, int calculate_cells_params);
void ndtpso_frame_free(ndtpso_frame* f);
/* NDTFrame::loadLaser (ndtframe.cpp:144) */
void ndtpso_frame_load_laser(ndtpso_frame* f, const float* ranges, int n, float angle_min, float angle_increment, float range_max);
/* NDTFrame::update (ndtframe.cpp:187) */
void ndtpso_frame_update(ndtpso_frame* f, const double* pose /* [3] */, ndtpso_frame* new_frame);
/* NDTFrame::build (ndtframe.cpp:68) */
void ndtpso_frame_build(ndtpso_frame* f);
/* NDTFrame::resetCells (ndtframe.cpp:208-212): drops every cell's points and window statistics; mean, Sigma^-1 and the
 * built flags stay as they are, like in the reference */
void ndtpso_frame_reset_cells(ndtpso_frame* f);
int ndtpso_frame_is_built(const ndtpso_frame* f);
/* the table cost_function reads: dense (one row per cell) or sparse (built cells only, after build) */
void ndtpso_frame_map_view(const ndtpso_frame* f, ndtpso_map_view* out);
void ndtpso_frame_sparse_map_view(const ndtpso_frame* f, ndtpso_map_view* out);
/* the points cost_function iterates (window slot 0 of every cell, cell order); returns their count */
int ndtpso_frame_scan_points(const ndtpso_frame* f, const double** out_xy);
int64_t ndtpso_frame_point_count(const ndtpso_frame* f);
/* NDTFrame::align (ndtframe.cpp:251): default PSOConfig, process-global rand() stream */
int ndtpso_frame_align(ndtpso_frame* ref_frame, const double* guess /* [3] */, ndtpso_frame* new_frame, double* out_pose /* [3] */);
/* the same with an explicit PSOConfig (conf != NULL) */
int ndtpso_frame_align_conf(ndtpso_frame* ref_frame, const double* guess, ndtpso_frame* new_frame, const ndtpso_pso_config* conf,
                            double* out_pose);
/* glir_pso_optimization (core.h:21-23, core.cpp:118-186): population PSO_POPULATION_SIZE, process-global rand() stream */
int ndtpso_frame_glir(ndtpso_frame* ref_frame, const double* guess /* [3] */, ndtpso_frame* new_frame, unsigned int iterations,
                      const double* deviation /* [3] */, double* out_pose /* [3] */);
/* cost_function (core.h:49) */
int ndtpso_frame_cost(ndtpso_frame* ref_frame, ndtpso_frame* new_frame, const double* pose /* [3] */, double* out_cost);
/* NDTFrame::addPose / NDTFrame::dumpMap (ndtframe.cpp:200-206,268-391): <filename>.pose.csv, .map.csv, .gnuplot */
void ndtpso_frame_add_pose(ndtpso_frame* f, double timestamp, const double* pose /* [3] */);
void ndtpso_frame_dump_map(ndtpso_frame* f, const char* filename);
/* best cost of the most recent align */
/* The map is mirrored in HBM (include/ndtpso_dframes.h) and ndtpso_frame_align / _update use the mirror: true for a frame
 * that has been filled through ndtpso_frame_update only, from its first align on (see shim/include/ndtpso_slam/ndtframe.h). */
int ndtpso_frame_device_resident(const ndtpso_frame* f);
/* bytes the most recent align / update of this frame moved host->device through the mirror (0 without one) */
void ndtpso_frame_last_h2d_bytes(const ndtpso_frame* f, int64_t* align_bytes, int64_t* update_bytes);
/* the device's copy of the table (builds it first; synchronises): mean [C][2], inv_cov [C][4], built [C]; NDTPSO_ERR_ARG without a mirror */
int ndtpso_frame_download_device_map(ndtpso_frame* f, double* mean, double* inv_cov, uint8_t* built);

/* The next n outputs of the process-global rand(), advancing it by exactly n draws: how the drop-in's align takes the
 * 3 + 3P + 6PI numbers the reference would draw (core.cpp:14,58-69,84) — from glibc's own state table, without n trips through
 * its lock; falls back to n calls of rand() wherever that is not possible (NDTPSO_SHIM_FAST_RAND=0 forces the fallback). */
void ndtpso_frame_draw_rand(int32_t* out, int64_t n);
/* What a failure of the device path does (pso_set_failure_handler of the drop-in's core.h; the reference's CPU path cannot fail).
 * 0 (default): the failing call reports an error (C++: throws std::runtime_error).  1: the failure is recorded for
 * ndtpso_frame_last_error, the call succeeds with its neutral result — align / glir return the caller's guess, cost returns 0 — and
 * ndtpso_frame_last_cost is NaN until the next successful call. */
void ndtpso_frame_set_failure_mode(int keep_going);
double ndtpso_frame_last_cost(void);
const char* ndtpso_frame_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
