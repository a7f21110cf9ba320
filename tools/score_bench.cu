// Phase-B microbenchmark library (tools only; not part of the product build).
// Builds the whole C ABI plus ndtpso_bench_score(), which times score_bench_kernel for one
// (points-per-thread, candidate-batch, launch-bounds, code variant) configuration.
#define NDTPSO_PHASE_TIMING 1
#include "../ndtpso_slam_b200/csrc/ndtpso_capi.cu"

namespace {
template <int NPT, int JB, int VAR, int MAXT, int MINB>
int run_cfg(ndtpso_batch* bt, int nw, int grid, int ncand, int reps, double* out_ms, int* out_regs, int* out_ctas_per_sm) {
  ndtpso_ctx* ctx = bt->ctx;
  auto kern = score_bench_kernel<NPT, JB, VAR, MAXT, MINB>;
  const int smem = round16(sliced_smem_bytes(ncand - 1, nw, 0, bt->max_table_smem));  // PW = nw (one CTA per problem)
  CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->max_smem_optin));
  cudaFuncAttributes fa;
  CUDA_TRY(ctx, cudaFuncGetAttributes(&fa, kern));
  *out_regs = fa.numRegs;
  CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(out_ctas_per_sm, kern, nw * 32, smem));
  double* d_out = nullptr;
  CUDA_TRY(ctx, cudaMalloc(&d_out, sizeof(double) * grid));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 1e30;
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0, ctx->stream);
    kern<<<grid, nw * 32, smem, ctx->stream>>>(bt->d_probs, bt->d_maps, bt->n, ncand, reps, d_out);
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it > 0) best = std::min(best, (double)ms);
  }
  cudaError_t e = cudaGetLastError();
  cudaFree(d_out);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (e != cudaSuccess) return fail(ctx, NDTPSO_ERR_CUDA, cudaGetErrorString(e));
  *out_ms = best;
  return NDTPSO_OK;
}

template <int NPT, int JB, int MAXT, int MINB>
int run_var(int var, ndtpso_batch* bt, int nw, int grid, int ncand, int reps, double* ms, int* regs, int* occ) {
  switch (var) {
    case 0: return run_cfg<NPT, JB, 0, MAXT, MINB>(bt, nw, grid, ncand, reps, ms, regs, occ);
    case 1: return run_cfg<NPT, JB, 1, MAXT, MINB>(bt, nw, grid, ncand, reps, ms, regs, occ);
    case 2: return run_cfg<NPT, JB, 2, MAXT, MINB>(bt, nw, grid, ncand, reps, ms, regs, occ);
    default: return run_cfg<NPT, JB, 3, MAXT, MINB>(bt, nw, grid, ncand, reps, ms, regs, occ);
  }
}
}  // namespace

// cfg: 0=(2,2,576,2) 1=(2,1,576,2) 2=(4,2,320,2) 3=(3,4,384,2) 4=(4,2,288,3) 5=(3,2,384,3) 6=(2,4,576,2) 7=(6,1,192,3) 8=(3,2,384,2)
extern "C" int ndtpso_bench_score(ndtpso_ctx* ctx, int32_t n, const ndtpso_problem* problems, int cfg, int var, int grid, int ncand, int reps,
                                  double* out_ms, int* out_info /* [4]: npt, warps, regs, ctas/SM */) {
  ndtpso_pso_config conf;
  ndtpso_pso_config_default(&conf);
  ndtpso_batch* bt = nullptr;
  int rc = ndtpso_batch_create(ctx, n, problems, &conf, &bt);
  if (rc) return rc;
  rc = launch_compact(bt);
  static const int npts[] = {2, 2, 4, 3, 4, 3, 2, 6, 3};
  const int npt = npts[cfg];
  const int nw = std::max(4, (bt->max_pts + 32 * npt - 1) / (32 * npt));
  int regs = 0, occ = 0;
  if (rc == NDTPSO_OK) switch (cfg) {
      case 0: rc = run_var<2, 2, 576, 2>(var, bt, nw, grid, ncand, reps, out_ms, &regs, &occ); break;
      case 1: rc = run_var<2, 1, 576, 2>(var, bt, nw, grid, ncand, reps, out_ms, &regs, &occ); break;
      case 2: rc = run_var<4, 2, 320, 2>(var, bt, nw, grid, ncand, reps, out_ms, &regs, &occ); break;
      case 3: rc = run_var<3, 4, 384, 2>(var, bt, nw, grid, ncand, reps, out_ms, &regs, &occ); break;
      case 4: rc = run_var<4, 2, 288, 3>(var, bt, nw, grid, ncand, reps, out_ms, &regs, &occ); break;
      case 5: rc = run_var<3, 2, 384, 3>(var, bt, nw, grid, ncand, reps, out_ms, &regs, &occ); break;
      case 6: rc = run_var<2, 4, 576, 2>(var, bt, nw, grid, ncand, reps, out_ms, &regs, &occ); break;
      case 7: rc = run_var<6, 1, 192, 3>(var, bt, nw, grid, ncand, reps, out_ms, &regs, &occ); break;
      default: rc = run_var<3, 2, 384, 2>(var, bt, nw, grid, ncand, reps, out_ms, &regs, &occ); break;
    }
  out_info[0] = npt;
  out_info[1] = nw;
  out_info[2] = regs;
  out_info[3] = occ;
  ndtpso_batch_destroy(bt);
  return rc;
}

// cycles summed over CTAs: [0] prologue+init, [1] phase A (+ barrier), [2] phase B (thread 0's own scoring), [3] wait at the barrier
// after B, [4] phase C; reset != 0 clears the counters first
extern "C" int ndtpso_bench_phase_cycles(unsigned long long* out, int reset) {
  unsigned long long z[8] = {0};
  if (reset) return cudaMemcpyToSymbol(ndtpso::g_phase_cycles, z, sizeof z) == cudaSuccess ? 0 : -2;
  return cudaMemcpyFromSymbol(out, ndtpso::g_phase_cycles, sizeof z) == cudaSuccess ? 0 : -2;
}
