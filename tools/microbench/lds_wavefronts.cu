// How many shared-memory wavefronts a warp-wide LDS.128 / LDS.64 / LDS.32 costs for a given address pattern (B200, sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o lds_wavefronts lds_wavefronts.cu
// Run under: ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed_op_shared_ld.sum ./lds_wavefronts
// One kernel launch per (width, pattern); every launch is one warp doing 1024 loads.  Patterns (record stride 32 bytes, like the
// screen's records):
//   0  every lane the same address
//   1  each quarter-warp (8 lanes) one address, four different records
//   2  runs of 5-6 consecutive lanes per record, 6 records spread at random (the screen's typical pattern)
//   3  32 different records, consecutive (stride 32 bytes)
//   4  32 different records at random positions
//   5  each half-warp one address
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned rec_of(int pattern, int lane, unsigned it) {
  const unsigned h = (it * 2654435761u) >> 20;  // changes every iteration so nothing is hoisted
  switch (pattern) {
    case 0: return h & 255u;
    case 1: return ((h + 37u * (lane >> 3)) & 255u);
    case 2: return ((h + 53u * ((lane * 6) >> 5)) & 255u);
    case 3: return (h + lane) & 255u;
    case 4: return ((h + 97u * lane * (lane + 3)) & 255u);
    default: return ((h + 71u * (lane >> 4)) & 255u);
  }
}

template <int WIDTH>
__global__ void probe(int pattern, float* out) {
  __shared__ __align__(16) float rec[256 * 8];
  for (int i = threadIdx.x; i < 256 * 8; i += 32) rec[i] = i * 0.001f;
  __syncthreads();
  const int lane = threadIdx.x;
  float acc = 0.f;
  for (unsigned it = 0; it < 1024; ++it) {
    const float* p = rec + 8 * rec_of(pattern, lane, it);
    if (WIDTH == 16) {
      const float4 v = *reinterpret_cast<const float4*>(p);
      acc += v.x + v.y + v.z + v.w;
    } else if (WIDTH == 8) {
      const float2 v = *reinterpret_cast<const float2*>(p + 4);
      acc += v.x + v.y;
    } else {
      acc += p[3];
    }
  }
  out[lane] = acc;
}

int main() {
  float* d;
  cudaMalloc(&d, 128);
  for (int pattern = 0; pattern < 6; ++pattern) {
    probe<16><<<1, 32>>>(pattern, d);
    probe<8><<<1, 32>>>(pattern, d);
    probe<4><<<1, 32>>>(pattern, d);
  }
  cudaDeviceSynchronize();
  printf("done: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
