// Grid-wide deterministic sum reduction used by the accumulation kernels (product code).
#pragma once
#include <cuda_runtime.h>

namespace cb {

constexpr int kReduceBlock = 256;  // every kernel using grid_reduce launches with this block size

// warp shuffle -> shared -> per-block partial -> the last block to finish sums the partials in a
// fixed order and resets the ticket counter. `counter` must be zero on entry.
template <int NV>
__device__ __forceinline__ void grid_reduce(double (&acc)[NV], double* __restrict__ partials,
                                            unsigned int* __restrict__ counter, double* __restrict__ result) {
  __shared__ double sm[NV][kReduceBlock / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    double v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sm[i][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double v = 0;
#pragma unroll
    for (int w = 0; w < kReduceBlock / 32; w++) v += sm[threadIdx.x][w];
    partials[(size_t)blockIdx.x * NV + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int ticket = atomicAdd(counter, 1u);
    is_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // Last block: value i is summed over blocks in a fixed order (lanes stride the blocks, then a
  // shuffle tree) so the result does not depend on block scheduling.
  for (int i = warp; i < NV; i += kReduceBlock / 32) {
    double v = 0;
    for (unsigned int b = lane; b < gridDim.x; b += 32) v += __ldcg(partials + (size_t)b * NV + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) result[i] = v;
  }
  if (threadIdx.x == 0) *counter = 0;
}


}  // namespace cb
