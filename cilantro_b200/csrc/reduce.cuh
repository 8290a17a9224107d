// Grid-wide deterministic sum reduction used by the accumulation kernels (product code).
#pragma once
#include "cb_internal.hpp"

namespace cb {

// Sums rows r0, r0 + 1, ..., r1 - 1 of a [rows][NV] table into out[NV] with all 256 threads, in a
// fixed order: thread t owns value i = t % IPAD and rows c, c + CH, ... (c = t / IPAD); the CH chunk
// sums of each value are then added in ascending chunk order.
template <int NV>
__device__ __forceinline__ void block_sum_rows(const double* __restrict__ table, unsigned int r0, unsigned int r1,
                                               double* __restrict__ out) {
  constexpr int IPAD = (NV > 16) ? 32 : 16;
  constexpr int CH = kReduceBlock / IPAD;
  __shared__ double red[CH][IPAD];
  const int i = threadIdx.x % IPAD, c = threadIdx.x / IPAD;
  double v = 0;
  if (i < NV)
    for (unsigned int b = r0 + c; b < r1; b += CH) v += __ldcg(table + (size_t)b * NV + i);
  red[c][i] = v;
  __syncthreads();
  if (threadIdx.x < NV) {
    double s = 0;
#pragma unroll
    for (int k = 0; k < CH; k++) s += red[k][threadIdx.x];
    out[threadIdx.x] = s;
  }
  __syncthreads();
}

// warp shuffle -> shared -> one row per block; the last block of each group of kReduceGroup blocks
// folds the group's rows into one group row; the last group folds the group rows into `result`.
// Two levels keep the serial tail short (a single last block summing thousands of rows was measured
// at ~30 us of a 100 us kernel). Summation order is fixed, so the result does not depend on block
// scheduling. All counters are zero on entry and are left zero.
template <int NV>
__device__ __forceinline__ void grid_reduce(double (&acc)[NV], const ReduceScratch& rs) {
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
    rs.partials[(size_t)blockIdx.x * NV + threadIdx.x] = v;
  }
  const unsigned int ngroups = (gridDim.x + kReduceGroup - 1) / kReduceGroup;
  const unsigned int g = blockIdx.x / kReduceGroup;
  const unsigned int g0 = g * kReduceGroup, g1 = min(gridDim.x, g0 + kReduceGroup);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(rs.counters + 1 + g, 1u) == (g1 - g0) - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  block_sum_rows<NV>(rs.partials, g0, g1, rs.gpartials + (size_t)g * NV);
  if (threadIdx.x == 0) rs.counters[1 + g] = 0;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(rs.counters, 1u) == ngroups - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  block_sum_rows<NV>(rs.gpartials, 0, ngroups, rs.result);
  if (threadIdx.x == 0) rs.counters[0] = 0;
}

}  // namespace cb
