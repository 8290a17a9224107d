// First and second moments of a point set (product code, sm_100a).
// Replaces the two serial passes of Covariance::operator() (core/covariance.hpp:64-76) and the
// rowwise().mean() of the ICP constructors (icp_single_transform_combined_metric.hpp:51-58) with one
// streaming pass over the packed xyz array: HBM-bound, 12 B/point.
#include "cb_internal.hpp"
#include "reduce.cuh"
#include "stats_kernels.cuh"
#include <algorithm>

namespace cb {

namespace {

// result[0] = n, [1..3] = sum (p - c), [4..9] = sum (p - c)(p - c)^T upper triangle (xx xy xz yy yz zz)
__global__ void __launch_bounds__(kReduceBlock) moments_kernel(const float* __restrict__ raw, size_t n, float cx,
                                                               float cy, float cz, const ReduceScratch rs) {
  double acc[kMomentValues];
#pragma unroll
  for (int i = 0; i < kMomentValues; i++) acc[i] = 0.0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const double x = (double)raw[3 * i] - (double)cx;
    const double y = (double)raw[3 * i + 1] - (double)cy;
    const double z = (double)raw[3 * i + 2] - (double)cz;
    acc[0] += 1.0;
    acc[1] += x;
    acc[2] += y;
    acc[3] += z;
    acc[4] += x * x;
    acc[5] += x * y;
    acc[6] += x * z;
    acc[7] += y * y;
    acc[8] += y * z;
    acc[9] += z * z;
  }
  grid_reduce<kMomentValues>(acc, rs);
}

}  // namespace

int launch_moments(cb_context* ctx, const float* d_raw, size_t n, const float* shift3) {
  int blocks = (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->sm_count * 4, (n + kReduceBlock - 1) / kReduceBlock));
  ReduceScratch rs;
  CB_TRY(get_reduce_scratch(ctx, blocks, kMomentValues, &rs));
  moments_kernel<<<blocks, kReduceBlock, 0, ctx->stream>>>(d_raw, n, shift3[0], shift3[1], shift3[2], rs);
  ctx->launches += 1;
  CB_CUDA(cudaGetLastError());
  return CB_OK;
}

}  // namespace cb
