// General-k nearest neighbours over the grid (product code, sm_100a).
// Replaces the batched KDTree::kNNInRadiusSearch / kNNSearch (core/kd_tree.hpp:215-318) for k <= 32:
// same shell sweep as nn_search.cuh, with a per-thread sorted list of the k best (d2, index) pairs
// whose worst entry plays the role of nanoflann's worstDist() (kd_tree.hpp:101).
#include "cb_internal.hpp"
#include "grid_sweep.cuh"
#include <algorithm>
#include <vector>

using namespace cb;

namespace {

constexpr int kMaxK = 32;
constexpr int kBlock = 128;

struct KBest {
  float d2[kMaxK];
  int idx[kMaxK];
};

// insertion keeping ascending (d2, idx); entries with d2 >= bound never enter. Ties on d2 are ordered
// by original index so that the result is independent of the visiting order.
template <int K>
__device__ __forceinline__ void kbest_insert(float (&bd)[K], int (&bi)[K], int k, int& count, float r, int pi) {
  // reject if not better than the current k-th
  if (count == k) {
    if (!(r < bd[k - 1] || (r == bd[k - 1] && pi < bi[k - 1]))) return;
  }
  int pos = (count < k) ? count : k - 1;
  while (pos > 0 && (bd[pos - 1] > r || (bd[pos - 1] == r && bi[pos - 1] > pi))) {
    bd[pos] = bd[pos - 1];
    bi[pos] = bi[pos - 1];
    --pos;
  }
  bd[pos] = r;
  bi[pos] = pi;
  if (count < k) ++count;
}

template <int K>
__global__ void __launch_bounds__(kBlock) knn_k_kernel(const GridView g, const float4* __restrict__ qry, uint32_t nq,
                                                       const Rigid T, int k, float max_d2, int* __restrict__ out_idx,
                                                       float* __restrict__ out_d2, uint32_t* __restrict__ out_cnt) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < nq; qi += stride) {
    const float4 s = __ldg(qry + qi);
    const int oi = __float_as_int(s.w);
    float qx, qy, qz;
    apply_rigid(T, s.x, s.y, s.z, qx, qy, qz);
    float bd[K];
    int bi[K];
    int count = 0;
    // bound = current worst admissible squared distance (strict): max_d2 until k found
    auto bound = [&]() { return (count == k) ? bd[k - 1] : max_d2; };
    auto scan = [&](uint32_t b, uint32_t e) {
      for (uint32_t j = b; j < e; ++j) {
        const float4 p = __ldg(g.pts + j);
        const float dx = __fsub_rn(qx, p.x), dy = __fsub_rn(qy, p.y), dz = __fsub_rn(qz, p.z);
        float r = __fmul_rn(dx, dx);
        r = __fadd_rn(r, __fmul_rn(dy, dy));
        r = __fadd_rn(r, __fmul_rn(dz, dz));
        if (r < max_d2) kbest_insert<K>(bd, bi, k, count, r, __float_as_int(p.w));
      }
    };
    grid_sweep(g, qx, qy, qz, bound, scan);
    for (int j = 0; j < k; j++) {
      out_idx[(size_t)oi * k + j] = (j < count) ? bi[j] : -1;
      out_d2[(size_t)oi * k + j] = (j < count) ? bd[j] : max_d2;
    }
    if (out_cnt) out_cnt[oi] = (uint32_t)count;
  }
}

}  // namespace

extern "C" int cb_knn_radius(cb_context* ctx, const cb_cloud* ref, const cb_cloud* qry, const float* T12, int k,
                             float max_d2, int64_t* idx, float* d2, uint32_t* counts) {
  CB_CHECK(ctx && ref && qry && idx && d2, CB_ERR_INVALID, "null argument");
  CB_CHECK(k >= 1 && k <= kMaxK, CB_ERR_UNSUPPORTED, "k must be in [1, 32]");
  CB_CHECK(ref->ctx == ctx && qry->ctx == ctx, CB_ERR_INVALID, "cloud belongs to another context");
  CB_CUDA(cudaSetDevice(ctx->device));
  CB_TRY(ensure_index(const_cast<cb_cloud*>(ref)));
  CB_TRY(ensure_index(const_cast<cb_cloud*>(qry)));
  const size_t nq = qry->n;
  if (nq == 0) return CB_OK;
  Rigid T;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T.r[i * 3 + j] = T12 ? T12[i * 4 + j] : (i == j ? 1.f : 0.f);
    T.t[i] = T12 ? T12[i * 4 + 3] : 0.f;
  }
  int* d_idx = nullptr;
  float* d_d2 = nullptr;
  uint32_t* d_cnt = nullptr;
  CB_CUDA(cudaMallocAsync(&d_idx, nq * k * sizeof(int), ctx->stream));
  CB_CUDA(cudaMallocAsync(&d_d2, nq * k * sizeof(float), ctx->stream));
  CB_CUDA(cudaMallocAsync(&d_cnt, nq * sizeof(uint32_t), ctx->stream));
  const int blocks = (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->sm_count * 8, (nq + kBlock - 1) / kBlock));
  const GridView g = grid_view(ref);
  if (k <= 4)
    knn_k_kernel<4><<<blocks, kBlock, 0, ctx->stream>>>(g, qry->d_pts, (uint32_t)nq, T, k, max_d2, d_idx, d_d2, d_cnt);
  else if (k <= 16)
    knn_k_kernel<16><<<blocks, kBlock, 0, ctx->stream>>>(g, qry->d_pts, (uint32_t)nq, T, k, max_d2, d_idx, d_d2, d_cnt);
  else
    knn_k_kernel<32><<<blocks, kBlock, 0, ctx->stream>>>(g, qry->d_pts, (uint32_t)nq, T, k, max_d2, d_idx, d_d2, d_cnt);
  ctx->launches += 1;
  CB_CUDA(cudaGetLastError());
  std::vector<int> h_idx(nq * k);
  CB_CUDA(cudaMemcpyAsync(h_idx.data(), d_idx, nq * k * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(cudaMemcpyAsync(d2, d_d2, nq * k * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (counts) CB_CUDA(cudaMemcpyAsync(counts, d_cnt, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CB_CUDA(cudaFreeAsync(d_idx, ctx->stream));
  CB_CUDA(cudaFreeAsync(d_d2, ctx->stream));
  CB_CUDA(cudaFreeAsync(d_cnt, ctx->stream));
  CB_CUDA(cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < nq * k; i++) idx[i] = h_idx[i] < 0 ? -1 : (int64_t)h_idx[i] + (int64_t)ref->index_offset;
  return CB_OK;
}
