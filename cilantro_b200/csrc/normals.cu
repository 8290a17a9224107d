// Per-point normal and curvature estimation (product code, sm_100a).
// Replaces NormalEstimation::compute_normals_*/compute_normals_curvature_* (core/normal_estimation.hpp:
// 279-332, 357-421) behind PointCloud::estimateNormals{KNN,Radius,KNNInRadius} (utilities/point_cloud.hpp:
// 294-420): one thread per point of the (cell-sorted) cloud runs the neighbourhood search over the
// cloud's own grid, accumulates mean and covariance of the neighbourhood in the reference's order and
// arithmetic (core/covariance.hpp:121-135), solves the 3x3 symmetric eigenproblem in registers and
// writes the eigenvector of the smallest eigenvalue, oriented towards the view point when one is set.
//
// kNN / kNN-in-radius: the k best (d2, position) pairs live in shared memory ([slot][thread], no bank
// conflicts), ascending (d2, original index) — the order KNNSearchResultAdaptor produces
// (core/kd_tree.hpp:77-99) whenever the neighbour distances are distinct; on exact ties the reference's
// order is its kd-tree traversal order, ours is lowest index first. Covariance: bit-exact in that case.
// Radius: two sweeps (mean, then covariance) in grid order; the reference sums in ascending distance,
// so this mode agrees to fp32 rounding, not bit for bit.
#include "cb_internal.hpp"
#include "grid_sweep.cuh"
#include <algorithm>
#include <cmath>

using namespace cb;

namespace {

constexpr int kBlock = 128;
constexpr int kMaxK = 128;  // k-best lists in shared memory: 8 B x k x 128 threads (dynamic above 48 KB)

// Cyclic Jacobi on a symmetric 3x3 (a = xx,xy,xz,yy,yz,zz). Eigenvalues ascending in w, v0 = unit
// eigenvector of w[0]. The matrix is scaled by its largest |entry| first, like
// SelfAdjointEigenSolver::compute, so tiny covariances (metric clouds in mm^2 .. m^2) keep their
// relative accuracy.
__device__ __forceinline__ void jacobi_rotate(float& app, float& aqq, float& apq, float& arp, float& arq, float& v0p,
                                              float& v0q, float& v1p, float& v1q, float& v2p, float& v2q) {
  if (fabsf(apq) < 1e-30f) return;
  const float theta = (aqq - app) / (2.f * apq);
  const float t = copysignf(1.f, theta) / (fabsf(theta) + sqrtf(theta * theta + 1.f));
  const float c = rsqrtf(t * t + 1.f);
  const float s = t * c;
  app -= t * apq;
  aqq += t * apq;
  apq = 0.f;
  const float rp = c * arp - s * arq, rq = s * arp + c * arq;
  arp = rp;
  arq = rq;
  float a, b;
  a = c * v0p - s * v0q; b = s * v0p + c * v0q; v0p = a; v0q = b;
  a = c * v1p - s * v1q; b = s * v1p + c * v1q; v1p = a; v1q = b;
  a = c * v2p - s * v2q; b = s * v2p + c * v2q; v2p = a; v2q = b;
}

__device__ __forceinline__ void sym3_smallest(const float (&cv)[6], float (&w)[3], float (&n)[3]) {
  float scale = fmaxf(fmaxf(fabsf(cv[0]), fabsf(cv[1])), fmaxf(fabsf(cv[2]), fabsf(cv[3])));
  scale = fmaxf(scale, fmaxf(fabsf(cv[4]), fabsf(cv[5])));
  if (!(scale > 0.f)) {  // zero matrix: eigenvectors = identity (and NaN input falls through as NaN below)
    w[0] = w[1] = w[2] = scale;
    n[0] = 1.f;
    n[1] = 0.f;
    n[2] = 0.f;
    return;
  }
  const float inv = 1.f / scale;
  float a00 = cv[0] * inv, a01 = cv[1] * inv, a02 = cv[2] * inv, a11 = cv[3] * inv, a12 = cv[4] * inv,
        a22 = cv[5] * inv;
  float v00 = 1.f, v01 = 0.f, v02 = 0.f, v10 = 0.f, v11 = 1.f, v12 = 0.f, v20 = 0.f, v21 = 0.f, v22 = 1.f;
#pragma unroll 1
  for (int sweep = 0; sweep < 8; ++sweep) {
    const float off = fabsf(a01) + fabsf(a02) + fabsf(a12);
    if (off < 1e-12f) break;
    jacobi_rotate(a00, a11, a01, a02, a12, v00, v01, v10, v11, v20, v21);  // (p,q,r) = (0,1,2)
    jacobi_rotate(a00, a22, a02, a01, a12, v00, v02, v10, v12, v20, v22);  // (0,2,1)
    jacobi_rotate(a11, a22, a12, a01, a02, v01, v02, v11, v12, v21, v22);  // (1,2,0)
  }
  // ascending eigenvalues; n = column of the smallest
  float l0 = a00, l1 = a11, l2 = a22;
  float nx = v00, ny = v10, nz = v20;
  if (l1 < l0) { nx = v01; ny = v11; nz = v21; }
  if (l2 < fminf(l0, l1)) { nx = v02; ny = v12; nz = v22; }
  float lo = fminf(l0, fminf(l1, l2)), hi = fmaxf(l0, fmaxf(l1, l2));
  float mid = (l0 + l1 + l2) - lo - hi;
  mid = fminf(fmaxf(mid, lo), hi);
  w[0] = lo * scale;
  w[1] = mid * scale;
  w[2] = hi * scale;
  const float rn = rsqrtf(nx * nx + ny * ny + nz * nz);
  n[0] = nx * rn;
  n[1] = ny * rn;
  n[2] = nz * rn;
}

struct NormalOut {
  float* raw_nrm;    // 3n, original order
  float4* nrm;       // n, cell-sorted
  float* curvature;  // n, original order, or nullptr
  float* cov6;       // 6n, original order, or nullptr
  float vx, vy, vz;  // view point
  int use_vp;
  int use_ref;  // orient by the cloud's current normals (takes precedence over the view point)
};

__device__ __forceinline__ void finish_point(const NormalOut& o, uint32_t qi, int oi, float px, float py, float pz,
                                             bool valid, const float (&cv)[6]) {
  const float nan = __int_as_float(0x7fc00000);
  float w[3] = {nan, nan, nan}, nv[3] = {nan, nan, nan};
  if (valid) {
    sym3_smallest(cv, w, nv);
    if (o.use_ref) {
      // eigenvectors().col(0).dot(ref_normals.col(i)) < 0 -> flip (normal_estimation.hpp:351-355)
      const float4 r = o.nrm[qi];
      const float d = __fadd_rn(__fmul_rn(nv[0], r.x), __fadd_rn(__fmul_rn(nv[1], r.y), __fmul_rn(nv[2], r.z)));
      if (d < 0.f) {
        nv[0] = -nv[0];
        nv[1] = -nv[1];
        nv[2] = -nv[2];
      }
    } else if (o.use_vp) {
      // eigenvectors().col(0).dot(view_point - p) < 0 -> flip (normal_estimation.hpp:325-329)
      const float ex = __fsub_rn(o.vx, px), ey = __fsub_rn(o.vy, py), ez = __fsub_rn(o.vz, pz);
      const float d = __fadd_rn(__fmul_rn(nv[0], ex), __fadd_rn(__fmul_rn(nv[1], ey), __fmul_rn(nv[2], ez)));
      if (d < 0.f) {
        nv[0] = -nv[0];
        nv[1] = -nv[1];
        nv[2] = -nv[2];
      }
    }
  }
  o.nrm[qi] = make_float4(nv[0], nv[1], nv[2], 0.f);
  o.raw_nrm[3 * (size_t)oi] = nv[0];
  o.raw_nrm[3 * (size_t)oi + 1] = nv[1];
  o.raw_nrm[3 * (size_t)oi + 2] = nv[2];
  if (o.curvature) o.curvature[oi] = valid ? w[0] / (w[0] + w[1] + w[2]) : nan;  // :389
  if (o.cov6) {
#pragma unroll
    for (int c = 0; c < 6; c++) o.cov6[6 * (size_t)oi + c] = valid ? cv[c] : nan;
  }
}

template <int K>
__global__ void __launch_bounds__(kBlock) normals_knn_kernel(const GridView g, int k, float max_d2, const NormalOut o) {
  extern __shared__ __align__(16) unsigned char knn_smem[];
  float(*sd)[kBlock] = reinterpret_cast<float(*)[kBlock]>(knn_smem);
  uint32_t(*sp)[kBlock] = reinterpret_cast<uint32_t(*)[kBlock]>(knn_smem + sizeof(float) * K * kBlock);
  const int t = threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t qi = blockIdx.x * blockDim.x + t; qi < g.n; qi += stride) {
    const float4 s = __ldg(g.pts + qi);
    const int oi = __float_as_int(s.w);
    int count = 0;
    float worst = max_d2;  // strict admission bound until the list is full, then the k-th distance
    auto orig = [&](uint32_t pos) { return __float_as_int(__ldg(&g.pts[pos].w)); };
    auto bound = [&]() { return worst; };
    auto scan = [&](uint32_t b, uint32_t e) {
      for (uint32_t j = b; j < e; ++j) {
        const float4 p = __ldg(g.pts + j);
        const float dx = __fsub_rn(s.x, p.x), dy = __fsub_rn(s.y, p.y), dz = __fsub_rn(s.z, p.z);
        float r = __fmul_rn(dx, dx);
        r = __fadd_rn(r, __fmul_rn(dy, dy));
        r = __fadd_rn(r, __fmul_rn(dz, dz));
        if (!(r < max_d2)) continue;
        const int pi = __float_as_int(p.w);
        if (count == k) {
          if (r > worst) continue;
          if (r == worst && pi > orig(sp[k - 1][t])) continue;
        }
        int pos = (count < k) ? count : k - 1;
        while (pos > 0) {
          const float pd = sd[pos - 1][t];
          if (pd > r || (pd == r && orig(sp[pos - 1][t]) > pi)) {
            sd[pos][t] = pd;
            sp[pos][t] = sp[pos - 1][t];
            --pos;
          } else {
            break;
          }
        }
        sd[pos][t] = r;
        sp[pos][t] = j;
        if (count < k) ++count;
        if (count == k) worst = sd[k - 1][t];
      }
    };
    grid_sweep(
        g, s.x, s.y, s.z, bound, scan,
        [&]() {
          count = 0;
          worst = max_d2;
        },
        (uint32_t)k);
    float cv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool valid = count >= 3;  // setMinValidSampleSize(points_.rows()), normal_estimation.hpp:27
    if (valid) {
      float mx = 0.f, my = 0.f, mz = 0.f;
      for (int j = 0; j < count; j++) {
        const float4 p = __ldg(g.pts + sp[j][t]);
        mx = __fadd_rn(mx, p.x);
        my = __fadd_rn(my, p.y);
        mz = __fadd_rn(mz, p.z);
      }
      const float inv = __fdiv_rn(1.0f, (float)count);
      mx = __fmul_rn(inv, mx);
      my = __fmul_rn(inv, my);
      mz = __fmul_rn(inv, mz);
      for (int j = 0; j < count; j++) {
        const float4 p = __ldg(g.pts + sp[j][t]);
        const float dx = __fsub_rn(p.x, mx), dy = __fsub_rn(p.y, my), dz = __fsub_rn(p.z, mz);
        cv[0] = __fadd_rn(cv[0], __fmul_rn(dx, dx));
        cv[1] = __fadd_rn(cv[1], __fmul_rn(dx, dy));
        cv[2] = __fadd_rn(cv[2], __fmul_rn(dx, dz));
        cv[3] = __fadd_rn(cv[3], __fmul_rn(dy, dy));
        cv[4] = __fadd_rn(cv[4], __fmul_rn(dy, dz));
        cv[5] = __fadd_rn(cv[5], __fmul_rn(dz, dz));
      }
      const float invm1 = __fdiv_rn(1.0f, (float)(count - 1));
#pragma unroll
      for (int c = 0; c < 6; c++) cv[c] = __fmul_rn(invm1, cv[c]);
    }
    finish_point(o, qi, oi, s.x, s.y, s.z, valid, cv);
  }
}

__global__ void __launch_bounds__(kBlock) normals_radius_kernel(const GridView g, float r2, const NormalOut o) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x; qi < g.n; qi += stride) {
    const float4 s = __ldg(g.pts + qi);
    const int oi = __float_as_int(s.w);
    auto bound = [&]() { return r2; };
    auto dist2 = [&](const float4& p) {
      const float dx = __fsub_rn(s.x, p.x), dy = __fsub_rn(s.y, p.y), dz = __fsub_rn(s.z, p.z);
      float r = __fmul_rn(dx, dx);
      r = __fadd_rn(r, __fmul_rn(dy, dy));
      return __fadd_rn(r, __fmul_rn(dz, dz));
    };
    int count = 0;
    float mx = 0.f, my = 0.f, mz = 0.f;
    grid_sweep(g, s.x, s.y, s.z, bound, [&](uint32_t b, uint32_t e) {
      for (uint32_t j = b; j < e; ++j) {
        const float4 p = __ldg(g.pts + j);
        if (dist2(p) < r2) {
          mx += p.x;
          my += p.y;
          mz += p.z;
          ++count;
        }
      }
    }, [&]() {
      count = 0;
      mx = my = mz = 0.f;
    }, 0u);
    float cv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const bool valid = count >= 3;
    if (valid) {
      const float inv = __fdiv_rn(1.0f, (float)count);
      mx = __fmul_rn(inv, mx);
      my = __fmul_rn(inv, my);
      mz = __fmul_rn(inv, mz);
      grid_sweep(g, s.x, s.y, s.z, bound, [&](uint32_t b, uint32_t e) {
        for (uint32_t j = b; j < e; ++j) {
          const float4 p = __ldg(g.pts + j);
          if (dist2(p) < r2) {
            const float dx = __fsub_rn(p.x, mx), dy = __fsub_rn(p.y, my), dz = __fsub_rn(p.z, mz);
            cv[0] = __fadd_rn(cv[0], __fmul_rn(dx, dx));
            cv[1] = __fadd_rn(cv[1], __fmul_rn(dx, dy));
            cv[2] = __fadd_rn(cv[2], __fmul_rn(dx, dz));
            cv[3] = __fadd_rn(cv[3], __fmul_rn(dy, dy));
            cv[4] = __fadd_rn(cv[4], __fmul_rn(dy, dz));
            cv[5] = __fadd_rn(cv[5], __fmul_rn(dz, dz));
          }
        }
      }, [&]() {
#pragma unroll
        for (int c = 0; c < 6; c++) cv[c] = 0.f;
      }, 0u);
      const float invm1 = __fdiv_rn(1.0f, (float)(count - 1));
#pragma unroll
      for (int c = 0; c < 6; c++) cv[c] = __fmul_rn(invm1, cv[c]);
    }
    finish_point(o, qi, oi, s.x, s.y, s.z, valid, cv);
  }
}

}  // namespace

extern "C" int cb_cloud_estimate_normals(cb_context* ctx, cb_cloud* cloud, int k, float radius2,
                                         const float* view_point3, int use_current_as_ref, float* normals,
                                         float* curvature, float* cov6, float* gpu_ms) {
  CB_CHECK(ctx && cloud, CB_ERR_INVALID, "null argument");
  CB_CHECK(cloud->ctx == ctx, CB_ERR_INVALID, "cloud belongs to another context");
  CB_CHECK(k >= 0 && k <= kMaxK, CB_ERR_UNSUPPORTED, "k must be in [0, 128] (0 = radius neighbourhood)");
  CB_CHECK(k > 0 || radius2 > 0.f, CB_ERR_INVALID, "need k > 0 and/or radius2 > 0");
  CB_CUDA(cudaSetDevice(ctx->device));
  if (gpu_ms) *gpu_ms = 0.f;
  const size_t n = cloud->n;
  if (n == 0) return CB_OK;
  CB_TRY(ensure_index(cloud));
  const bool use_ref = use_current_as_ref && cloud->d_nrm;  // like PointCloud: only when normals exist
  if (!cloud->d_raw_nrm) CB_CUDA(cudaMallocAsync(&cloud->d_raw_nrm, 3 * n * sizeof(float), ctx->stream));
  if (!cloud->d_nrm) CB_CUDA(cudaMallocAsync(&cloud->d_nrm, n * sizeof(float4), ctx->stream));
  float* d_curv = nullptr;
  float* d_cov = nullptr;
  if (curvature) CB_CUDA(cudaMallocAsync(&d_curv, n * sizeof(float), ctx->stream));
  if (cov6) CB_CUDA(cudaMallocAsync(&d_cov, 6 * n * sizeof(float), ctx->stream));
  NormalOut o;
  o.raw_nrm = cloud->d_raw_nrm;
  o.nrm = cloud->d_nrm;
  o.curvature = d_curv;
  o.cov6 = d_cov;
  o.use_vp = view_point3 && std::isfinite(view_point3[0]) && std::isfinite(view_point3[1]) &&
             std::isfinite(view_point3[2]);  // view_point_.allFinite(), normal_estimation.hpp:283
  o.use_ref = use_ref ? 1 : 0;
  o.vx = o.use_vp ? view_point3[0] : 0.f;
  o.vy = o.use_vp ? view_point3[1] : 0.f;
  o.vz = o.use_vp ? view_point3[2] : 0.f;
  const float max_d2 = radius2 > 0.f ? radius2 : 3.402823466e38f;
  const GridView g = grid_view(cloud);
  const int blocks = (int)std::max<size_t>(1, std::min<size_t>((size_t)ctx->sm_count * 8, (n + kBlock - 1) / kBlock));
  ScopedEvents ev;
  if (gpu_ms) {
    CB_TRY(ev.create());
    CB_CUDA(cudaEventRecord(ev.e0, ctx->stream));
  }
  if (k == 0)
    normals_radius_kernel<<<blocks, kBlock, 0, ctx->stream>>>(g, max_d2, o);
  else if (k <= 8)
    normals_knn_kernel<8><<<blocks, kBlock, 8 * kBlock * 8, ctx->stream>>>(g, k, max_d2, o);
  else if (k <= 16)
    normals_knn_kernel<16><<<blocks, kBlock, 16 * kBlock * 8, ctx->stream>>>(g, k, max_d2, o);
  else if (k <= 32)
    normals_knn_kernel<32><<<blocks, kBlock, 32 * kBlock * 8, ctx->stream>>>(g, k, max_d2, o);
  else if (k <= 64) {
    // (per call, not once per process: function attributes belong to the device the context is on)
    CB_CUDA(cudaFuncSetAttribute(normals_knn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * kBlock * 8));
    normals_knn_kernel<64><<<blocks, kBlock, 64 * kBlock * 8, ctx->stream>>>(g, k, max_d2, o);
  } else {
    CB_CUDA(cudaFuncSetAttribute(normals_knn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * kBlock * 8));
    normals_knn_kernel<128><<<blocks, kBlock, 128 * kBlock * 8, ctx->stream>>>(g, k, max_d2, o);
  }
  ctx->launches += 1;
  if (gpu_ms) CB_CUDA(cudaEventRecord(ev.e1, ctx->stream));
  CB_CUDA(cudaGetLastError());
  if (normals)
    CB_CUDA(cudaMemcpyAsync(normals, cloud->d_raw_nrm, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (curvature)
    CB_CUDA(cudaMemcpyAsync(curvature, d_curv, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (cov6) CB_CUDA(cudaMemcpyAsync(cov6, d_cov, 6 * n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  if (d_curv) CB_CUDA(cudaFreeAsync(d_curv, ctx->stream));
  if (d_cov) CB_CUDA(cudaFreeAsync(d_cov, ctx->stream));
  CB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (gpu_ms) CB_CUDA(cudaEventElapsedTime(gpu_ms, ev.e0, ev.e1));
  return CB_OK;
}
