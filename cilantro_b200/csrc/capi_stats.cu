// C ABI: covariance / PCA (product code).
#include "cb_internal.hpp"
#include "stats_kernels.cuh"
#include "host_solve.hpp"
#include <cmath>
#include <cstring>
#include <limits>

using namespace cb;

// Mean and (n-1)-normalised covariance. The reference does two serial passes (mean, then centred
// outer products, core/covariance.hpp:64-76); here: a tiny first launch over <= 4096 points gives a
// pivot close to the mean, then ONE streaming pass accumulates sum(p - c) and sum (p - c)(p - c)^T in
// double, which is combined exactly: cov = (S2 - S1 S1^T / n) / (n - 1), mean = c + S1 / n.
static int mean_cov_double(cb_context* ctx, const cb_cloud* pts, double* mean3, double* cov9, double* n_out) {
  CB_CUDA(cudaSetDevice(ctx->device));
  const float zero[3] = {0, 0, 0};
  double m[kMomentValues];
  float pivot[3] = {0, 0, 0};
  if (pts->n > 0) {
    const size_t head = pts->n < 4096 ? pts->n : 4096;
    CB_TRY(launch_moments(ctx, pts->d_raw, head, zero));
    CB_TRY(fetch_result(ctx, kMomentValues, false, m));
    for (int r = 0; r < 3; r++) pivot[r] = (float)(m[1 + r] / m[0]);
  }
  if (ctx->world > 1) {
    // every rank must use the same pivot: average the non-empty ranks' pivots (4-value all-reduce)
    const double wgt = pts->n > 0 ? 1.0 : 0.0;
    double pv[4] = {pivot[0] * wgt, pivot[1] * wgt, pivot[2] * wgt, wgt};
    std::memcpy(ctx->h_result, pv, sizeof(pv));
    CB_CUDA(cudaMemcpyAsync(ctx->d_result, ctx->h_result, sizeof(pv), cudaMemcpyHostToDevice, ctx->stream));
    CB_TRY(fetch_result(ctx, 4, true, pv));
    for (int r = 0; r < 3; r++) pivot[r] = pv[3] > 0 ? (float)(pv[r] / pv[3]) : 0.f;
  }
  CB_TRY(launch_moments(ctx, pts->d_raw, pts->n, pivot));
  CB_TRY(fetch_result(ctx, kMomentValues, true, m));
  const double n = m[0];
  *n_out = n;
  if (n < 2.0) return CB_OK;
  const double s1[3] = {m[1], m[2], m[3]};
  const double s2[3][3] = {{m[4], m[5], m[6]}, {m[5], m[7], m[8]}, {m[6], m[8], m[9]}};
  for (int r = 0; r < 3; r++) mean3[r] = (double)pivot[r] + s1[r] / n;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) cov9[r * 3 + c] = (s2[r][c] - s1[r] * s1[c] / n) / (n - 1.0);
  return CB_OK;
}

extern "C" {

int cb_mean_cov(cb_context* ctx, const cb_cloud* pts, float* mean3, float* cov9) {
  CB_CHECK(ctx && pts && mean3 && cov9, CB_ERR_INVALID, "null argument");
  double mu[3], cov[9], n = 0;
  CB_TRY(mean_cov_double(ctx, pts, mu, cov, &n));
  if (n < 2.0) {  // covariance.hpp:35-38 (min_sample_size_ = 2)
    for (int i = 0; i < 3; i++) mean3[i] = std::numeric_limits<float>::quiet_NaN();
    for (int i = 0; i < 9; i++) cov9[i] = std::numeric_limits<float>::quiet_NaN();
    return 0;
  }
  for (int i = 0; i < 3; i++) mean3[i] = (float)mu[i];
  for (int i = 0; i < 9; i++) cov9[i] = (float)cov[i];
  return 1;
}

int cb_pca(cb_context* ctx, const cb_cloud* pts, float* mean3, float* cov9, float* evals3, float* evecs9) {
  CB_CHECK(ctx && pts && mean3 && cov9 && evals3 && evecs9, CB_ERR_INVALID, "null argument");
  double mu[3], cov[9], n = 0;
  CB_TRY(mean_cov_double(ctx, pts, mu, cov, &n));
  if (n < 2.0) {
    for (int i = 0; i < 3; i++) mean3[i] = evals3[i] = std::numeric_limits<float>::quiet_NaN();
    for (int i = 0; i < 9; i++) cov9[i] = evecs9[i] = std::numeric_limits<float>::quiet_NaN();
    return 0;
  }
  for (int i = 0; i < 3; i++) mean3[i] = (float)mu[i];
  double covf[9];
  for (int i = 0; i < 9; i++) {
    cov9[i] = (float)cov[i];
    covf[i] = (double)cov9[i];  // the eigen-solver sees the fp32 covariance, as in the reference (:77)
  }
  pca_from_cov(covf, evals3, evecs9);
  return 1;
}

}  // extern "C"
