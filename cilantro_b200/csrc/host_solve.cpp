// Host-only O(1) solves of the ICP iteration (product code; no device, no oracle).
#include "cb_internal.hpp"
#include "host_linalg.hpp"
#include "host_solve.hpp"
#include <cstring>
#include <cmath>

namespace cb {

void t34_identity(float* T) {
  for (int i = 0; i < 12; i++) T[i] = 0.f;
  T[0] = T[5] = T[10] = 1.f;
}

// estimateTransformPointToPointMetric from the reduced moments (transform_estimation.hpp:25-47):
//   mu_d = sum d / n, mu_q = sum q / n, sigma = (1/n) sum (d - mu_d)(q - mu_q)^T
//        = (sum d q^T) / n - mu_d mu_q^T,   R = U V^T (reflection: last column), t = mu_d - R mu_q
bool kabsch_from_moments(const double* s, float* T) {
  const double n = s[0];
  if (!(n > 0.0)) {  // :20-23
    t34_identity(T);
    return false;
  }
  double mud[3], muq[3];
  for (int r = 0; r < 3; r++) {
    mud[r] = s[1 + r] / n;
    muq[r] = s[4 + r] / n;
  }
  la::Mat3 sigma;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) sigma.m[r][c] = s[7 + r * 3 + c] / n - mud[r] * muq[c];
  const la::Svd svd = la::svd_hestenes(sigma);
  const la::Mat3 R = la::rotation_from_svd(svd, 2);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) T[r * 4 + c] = (float)R.m[r][c];
  for (int r = 0; r < 3; r++) {
    double t = mud[r];
    for (int c = 0; c < 3; c++) t -= (double)T[r * 4 + c] * muq[c];
    T[r * 4 + 3] = (float)t;
  }
  return n >= 3.0;  // :47
}

// One Gauss-Newton update (transform_estimation.hpp:346-357):
//   d_theta = AtA^-1 Atb ; theta = atan(|w|) ; Ra = AngleAxis(theta, w/|w|) ; ta = cos(theta) v
//   T_out = Ra * Translation(ta) * Ra * T_in
bool gauss_newton_update(const double* s28, const float* Tin, float* Tout, float* dtheta_norm) {
  double A[36], b[6], x[6];
  int k = 1;
  for (int r = 0; r < 6; r++)
    for (int c = r; c < 6; c++) {
      A[r * 6 + c] = s28[k];
      A[c * 6 + r] = s28[k];
      ++k;
    }
  for (int r = 0; r < 6; r++) b[r] = s28[22 + r];
  const bool ok = la::solve6(A, b, x);
  float dth[6];
  for (int i = 0; i < 6; i++) dth[i] = (float)x[i];  // the reference holds d_theta in fp32
  const double na = std::sqrt((double)dth[0] * dth[0] + (double)dth[1] * dth[1] + (double)dth[2] * dth[2]);
  const double theta = std::atan(na);
  double ax[3] = {0, 0, 0};
  if (na > 0.0)
    for (int i = 0; i < 3; i++) ax[i] = dth[i] / na;
  const double c = std::cos(theta), sn = std::sin(theta), k1 = 1.0 - c;
  la::Mat3 Ra;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Ra.m[i][j] = k1 * ax[i] * ax[j] + (i == j ? c : 0.0);
  Ra.m[0][1] -= sn * ax[2];
  Ra.m[0][2] += sn * ax[1];
  Ra.m[1][0] += sn * ax[2];
  Ra.m[1][2] -= sn * ax[0];
  Ra.m[2][0] -= sn * ax[1];
  Ra.m[2][1] += sn * ax[0];
  la::Mat3 L;
  double t0[3];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) L.m[i][j] = Tin[i * 4 + j];
    t0[i] = Tin[i * 4 + 3];
  }
  const la::Mat3 RR = la::mul(Ra, Ra);
  const la::Mat3 Lo = la::mul(RR, L);
  double t1[3], t2[3];
  for (int i = 0; i < 3; i++)
    t1[i] = Ra.m[i][0] * t0[0] + Ra.m[i][1] * t0[1] + Ra.m[i][2] * t0[2] + c * (double)dth[3 + i];
  for (int i = 0; i < 3; i++) t2[i] = Ra.m[i][0] * t1[0] + Ra.m[i][1] * t1[1] + Ra.m[i][2] * t1[2];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) Tout[i * 4 + j] = (float)Lo.m[i][j];
    Tout[i * 4 + 3] = (float)t2[i];
  }
  double nn = 0;
  for (int i = 0; i < 6; i++) nn += (double)dth[i] * dth[i];
  if (dtheta_norm) *dtheta_norm = (float)std::sqrt(nn);
  return ok;
}

// tform = Translation(dst_mean) * tform * Translation(-src_mean)   (transform_estimation.hpp:361/365)
void uncenter(float* T, const float* dst_mean, const float* src_mean) {
  for (int i = 0; i < 3; i++) {
    double t = T[i * 4 + 3];
    for (int k = 0; k < 3; k++) t -= (double)T[i * 4 + k] * (double)src_mean[k];
    T[i * 4 + 3] = (float)(t + (double)dst_mean[i]);
  }
}

// LinearTransform::rotation() (core/space_transformations.hpp:43-51) on T's linear part
void reorthonormalize(float* T) {
  la::Mat3 A;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) A.m[i][j] = T[i * 4 + j];
  const la::Mat3 R = la::rotation_from_svd(la::svd_hestenes(A), 0);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) T[i * 4 + j] = (float)R.m[i][j];
}

void compose(const float* A, const float* B, float* out) {
  float r[12];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += (double)A[i * 4 + k] * (double)B[k * 4 + j];
      r[i * 4 + j] = (float)s;
    }
    double s = A[i * 4 + 3];
    for (int k = 0; k < 3; k++) s += (double)A[i * 4 + k] * (double)B[k * 4 + 3];
    r[i * 4 + 3] = (float)s;
  }
  std::memcpy(out, r, sizeof(r));
}

// sqrt(|R - I|_F^2 + |t|^2)   (icp_single_transform_combined_metric.hpp:214-216)
float update_norm(const float* T) {
  float dn = 0.f;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) {
      const float e = T[r * 4 + c] - (r == c ? 1.f : 0.f);
      dn += e * e;
    }
    dn += T[r * 4 + 3] * T[r * 4 + 3];
  }
  return std::sqrt(dn);
}

void apply_point(const float* T, const float* p, float* q) {
  for (int r = 0; r < 3; r++) q[r] = (T[r * 4] * p[0] + (T[r * 4 + 1] * p[1] + T[r * 4 + 2] * p[2])) + T[r * 4 + 3];
}

// Symmetric 3x3 eigen-decomposition for PCA: eigenvalues descending, right-handed eigenvector
// frame (principal_component_analysis.hpp:76-84). cov is PSD, so its SVD is its eigensystem.
void pca_from_cov(const double* cov9, float* evals3, float* evecs9) {
  la::Mat3 C;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) C.m[i][j] = cov9[i * 3 + j];
  const la::Svd s = la::svd_hestenes(C);
  la::Mat3 E = s.V;
  if (la::det(E) < 0.0)
    for (int i = 0; i < 3; i++) E.m[i][2] = -E.m[i][2];
  for (int j = 0; j < 3; j++) {
    // Rayleigh quotient recovers the signed eigenvalue (tiny negatives from rounding stay negative)
    double lam = 0;
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) lam += s.V.m[a][j] * C.m[a][b] * s.V.m[b][j];
    evals3[j] = (float)lam;
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) evecs9[i * 3 + j] = (float)E.m[i][j];
}

}  // namespace cb

// ---- exported host-only helpers ------------------------------------------------------------------
extern "C" {

int cb_solve_kabsch_moments(const double* sums16, float* T12) {
  if (!sums16 || !T12) return CB_ERR_INVALID;
  return cb::kabsch_from_moments(sums16, T12) ? 1 : 0;
}

int cb_solve_gauss_newton(const double* sums28, const float* T_in12, float* T_out12, float* dtheta_norm) {
  if (!sums28 || !T_in12 || !T_out12) return CB_ERR_INVALID;
  float tmp[12];
  cb::gauss_newton_update(sums28, T_in12, tmp, dtheta_norm);
  std::memcpy(T_out12, tmp, sizeof(tmp));
  return CB_OK;
}

int cb_solve_rotation(const float* L9, float* R9) {
  if (!L9 || !R9) return CB_ERR_INVALID;
  float T[12];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) T[i * 4 + j] = L9[i * 3 + j];
    T[i * 4 + 3] = 0.f;
  }
  cb::reorthonormalize(T);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R9[i * 3 + j] = T[i * 4 + j];
  return CB_OK;
}

int cb_compose(const float* A12, const float* B12, float* out12) {
  if (!A12 || !B12 || !out12) return CB_ERR_INVALID;
  cb::compose(A12, B12, out12);
  return CB_OK;
}

}  // extern "C"
