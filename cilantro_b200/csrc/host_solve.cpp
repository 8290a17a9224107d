// Host entry points of the O(1) solves of the ICP iteration (product code; no device, no oracle). The
// arithmetic lives in solve_core.hpp, which the device-resident loop (icp_loop.cu) compiles for the GPU too.
#include "cb_internal.hpp"
#include "solve_core.hpp"
#include "host_solve.hpp"
#include <cstring>
#include <cmath>

namespace cb {

void t34_identity(float* T) { sc::t34_identity(T); }
bool kabsch_from_moments(const double* s, float* T) { return sc::kabsch_from_moments(s, nullptr, nullptr, T); }
bool kabsch_from_pivoted_moments(const double* s, const float* pd, const float* pq, float* T) {
  return sc::kabsch_from_moments(s, pd, pq, T);
}
bool gauss_newton_update(const double* s28, const float* Tin, float* Tout, float* dtheta_norm) {
  return sc::gauss_newton_update(s28, Tin, Tout, dtheta_norm);
}
void uncenter(float* T, const float* dst_mean, const float* src_mean) { sc::uncenter(T, dst_mean, src_mean); }
void reorthonormalize(float* T) { sc::reorthonormalize(T); }
void compose(const float* A, const float* B, float* out) { sc::compose(A, B, out); }
float update_norm(const float* T) { return sc::update_norm(T); }
void apply_point(const float* T, const float* p, float* q) { sc::apply_point(T, p, q); }

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
