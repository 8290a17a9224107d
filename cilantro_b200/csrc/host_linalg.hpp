// Product host-side O(1) linear algebra (double precision): the solves that the reference performs
// with Eigen on the host once per ICP iteration stay on the host here too.
//   3x3 SVD              one-sided (Hestenes) Jacobi       <- Eigen::JacobiSVD
//                                                             (registration/transform_estimation.hpp:36-44,
//                                                              core/space_transformations.hpp:43-51)
//   6x6 symmetric solve  Gaussian elimination, partial pivoting <- AtA.ldlt().solve(Atb) (:346)
//   3x3 symmetric eigen  via the SVD of the PSD covariance  <- Eigen::SelfAdjointEigenSolver
//                                                             (core/principal_component_analysis.hpp:77)
#pragma once
#include <cmath>
#include <algorithm>
#include <utility>

namespace cb {
namespace la {

struct Mat3 {
  double m[3][3];
  static Mat3 identity() {
    Mat3 r{};
    r.m[0][0] = r.m[1][1] = r.m[2][2] = 1.0;
    return r;
  }
};

inline Mat3 mul(const Mat3& a, const Mat3& b) {
  Mat3 r{};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}

inline Mat3 transpose(const Mat3& a) {
  Mat3 r{};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i];
  return r;
}

inline double det(const Mat3& a) {
  return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) -
         a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0]) +
         a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}

// A = U diag(sv) V^T, sv descending. U's columns belonging to (numerically) zero singular values
// are completed to a right- or left-handed frame as needed by the callers below, which only use
// u0, u1 and u0 x u1.
struct Svd {
  Mat3 U, V;
  double sv[3];
};

inline Svd svd_hestenes(const Mat3& A) {
  double G[3][3], V[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      G[i][j] = A.m[i][j];
      V[i][j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 60; ++sweep) {
    bool rotated = false;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < 3; i++) {
          alpha += G[i][p] * G[i][p];
          beta += G[i][q] * G[i][q];
          gamma += G[i][p] * G[i][q];
        }
        if (gamma == 0.0 || std::fabs(gamma) <= 1e-17 * std::sqrt(alpha * beta)) continue;
        rotated = true;
        const double zeta = (beta - alpha) / (2.0 * gamma);
        const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
        for (int i = 0; i < 3; i++) {
          const double gp = G[i][p], gq = G[i][q];
          G[i][p] = c * gp - s * gq;
          G[i][q] = s * gp + c * gq;
          const double vp = V[i][p], vq = V[i][q];
          V[i][p] = c * vp - s * vq;
          V[i][q] = s * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  double nrm[3];
  int order[3] = {0, 1, 2};
  for (int j = 0; j < 3; j++) nrm[j] = std::sqrt(G[0][j] * G[0][j] + G[1][j] * G[1][j] + G[2][j] * G[2][j]);
  std::sort(order, order + 3, [&](int x, int y) { return nrm[x] > nrm[y]; });
  Svd r;
  for (int j = 0; j < 3; j++) {
    const int s = order[j];
    r.sv[j] = nrm[s];
    for (int i = 0; i < 3; i++) {
      r.V.m[i][j] = V[i][s];
      r.U.m[i][j] = (nrm[s] > 0.0) ? G[i][s] / nrm[s] : 0.0;
    }
  }
  // complete U where singular values vanish (rank-deficient input)
  const double tiny = 1e-14 * std::max(r.sv[0], 1e-300);
  if (r.sv[0] <= 0.0) {
    r.U = Mat3::identity();
  } else {
    if (r.sv[1] <= tiny) {  // pick any unit vector orthogonal to u0
      int k = 0;
      for (int i = 1; i < 3; i++)
        if (std::fabs(r.U.m[i][0]) < std::fabs(r.U.m[k][0])) k = i;
      double e[3] = {0, 0, 0};
      e[k] = 1.0;
      double d = r.U.m[k][0];
      double v[3] = {e[0] - d * r.U.m[0][0], e[1] - d * r.U.m[1][0], e[2] - d * r.U.m[2][0]};
      double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
      for (int i = 0; i < 3; i++) r.U.m[i][1] = v[i] / n;
    }
    if (r.sv[2] <= tiny) {  // u2 = u0 x u1 (sign irrelevant to the callers)
      r.U.m[0][2] = r.U.m[1][0] * r.U.m[2][1] - r.U.m[2][0] * r.U.m[1][1];
      r.U.m[1][2] = r.U.m[2][0] * r.U.m[0][1] - r.U.m[0][0] * r.U.m[2][1];
      r.U.m[2][2] = r.U.m[0][0] * r.U.m[1][1] - r.U.m[1][0] * r.U.m[0][1];
    }
  }
  return r;
}

// U V^T with the reflection repaired by negating column `flip_col` of U when det(U V) < 0:
// flip_col = 2 is the Kabsch rule (transform_estimation.hpp:38-41), flip_col = 0 is
// LinearTransform::rotation() (space_transformations.hpp:45-48).
inline Mat3 rotation_from_svd(const Svd& s, int flip_col) {
  Mat3 U = s.U;
  if (det(U) * det(s.V) < 0.0)
    for (int i = 0; i < 3; i++) U.m[i][flip_col] = -U.m[i][flip_col];
  return mul(U, transpose(s.V));
}

// Solve the 6x6 system A x = b (A symmetric, given full row-major). Returns false when singular
// to working precision (x is then the least-damaged elimination result, like a failed LDLT).
inline bool solve6(const double* A_in, const double* b_in, double* x) {
  const int n = 6;
  double M[6][7];
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) M[i][j] = A_in[i * n + j];
    M[i][n] = b_in[i];
  }
  bool ok = true;
  for (int k = 0; k < n; k++) {
    int piv = k;
    for (int i = k + 1; i < n; i++)
      if (std::fabs(M[i][k]) > std::fabs(M[piv][k])) piv = i;
    if (piv != k)
      for (int j = 0; j <= n; j++) std::swap(M[k][j], M[piv][j]);
    const double d = M[k][k];
    if (d == 0.0 || !(d == d)) {
      ok = false;
      continue;
    }
    for (int i = k + 1; i < n; i++) {
      const double f = M[i][k] / d;
      if (f == 0.0) continue;
      for (int j = k; j <= n; j++) M[i][j] -= f * M[k][j];
    }
  }
  for (int i = n - 1; i >= 0; i--) {
    double s = M[i][n];
    for (int j = i + 1; j < n; j++) s -= M[i][j] * x[j];
    x[i] = (M[i][i] != 0.0) ? s / M[i][i] : 0.0;
  }
  return ok;
}

}  // namespace la
}  // namespace cb
