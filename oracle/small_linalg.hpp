// ORACLE (test infrastructure, NOT product code).
//
// Tiny dense linear algebra used by the CPU restatement in cilantro_oracle.cpp.
// It stands in for the Eigen3 calls the reference makes at O(1)-size solve sites
// (Eigen3 is an un-vendored, unpinned dependency of the reference — CMakeLists.txt:7 —
// and is absent from this image; "parity unpinned" at that boundary):
//   * Eigen::JacobiSVD 3x3            registration/transform_estimation.hpp:36-44
//                                     core/space_transformations.hpp:43-51
//   * Matrix<6,6>::ldlt().solve()     registration/transform_estimation.hpp:346,718
//   * Eigen::SelfAdjointEigenSolver   core/principal_component_analysis.hpp:76-84
// Everything here is double precision; callers round to float where the reference stores float.
//
// Method notes (deliberately different from the product's host solver in
// cilantro_b200/csrc/host_linalg.hpp, so that the two validate each other):
//   SVD:  eigen-decomposition of A^T A by cyclic Jacobi, U recovered from A V / sigma.
//   6x6:  Cholesky-free symmetric LDL^T with diagonal pivoting.
#pragma once
#include <cmath>
#include <cstring>
#include <algorithm>

namespace orc {

struct M3 {
  double a[3][3];
};

inline M3 m3_identity() {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.a[i][j] = (i == j) ? 1.0 : 0.0;
  return r;
}

inline M3 m3_mul(const M3& x, const M3& y) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += x.a[i][k] * y.a[k][j];
      r.a[i][j] = s;
    }
  return r;
}

inline M3 m3_transpose(const M3& x) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r.a[i][j] = x.a[j][i];
  return r;
}

inline double m3_det(const M3& m) {
  return m.a[0][0] * (m.a[1][1] * m.a[2][2] - m.a[1][2] * m.a[2][1]) -
         m.a[0][1] * (m.a[1][0] * m.a[2][2] - m.a[1][2] * m.a[2][0]) +
         m.a[0][2] * (m.a[1][0] * m.a[2][1] - m.a[1][1] * m.a[2][0]);
}

// Cyclic Jacobi for a symmetric 3x3. On return: s = V diag(w) V^T, eigenvalues ASCENDING
// (the order Eigen::SelfAdjointEigenSolver reports), V columns orthonormal.
inline void sym3_eigen(const M3& s_in, double w[3], M3& V) {
  M3 s = s_in;
  V = m3_identity();
  for (int sweep = 0; sweep < 64; sweep++) {
    double off = s.a[0][1] * s.a[0][1] + s.a[0][2] * s.a[0][2] + s.a[1][2] * s.a[1][2];
    double diag = s.a[0][0] * s.a[0][0] + s.a[1][1] * s.a[1][1] + s.a[2][2] * s.a[2][2];
    if (off <= 1e-60 || off <= 1e-34 * diag) break;
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        double apq = s.a[p][q];
        if (apq == 0.0) continue;
        double tau = (s.a[q][q] - s.a[p][p]) / (2.0 * apq);
        double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        double c = 1.0 / std::sqrt(1.0 + t * t), sn = t * c;
        // S <- J^T S J with J = [[c, sn], [-sn, c]] on (p,q)
        for (int k = 0; k < 3; k++) {
          double skp = s.a[k][p], skq = s.a[k][q];
          s.a[k][p] = c * skp - sn * skq;
          s.a[k][q] = sn * skp + c * skq;
        }
        for (int k = 0; k < 3; k++) {
          double spk = s.a[p][k], sqk = s.a[q][k];
          s.a[p][k] = c * spk - sn * sqk;
          s.a[q][k] = sn * spk + c * sqk;
        }
        for (int k = 0; k < 3; k++) {
          double vkp = V.a[k][p], vkq = V.a[k][q];
          V.a[k][p] = c * vkp - sn * vkq;
          V.a[k][q] = sn * vkp + c * vkq;
        }
      }
  }
  int idx[3] = {0, 1, 2};
  double ev[3] = {s.a[0][0], s.a[1][1], s.a[2][2]};
  std::sort(idx, idx + 3, [&](int x, int y) { return ev[x] < ev[y]; });
  M3 Vs;
  for (int j = 0; j < 3; j++) {
    w[j] = ev[idx[j]];
    for (int k = 0; k < 3; k++) Vs.a[k][j] = V.a[k][idx[j]];
  }
  V = Vs;
}

// Thin description of the SVD that both reference call sites consume: singular values
// descending, V orthonormal, u1/u2 the two leading left singular vectors and c = u1 x u2.
// The third left singular vector of any full SVD is u3 = +-c with det(U) = +-1 accordingly,
// which is all that the two "reflection fix" rules below need.
struct Svd3 {
  double sv[3];
  M3 V;
  double u1[3], u2[3], c[3];
  double detV;
};

inline void normalize3(double v[3]) {
  double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (n > 0) {
    v[0] /= n;
    v[1] /= n;
    v[2] /= n;
  }
}

inline void any_orthogonal(const double a[3], double out[3]) {
  int k = 0;
  if (std::fabs(a[1]) < std::fabs(a[k])) k = 1;
  if (std::fabs(a[2]) < std::fabs(a[k])) k = 2;
  double e[3] = {0, 0, 0};
  e[k] = 1.0;
  out[0] = a[1] * e[2] - a[2] * e[1];
  out[1] = a[2] * e[0] - a[0] * e[2];
  out[2] = a[0] * e[1] - a[1] * e[0];
  normalize3(out);
}

inline Svd3 svd3(const M3& A) {
  Svd3 r;
  M3 AtA = m3_mul(m3_transpose(A), A);
  double w[3];
  M3 Vasc;
  sym3_eigen(AtA, w, Vasc);
  // descending order
  for (int j = 0; j < 3; j++) {
    r.sv[j] = std::sqrt(std::max(0.0, w[2 - j]));
    for (int k = 0; k < 3; k++) r.V.a[k][j] = Vasc.a[k][2 - j];
  }
  r.detV = m3_det(r.V);
  double scale = r.sv[0];
  auto col = [&](int j, double out[3]) {
    for (int i = 0; i < 3; i++)
      out[i] = A.a[i][0] * r.V.a[0][j] + A.a[i][1] * r.V.a[1][j] + A.a[i][2] * r.V.a[2][j];
  };
  if (scale <= 0.0) {  // A == 0: any orthonormal frame is a valid U
    r.u1[0] = 1; r.u1[1] = 0; r.u1[2] = 0;
    r.u2[0] = 0; r.u2[1] = 1; r.u2[2] = 0;
  } else {
    col(0, r.u1);
    normalize3(r.u1);
    if (r.sv[1] > 1e-13 * scale) {
      col(1, r.u2);
      // re-orthogonalise against u1 (guards the loss of accuracy of A v / sigma for small sigma)
      double d = r.u2[0] * r.u1[0] + r.u2[1] * r.u1[1] + r.u2[2] * r.u1[2];
      for (int i = 0; i < 3; i++) r.u2[i] -= d * r.u1[i];
      normalize3(r.u2);
    } else {
      any_orthogonal(r.u1, r.u2);
    }
  }
  r.c[0] = r.u1[1] * r.u2[2] - r.u1[2] * r.u2[1];
  r.c[1] = r.u1[2] * r.u2[0] - r.u1[0] * r.u2[2];
  r.c[2] = r.u1[0] * r.u2[1] - r.u1[1] * r.u2[0];
  return r;
}

// R = U V^T with the reference's Kabsch reflection rule: if det(U V) < 0 negate the LAST
// column of U (registration/transform_estimation.hpp:38-44). For any full SVD this equals
// [u1, u2, (u1 x u2) det(V)] V^T.
inline M3 kabsch_rotation_from_sigma(const M3& sigma) {
  Svd3 s = svd3(sigma);
  M3 U;
  for (int i = 0; i < 3; i++) {
    U.a[i][0] = s.u1[i];
    U.a[i][1] = s.u2[i];
    U.a[i][2] = s.c[i] * s.detV;
  }
  return m3_mul(U, m3_transpose(s.V));
}

// LinearTransform::rotation() (core/space_transformations.hpp:43-51): as above but the fix
// negates column 0 of U. With Ut = [u1, u2, u3], u3 = (u1 x u2) det(Ut):
//   det(Ut V) >= 0  ->  det(Ut) = det(V)   -> R = [ u1, u2,  c det(V)] V^T
//   det(Ut V) <  0  ->  det(Ut) = -det(V)  -> R = [-u1, u2, -c det(V)] V^T
// The sign of det(Ut V) equals the sign of det(A) when A is non-singular.
inline M3 nearest_rotation_col0_rule(const M3& A) {
  Svd3 s = svd3(A);
  bool reflect = m3_det(A) < 0.0;
  M3 U;
  for (int i = 0; i < 3; i++) {
    U.a[i][0] = reflect ? -s.u1[i] : s.u1[i];
    U.a[i][1] = s.u2[i];
    U.a[i][2] = (reflect ? -1.0 : 1.0) * s.c[i] * s.detV;
  }
  return m3_mul(U, m3_transpose(s.V));
}

// Solve the symmetric system A x = b (6x6) by LDL^T with symmetric diagonal pivoting —
// the factorisation Eigen's .ldlt() performs (transform_estimation.hpp:346).
inline void ldlt6_solve(const double Ain[36], const double bin[6], double x[6]) {
  const int n = 6;
  double A[6][6];
  double b[6];
  int perm[6];
  for (int i = 0; i < n; i++) {
    perm[i] = i;
    b[i] = bin[i];
    for (int j = 0; j < n; j++) A[i][j] = Ain[i * n + j];
  }
  double L[6][6] = {{0}};
  double D[6] = {0};
  for (int k = 0; k < n; k++) {
    int piv = k;
    double best = std::fabs(A[k][k]);
    for (int i = k + 1; i < n; i++)
      if (std::fabs(A[i][i]) > best) {
        best = std::fabs(A[i][i]);
        piv = i;
      }
    if (piv != k) {
      for (int j = 0; j < n; j++) std::swap(A[k][j], A[piv][j]);
      for (int i = 0; i < n; i++) std::swap(A[i][k], A[i][piv]);
      for (int j = 0; j < k; j++) std::swap(L[k][j], L[piv][j]);
      std::swap(perm[k], perm[piv]);
    }
    D[k] = A[k][k];
    L[k][k] = 1.0;
    for (int i = k + 1; i < n; i++) L[i][k] = (D[k] != 0.0) ? A[i][k] / D[k] : 0.0;
    for (int i = k + 1; i < n; i++)
      for (int j = k + 1; j < n; j++) A[i][j] -= L[i][k] * D[k] * L[j][k];
  }
  double y[6], z[6];
  for (int i = 0; i < n; i++) {
    double s = b[perm[i]];
    for (int j = 0; j < i; j++) s -= L[i][j] * y[j];
    y[i] = s;
  }
  for (int i = 0; i < n; i++) z[i] = (D[i] != 0.0) ? y[i] / D[i] : 0.0;
  double xp[6];
  for (int i = n - 1; i >= 0; i--) {
    double s = z[i];
    for (int j = i + 1; j < n; j++) s -= L[j][i] * xp[j];
    xp[i] = s;
  }
  for (int i = 0; i < n; i++) x[perm[i]] = xp[i];
}

}  // namespace orc
