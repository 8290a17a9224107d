// Type-check of the Eigen adapters of b200_shims.hpp (compiled with -I tests/cpp/eigen_stub where Eigen3 is absent,
// against the real Eigen where it is installed). Host-only: no device call is made.
#include <cilantro/b200_shims.hpp>
#ifndef CILANTRO_B200_HAS_EIGEN
#error "the Eigen adapters were not enabled"
#endif

int main() {
  Eigen::Matrix<float, 3, Eigen::Dynamic> pts(3, 4);
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 3; r++) pts(r, c) = (float)(10 * c + r);
  cilantro::ConstVectorSetMatrixMap3f view(pts);  // what every shim constructor takes
  if (view.cols() != 4 || view.col(2)[1] != 21.f) return 1;
  Eigen::Map<const Eigen::Matrix<float, 3, Eigen::Dynamic>> m(pts.data(), 3, 4);
  cilantro::ConstVectorSetMatrixMap3f view2(m);
  if (view2.data() != pts.data()) return 2;
  std::vector<Eigen::Vector3f> vv(2);
  vv[1](2) = 5.f;
  cilantro::ConstVectorSetMatrixMap3f view3(vv);
  if (view3.cols() != 2 || view3.col(1)[2] != 5.f) return 3;
  cilantro::VectorSet3f owned(pts);
  Eigen::Matrix<float, 3, Eigen::Dynamic> back = owned;
  if (back.cols() != 4 || back(2, 3) != 32.f) return 4;
  Eigen::Transform<float, 3, Eigen::Isometry> T = Eigen::Transform<float, 3, Eigen::Isometry>::Identity();
  T.translation()(1) = 2.f;
  T.linear()(0, 1) = -1.f;
  cilantro::RigidTransform3f R(T);
  if (R.translation(1) != 2.f || R.linear(0, 1) != -1.f) return 5;
  Eigen::Transform<float, 3, Eigen::Isometry> T2 = R;
  if (T2.translation()(1) != 2.f || T2.linear()(0, 1) != -1.f || T2.linear()(2, 2) != 1.f) return 6;
  return 0;
}
