// C++ drop-in test: the calls of the reference's example programs, written against the cilantro names
// (include/cilantro/...), running on the GPU through libcilantro_b200.so.
//   examples/kd_tree.cpp:6-19                      -> known answer 0,3 / 0.18,0.38
//   examples/principal_component_analysis.cpp:6-16 -> analytic covariance of a 1 x 100 x 1000 box
//   examples/rigid_icp.cpp:116-136                 -> SimpleCombinedMetricRigidICP3f with fluent setters
//   examples/kmeans.cpp:30-59                      -> KMeans3f<>
//   examples/ransac_transform_estimator.cpp:79-86  -> RigidTransformRANSACEstimator3f<>
// Exit code 0 = all checks passed. Built and run by tests/test_cpp_shims.py.
#include <cilantro/clustering/kmeans.hpp>
#include <cilantro/core/grid_downsampler.hpp>
#include <cilantro/core/kd_tree.hpp>
#include <cilantro/core/normal_estimation.hpp>
#include <cilantro/core/principal_component_analysis.hpp>
#include <cilantro/model_estimation/ransac_transform_estimator.hpp>
#include <cilantro/registration/icp_common_instances.hpp>
#include <cilantro/utilities/point_cloud.hpp>
#include <cilantro/utilities/timer.hpp>

#include <cstdio>
#include <random>

#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::printf("CHECK failed at %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      return 1;                                                            \
    }                                                                      \
  } while (0)

static cilantro::RigidTransform3f make_pose(float ax, float ay, float az, float angle, float tx, float ty, float tz) {
  const float n = std::sqrt(ax * ax + ay * ay + az * az);
  ax /= n; ay /= n; az /= n;
  const float c = std::cos(angle), s = std::sin(angle), k = 1.f - c;
  cilantro::RigidTransform3f T;
  T.linear(0, 0) = c + k * ax * ax;      T.linear(0, 1) = k * ax * ay - s * az; T.linear(0, 2) = k * ax * az + s * ay;
  T.linear(1, 0) = k * ay * ax + s * az; T.linear(1, 1) = c + k * ay * ay;      T.linear(1, 2) = k * ay * az - s * ax;
  T.linear(2, 0) = k * az * ax - s * ay; T.linear(2, 1) = k * az * ay + s * ax; T.linear(2, 2) = c + k * az * az;
  T.translation(0) = tx; T.translation(1) = ty; T.translation(2) = tz;
  return T;
}

static float frob(const cilantro::RigidTransform3f& a, const cilantro::RigidTransform3f& b) {
  float s = 0;
  for (int i = 0; i < 12; i++) s += (a.data()[i] - b.data()[i]) * (a.data()[i] - b.data()[i]);
  return std::sqrt(s);
}

int main() {
  // ---- kd_tree.cpp ---------------------------------------------------------------------------------
  {
    std::vector<cilantro::Vector3f> points;
    points.emplace_back(0, 0, 0); points.emplace_back(1, 0, 0); points.emplace_back(0, 1, 0); points.emplace_back(0, 0, 1);
    points.emplace_back(0, 1, 1); points.emplace_back(1, 0, 1); points.emplace_back(1, 1, 0); points.emplace_back(1, 1, 1);
    cilantro::KDTree3f<> tree(points);
    cilantro::NeighborSet<float> nn = tree.kNNInRadiusSearch(cilantro::Vector3f(0.1f, 0.1f, 0.4f), 2, 1.001f);
    CHECK(nn.size() == 2);
    CHECK(nn[0].index == 0 && nn[1].index == 3);
    CHECK(std::fabs(nn[0].value - 0.18f) < 1e-6f && std::fabs(nn[1].value - 0.38f) < 1e-6f);
    CHECK(tree.nearestNeighborSearch(cilantro::Vector3f(0.9f, 0.9f, 0.8f)).index == 7);
    // radiusSearch / search(spec): squared radius, ascending distance, equal distances by index
    cilantro::NeighborSet<float> rn = tree.radiusSearch(cilantro::Vector3f(0.1f, 0.1f, 0.4f), 1.0f);
    CHECK(rn.size() == 4 && rn[0].index == 0 && rn[1].index == 3 && rn[2].index == 1 && rn[3].index == 2);
    CHECK(rn[2].value == rn[3].value && std::fabs(rn[2].value - 0.98f) < 1e-6f);
    CHECK(tree.search(cilantro::Vector3f(0.1f, 0.1f, 0.4f), cilantro::RadiusNeighborhoodSpecification<float>(0.2f)).size() == 1);
    CHECK(tree.search(cilantro::Vector3f(0.1f, 0.1f, 0.4f), cilantro::KNNNeighborhoodSpecification<>(3)).size() == 3);
    CHECK(tree.search(cilantro::Vector3f(0.1f, 0.1f, 0.4f),
                      cilantro::KNNInRadiusNeighborhoodSpecification<float>(5, 0.5f)).size() == 2);
  }
  // ---- principal_component_analysis.cpp ----------------------------------------------------------------
  {
    std::vector<cilantro::Vector3f> pts;
    for (float x : {0.f, 1.f}) for (float y : {0.f, 100.f}) for (float z : {0.f, 1000.f}) pts.emplace_back(x, y, z);
    cilantro::PrincipalComponentAnalysis3f pca(pts);
    CHECK(std::fabs(pca.getDataMean()[0] - 0.5f) < 1e-5f && std::fabs(pca.getDataMean()[2] - 500.f) < 1e-3f);
    CHECK(std::fabs(pca.getEigenValues()[0] - 8.f * 250000.f / 7.f) < 1.f);
    CHECK(std::fabs(pca.getEigenValues()[2] - 8.f * 0.25f / 7.f) < 1e-5f);
    CHECK(std::fabs(std::fabs(pca.getEigenVectors()[2 * 3 + 0]) - 1.f) < 1e-5f);  // first eigenvector = +-e_z
    // project / reconstruct (core/principal_component_analysis.hpp:46-72): full-rank round trip, rank-1 keeps z only
    std::vector<float> pr3 = pca.project(pts, 3);
    cilantro::VectorSet3f back = pca.reconstruct(pr3, 3);
    for (size_t i = 0; i < pts.size(); i++)
      for (int r = 0; r < 3; r++) CHECK(std::fabs(back(r, i) - pts[i][r]) < 2e-3f);
    std::vector<float> pr1 = pca.project<1>(pts);
    CHECK(pr1.size() == pts.size());
    cilantro::VectorSet3f back1 = pca.reconstruct(pr1, 1);
    for (size_t i = 0; i < pts.size(); i++) {
      CHECK(std::fabs(std::fabs(pr1[i]) - 500.f) < 1e-2f);
      CHECK(std::fabs(back1(2, i) - pts[i][2]) < 1e-2f && std::fabs(back1(0, i) - 0.5f) < 1e-3f);
    }
    // subset constructor: the four points with z = 0 -> no variance along z
    std::vector<size_t> subset = {0, 2, 4, 6};
    cilantro::PrincipalComponentAnalysis3f pca_sub(pts, subset);
    CHECK(std::fabs(pca_sub.getDataMean()[2]) < 1e-6f && std::fabs(pca_sub.getEigenValues()[2]) < 1e-3f);
  }
  // ---- rigid_icp.cpp recipe on a synthetic surface ------------------------------------------------------
  std::mt19937 rng(7);
  std::uniform_real_distribution<float> U(-1.f, 1.f);
  cilantro::PointCloud3f dst, src;
  const size_t N = 20000;
  dst.points.resize(3, N);
  dst.normals.resize(3, N);
  for (size_t i = 0; i < N; i++) {
    const float u = U(rng), v = U(rng);
    const float z = 0.3f * std::sin(2 * u) * std::cos(1.5f * v);
    const float gx = 0.6f * std::cos(2 * u) * std::cos(1.5f * v), gy = -0.45f * std::sin(2 * u) * std::sin(1.5f * v);
    const float nn = std::sqrt(gx * gx + gy * gy + 1.f);
    dst.points.setCol(i, {u, v, z});
    dst.normals.setCol(i, {-gx / nn, -gy / nn, 1.f / nn});
  }
  const cilantro::RigidTransform3f tf_ref = make_pose(0.3f, -1.f, 0.5f, 0.1f, -0.20f, -0.05f, 0.10f);
  src = dst;
  for (size_t i = 0; i < N; i++) {
    cilantro::Vector3f p = src.points.col(i);
    for (int r = 0; r < 3; r++) p[r] += 0.004f * U(rng);
    src.points.setCol(i, p);
  }
  src.transform(tf_ref);
  // ---- normal_estimation.cpp: NormalEstimation3f / PointCloud3f::estimateNormalsKNN ------------------------
  {
    cilantro::KDTree3f<> tree(dst.points);
    cilantro::NormalEstimation3f ne(tree);
    ne.setViewPoint(cilantro::Vector3f(0.f, 0.f, 10.f));
    cilantro::VectorSet3f est = ne.getNormalsKNN(12);
    std::vector<float> curv;
    ne.estimateCurvatureKNNInRadius(curv, 12, 0.05f * 0.05f);
    CHECK(est.cols() == N && curv.size() == N);
    double acc = 0;
    size_t bad = 0;
    for (size_t i = 0; i < N; i++) {
      const cilantro::Vector3f a = est.col(i), b = dst.normals.col(i);
      const float d = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
      acc += d;
      bad += (d < 0.9f);
      CHECK(std::fabs(a.norm() - 1.f) < 1e-4f);
      CHECK(std::isnan(curv[i]) || (curv[i] > -1e-5f && curv[i] < 0.34f));
    }
    std::printf("normals: mean dot with analytic %.4f, %zu of %zu below 0.9\n", acc / N, bad, N);
    CHECK(acc / N > 0.98 && bad < N / 50);
    // PointCloud API: view point = origin; then the current normals as the reference keep their side
    cilantro::PointCloud3f pc;
    pc.points = dst.points;
    pc.normals = dst.normals;
    pc.estimateNormalsKNN(tree, 12, /*use_current_as_ref=*/true);
    CHECK(pc.hasNormals());
    size_t differ = 0;  // same eigenvector; the side may differ only where it is (nearly) tangent to both rules
    for (size_t i = 0; i < N; i++) {
      const cilantro::Vector3f a = pc.normals.col(i), b = est.col(i);
      const bool same = std::fabs(a[0] - b[0]) < 1e-6f && std::fabs(a[1] - b[1]) < 1e-6f && std::fabs(a[2] - b[2]) < 1e-6f;
      const bool flipped = std::fabs(a[0] + b[0]) < 1e-6f && std::fabs(a[1] + b[1]) < 1e-6f && std::fabs(a[2] + b[2]) < 1e-6f;
      CHECK(same || flipped);
      differ += flipped;
    }
    CHECK(differ < N / 100);
    // gridDownsample (the first step of examples/normal_estimation.cpp): fewer points, unit normals, and the
    // serial build's order (parallel = false) holds the same bins as the map order
    cilantro::PointCloud3f ds = pc.gridDownsampled(0.05f);
    cilantro::PointCloud3f ds_serial = pc.gridDownsampled(0.05f, 1, false);
    std::printf("gridDownsample(0.05): %zu -> %zu points\n", pc.size(), ds.size());
    CHECK(ds.size() > 100 && ds.size() < N / 4 && ds.hasNormals() && !ds.hasColors());
    CHECK(ds_serial.size() == ds.size());
    double sx = 0, sx2 = 0;
    for (size_t i = 0; i < ds.size(); i++) {
      CHECK(std::fabs(ds.normals.col(i).norm() - 1.f) < 1e-5f);
      sx += ds.points(0, i);
      sx2 += ds_serial.points(0, i);
      if (i > 0) CHECK(std::floor(ds.points(0, i) * 20.f) >= std::floor(ds.points(0, i - 1) * 20.f) - 1.f);
    }
    CHECK(std::fabs(sx - sx2) < 1e-2);
    cilantro::VectorSet3f only_pts = cilantro::PointsGridDownsampler3f(pc.points, 0.05f).getDownsampledPoints(3);
    CHECK(only_pts.cols() > 0 && only_pts.cols() <= ds.size());
    pc.gridDownsample(0.05f);
    CHECK(pc.size() == ds.size());
  }
  {
    cilantro::Timer timer;
    timer.start();
    cilantro::SimpleCombinedMetricRigidICP3f icp(dst.points, dst.normals, src.points);
    icp.setMaxNumberOfOptimizationStepIterations(1).setPointToPointMetricWeight(0.0f).setPointToPlaneMetricWeight(1.0f);
    icp.correspondenceSearchEngine().setMaxDistance(0.1f * 0.1f);
    icp.setConvergenceTolerance(1e-4f).setMaxNumberOfIterations(30);
    cilantro::RigidTransform3f tf_est = icp.estimate().getTransform();
    timer.stop();
    std::printf("combined ICP: %zu iterations, converged %d, %.2f ms, |T - tf_ref^-1|_F = %.2e\n",
                icp.getNumberOfPerformedIterations(), (int)icp.hasConverged(), timer.getElapsedTime(),
                frob(tf_est, tf_ref.inverse()));
    CHECK(icp.hasConverged() && icp.getNumberOfPerformedIterations() < 30);
    CHECK(frob(tf_est, tf_ref.inverse()) < 2e-2f);
    auto residuals = icp.getResiduals();
    CHECK(residuals.size() == N);
    CHECK(icp.getCorrespondences().size() > N / 2);
    // the general class with caller-owned correspondence weight evaluators (icp_single_transform_combined_metric.hpp:9-101):
    // RBF weights on both terms; a very wide kernel must reproduce the unity-weight result, a narrow one still converges
    cilantro::RBFKernelWeightEvaluator<float, float, true> w_pt_eval(1e3f), w_pl_eval(1e3f);
    cilantro::CombinedMetricRigidICP3f<cilantro::RBFKernelWeightEvaluator<float, float, true>,
                                       cilantro::RBFKernelWeightEvaluator<float, float, true>>
        icp_w(dst.points, dst.normals, src.points, w_pt_eval, w_pl_eval);
    icp_w.setMaxNumberOfOptimizationStepIterations(1).setPointToPointMetricWeight(0.0f).setPointToPlaneMetricWeight(1.0f);
    icp_w.correspondenceSearchEngine().setMaxDistance(0.1f * 0.1f);
    icp_w.setConvergenceTolerance(1e-4f).setMaxNumberOfIterations(30);
    CHECK(frob(icp_w.estimate().getTransform(), tf_est) < 1e-5f);
    icp_w.pointToPlaneCorrespondenceWeightEvaluator().setSigma(0.02f);
    icp_w.estimate();
    std::printf("combined ICP, RBF(0.02) plane weights: %zu iterations, |T - tf_ref^-1|_F = %.2e\n",
                icp_w.getNumberOfPerformedIterations(), frob(icp_w.getTransform(), tf_ref.inverse()));
    CHECK(icp_w.hasConverged() && frob(icp_w.getTransform(), tf_ref.inverse()) < 2e-2f);
    CHECK(frob(icp_w.getTransform(), tf_est) > 0.f);  // the weights do change the estimate
  }
  {
    cilantro::SimplePointToPointMetricRigidICP3f icp(dst.points, src.points);
    icp.correspondenceSearchEngine().setMaxDistance(0.3f * 0.3f);
    icp.setMaxNumberOfIterations(60).setConvergenceTolerance(1e-5f);
    icp.estimate();
    std::printf("p2p ICP: %zu iterations, |T - tf_ref^-1|_F = %.2e\n", icp.getNumberOfPerformedIterations(),
                frob(icp.getTransform(), tf_ref.inverse()));
    CHECK(frob(icp.getTransform(), tf_ref.inverse()) < 5e-2f);
    // non-default engine modes (correspondence_search_kd_tree.hpp:237-285): reciprocal pairs, closest 80 %
    cilantro::SimplePointToPointMetricRigidICP3f icp2(dst.points, src.points);
    icp2.correspondenceSearchEngine()
        .setMaxDistance(0.3f * 0.3f)
        .setSearchDirection(cilantro::CorrespondenceSearchDirection::BOTH)
        .setRequireReciprocality(true)
        .setInlierFraction(0.8);
    icp2.setInitialTransform(icp.getTransform()).setMaxNumberOfIterations(10).setConvergenceTolerance(1e-6f).estimate();
    const auto& corr = icp2.getCorrespondences();
    std::printf("reciprocal + fraction ICP: %zu iterations, %zu pairs, |T - tf_ref^-1|_F = %.2e\n",
                icp2.getNumberOfPerformedIterations(), corr.size(), frob(icp2.getTransform(), tf_ref.inverse()));
    CHECK(corr.size() > N / 4 && corr.size() <= N);
    for (size_t i = 1; i < corr.size(); i++) CHECK(corr[i - 1].value <= corr[i].value);  // sorted by the fraction filter
    CHECK(frob(icp2.getTransform(), tf_ref.inverse()) < 5e-2f);
    cilantro::SimplePointToPointMetricRigidICP3f icp3(dst.points, src.points);
    icp3.correspondenceSearchEngine().setMaxDistance(0.3f * 0.3f).setOneToOne(true);
    icp3.setInitialTransform(icp.getTransform()).setMaxNumberOfIterations(3).estimate();
    const auto& c3 = icp3.getCorrespondences();
    CHECK(!c3.empty());
    for (size_t i = 1; i < c3.size(); i++) CHECK(c3[i - 1].indexInFirst < c3[i].indexInFirst);  // one pair per dst point
  }
  // ---- kmeans.cpp ------------------------------------------------------------------------------------------
  {
    cilantro::KMeans3f<> kmc(dst.points);
    kmc.cluster(250, 100, std::numeric_limits<float>::epsilon(), false, /*seed=*/1234);
    const auto& cpi = kmc.getClusterToPointIndicesMap();
    size_t total = 0, mins = N, maxs = 0;
    for (const auto& c : cpi) { total += c.size(); mins = std::min(mins, c.size()); maxs = std::max(maxs, c.size()); }
    std::printf("kmeans: %zu iterations, cluster sizes %zu..%zu\n", kmc.getNumberOfPerformedIterations(), mins, maxs);
    CHECK(cpi.size() == 250 && total == N && mins > 0);
    CHECK(kmc.getPointToClusterIndexMap().size() == N);
  }
  // ---- ransac_transform_estimator.cpp ------------------------------------------------------------------------
  {
    const size_t M = 1000;
    cilantro::VectorSet3f d(3, M), s(3, M);
    cilantro::CorrespondenceSet<float> corr(M);
    for (size_t i = 0; i < M; i++) {
      cilantro::Vector3f p(U(rng), U(rng), U(rng));
      s.setCol(i, p);
      if (i % 4 == 0) {
        cilantro::Vector3f q = tf_ref * p;
        for (int r = 0; r < 3; r++) q[r] += 0.002f * U(rng);
        d.setCol(i, q);
      } else {
        d.setCol(i, {U(rng), U(rng), U(rng)});
      }
      corr[i] = {i, i, 0.f};
    }
    cilantro::RigidTransformRANSACEstimator3f<> te(d, s, corr);
    te.setMaxInlierResidual(0.01f).setTargetInlierCount((size_t)(0.50 * M)).setMaxNumberOfIterations(250).setReEstimationStep(true);
    te.setRandomSeed(99);
    cilantro::RigidTransform3f tf_est = te.estimate().getModel();
    std::printf("ransac: %zu iterations, %zu inliers, |T - tf_ref|_F = %.2e\n", te.getNumberOfPerformedIterations(),
                te.getNumberOfInliers(), frob(tf_est, tf_ref));
    CHECK(te.getNumberOfPerformedIterations() == 250);  // 25 % true correspondences: the 50 % target is unreachable
    CHECK(te.getNumberOfInliers() > 200 && te.getNumberOfInliers() < 300);
    CHECK(frob(tf_est, tf_ref) < 5e-3f);
    CHECK(te.getModelResiduals().size() == M);
    // (dst, src, dst_ind, src_ind) constructor (ransac_transform_estimator.hpp:46-59): the same pairs by index vectors
    std::vector<size_t> di(M), si(M);
    for (size_t i = 0; i < M; i++) di[i] = si[i] = i;
    cilantro::RigidTransformRANSACEstimator3f<> te2(d, s, di, si);
    te2.setMaxInlierResidual(0.01f).setTargetInlierCount((size_t)(0.50 * M)).setMaxNumberOfIterations(250).setReEstimationStep(true);
    te2.setRandomSeed(99);
    CHECK(frob(te2.estimate().getModel(), tf_est) == 0.f && te2.getNumberOfInliers() == te.getNumberOfInliers());
  }
  std::printf("all C++ shim checks passed\n");
  return 0;
}
