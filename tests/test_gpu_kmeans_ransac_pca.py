"""GPU parity: k-means assignment / Lloyd loop, RANSAC scoring, covariance / PCA vs the oracle."""
import numpy as np
import pytest

from cilantro_b200 import synth
from conftest import frob

pytestmark = pytest.mark.gpu


# ---- k-means ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,k", [(50000, 64), (20000, 1024), (3000, 2500)])
def test_kmeans_assign_labels_bitexact(cb, ctx, orc, n, k):
    pts, cent = synth.kmeans_data(n, k, seed=5)
    labels, sums, counts = cb.kmeans_assign(ctx, cb.Cloud(ctx, pts), cent)
    want, _ = orc.kmeans_assign(pts, cent)
    assert np.array_equal(labels, want.astype(np.int64))
    assert np.array_equal(counts, np.bincount(labels, minlength=k))
    ref_sums = np.zeros((k, 3))
    np.add.at(ref_sums, labels, pts.astype(np.float64))
    assert np.allclose(sums, ref_sums, rtol=1e-12, atol=1e-9)


def test_kmeans_assign_ties_lowest_cluster_wins(cb, ctx, orc):
    rng = np.random.default_rng(1)
    cent = rng.random((10, 3), dtype=np.float32)
    cent = np.vstack([cent, cent])  # duplicate centroids: every point ties between j and j + 10
    pts = rng.random((5000, 3), dtype=np.float32)
    labels, _, _ = cb.kmeans_assign(ctx, cb.Cloud(ctx, pts), cent)
    want, _ = orc.kmeans_assign(pts, cent)
    assert labels.max() < 10 and np.array_equal(labels, want.astype(np.int64))


def test_kmeans_lloyd_matches_oracle(cb, ctx, orc):
    pts, cent0 = synth.kmeans_data(40000, 32, seed=7)
    res = cb.kmeans_cluster(ctx, cb.Cloud(ctx, pts), cent0, max_iter=10, tol=0.0)
    oc, ol, oit = orc.kmeans(pts, cent0, max_iter=10, tol=0.0)
    assert res["iterations"] == oit == 10
    # centroids: the oracle sums in fp32 serial order (like the reference), the device in double
    assert np.abs(res["centroids"] - oc).max() < 5e-5
    flips = (res["labels"] != ol).sum()
    assert flips <= 40, f"{flips} label differences (boundary points only)"
    # per-step parity is exact: feeding the oracle's centroids reproduces the oracle's labels
    labels, _, _ = cb.kmeans_assign(ctx, cb.Cloud(ctx, pts), oc)
    want, _ = orc.kmeans_assign(pts, oc)
    assert np.array_equal(labels, want.astype(np.int64))


def test_kmeans_converges_and_stops_like_reference(cb, ctx, orc):
    rng = np.random.default_rng(3)
    blobs = np.vstack([rng.normal(c, 0.01, (2000, 3)) for c in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1))]).astype(np.float32)
    cent0 = blobs[[0, 2000, 4000, 6000]].copy()
    res = cb.kmeans_cluster(ctx, cb.Cloud(ctx, blobs), cent0, max_iter=100)
    oc, ol, oit = orc.kmeans(blobs, cent0, max_iter=100)
    assert res["iterations"] == oit and oit < 10
    assert np.array_equal(res["labels"], ol)
    assert np.abs(res["centroids"] - oc).max() < 1e-5


def test_kmeans_empty_cluster_repair(cb, ctx, orc):
    rng = np.random.default_rng(9)
    pts = rng.random((5000, 3), dtype=np.float32)
    cent0 = pts[:8].copy()
    cent0[7] = [50, 50, 50]  # a centroid no point will choose -> empty cluster in iteration 1
    res = cb.kmeans_cluster(ctx, cb.Cloud(ctx, pts), cent0, max_iter=1, tol=0.0)
    oc, ol, oit = orc.kmeans(pts, cent0, max_iter=1, tol=0.0)
    assert np.array_equal(res["labels"], ol)
    assert np.abs(res["centroids"] - oc).max() < 1e-5


def test_kmeans_seed_indices_match_oracle(cb, orc):
    a = cb.kmeans_seed_indices(10000, 250, 1234)
    b = orc.kmeans_seed_indices(10000, 250, 1234)
    assert np.array_equal(a, b) and len(set(a.tolist())) == 250


# ---- RANSAC ----------------------------------------------------------------------------------------
def test_ransac_score_counts_bitexact(cb, ctx, orc):
    dst, src, T_ref, inl = synth.ransac_pairs(100000, 0.3, seed=2)
    samples = orc.ransac_samples(dst.shape[0], 3, 300, seed=77)
    T_h = orc.ransac_fit_samples(dst, src, samples)
    T_h[0] = T_ref.astype(np.float32)  # make sure one good hypothesis is in the batch
    for thresh in (0.01, 0.003, 0.0):
        got = cb.ransac_score(ctx, cb.Cloud(ctx, dst), cb.Cloud(ctx, src), T_h, thresh)
        want = orc.ransac_score(dst, src, T_h, thresh)
        assert np.array_equal(got, want), (thresh, np.abs(got.astype(int) - want.astype(int)).max())
    assert got.dtype == np.uint32
    big = cb.ransac_score(ctx, cb.Cloud(ctx, dst), cb.Cloud(ctx, src), T_h, 0.01)
    assert abs(int(big[0]) - int(inl.sum())) < 0.02 * inl.sum()


def test_ransac_residuals_bitexact(cb, ctx, orc):
    dst, src, T_ref, _ = synth.ransac_pairs(20000, 0.4, seed=3)
    T = T_ref.astype(np.float32)
    res, inl = cb.ransac_residuals(ctx, cb.Cloud(ctx, dst), cb.Cloud(ctx, src), T, 0.01)
    want = orc.ransac_score(dst, src, T[None], 0.01)[0]
    assert inl.size == want
    # residual values: restate with numpy in the contract order
    q = orc.transform_points(T, src)
    e = q - dst
    x = e[:, 0] * e[:, 0] + (e[:, 1] * e[:, 1] + e[:, 2] * e[:, 2])
    assert np.array_equal(res.view(np.uint32), np.sqrt(x).view(np.uint32))


def test_ransac_full_loop_matches_oracle(cb, ctx, orc):
    dst, src, T_ref, inl = synth.ransac_pairs(20000, 0.3, seed=4)
    kw = dict(seed=99, max_iter=400, thresh=0.01, inlier_count_thresh=int(0.27 * 20000), re_estimate=True)
    got = cb.ransac_rigid(ctx, cb.Cloud(ctx, dst), cb.Cloud(ctx, src), **kw)
    want = orc.ransac_rigid(dst, src, **kw)
    # same hypothesis sequence, same early exit (hypothesis transforms differ by ~1e-7 between the two
    # host Kabsch implementations, which can move a count by a few at the threshold)
    assert got["iterations"] == want["iterations"] < 400
    assert got["best_iteration"] == want["best_iteration"]
    assert abs(got["num_inliers"] - want["num_inliers"]) <= 3
    assert frob(got["T"], want["T"]) < 1e-5
    assert frob(got["T"], T_ref) < 1e-3
    assert abs(got["num_inliers"] - inl.sum()) < 0.02 * inl.sum()


def test_ransac_sample_sequence_matches_oracle_when_no_early_exit(cb, ctx, orc):
    dst, src, _, _ = synth.ransac_pairs(5000, 0.2, seed=6)
    kw = dict(seed=5, max_iter=250, thresh=0.01, inlier_count_thresh=5000, re_estimate=False)
    got = cb.ransac_rigid(ctx, cb.Cloud(ctx, dst), cb.Cloud(ctx, src), **kw)
    want = orc.ransac_rigid(dst, src, **kw)
    assert got["iterations"] == want["iterations"] == 250
    assert got["best_iteration"] == want["best_iteration"]
    assert abs(got["num_inliers"] - want["num_inliers"]) <= 3


# ---- covariance / PCA --------------------------------------------------------------------------------
def test_pca_example_known_answer(cb, ctx):
    """examples/principal_component_analysis.cpp:6-16: corners of a 1 x 100 x 1000 box."""
    box = np.array([[x, y, z] for x in (0, 1) for y in (0, 100) for z in (0, 1000)], np.float32)
    r = cb.pca(ctx, cb.Cloud(ctx, box))
    assert r["ok"]
    assert np.allclose(r["mean"], [0.5, 50, 500])
    assert np.allclose(np.diag(r["cov"]), [8 * 0.25 / 7, 8 * 2500 / 7, 8 * 250000 / 7], rtol=1e-6)
    assert np.allclose(r["eigenvalues"], [8 * 250000 / 7, 8 * 2500 / 7, 8 * 0.25 / 7], rtol=1e-6)
    E = r["eigenvectors"]
    assert np.allclose(np.abs(E), [[0, 0, 1], [0, 1, 0], [1, 0, 0]], atol=1e-6)
    assert np.linalg.det(E.astype(np.float64)) > 0


def test_mean_cov_and_pca_match_oracle(cb, ctx, orc):
    rng = np.random.default_rng(8)
    A = rng.normal(size=(3, 3))
    pts = (rng.normal(size=(200000, 3)) @ A.T + [3.0, -2.0, 10.0]).astype(np.float32)
    got = cb.pca(ctx, cb.Cloud(ctx, pts))
    want = orc.pca(pts, accum_double=True)
    assert np.allclose(got["mean"], want["mean"], rtol=1e-6, atol=1e-6)
    assert np.allclose(got["cov"], want["cov"], rtol=1e-5, atol=1e-6)
    assert np.allclose(got["eigenvalues"], want["eigenvalues"], rtol=1e-5)
    for j in range(3):  # eigenvectors up to sign
        d = abs(float(got["eigenvectors"][:, j] @ want["eigenvectors"][:, j]))
        assert d > 1 - 1e-5
    assert np.linalg.det(got["eigenvectors"].astype(np.float64)) > 0
    # the reference's fp32 serial accumulation differs from the exact value by more than we do
    f32 = orc.pca(pts, accum_double=False)
    assert np.abs(got["cov"] - want["cov"]).max() <= np.abs(f32["cov"] - want["cov"]).max() + 1e-6


def test_mean_cov_too_few_points_is_nan(cb, ctx):
    mean, cov, ok = cb.mean_cov(ctx, cb.Cloud(ctx, np.array([[1, 2, 3]], np.float32)))
    assert not ok and np.isnan(mean).all() and np.isnan(cov).all()
