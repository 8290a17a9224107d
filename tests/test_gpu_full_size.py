"""GPU: BASELINE.json configs 3, 4 and 5 at FULL size, checked through size-independent properties (the oracle
would need minutes to hours at these sizes). Config 2 (1 M p2p) has its full-size test in test_gpu_icp.py.

  config 3  10 M -> 10 M combined-metric ICP: the estimate inverts the generating pose, is a fixed point, and a
            random sample of the correspondences is bit-exact against brute force over all 10 M points
  config 4  k-means 50 M x K = 1024: every sampled point's label is its arg-min centroid (bit-exact contract
            arithmetic, lowest index on ties), counts sum to N, sums equal the per-cluster coordinate sums
  config 5  RANSAC scoring 5 M pairs: inlier counts of sampled hypotheses equal numpy's count on all 5 M pairs
"""
import numpy as np
import pytest

from cilantro_b200 import synth
from conftest import frob

pytestmark = pytest.mark.gpu


def test_config3_icp_combined_10m(cb, ctx, orc):
    n = 10_000_000
    dst, src, nrm, T_ref = synth.icp_pair(n, seed=1, noise=0.0005, with_normals=True)
    d, s = cb.Cloud(ctx, dst, nrm), cb.Cloud(ctx, src)
    icp = cb.Icp(ctx, d, s)
    max_d2 = np.float32((2.0 * n ** (-1.0 / 3.0)) ** 2)
    kw = dict(metric="combined", tol=0.0, max_d2=max_d2, w_pt=0.1, w_pl=1.0)
    res = icp.estimate(max_iter=12, **kw)
    assert res["iterations"] == 12 and res["num_corr"] == n
    assert frob(res["T"], T_ref) < 1e-5, frob(res["T"], T_ref)  # BASELINE: transforms within 1e-5
    again = icp.estimate(max_iter=1, T_init=res["T"], **kw)
    assert frob(again["T"], res["T"]) < 1e-6
    # correspondences of the final search, sampled: brute force over all destination points must agree exactly
    first, second, value = icp.correspondences()
    assert first.size == n and np.array_equal(second, np.arange(n))
    rng = np.random.default_rng(0)
    pick = rng.choice(n, 600, replace=False)
    # `again` searched with T_init = res["T"]
    q = orc.transform_points(res["T"], src[pick])
    oi, od = orc.BruteKnn(dst).query(q, max_d2)
    assert np.array_equal(first[pick], oi)
    assert np.array_equal(value[pick].view(np.uint32), od.view(np.uint32))


def test_config4_kmeans_50m_k1024(cb, ctx):
    n, k = 50_000_000, 1024
    pts, cent = synth.kmeans_data(n, k, seed=1)
    cloud = cb.Cloud(ctx, pts)
    labels, sums, counts = cb.kmeans_assign(ctx, cloud, cent)
    assert labels.shape == (n,) and counts.sum() == n
    assert np.array_equal(counts, np.bincount(labels, minlength=k))
    rng = np.random.default_rng(1)
    pick = rng.choice(n, 20000, replace=False)
    p = pts[pick]
    # contract arithmetic: d = c - p, d2 = dx^2 + (dy^2 + dz^2), strict <, lowest index wins
    dx = cent[None, :, 0] - p[:, None, 0]
    dy = cent[None, :, 1] - p[:, None, 1]
    dz = cent[None, :, 2] - p[:, None, 2]
    d2 = dx * dx + (dy * dy + dz * dz)
    assert d2.dtype == np.float32
    assert np.array_equal(labels[pick], np.argmin(d2, axis=1))
    # per-cluster sums (double on the device): check a few clusters against numpy over all 50 M points
    for j in (0, 511, 1023):
        member = labels == j
        assert np.allclose(sums[j], pts[member].astype(np.float64).sum(axis=0), rtol=1e-12, atol=1e-7)
    # one Lloyd step moves every centroid to the mean of its members
    res = cb.kmeans_cluster(ctx, cloud, cent, max_iter=1, tol=0.0, want_labels=False)
    want = sums / np.maximum(counts, 1)[:, None]
    assert np.abs(res["centroids"] - want).max() < 1e-6


def test_config5_ransac_scoring_5m(cb, ctx, orc):
    n = 5_000_000
    dst, src, T_ref, inl = synth.ransac_pairs(n, 0.3, seed=1)
    samples = orc.ransac_samples(n, 3, 1000, seed=7)
    T_h = orc.ransac_fit_samples(dst, src, samples)
    T_h[0] = T_ref.astype(np.float32)
    thresh = 0.01
    got = cb.ransac_score(ctx, cb.Cloud(ctx, dst), cb.Cloud(ctx, src), T_h, thresh)
    assert got.shape == (1000,) and got.dtype == np.uint32
    assert abs(int(got[0]) - int(inl.sum())) < 0.01 * inl.sum()  # the generating pose explains the inliers
    for h in (0, 1, 499, 999):
        q = orc.transform_points(T_h[h], src)
        e = q - dst
        x = e[:, 0] * e[:, 0] + (e[:, 1] * e[:, 1] + e[:, 2] * e[:, 2])
        assert int(got[h]) == int((np.sqrt(x) <= np.float32(thresh)).sum()), h
