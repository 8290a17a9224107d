"""GPU: BASELINE.json configs 2-5 at FULL size. Configs 2 and 3 (ICP at 1 M and 10 M) run against the ORACLE at full
size (the reference's own nanoflann answers the queries; bottom of the file) and through size-independent properties;
configs 4 and 5 through properties checked with numpy on samples / whole arrays.

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


# ---- full-size ORACLE parity (reference nanoflann drives the restated ICP loop) -------------------------------------
# the named configs were only property-tested. The reference's own kd-tree answers 1 M queries in
# ~20 ms and 10 M in ~1 s per iteration on the box's host cores, so the oracle runs the whole ICP at full size.
def _list_equal_or_tie(first, value, o_first, o_value):
    """Same pairs; where the index differs the squared distance must be bit-equal (an exact tie: nanoflann keeps the
    first point its traversal meets, the grid keeps the lowest index)."""
    diff = np.flatnonzero(first != o_first)
    assert np.array_equal(value.view(np.uint32), o_value.view(np.uint32))
    return diff.size


def _full_size_parity(cb, ctx, orc, n, iters, kw, with_normals, noise):
    assert orc.have_ref(), "oracle/_ref (reference nanoflann) must be built: python -c 'import oracle; oracle.build()'"
    dst, src, nrm, T_ref = synth.icp_pair(n, seed=1, noise=noise, with_normals=with_normals)
    icp = cb.Icp(ctx, cb.Cloud(ctx, dst, nrm), cb.Cloud(ctx, src))
    knn = orc.RefKnn(dst)
    res = icp.estimate(max_iter=iters, tol=0.0, **kw)
    # oracle with double accumulation (the GPU accumulates in double; the fp32 serial sums of the reference's
    # deterministic build are checked separately below with the tolerance their rounding needs)
    ref = orc.icp(dst, src, knn, dst_n=nrm, max_iter=iters, tol=0.0, accum_double=True, **kw)
    assert res["iterations"] == ref["iterations"] == iters
    assert res["num_corr"] == ref["num_corr"], (res["num_corr"], ref["num_corr"])
    err = frob(res["T"], ref["T"])
    assert err < 1e-5, err  # north_star: final rigid transforms within 1e-5 Frobenius
    ref32 = orc.icp(dst, src, knn, dst_n=nrm, max_iter=iters, tol=0.0, accum_double=False, **kw)
    err32 = frob(res["T"], ref32["T"])
    assert ref32["num_corr"] == res["num_corr"]
    assert err32 < 1e-4, err32  # fp32 serial sums over >= 1e6 terms: rounding of the REFERENCE's accumulation
    # the correspondence list of one more search from the GPU's own estimate, against the reference kd-tree
    one = icp.estimate(max_iter=1, tol=0.0, T_init=res["T"], **kw)
    first, second, value = icp.correspondences()
    o1, o2, ov = orc.find_correspondences(res["T"], src, knn, kw["max_d2"])
    assert first.shape == o1.shape and np.array_equal(second, o2)
    ties = _list_equal_or_tie(first, value, o1, ov)
    assert ties <= 8, ties
    assert one["num_corr"] == o1.size
    return err, err32, ties


def test_config2_icp_p2p_1m_oracle_parity(cb, ctx, orc):
    err, err32, ties = _full_size_parity(cb, ctx, orc, 1_000_000, 15,
                                         dict(metric="p2p", max_d2=np.float32(0.02 ** 2)), False, 0.001)
    print(f"config 2 @1M: |T_gpu - T_oracle(double)|_F = {err:.2e}, vs fp32-serial oracle {err32:.2e}, index ties {ties}")


def test_config3_icp_combined_10m_oracle_parity(cb, ctx, orc):
    err, err32, ties = _full_size_parity(cb, ctx, orc, 10_000_000, 10,
                                         dict(metric="combined", max_d2=np.float32(0.01 ** 2), w_pt=0.1, w_pl=1.0),
                                         True, 0.001)
    print(f"config 3 @10M: |T_gpu - T_oracle(double)|_F = {err:.2e}, vs fp32-serial oracle {err32:.2e}, index ties {ties}")
