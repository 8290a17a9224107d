"""GPU parity of the device-resident ICP loop (icp_loop.cu): against the host-driven loop, against the oracle, and
the exactness of its per-query cache (a match kept without a search must be what the full search returns).

Tolerances: transforms 1e-5 Frobenius against the oracle (BASELINE.json north_star), 1e-6 between the two loops
(they reduce different but equivalent moment sets: raw vs pivoted Kabsch sums); integer outputs exact.
"""
import numpy as np
import pytest

from cilantro_b200 import synth
from conftest import frob

pytestmark = pytest.mark.gpu


def _pair(n, seed, normals=True):
    return synth.icp_pair(n, seed=seed, noise=0.002, with_normals=normals)


@pytest.mark.parametrize("timing", [1, 0])
@pytest.mark.parametrize("metric,kw", [("p2p", {}), ("combined", dict(w_pt=0.1, w_pl=1.0)), ("combined", dict(w_pt=0.0, w_pl=1.0))])
def test_device_loop_matches_host_loop_every_iteration(cb, ctx, metric, kw, timing):
    """timing = 0 is the production setting: no event records between the kernels, i.e. the launches of consecutive
    iterations are adjacent in the stream and programmatic dependent launch is in effect across iterations too."""
    dst, src, nrm, _ = _pair(60000, 5)
    icp = cb.Icp(ctx, cb.Cloud(ctx, dst, nrm), cb.Cloud(ctx, src))
    max_d2 = np.float32(0.03**2)
    for k in (1, 2, 3, 5, 9, 14, 23):
        a = icp.estimate(metric=metric, max_iter=k, tol=0.0, max_d2=max_d2, host_loop=True, **kw)
        b = icp.estimate(metric=metric, max_iter=k, tol=0.0, max_d2=max_d2, host_loop=False, timing=timing, **kw)
        assert a["iterations"] == b["iterations"] == k
        assert a["num_corr"] == b["num_corr"], (k, a["num_corr"], b["num_corr"])
        assert frob(a["T"], b["T"]) < 1e-6, (k, frob(a["T"], b["T"]))


def test_device_loop_convergence_count_and_flags(cb, ctx):
    dst, src, nrm, T_ref = _pair(40000, 6)
    icp = cb.Icp(ctx, cb.Cloud(ctx, dst, nrm), cb.Cloud(ctx, src))
    kw = dict(metric="combined", w_pt=0.1, w_pl=1.0, max_iter=40, tol=1e-6, max_d2=np.float32(0.03**2))
    a = icp.estimate(host_loop=True, **kw)
    b = icp.estimate(host_loop=False, **kw)
    assert a["iterations"] == b["iterations"] < 40 and a["converged"] and b["converged"]
    assert frob(a["T"], b["T"]) < 1e-6 and abs(a["last_delta"] - b["last_delta"]) < 1e-7
    # max_iter = 0: the loop body never runs (icp_base.hpp:76)
    z = icp.estimate(metric="p2p", max_iter=0, tol=0.0, max_d2=np.float32(0.03**2))
    assert z["iterations"] == 0 and np.array_equal(z["T"], np.hstack([np.eye(3), np.zeros((3, 1))]).astype(np.float32))


@pytest.mark.parametrize("metric,kw", [("p2p", {}), ("combined", dict(w_pt=0.1, w_pl=1.0))])
def test_device_loop_against_oracle(cb, ctx, orc, metric, kw):
    dst, src, nrm, _ = _pair(30000, 7)
    icp = cb.Icp(ctx, cb.Cloud(ctx, dst, nrm), cb.Cloud(ctx, src))
    max_d2 = np.float32(0.03**2)
    got = icp.estimate(metric=metric, max_iter=8, tol=0.0, max_d2=max_d2, **kw)
    want = orc.icp(dst, src, orc.BruteKnn(dst), metric=metric, dst_n=nrm if metric == "combined" else None, max_iter=8,
                   tol=0.0, max_d2=max_d2, **kw)
    assert got["num_corr"] == want["num_corr"]
    assert frob(got["T"], want["T"]) < 1e-5


@pytest.mark.parametrize("n,max_d,iters", [(50000, 0.03, 6), (50000, 0.004, 6), (200000, 0.02, 10)])
def test_cached_matches_are_the_exact_nearest_neighbours(cb, ctx, orc, n, max_d, iters):
    """After k iterations most queries were never searched again: what the cache holds for them must still be the
    unique nearest neighbour under the last searched transform (and the flagged ones were searched this iteration)."""
    dst, src, _, _ = _pair(n, 8, normals=False)
    icp = cb.Icp(ctx, cb.Cloud(ctx, dst), cb.Cloud(ctx, src))
    max_d2 = np.float32(max_d**2)
    icp.estimate(metric="p2p", max_iter=2, tol=0.0, max_d2=max_d2)
    assert 0 < icp.loop_cache()[2] < n  # the second iteration still searches some queries, but no longer all of them
    r = icp.estimate(metric="p2p", max_iter=iters, tol=0.0, max_d2=max_d2, timing=0)  # production setting (see above)
    T_search, near, searched = icp.loop_cache()
    assert searched < n // 4, searched  # the cache is doing its job by now
    q = orc.transform_points(T_search, src)
    idx, d2 = orc.RefKnn(dst).query(q, np.float32(3.0e38)) if orc.have_ref() else orc.BruteKnn(dst).query(q, np.float32(3.0e38))
    inside = d2 < max_d2
    # within the radius the cache must name the exact nearest neighbour
    assert np.array_equal(near[inside], idx[inside])
    # outside it, the cache may know the nearest candidate or nothing; if it names one, it is the nearest point
    named = (~inside) & (near >= 0)
    assert np.array_equal(near[named], idx[named])
    assert r["num_corr"] == int(inside.sum())


def test_loop_reuses_object_and_restarts_cold(cb, ctx):
    """Two estimate() calls on the same object with different starting transforms: the cache of the first must not
    leak into the second."""
    dst, src, _, T_ref = _pair(40000, 9, normals=False)
    icp = cb.Icp(ctx, cb.Cloud(ctx, dst), cb.Cloud(ctx, src))
    kw = dict(metric="p2p", max_iter=6, tol=0.0, max_d2=np.float32(0.03**2))
    a = icp.estimate(**kw)
    T0 = (0.5 * np.asarray(T_ref) + 0.5 * np.hstack([np.eye(3), np.zeros((3, 1))])).astype(np.float32)
    b = icp.estimate(T_init=T0, **kw)
    c = icp.estimate(T_init=T0, host_loop=True, **kw)
    assert b["num_corr"] == c["num_corr"] and frob(b["T"], c["T"]) < 1e-6
    a2 = icp.estimate(**kw)
    assert np.array_equal(a["T"], a2["T"]) and a["num_corr"] == a2["num_corr"]  # deterministic, run to run


@pytest.mark.parametrize("host_loop", [False, True])
@pytest.mark.parametrize("sig", [dict(pt_rbf_sigma=0.002, pl_rbf_sigma=0.003), dict(pl_rbf_sigma=0.0015), dict(pt_rbf_sigma=0.5)])
def test_rbf_correspondence_weights_against_oracle(cb, ctx, orc, host_loop, sig):
    """RBFKernelWeightEvaluator on the point-to-point / point-to-plane terms (common_pair_evaluators.hpp:46-79,
    transform_estimation.hpp:302-304, :331-333): both loops against the oracle's restatement."""
    dst, src, nrm, _ = _pair(30000, 11)
    icp = cb.Icp(ctx, cb.Cloud(ctx, dst, nrm), cb.Cloud(ctx, src))
    kw = dict(metric="combined", w_pt=0.2, w_pl=1.0, max_iter=7, tol=0.0, max_d2=np.float32(0.03**2), **sig)
    got = icp.estimate(host_loop=host_loop, **kw)
    want = orc.icp(dst, src, orc.BruteKnn(dst), dst_n=nrm, **kw)
    plain = orc.icp(dst, src, orc.BruteKnn(dst), dst_n=nrm, **{k: v for k, v in kw.items() if "rbf" not in k})
    assert got["num_corr"] == want["num_corr"]
    assert frob(got["T"], want["T"]) < 1e-5
    if min(sig.values()) < 0.1:
        assert frob(want["T"], plain["T"]) > 2e-7  # the weights matter in this configuration


def test_rbf_weights_with_inner_gauss_newton_steps_and_engine_mode(cb, ctx, orc):
    """Inner Gauss-Newton passes re-use the stored correspondence values; the pair-list path (non-default engine
    mode) carries them in the list."""
    dst, src, nrm, _ = _pair(20000, 12)
    icp = cb.Icp(ctx, cb.Cloud(ctx, dst, nrm), cb.Cloud(ctx, src))
    knn = orc.BruteKnn(dst)
    base = dict(metric="combined", w_pt=0.1, w_pl=1.0, max_iter=5, tol=0.0, max_d2=np.float32(0.03**2), pt_rbf_sigma=0.003,
                pl_rbf_sigma=0.003)
    for extra in (dict(max_opt_iter=3, opt_tol=0.0), dict(inlier_fraction=0.8), dict(one_to_one=True)):
        got = icp.estimate(**base, **extra)
        want = orc.icp(dst, src, knn, dst_n=nrm, **base, **extra)
        assert got["num_corr"] == want["num_corr"], extra
        assert frob(got["T"], want["T"]) < 1e-5, (extra, frob(got["T"], want["T"]))


def test_non_converging_run_is_handed_over_to_the_host_loop(cb, ctx):
    """A pose outside ICP's basin (SURVEY 8d's fixed 0.02 rad on a dense cloud): a third of the queries is searched
    again in every iteration, the exclusion cache does not pay, and after its first batch the device loop hands the run
    over to the host-driven loop - same iterations, same correspondences, same transform as the host loop alone."""
    n = 300000
    dst, src, _, _ = synth.icp_pair(n, seed=4, noise=0.001, T_ref=synth.t_ref_default())
    icp = cb.Icp(ctx, cb.Cloud(ctx, dst), cb.Cloud(ctx, src))
    kw = dict(metric="p2p", max_iter=9, tol=0.0, max_d2=np.float32(0.02**2))
    a = icp.estimate(host_loop=True, **kw)
    b = icp.estimate(host_loop=False, **kw)
    assert a["iterations"] == b["iterations"] == 9 and a["num_corr"] == b["num_corr"]
    assert frob(a["T"], b["T"]) < 1e-6
    assert len(b["iter_ms"]) == 9 and np.all(b["iter_ms"] > 0)
    with pytest.raises(cb.CbError):  # the run did not END on the device loop: there is no loop cache to inspect
        icp.loop_cache()
    # a converging run on the same object afterwards stays on the device loop
    dst2, src2, _, _ = _pair(n, 4, normals=False)
    icp2 = cb.Icp(ctx, cb.Cloud(ctx, dst2), cb.Cloud(ctx, src2))
    icp2.estimate(**kw)
    assert icp2.loop_cache()[2] < n // 4


def test_clouds_far_from_the_origin(cb, ctx):
    """Kabsch moments are taken about the pivots (dst mean, T * src mean) in both loops: a cloud 2000 units away from
    the origin registers as well as the same cloud at the origin (raw second moments would cancel ~1e7 against ~0.1)."""
    dst, src, _, T_ref = _pair(60000, 13, normals=False)
    off = np.array([2000.0, -1500.0, 800.0])
    # x_dst = R x_src + t  ->  (x_dst + off) = R (x_src + off) + (t + off - R off)
    R, t = np.asarray(T_ref)[:, :3], np.asarray(T_ref)[:, 3]
    T_far = np.hstack([R, (t + off - R @ off)[:, None]])
    dst_f, src_f = (dst + off).astype(np.float32), (src + off).astype(np.float32)
    icp = cb.Icp(ctx, cb.Cloud(ctx, dst_f), cb.Cloud(ctx, src_f))
    kw = dict(metric="p2p", max_iter=12, tol=0.0, max_d2=np.float32(0.03**2))
    a = icp.estimate(host_loop=True, **kw)
    b = icp.estimate(host_loop=False, **kw)
    assert a["num_corr"] == b["num_corr"] and np.abs(a["T"][:, :3] - b["T"][:, :3]).max() < 1e-6
    # rotation recovered as well as the same pair registers at the origin (noise 0.002, 60 k points: ~1e-5), given the
    # fp32 coordinate spacing of 1.2e-4 at |x| ~ 2000
    o = cb.Icp(ctx, cb.Cloud(ctx, dst), cb.Cloud(ctx, src)).estimate(**kw)
    err_far, err_origin = np.abs(b["T"][:, :3] - R).max(), np.abs(o["T"][:, :3] - R).max()
    assert err_far < max(1e-4, 3 * err_origin), (err_far, err_origin)
    assert np.abs(b["T"][:, 3] - T_far[:, 3]).max() < 0.3
