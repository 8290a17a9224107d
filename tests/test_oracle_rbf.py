"""Pins the oracle's combined-metric Gauss-Newton step WITH correspondence weight evaluators against an independent
numpy restatement of registration/transform_estimation.hpp:238-367 (the formulas written out again, in float64):
UnityWeightEvaluator and RBFKernelWeightEvaluator<float, float, true> (core/common_pair_evaluators.hpp:29-79), called as
evaluator(indexInFirst, indexInSecond, value) with value = the correspondence's squared distance (:302-304, :331-333).
CPU only."""
import numpy as np
import pytest


def _angle_axis(axis, theta):
    k = np.asarray(axis, np.float64)
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(theta) * K + (1 - np.cos(theta)) * (K @ K)


def _gn_step_numpy(dst, nrm, src, w_pt, w_pl, max_d2, pt_sigma, pl_sigma):
    """One ICP iteration from the identity, one Gauss-Newton step: returns the 3x4 transform."""
    dst64, src64, nrm64 = dst.astype(np.float64), src.astype(np.float64), nrm.astype(np.float64)
    # correspondences (SECOND_TO_FIRST): nearest dst point of every src point, kept iff d2 < max_d2
    d2 = ((src64[:, None, :] - dst64[None, :, :]) ** 2).sum(-1)
    nn = d2.argmin(1)
    val = d2[np.arange(len(src)), nn]
    keep = val < max_d2
    dm, sm = dst64.mean(0), src64.mean(0)
    AtA, Atb = np.zeros((6, 6)), np.zeros(6)

    def weight(sigma, v):
        if sigma is None:
            return 1.0
        coeff = np.float32(-0.5) / (np.float32(sigma) * np.float32(sigma))
        return float(np.exp(np.float64(coeff) * v))

    for j in np.nonzero(keep)[0]:
        i = nn[j]
        d, s = dst64[i] - dm, src64[j] - sm
        v, e = d + s, d - s
        if w_pt > 0:
            w = w_pt * weight(pt_sigma, val[j])
            E = np.zeros((6, 3))
            E[0, 1], E[0, 2], E[1, 2] = -v[2], v[1], -v[0]
            E[1, 0], E[2, 0], E[2, 1] = -E[0, 1], -E[0, 2], -E[1, 2]
            E[3:, :] = np.eye(3)
            AtA += w * (E @ E.T)
            Atb += w * (E @ e)
        if w_pl > 0:
            w = w_pl * weight(pl_sigma, val[j])
            n = nrm64[i]
            a = np.concatenate([np.cross(v, n), n])
            AtA += w * np.outer(a, a)
            Atb += w * float(n @ e) * a
    x = np.linalg.solve(AtA, Atb)
    na = np.linalg.norm(x[:3])
    theta = np.arctan(na)
    Ra = _angle_axis(x[:3] / na, theta) if na > 0 else np.eye(3)
    ta = np.cos(theta) * x[3:]
    L, t = Ra @ Ra, Ra @ ta                      # Ra * Translation(ta) * Ra
    t = dm + t - L @ sm                          # Translation(dst_mean) * tform * Translation(-src_mean)
    U, _, Vt = np.linalg.svd(L)                  # rotation() of the update (icp_single_transform_combined_metric.hpp:207-211)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        U[:, 0] = -U[:, 0]
    return np.hstack([U @ Vt, t[:, None]]), int(keep.sum())


_SIGMAS = [(None, None), (0.05, None), (None, 0.03), (0.04, 0.06)]
_WEIGHTS = [(0.0, 1.0), (0.3, 1.0), (1.0, 0.0)]
# an evaluator on a switched-off term (metric weight 0) is never read: those combinations are not cases
_CASES = [(w, s) for w in _WEIGHTS for s in _SIGMAS
          if not ((w[0] == 0.0 and s[0] is not None) or (w[1] == 0.0 and s[1] is not None))]


@pytest.mark.parametrize("weights,sig", _CASES)
def test_weighted_gauss_newton_step_matches_numpy(orc, sig, weights):
    w_pt, w_pl = weights
    rng = np.random.default_rng(17)
    dst = rng.random((400, 3)).astype(np.float32)
    g = rng.standard_normal((400, 3))
    nrm = (g / np.linalg.norm(g, axis=1, keepdims=True)).astype(np.float32)
    ang = 0.02
    R = _angle_axis(np.array([0.3, -0.5, 0.81]) / np.linalg.norm([0.3, -0.5, 0.81]), ang)
    src = ((dst.astype(np.float64) - np.array([0.01, -0.02, 0.015])) @ R + 0.004 * rng.standard_normal((400, 3))).astype(np.float32)
    max_d2 = np.float32(0.08**2)
    want, n_want = _gn_step_numpy(dst, nrm, src, w_pt, w_pl, float(max_d2), sig[0], sig[1])
    got = orc.icp(dst, src, orc.BruteKnn(dst), metric="combined", dst_n=nrm, max_iter=1, tol=0.0, max_d2=max_d2, w_pt=w_pt,
                  w_pl=w_pl, pt_rbf_sigma=sig[0], pl_rbf_sigma=sig[1], accum_double=True)
    assert got["num_corr"] == n_want
    assert np.abs(got["T"].astype(np.float64) - want).max() < 2e-6, np.abs(got["T"] - want).max()
    if sig != (None, None):  # and the evaluators do change the step
        plain = orc.icp(dst, src, orc.BruteKnn(dst), metric="combined", dst_n=nrm, max_iter=1, tol=0.0, max_d2=max_d2,
                        w_pt=w_pt, w_pl=w_pl, accum_double=True)
        assert np.abs(plain["T"] - got["T"]).max() > 1e-6
