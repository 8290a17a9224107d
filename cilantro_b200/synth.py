"""Seeded synthetic workloads; shared by tests/ and bench.py. numpy only."""
import numpy as np


def rigid_from_axis_angle(axis, angle, t):
    axis = np.asarray(axis, np.float64)
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)
    T = np.zeros((3, 4), np.float64)
    T[:, :3] = R
    T[:, 3] = np.asarray(t, np.float64)
    return T


def invert(T):
    T = np.asarray(T, np.float64)
    R, t = T[:, :3], T[:, 3]
    out = np.zeros((3, 4), np.float64)
    out[:, :3] = R.T
    out[:, 3] = -R.T @ t
    return out


def apply(T, pts):
    T = np.asarray(T, np.float64)
    return (np.asarray(pts, np.float64) @ T[:, :3].T + T[:, 3]).astype(np.float32)


def t_ref_default():
    """AngleAxis(0.02 rad about (1,1,1)/sqrt(3)), t = (0.01, -0.005, 0.008)."""
    return rigid_from_axis_angle([1, 1, 1], 0.02, [0.01, -0.005, 0.008])


def t_ref_for(n):
    """Generating pose for an n-point uniform cloud in the unit cube.

    SURVEY.md §8(d) proposes AngleAxis(0.02 rad, (1,1,1)/sqrt 3), t = (0.01,-0.005,0.008) for every
    size. Measured here with the oracle: at 1 M points (mean spacing 0.01) that offset (up to 0.03 at
    the cube corners) is outside ICP's basin of convergence on a dense uniform cloud — nearest
    neighbours are almost all wrong matches and BOTH the reference path and ours stall at
    |T - T_ref|_F = 3e-2. The work per iteration is unchanged, but the run is useless as a
    correctness check, so the pose is scaled with the point spacing s = n^(-1/3): angle =
    min(0.02, s/4), translation scaled by the same factor. For n <= 2000 this is SURVEY's pose."""
    s = float(n) ** (-1.0 / 3.0)
    f = min(1.0, (s / 4.0) / 0.02)
    return rigid_from_axis_angle([1, 1, 1], 0.02 * f, [0.01 * f, -0.005 * f, 0.008 * f])


def icp_pair(n, seed=1, noise=0.001, with_normals=False, n_src=None, T_ref=None):
    """dst uniform in [0,1)^3; src = T_ref^-1 dst + uniform noise in +-noise.

    Returns dst (n,3), src (n_src,3), dst_normals or None, T_ref (3,4 float64): the transform ICP
    should recover (src -> dst)."""
    rng = np.random.default_rng(seed)
    dst = rng.random((n, 3), dtype=np.float32)
    if T_ref is None:
        T_ref = t_ref_for(n)
    m = n if n_src is None else n_src
    base = dst[:m] if m <= n else rng.random((m, 3), dtype=np.float32)
    src = apply(invert(T_ref), base)
    src = (src + (rng.random((m, 3), dtype=np.float32) - 0.5) * np.float32(2 * noise)).astype(np.float32)
    nrm = None
    if with_normals:
        g = rng.standard_normal((n, 3)).astype(np.float32)
        nrm = (g / np.linalg.norm(g, axis=1, keepdims=True)).astype(np.float32)
    return dst, src, nrm, T_ref


def surface_cloud(n, seed=1, noise=0.0005):
    """A scanned-surface stand-in: the sheet z = 0.5 + 0.1 sin(6x) cos(5y) over [0,1)^2, sampled uniformly in
    (x, y) with Gaussian noise along z. Returns points (n,3) float32 and the analytic unit normals (n,3), +z side."""
    rng = np.random.default_rng(seed)
    xy = rng.random((n, 2))
    x, y = xy[:, 0], xy[:, 1]
    z = 0.5 + 0.1 * np.sin(6 * x) * np.cos(5 * y) + noise * rng.standard_normal(n)
    g = np.stack([-0.6 * np.cos(6 * x) * np.cos(5 * y), 0.5 * np.sin(6 * x) * np.sin(5 * y), np.ones(n)], axis=1)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    return np.stack([x, y, z], axis=1).astype(np.float32), g.astype(np.float32)


def rigid_icp_example_pair(points, normals, seed=1):
    """The input recipe of the reference's examples/rigid_icp.cpp:25-65 on a loaded scan (BASELINE config 1):
    src = dst + 0.01 * U(-1,1)^3 (normals + 0.02 * U, re-normalised), dst keeps only x > -0.4, then
    src <- tf_ref * src with tf_ref = Rz(-0.1) Ry(0.1) Rx(-0.1), t = (-0.20, -0.05, 0.10).
    Returns dst_p, dst_n, src_p, src_n, tf_ref (3x4 float64); ICP should recover tf_ref^-1."""
    rng = np.random.default_rng(seed)
    p = np.asarray(points, np.float32)
    n = np.asarray(normals, np.float32)
    src_p = (p + np.float32(0.01) * (rng.random(p.shape, dtype=np.float32) * 2 - 1)).astype(np.float32)
    src_n = n + np.float32(0.02) * (rng.random(n.shape, dtype=np.float32) * 2 - 1)
    src_n = (src_n / np.linalg.norm(src_n, axis=1, keepdims=True)).astype(np.float32)
    keep = p[:, 0] > np.float32(-0.4)
    dst_p, dst_n = np.ascontiguousarray(p[keep]), np.ascontiguousarray(n[keep])

    def rot(axis, a):
        return np.asarray(rigid_from_axis_angle(axis, a, [0, 0, 0]))[:, :3]

    R = rot([0, 0, 1], -0.1) @ rot([0, 1, 0], 0.1) @ rot([1, 0, 0], -0.1)
    tf_ref = np.hstack([R, np.array([[-0.20], [-0.05], [0.10]])])
    src_p = apply(tf_ref, src_p).astype(np.float32)
    src_n = (src_n.astype(np.float64) @ R.T).astype(np.float32)
    return dst_p, dst_n, src_p, src_n, tf_ref


def kmeans_data(n, k, seed=1):
    """uniform [0,1)^3 points; initial centroids = first k points of a seeded shuffle (config 4)."""
    rng = np.random.default_rng(seed)
    pts = rng.random((n, 3), dtype=np.float32)
    idx = rng.permutation(n)[:k]
    return pts, pts[idx].copy()


def ransac_pairs(n, inlier_frac=0.3, seed=1, sigma=0.002):
    """src uniform [0,1)^3; dst = T_ref src + N(0, sigma^2) for inliers, uniform otherwise (config 5)."""
    rng = np.random.default_rng(seed)
    src = rng.random((n, 3), dtype=np.float32)
    T_ref = t_ref_default()
    dst = apply(T_ref, src) + (rng.standard_normal((n, 3)) * sigma).astype(np.float32)
    out = rng.random(n) >= inlier_frac
    dst[out] = rng.random((int(out.sum()), 3), dtype=np.float32)
    return dst.astype(np.float32), src, T_ref, ~out


def frobenius(Ta, Tb):
    return float(np.linalg.norm(np.asarray(Ta, np.float64) - np.asarray(Tb, np.float64)))
