"""CPU: the voxel-grid downsampling oracle (orc_grid_downsample) against hand-derivable answers.

Reference behaviour: core/grid_accumulator.hpp:117-126 (bin = floor(p / bin)), :9-39 (lexicographic bin order),
:187-199 (serial accumulation), core/grid_downsampler.hpp:20-37 / :97-105 (mean, normalised normal mean),
core/common_accumulators.hpp:122-131 (sign-consistent normal sum).
"""
import numpy as np


def test_known_bins_orders_and_min_points(orc):
    pts = np.array([
        [0.9, 0.1, 0.1],    # bin (1,0,0)   first seen 1st
        [-0.1, 0.1, 0.1],   # bin (-1,0,0)  first seen 2nd
        [0.6, 0.2, 0.3],    # bin (1,0,0)
        [0.1, 0.7, 0.2],    # bin (0,1,0)   first seen 3rd
        [0.2, 0.6, 0.9],    # bin (0,1,1)
        [0.7, 0.4, 0.4],    # bin (1,0,0)
        [0.1, 0.1, -0.2],   # bin (0,0,-1)
    ], np.float32)
    p0, _, _ = orc.grid_downsample(pts, 0.5, order=0)
    # map order: (-1,0,0) (0,0,-1) (0,1,0) (0,1,1) (1,0,0)
    want = np.array([pts[1], pts[6], pts[3], pts[4],
                     np.float32(1.0) / np.float32(3) * ((pts[0] + pts[2]) + pts[5])], np.float32)
    assert np.array_equal(p0.view(np.uint32), want.view(np.uint32))
    p1, _, _ = orc.grid_downsample(pts, 0.5, order=1)
    assert np.array_equal(p1, want[[4, 0, 2, 3, 1]])  # first-occurrence order
    p2, _, _ = orc.grid_downsample(pts, 0.5, min_points=2)
    assert p2.shape == (1, 3) and np.array_equal(p2[0], want[4])
    assert orc.grid_downsample(pts, 0.5, min_points=4)[0].shape == (0, 3)
    assert orc.grid_downsample(pts[:0], 0.5)[0].shape == (0, 3)


def test_normals_sign_consistency_and_colors(orc):
    pts = np.array([[0.1, 0.1, 0.1], [0.2, 0.2, 0.2], [0.3, 0.3, 0.3], [0.8, 0.8, 0.8]], np.float32)
    nrm = np.array([[0, 0, 1], [0, 0, -1], [0, 1, 0], [1, 0, 0]], np.float32)
    col = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0.5, 0.5, 0.5]], np.float32)
    p, n, c = orc.grid_downsample(pts, 0.5, normals=nrm, colors=col)
    assert p.shape == (2, 3)
    # bin 0: normal sum = (0,0,1) - (0,0,-1) [flipped] + (0,1,0) = (0,1,2) -> normalised
    assert np.allclose(n[0], np.array([0, 1, 2]) / np.sqrt(5), atol=1e-7)
    assert np.allclose(n[1], [1, 0, 0])
    assert np.allclose(c[0], [1 / 3, 1 / 3, 1 / 3], atol=1e-7) and np.allclose(c[1], [0.5, 0.5, 0.5])
    assert np.allclose(p[0], [0.2, 0.2, 0.2], atol=1e-7)


def test_parallel_build_agrees_with_serial_to_rounding(orc):
    # order = 2 restates the reference's default OpenMP build (per-thread maps merged in arrival order)
    rng = np.random.default_rng(9)
    pts = rng.random((200000, 3), dtype=np.float32)
    g = rng.standard_normal((200000, 3)).astype(np.float32)
    g = 0.2 * g + np.array([0, 0, 1], np.float32)  # one-sided: the sign rule never flips, the sums commute
    nrm = (g / np.linalg.norm(g, axis=1, keepdims=True)).astype(np.float32)
    s = orc.grid_downsample(pts, 0.05, normals=nrm, order=0)
    p = orc.grid_downsample(pts, 0.05, normals=nrm, order=2)
    assert s[0].shape == p[0].shape
    assert np.allclose(s[0], p[0], atol=1e-5)
    assert np.all(np.sum(s[1] * p[1], axis=1) > 0.99)


def test_random_cloud_against_numpy_grouping(orc):
    rng = np.random.default_rng(4)
    pts = (rng.random((20000, 3), dtype=np.float32) * 2 - 1).astype(np.float32)
    b = np.float32(0.13)
    out, _, _ = orc.grid_downsample(pts, b)
    inv = np.float32(1.0) / b
    key = np.floor(pts * inv).astype(np.int64)
    uniq, inverse, counts = np.unique(key, axis=0, return_inverse=True, return_counts=True)  # lexicographic rows
    assert out.shape[0] == uniq.shape[0]
    means = np.zeros((uniq.shape[0], 3))
    np.add.at(means, inverse.ravel(), pts.astype(np.float64))
    means /= counts[:, None]
    assert np.allclose(out, means, atol=2e-6)
