"""GPU parity: cb_grid_downsample / cb_cloud_grid_downsample vs the oracle — bit-exact points, normals, colours
(byte-for-byte: the device replays the reference's serial per-bin accumulation in point-index order), in both bin
orders, plus the size-independent properties of the operation at a size the oracle would take too long for."""
import numpy as np
import pytest

from cilantro_b200 import synth

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _same(got, want):
    for g, w in zip(got, want):
        assert (g is None) == (w is None)
        if g is not None:
            assert g.shape == w.shape, (g.shape, w.shape)
            assert np.array_equal(_bits(g), _bits(w)), f"{(_bits(g) != _bits(w)).sum()} differing words"


@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("n,bin_size", [(50000, 0.05), (200000, 0.013), (30000, 0.5), (4097, 0.001)])
def test_points_only_bitexact(cb, ctx, orc, n, bin_size, order):
    rng = np.random.default_rng(n)
    pts = (rng.random((n, 3), dtype=np.float32) * np.float32(2) - np.float32(0.7)).astype(np.float32)
    _same(cb.grid_downsample(ctx, pts, bin_size, order=order), orc.grid_downsample(pts, bin_size, order=order))


@pytest.mark.parametrize("order", [0, 1])
def test_normals_colors_and_min_points_bitexact(cb, ctx, orc, order):
    pts, nrm = synth.surface_cloud(120000, seed=3, noise=0.002)
    rng = np.random.default_rng(1)
    flip = rng.random(pts.shape[0]) < 0.3
    nrm = nrm.copy()
    nrm[flip] *= -1  # inconsistent input orientation exercises the sign rule
    col = rng.random((pts.shape[0], 3), dtype=np.float32)
    for minp in (1, 2, 5):
        got = cb.grid_downsample(ctx, pts, 0.02, normals=nrm, colors=col, min_points=minp, order=order)
        want = orc.grid_downsample(pts, 0.02, normals=nrm, colors=col, min_points=minp, order=order)
        _same(got, want)
    _same(cb.grid_downsample(ctx, pts, 0.02, normals=nrm, order=order), orc.grid_downsample(pts, 0.02, normals=nrm, order=order))
    _same(cb.grid_downsample(ctx, pts, 0.02, colors=col, order=order), orc.grid_downsample(pts, 0.02, colors=col, order=order))


def test_edge_cases(cb, ctx, orc):
    # empty input, one point, all points in one bin, negative coordinates across the origin, duplicates
    assert cb.grid_downsample(ctx, np.zeros((0, 3), np.float32), 0.1)[0].shape == (0, 3)
    one = np.array([[0.3, -0.2, 5.0]], np.float32)
    _same(cb.grid_downsample(ctx, one, 0.1), orc.grid_downsample(one, 0.1))
    rng = np.random.default_rng(2)
    blob = (rng.random((7000, 3), dtype=np.float32) * np.float32(0.09)).astype(np.float32)
    got = cb.grid_downsample(ctx, blob, 0.1)
    assert got[0].shape == (1, 3)
    _same(got, orc.grid_downsample(blob, 0.1))
    cross = np.vstack([blob - np.float32(0.045), blob[:100], blob[:100]]).astype(np.float32)
    for order in (0, 1):
        _same(cb.grid_downsample(ctx, cross, 0.03, order=order), orc.grid_downsample(cross, 0.03, order=order))
    assert cb.grid_downsample(ctx, blob, 0.1, min_points=7001)[0].shape == (0, 3)
    # far from the origin with a small bin: large bin coordinates, still exact
    far = (blob + np.float32(1000.0)).astype(np.float32)
    _same(cb.grid_downsample(ctx, far, 0.004), orc.grid_downsample(far, 0.004))
    with pytest.raises(cb.CbError):
        cb.grid_downsample(ctx, blob, 0.0)


def test_device_resident_pipeline(cb, ctx, orc):
    # gridDownsample -> estimateNormalsKNN on the device; the downsampled cloud never visits the host in between
    pts, _ = synth.surface_cloud(150000, seed=5, noise=0.0005)
    cloud = cb.Cloud(ctx, pts)
    ds = cloud.grid_downsample(0.01)
    want, _, _ = orc.grid_downsample(pts, 0.01)
    assert ds.n == want.shape[0]
    assert np.array_equal(_bits(ds.download()), _bits(want))
    vp = [0.5, 0.5, 10.0]
    got = ds.estimate_normals(k=8, view_point=vp, want_cov=True)
    ref = orc.estimate_normals(want, orc.make_knn(want), k=8, view_point=vp)
    assert np.array_equal(_bits(got["cov6"]), _bits(ref[2]))
    # a cloud with normals downsamples them too
    ds2 = ds.grid_downsample(0.03, min_points=2, order=1)
    p2, n2 = ds2.download(normals=True)
    w2 = orc.grid_downsample(want, 0.03, normals=got["normals"], min_points=2, order=1)
    assert np.array_equal(_bits(p2), _bits(w2[0])) and np.array_equal(_bits(n2), _bits(w2[1]))


def test_large_cloud_properties(cb, ctx):
    # 8 M points (oracle: minutes): occupied-bin count, bin membership and centroid location are checked with numpy
    n = 8_000_000
    rng = np.random.default_rng(11)
    pts = rng.random((n, 3), dtype=np.float32)
    b = np.float32(0.01)
    out, _, _ = cb.grid_downsample(ctx, pts, float(b))
    inv = np.float32(1.0) / b
    key = np.floor(pts * inv).astype(np.int64)
    flat = (key[:, 0] * 101 + key[:, 1]) * 101 + key[:, 2]
    uniq, counts = np.unique(flat, return_counts=True)
    assert out.shape[0] == uniq.shape[0]
    okey = np.floor(out * inv).astype(np.int64)
    oflat = (okey[:, 0] * 101 + okey[:, 1]) * 101 + okey[:, 2]
    # every centroid lies in its own bin (up to rounding at a face) and bins come out in ascending order
    assert np.mean(oflat == uniq) > 0.999
    sums = np.zeros((uniq.shape[0], 3))
    np.add.at(sums, np.searchsorted(uniq, flat), pts.astype(np.float64))
    assert np.allclose(out, sums / counts[:, None], atol=5e-6)
    # idempotence: downsampling the centroids with the same bins keeps one point per bin
    again, _, _ = cb.grid_downsample(ctx, out, float(b))
    assert again.shape[0] >= 0.999 * out.shape[0]
