"""GPU parity: grid nearest-neighbour search vs the oracle (bit-exact indices and squared distances).

Calls go through the C ABI (cb_knn1_radius / cb_find_correspondences / cb_knn_radius). Oracle:
orc.BruteKnn (restatement; lowest index wins exact ties — the same rule as the CUDA path) and,
where oracle/_ref exists, the reference's own nanoflann (ties may pick another index with a
bit-equal d2).
"""
import numpy as np
import pytest

from cilantro_b200 import synth

pytestmark = pytest.mark.gpu
FMAX = float(np.finfo(np.float32).max)


def _check_exact(cb, ctx, orc, dst, qry, T, max_d2):
    ref = cb.Cloud(ctx, dst)
    q = cb.Cloud(ctx, qry)
    idx, d2 = cb.knn1_radius(ctx, ref, q, T, max_d2)
    qt = orc.transform_points(T if T is not None else orc.identity(), qry)
    oi, od = orc.BruteKnn(dst).query(qt, max_d2)
    assert np.array_equal(idx, oi), f"{(idx != oi).sum()} index mismatches"
    assert np.array_equal(d2.view(np.uint32), od.view(np.uint32)), "squared distances differ bitwise"
    return idx, d2


@pytest.mark.parametrize("n,m,max_d2", [(20000, 20000, 0.02**2), (50000, 7777, 0.05**2), (3000, 9000, FMAX)])
def test_knn1_matches_oracle_bitexact(cb, ctx, orc, n, m, max_d2):
    dst, src, _, T_ref = synth.icp_pair(n, seed=3, noise=0.002, n_src=min(m, n))
    if m > n:
        src = np.random.default_rng(5).random((m, 3), dtype=np.float32)
    T = (T_ref * 0.9 + 0.1 * np.hstack([np.eye(3), np.zeros((3, 1))])).astype(np.float32)
    _check_exact(cb, ctx, orc, dst, src, T, max_d2)


def test_knn1_identity_transform_none(cb, ctx, orc):
    rng = np.random.default_rng(11)
    dst = rng.random((10000, 3), dtype=np.float32)
    qry = rng.random((5000, 3), dtype=np.float32)
    _check_exact(cb, ctx, orc, dst, qry, None, 0.03**2)


def test_knn1_edge_cases(cb, ctx, orc):
    rng = np.random.default_rng(7)
    dst = rng.random((5000, 3), dtype=np.float32)
    # queries far outside the bounding box, on its faces, and NaN-free extremes
    qry = np.vstack([
        rng.random((200, 3), dtype=np.float32) * 4 - 2,
        np.array([[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [-10, 0.5, 0.5], [0.5, 20, 0.5], [5, 5, -5]], np.float32),
    ]).astype(np.float32)
    for max_d2 in (0.05**2, 1.0, FMAX):
        _check_exact(cb, ctx, orc, dst, qry, None, max_d2)
    # tiny radius: almost nothing qualifies
    idx, d2 = _check_exact(cb, ctx, orc, dst, qry, None, 1e-12)
    assert (idx >= 0).sum() == 0 and np.all(d2 == np.float32(1e-12))


def test_knn1_ties_pick_lowest_index(cb, ctx, orc):
    # duplicated reference points: exact ties everywhere
    rng = np.random.default_rng(2)
    base = rng.random((500, 3), dtype=np.float32)
    dst = np.vstack([base, base, base])[rng.permutation(1500)]
    qry = base + np.float32(1e-3)
    idx, _ = _check_exact(cb, ctx, orc, dst, qry, None, 0.1)
    # the winner is the lowest index among the coincident copies
    for i in range(0, 500, 50):
        same = np.where((dst == dst[idx[i]]).all(axis=1))[0]
        assert idx[i] == same.min()


def test_knn1_degenerate_clouds(cb, ctx, orc):
    rng = np.random.default_rng(4)
    plane = rng.random((4000, 3), dtype=np.float32)
    plane[:, 2] = 0.25  # planar reference set
    line = np.zeros((300, 3), np.float32)
    line[:, 0] = np.linspace(0, 1, 300)
    single = np.array([[0.3, 0.3, 0.3]], np.float32)
    coincident = np.repeat(single, 100, axis=0)
    qry = rng.random((1000, 3), dtype=np.float32)
    for dst in (plane, line, single, coincident):
        _check_exact(cb, ctx, orc, dst, qry, None, FMAX)
        _check_exact(cb, ctx, orc, dst, qry, None, 0.1**2)


def test_knn1_empty_inputs(cb, ctx):
    empty = cb.Cloud(ctx, np.zeros((0, 3), np.float32))
    pts = cb.Cloud(ctx, np.random.default_rng(0).random((100, 3), dtype=np.float32))
    idx, d2 = cb.knn1_radius(ctx, empty, pts, None, 1.0)
    assert np.all(idx == -1) and np.all(d2 == 1.0)
    idx, d2 = cb.knn1_radius(ctx, pts, empty, None, 1.0)
    assert idx.shape == (0,)
    i1, i2, v = cb.find_correspondences(ctx, empty, pts, None, 1.0)
    assert i1.size == 0


def test_find_correspondences_matches_oracle(cb, ctx, orc):
    dst, src, _, T_ref = synth.icp_pair(30000, seed=9, noise=0.004)
    T = T_ref.astype(np.float32)
    max_d2 = np.float32(0.004**2)  # many queries have no neighbour inside the radius
    ref, q = cb.Cloud(ctx, dst), cb.Cloud(ctx, src)
    i1, i2, v = cb.find_correspondences(ctx, ref, q, T, max_d2)
    o1, o2, ov = orc.find_correspondences(T, src, orc.BruteKnn(dst), max_d2)
    assert 0 < i1.size < src.shape[0]
    assert np.array_equal(i1, o1) and np.array_equal(i2, o2)
    assert np.array_equal(v.view(np.uint32), ov.view(np.uint32))
    assert np.all(np.diff(i2) > 0), "correspondences must be compacted in query order"


def test_knn1_vs_reference_nanoflann_250k(cb, ctx, orc):
    """Against the reference's own kd-tree: equal index, or an exact tie (bit-equal d2)."""
    if not orc.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    dst, src, _, T_ref = synth.icp_pair(250000, seed=1, noise=0.001)
    T = T_ref.astype(np.float32)
    max_d2 = np.float32(0.02**2)
    idx, d2 = cb.knn1_radius(ctx, cb.Cloud(ctx, dst), cb.Cloud(ctx, src), T, max_d2)
    ri, rd = orc.RefKnn(dst).query(orc.transform_points(T, src), max_d2)
    assert np.array_equal(d2.view(np.uint32), rd.view(np.uint32))
    diff = idx != ri
    assert diff.sum() <= 5, f"{diff.sum()} index differences (only exact ties may differ)"


def test_knn_k_matches_numpy(cb, ctx):
    rng = np.random.default_rng(13)
    dst = rng.random((4000, 3), dtype=np.float32)
    qry = rng.random((300, 3), dtype=np.float32)
    ref, q = cb.Cloud(ctx, dst), cb.Cloud(ctx, qry)
    for k, r2 in ((2, FMAX), (8, 0.08**2), (20, FMAX)):
        idx, d2, cnt = cb.knn_radius(ctx, ref, q, k, None, r2)
        # numpy restatement with the same fp32 arithmetic order
        dx = qry[:, None, 0] - dst[None, :, 0]
        dy = qry[:, None, 1] - dst[None, :, 1]
        dz = qry[:, None, 2] - dst[None, :, 2]
        D = (dx * dx + dy * dy) + dz * dz
        order = np.lexsort((np.broadcast_to(np.arange(dst.shape[0]), D.shape), D), axis=1)[:, :k]
        for i in range(qry.shape[0]):
            ok = D[i, order[i]] < r2
            want = order[i][ok]
            assert cnt[i] == want.size
            assert np.array_equal(idx[i, : want.size], want)
            assert np.array_equal(d2[i, : want.size], D[i, want])
            assert np.all(idx[i, want.size:] == -1)


def test_kd_tree_example_known_answer(cb, ctx):
    """examples/kd_tree.cpp:6-19: unit-cube corners, query (0.1,0.1,0.4), k=2, r2=1.001 -> 0,3 / 0.18,0.38."""
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [0, 1, 1], [1, 0, 1], [1, 1, 0], [1, 1, 1]], np.float32)
    ref = cb.Cloud(ctx, pts)
    q = cb.Cloud(ctx, np.array([[0.1, 0.1, 0.4]], np.float32))
    idx, d2, cnt = cb.knn_radius(ctx, ref, q, 2, None, 1.001)
    assert cnt[0] == 2 and list(idx[0]) == [0, 3]
    assert np.allclose(d2[0], [0.18, 0.38], rtol=1e-6)
