"""Multi-GPU equivalence (needs >= 2 GPUs on the box; skipped otherwise): source points sharded over 2
ranks + the library's own NCCL all-reduce must give the single-GPU transform (identical correspondences,
moments equal up to summation order) and identical results on both ranks."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from cilantro_b200 import capi, dist as cdist, synth

    cdist.init_process_group(backend="nccl")
    ctx = capi.Context(rank)
    assert cdist.attach_comm(ctx) == (rank, world)
    dst, src, nrm, T_ref = synth.icp_pair(200000, seed=5, noise=0.001, with_normals=True)
    lo, hi = cdist.shard_bounds(src.shape[0], rank, world)
    # the replicated destination cloud from its blocks (each rank uploads half, NVLink exchange): bit-identical to a
    # plain upload of the whole cloud
    dlo, dhi = cdist.shard_bounds(dst.shape[0], rank, world)
    d_dst = capi.Cloud.replicated(ctx, dst[dlo:dhi], nrm[dlo:dhi], dlo, dst.shape[0])
    got_p, got_n = d_dst.download(normals=True)
    assert np.array_equal(got_p.view(np.uint32), dst.view(np.uint32)) and np.array_equal(got_n.view(np.uint32), nrm.view(np.uint32))
    d_src = capi.Cloud(ctx, src[lo:hi], None, index_offset=lo)
    icp = capi.Icp(ctx, d_dst, d_src)
    out = {}
    for name, kw in (("p2p", dict(metric="p2p")), ("comb", dict(metric="combined", w_pt=0.1, w_pl=1.0))):
        r = icp.estimate(max_iter=8, tol=0.0, max_d2=np.float32(0.03**2), **kw)
        out[name + "_T"] = r["T"]
        out[name + "_n"] = np.array([r["num_corr"]])
    # non-default correspondence-engine modes with sharded source points: every rank works on the replicated cloud
    for name, kw in (("eng_frac", dict(metric="p2p", inlier_fraction=0.8, one_to_one=True)),
                     ("eng_both", dict(metric="combined", w_pt=0.1, w_pl=1.0, search_dir="both", require_reciprocal=True))):
        r = icp.estimate(max_iter=4, tol=0.0, max_d2=np.float32(0.03**2), **kw)
        f, s2, v = icp.correspondences()
        out[name + "_T"] = r["T"]
        out[name + "_n"] = np.array([r["num_corr"]])
        out[name + "_f"], out[name + "_s"], out[name + "_v"] = f, s2, v
    # k-means: points sharded, one all-reduce of K x 4 sums per iteration
    pts, cent0 = synth.kmeans_data(300000, 64, seed=3)
    plo, phi = cdist.shard_bounds(pts.shape[0], rank, world)
    km = capi.kmeans_cluster(ctx, capi.Cloud(ctx, pts[plo:phi]), cent0, max_iter=6, tol=0.0)
    out["km_cent"] = km["centroids"]
    out["km_labels"] = km["labels"]
    # PCA over the sharded cloud
    p = capi.pca(ctx, capi.Cloud(ctx, pts[plo:phi]))
    out["pca_cov"] = p["cov"]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **out)
    ctx.close()
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_sharded_equals_single_gpu(cb, ctx, tmp_path):
    import torch
    import torch.multiprocessing as mp

    from cilantro_b200 import synth

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    dst, src, nrm, T_ref = synth.icp_pair(200000, seed=5, noise=0.001, with_normals=True)
    icp = cb.Icp(ctx, cb.Cloud(ctx, dst, nrm), cb.Cloud(ctx, src))
    for name, kw in (("p2p", dict(metric="p2p")), ("comb", dict(metric="combined", w_pt=0.1, w_pl=1.0))):
        single = icp.estimate(max_iter=8, tol=0.0, max_d2=np.float32(0.03**2), **kw)
        assert np.array_equal(r0[name + "_T"], r1[name + "_T"]), "ranks must agree bit for bit"
        assert r0[name + "_n"][0] == single["num_corr"]
        assert np.abs(r0[name + "_T"] - single["T"]).max() < 2e-7
    for name, kw in (("eng_frac", dict(metric="p2p", inlier_fraction=0.8, one_to_one=True)),
                     ("eng_both", dict(metric="combined", w_pt=0.1, w_pl=1.0, search_dir="both", require_reciprocal=True))):
        single = icp.estimate(max_iter=4, tol=0.0, max_d2=np.float32(0.03**2), **kw)
        f, s2, v = icp.correspondences()
        for r in (r0, r1):  # both ranks hold the global list of the single-GPU run, and its transform bit for bit
            assert np.array_equal(r[name + "_T"], single["T"]) and r[name + "_n"][0] == single["num_corr"]
            assert np.array_equal(r[name + "_f"], f) and np.array_equal(r[name + "_s"], s2)
            assert np.array_equal(r[name + "_v"].view(np.uint32), v.view(np.uint32))
    pts, cent0 = synth.kmeans_data(300000, 64, seed=3)
    km = cb.kmeans_cluster(ctx, cb.Cloud(ctx, pts), cent0, max_iter=6, tol=0.0)
    assert np.array_equal(r0["km_cent"], r1["km_cent"])
    assert np.abs(r0["km_cent"] - km["centroids"]).max() < 1e-6
    labels = np.concatenate([r0["km_labels"], r1["km_labels"]])
    assert (labels != km["labels"]).sum() <= 3
    p = cb.pca(ctx, cb.Cloud(ctx, pts))
    assert np.allclose(r0["pca_cov"], p["cov"], rtol=1e-6, atol=1e-8) and np.array_equal(r0["pca_cov"], r1["pca_cov"])


def _worker_dead_peer(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), CB_EXCHANGE_TIMEOUT_MS="400")
    import time

    from cilantro_b200 import capi, dist as cdist, synth

    cdist.init_process_group(backend="nccl")
    ctx = capi.Context(rank)
    assert cdist.attach_comm(ctx) == (rank, world)
    dst, src, _, _ = synth.icp_pair(50000, seed=6, noise=0.001)
    lo, hi = cdist.shard_bounds(src.shape[0], rank, world)
    icp = capi.Icp(ctx, capi.Cloud(ctx, dst), capi.Cloud(ctx, src[lo:hi], None, index_offset=lo))
    kw = dict(metric="p2p", max_iter=5, tol=0.0, max_d2=np.float32(0.03**2))
    ok = icp.estimate(**kw)  # both ranks take part: fine
    status = "ok"
    t0 = time.perf_counter()
    if rank == 0:
        # rank 1 never issues its passes: the in-kernel wait for its row must give up after CB_EXCHANGE_TIMEOUT_MS and
        # the call must come back with an error instead of hanging (reduce.cuh, exchange_rows)
        try:
            icp.estimate(**kw)
            status = "returned without an error"
        except capi.CbError as e:
            status = "error: " + str(e)
    dt = time.perf_counter() - t0
    with open(os.path.join(out_dir, f"dead{rank}.txt"), "w") as f:
        f.write(f"{status}\n{dt}\n{ok['iterations']}\n")
    import torch.distributed as dist

    dist.barrier()
    dist.destroy_process_group()


def test_dead_peer_times_out_instead_of_hanging(cb, tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker_dead_peer, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    status, dt, iters = open(tmp_path / "dead0.txt").read().split("\n")[:3]
    assert int(iters) == 5
    assert status.startswith("error") and "peer rank" in status, status
    assert float(dt) < 10.0, dt
