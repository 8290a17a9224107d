"""N > 1 host logic on CPU: world_size-2 gloo. Source points are sharded in contiguous blocks, each
rank reduces ITS correspondences to the 16 Kabsch moments / 28 normal-equation values (here with
numpy standing in for the accumulation kernel), one all-reduce sums them, and every rank solves
redundantly with the library's host solver -> identical transform on every rank, equal to the
single-process result. Also covers shard_bounds and the unique-id broadcast used by attach_comm.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _moments(dst, q, i1, i2):
    d, s = dst[i1].astype(np.float64), q[i2].astype(np.float64)
    out = np.zeros(16)
    out[0] = len(i1)
    out[1:4] = d.sum(0)
    out[4:7] = s.sum(0)
    out[7:] = (d.T @ s).reshape(-1)
    return out


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist

    import oracle
    from cilantro_b200 import capi, dist as cdist, synth

    r, w, _ = cdist.init_process_group(backend="gloo")
    assert (r, w) == (rank, world)
    # unique-id style broadcast
    payload = cdist.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 0)
    assert payload == bytes(range(128))

    dst, src, _, T_ref = synth.icp_pair(6000, seed=17, noise=0.001)
    T = np.hstack([np.eye(3), np.zeros((3, 1))]).astype(np.float32)
    max_d2 = np.float32(0.05**2)
    lo, hi = cdist.shard_bounds(src.shape[0], rank, world)
    knn = oracle.BruteKnn(dst)
    T_iter = None
    for _ in range(3):  # three ICP iterations, each: shard-local moments -> all-reduce -> redundant solve
        q = oracle.transform_points(T, src[lo:hi])
        i1, i2, _v = oracle.find_correspondences(T, src[lo:hi], knn, max_d2)
        total = cdist.allreduce_sum_f64(_moments(dst, q, i1, i2))
        T_iter, ok = capi.solve_kabsch_moments(total)
        R = capi.solve_rotation(T_iter[:, :3])
        T_iter[:, :3] = R
        T = capi.compose(T_iter, T)
    np.save(os.path.join(out_dir, f"T_rank{rank}.npy"), T)
    np.save(os.path.join(out_dir, f"n_rank{rank}.npy"), np.array([total[0]]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_exactly():
    from cilantro_b200.dist import shard_bounds

    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_sharded_icp_equals_single_process(tmp_path, orc, cb):
    from cilantro_b200 import synth

    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    T0 = np.load(tmp_path / "T_rank0.npy")
    T1 = np.load(tmp_path / "T_rank1.npy")
    assert np.array_equal(T0, T1), "ranks must hold bit-identical transforms (redundant solve)"
    # single-process reference with the same pipeline
    dst, src, _, T_ref = synth.icp_pair(6000, seed=17, noise=0.001)
    T = np.hstack([np.eye(3), np.zeros((3, 1))]).astype(np.float32)
    knn = orc.BruteKnn(dst)
    for _ in range(3):
        q = orc.transform_points(T, src)
        i1, i2, _v = orc.find_correspondences(T, src, knn, np.float32(0.05**2))
        Ti, _ = cb.solve_kabsch_moments(_moments(dst, q, i1, i2))
        Ti[:, :3] = cb.solve_rotation(Ti[:, :3])
        T = cb.compose(Ti, T)
    assert np.abs(T0 - T).max() < 1e-6
    assert np.load(tmp_path / "n_rank0.npy")[0] == len(i1)
    # and the oracle's own ICP agrees
    ref = orc.icp(dst, src, knn, metric="p2p", max_iter=3, tol=0.0, max_d2=np.float32(0.05**2))
    assert np.linalg.norm(T0.astype(np.float64) - ref["T"]) < 1e-5


class _FakeCtx:
    """Stands in for capi.Context in attach_comm (no GPU here): records the calls, fails the peer mapping on demand."""

    def __init__(self, capi, fail_attach):
        self.capi, self.fail_attach, self.calls = capi, fail_attach, []

    def init_comm(self, uid, rank, world):
        assert len(uid) == 128
        self.calls.append(("init_comm", rank, world))

    def ipc_handle(self):
        return bytes(64)

    def ipc_attach(self, handles):
        self.calls.append(("ipc_attach", len(handles)))
        if self.fail_attach:
            raise self.capi.CbError("no peer access")

    def ipc_detach(self):
        self.calls.append(("ipc_detach",))


def _attach_worker(rank, world, port, out_dir, failing_rank):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    os.environ.pop("CB_NO_FUSED_EXCHANGE", None)
    os.environ.pop("CB_TEST_IPC_FAIL_RANK", None)
    import torch.distributed as dist

    from cilantro_b200 import capi, dist as cdist

    cdist.init_process_group(backend="gloo")
    capi.comm_unique_id = lambda: bytes(128)  # the real one asks NCCL; the agreement logic is what is under test
    ctx = _FakeCtx(capi, fail_attach=(rank == failing_rank))
    r, w = cdist.attach_comm(ctx)
    assert (r, w) == (rank, world)
    with open(os.path.join(out_dir, f"calls_{failing_rank}_{rank}.txt"), "w") as f:
        f.write(repr(ctx.calls))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("failing_rank", [-1, 1])
def test_attach_comm_ranks_agree_on_the_exchange_path(tmp_path, failing_rank):
    """If the peer mapping of the fused exchange fails on ANY rank, EVERY rank must go back to the NCCL path
    (cb_comm_ipc_detach) — a rank left on the fused path would wait for rows that never come."""
    world = 2
    mp.spawn(_attach_worker, args=(world, _free_port(), str(tmp_path), failing_rank), nprocs=world, join=True)
    for rank in range(world):
        calls = eval(open(tmp_path / f"calls_{failing_rank}_{rank}.txt").read())
        names = [c[0] for c in calls]
        assert names[:2] == ["init_comm", "ipc_attach"] and calls[1][1] == 64 * world
        assert ("ipc_detach" in names) == (failing_rank >= 0), (rank, names)


def _optout_worker(rank, world, port, out_dir, optout_rank):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    os.environ.pop("CB_TEST_IPC_FAIL_RANK", None)
    os.environ.pop("CB_NO_FUSED_EXCHANGE", None)
    if rank == optout_rank:
        os.environ["CB_NO_FUSED_EXCHANGE"] = "1"
    import torch.distributed as dist

    from cilantro_b200 import capi, dist as cdist

    cdist.init_process_group(backend="gloo")
    capi.comm_unique_id = lambda: bytes(128)
    ctx = _FakeCtx(capi, fail_attach=False)
    cdist.attach_comm(ctx)
    with open(os.path.join(out_dir, f"optout_{rank}.txt"), "w") as f:
        f.write(repr(ctx.calls))
    dist.barrier()
    dist.destroy_process_group()


def test_attach_comm_one_rank_opting_out_keeps_every_rank_on_nccl(tmp_path):
    """CB_NO_FUSED_EXCHANGE set on ONE rank only: no rank may map the peer tables (a rank on the fused path would wait
    for a row its peer sends through NCCL instead) - the setting is agreed on before anyone attaches."""
    world = 2
    mp.spawn(_optout_worker, args=(world, _free_port(), str(tmp_path), 1), nprocs=world, join=True)
    for rank in range(world):
        names = [c[0] for c in eval(open(tmp_path / f"optout_{rank}.txt").read())]
        assert names == ["init_comm"], (rank, names)
