"""Multi-GPU plumbing: one process per GPU (torchrun), source points sharded in contiguous blocks,
destination cloud replicated, ONE small all-reduce per iteration inside the library (NCCL, its own
communicator). torch.distributed is used only for the rendezvous: broadcasting the 128-byte NCCL
unique id and barriers / max-over-ranks timing in bench.py.
"""
import os

import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous block of rank `rank` when n items are split over `world` ranks (sizes differ by <= 1)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend=None):
    """Join the torchrun rendezvous (MASTER_ADDR/PORT from the environment; 127.0.0.1 by contract)."""
    import torch
    import torch.distributed as dist

    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def broadcast_bytes(payload, src=0):
    """Broadcast a bytes object from rank `src` over the default process group (any backend)."""
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return payload
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def attach_comm(ctx):
    """Create the library's NCCL communicator on `ctx` across the torch.distributed world."""
    import torch.distributed as dist

    from . import capi

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0, 1
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = capi.comm_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 0)
    ctx.init_comm(uid, rank, world)
    # fused exchange over NVLink peer memory: all-gather the CUDA-IPC handles, map the peers' tables.
    # All ranks must take the same path, so the outcome is agreed on before anyone proceeds.
    # every rank must take the same path: if ANY rank was started with CB_NO_FUSED_EXCHANGE, nobody maps the tables and
    # all ranks reduce through NCCL (a rank waiting in the fused exchange for a peer that went to NCCL would only time out)
    want_fused = min_over_ranks(0 if os.environ.get("CB_NO_FUSED_EXCHANGE") is not None else 1)
    if want_fused:
        handles = [None] * world
        dist.all_gather_object(handles, ctx.ipc_handle())
        ok = 1
        try:
            if os.environ.get("CB_TEST_IPC_FAIL_RANK") == str(rank):  # test hook: pretend this rank cannot map its peers
                raise capi.CbError("simulated cudaIpcOpenMemHandle failure")
            ctx.ipc_attach(b"".join(handles))
        except capi.CbError:
            ok = 0
        if min_over_ranks(ok) == 0:
            # no peer mapping somewhere (e.g. no P2P between two of the GPUs): every rank goes back to the NCCL
            # all-reduce path, together
            ctx.ipc_detach()
            if rank == 0:
                import sys

                print("cilantro_b200: fused NVLink exchange unavailable on some rank, using ncclAllReduce",
                      file=sys.stderr)
    dist.barrier()
    return rank, world


def min_over_ranks(x):
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([float(x)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return t.item()


def allreduce_sum_f64(values):
    """Sum a small float64 vector over the default process group (host-side logic / CPU tests)."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(values, np.float64).copy())
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t = t.cpu()
    return t.numpy()


def max_over_ranks(x):
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
