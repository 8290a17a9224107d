#!/usr/bin/env python
"""Benchmark of the rigid-ICP hot path (BASELINE.json metric: ICP iterations/s and correspondences/s).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

A "step" is one full ICP iteration (transform + radius-bounded 1-NN over all source points +
moment accumulation + reduction + host solve) on the named workload. At N = 1 the workload is
BASELINE.json configs[1]: 1 M -> 1 M synthetic uniform clouds, point-to-point metric, k = 1.
For N > 1 (one process per GPU under torchrun) the job is weak-scaled: every rank owns 1 M source
points, the destination cloud (N M points) is replicated, and the only exchange per iteration is the
all-reduce of the 16 Kabsch moments, fused into the accumulation kernel's epilogue (rows written straight
into the peers' tables over NVLink peer memory; CB_NO_FUSED_EXCHANGE=1 selects the ncclAllReduce path).
`value` counts correspondences (source points processed) per second over ALL ranks; iterations/s is
reported next to it.

Timing: W untimed warm-up iterations, then K timed iterations bracketed by barrier + synchronize;
each iteration is timed on the device with CUDA events inside the library (cb_icp_estimate), with an
L2 flush (256 MiB memset) before every iteration OUTSIDE the event bracket; max over ranks.
The reference arm (--impl reference) times cilantro's own CPU path: the reference's vendored
nanoflann compiled in place (oracle/_ref) driving the Eigen-free restatement of its ICP loop
(oracle/), on all host cores. Nothing here reads /root/reference at run time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (points per GPU, metric, with normals, max_d2, icp kwargs)
    "icp_p2p_1m": dict(n=1_000_000, metric="p2p", normals=False, max_d2=0.02 ** 2, kw={}),
    "icp_combined_10m": dict(n=10_000_000, metric="combined", normals=True, max_d2=0.01 ** 2,
                             kw=dict(w_pt=0.1, w_pl=1.0)),
    "icp_p2p_100k": dict(n=100_000, metric="p2p", normals=False, max_d2=0.05 ** 2, kw={}),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(workload)
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_inputs(w, rank, world, pinned=False):
    """Seeded synthetic pair: dst = world * n points (replicated), this rank's src shard = n points."""
    from cilantro_b200 import synth
    from cilantro_b200.dist import shard_bounds

    n_total = w["n"] * world
    dst, src, nrm, T_ref = synth.icp_pair(n_total, seed=1, noise=0.001, with_normals=w["normals"])
    lo, hi = shard_bounds(n_total, rank, world)
    src = np.ascontiguousarray(src[lo:hi])
    if pinned:
        import torch

        def pin(a):
            if a is None:
                return None
            t = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
            t.numpy()[...] = a
            return t

        keep = [pin(dst), pin(src), pin(nrm)]
        return keep[0].numpy(), keep[1].numpy(), (keep[2].numpy() if keep[2] is not None else None), T_ref, lo, keep
    return dst, src, nrm, T_ref, lo, None


def cpu_reference_run(w, steps, warmup, dst, src, nrm, build_in_timed_region):
    """cilantro's CPU path on the host cores: reference nanoflann (oracle/_ref) + restated ICP loop."""
    import oracle

    oracle.build()
    kind = "reference" if oracle.have_ref() else "port"
    mk = (lambda: oracle.RefKnn(dst)) if oracle.have_ref() else (lambda: oracle.BruteKnn(dst))
    kw = dict(metric=w["metric"], dst_n=nrm, tol=0.0, max_d2=np.float32(w["max_d2"]), parallel=True, **w["kw"])
    t0 = time.perf_counter()
    knn = mk()
    t_build = time.perf_counter() - t0
    # "all the host threads it can use": the kd-tree sweep is latency-bound and was measured 3x slower
    # with every hyper-thread (128) than with one thread per core (64) on the B200 host, so probe both
    # on one iteration and give the reference the better setting
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    best = None
    for nt in sorted({ncpu, max(1, ncpu // 2)}):
        oracle.set_num_threads(nt)
        t0 = time.perf_counter()
        oracle.icp(dst, src, knn, max_iter=1, **kw)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, nt)
    oracle.set_num_threads(best[1])
    cores = best[1]
    if warmup > 0:
        oracle.icp(dst, src, knn, max_iter=warmup, **kw)
    t0 = time.perf_counter()
    if build_in_timed_region:
        knn = mk()  # SimpleICP::estimate() builds the kd-tree lazily inside the first iteration
    r = oracle.icp(dst, src, knn, max_iter=steps, **kw)
    dt = time.perf_counter() - t0
    return dict(kind=kind, cores=cores, seconds=dt, build_s=t_build, iters=r["iterations"], t_knn_s=r["t_knn_s"],
                t_est_s=r["t_est_s"], T=r["T"])


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = args.gpus
    dst, src, nrm, T_ref, lo, _ = make_inputs(w, 0, world)
    # bounded sample: the per-rank shard of the same workload (1 M queries into the full dst cloud)
    r = cpu_reference_run(w, args.steps, args.warmup, dst, src, nrm, build_in_timed_region=True)
    its = r["iters"] / r["seconds"]
    value = its * src.shape[0]
    line = {
        "impl": "reference",
        "metric": "icp_correspondences_per_sec", "value": value, "unit": "correspondences/s",
        "iterations_per_sec": its,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["seconds"] / max(r["iters"], 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, w, world, src.shape[0], dst.shape[0]),
        "cpu_baseline": {
            "value": value, "unit": "correspondences/s", "cores": r["cores"], "kind": r["kind"],
            "sample": (f"{r['iters']} ICP iterations of one rank's shard ({src.shape[0]} queries into {dst.shape[0]} "
                       f"reference points), kd-tree build ({r['build_s']:.2f} s) inside the timed region; "
                       f"kNN {r['t_knn_s']:.2f} s + estimate {r['t_est_s']:.2f} s; kNN = cilantro's vendored nanoflann "
                       "compiled in place, ICP loop = Eigen-free restatement (Eigen3 absent from the image)"),
        },
        "e2e": {"value": value, "unit": "correspondences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, w, world, n_src_rank, n_dst):
    return {
        "workload": (f"{args.workload}: rigid ICP, {w['metric']} metric, k=1, {n_src_rank} source points per GPU x {world} "
                     f"GPU(s) against {n_dst} destination points (uniform random in the unit cube, seed 1, "
                     f"noise +-0.001), max_distance^2={w['max_d2']:g}, fixed iteration count (tol=0)"),
        "parallelism": (f"src sharded x{world}, dst replicated, one 16-value (p2p) / 28-value all-reduce per iteration, "
                        + ("ncclAllReduce" if os.environ.get("CB_NO_FUSED_EXCHANGE") else
                           "fused into the kernel epilogue over NVLink peer memory")),
        "l2": ("NOT flushed (--no-flush experiment; inputs smaller than L2 stay resident)" if getattr(args, "no_flush", False)
               else "flushed before every timed iteration (256 MiB memset outside the CUDA-event bracket)"),
    }


def run_ours(args, w):
    import torch

    from cilantro_b200 import capi, dist as cdist, synth

    rank, world, local = cdist.init_process_group()
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: cilantro_b200 has no CPU fallback")
    capi.lib()
    ctx = capi.Context(local)
    cdist.attach_comm(ctx)
    dst, src, nrm, T_ref, lo, keep = make_inputs(w, rank, world, pinned=True)
    n_src, n_dst = src.shape[0], dst.shape[0]
    max_d2 = np.float32(w["max_d2"])
    kw = dict(metric=w["metric"], tol=0.0, max_d2=max_d2, **w["kw"])

    # ---- value: inputs already resident in HBM, index built ------------------------------------------
    d_dst = capi.Cloud(ctx, dst, nrm)
    d_src = capi.Cloud(ctx, src, None, index_offset=lo)
    gi = d_dst.grid_info()
    d_src.grid_info()
    icp = capi.Icp(ctx, d_dst, d_src)
    if args.warmup > 0:
        icp.estimate(max_iter=args.warmup, flush_l2=not args.no_flush, **kw)
    barrier(world)
    ctx.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.kernel_launches()
    t0 = time.perf_counter()
    res = icp.estimate(max_iter=args.steps, flush_l2=not args.no_flush, **kw)
    ctx.synchronize()
    barrier(world)
    wall = time.perf_counter() - t0
    launches = ctx.kernel_launches() - launches0
    # keep the GPUs busy a little longer for the clock sampler on very short runs; EVERY rank runs
    # it (the iteration contains a collective), the decision is taken on the max-over-ranks wall time
    clocks = None
    # the same K iterations again with the events around the search kernel only (timing=2): the
    # kernel's average duration for the roofline (one bracket per iteration at a time, see cb_icp_params)
    resk = icp.estimate(max_iter=args.steps, flush_l2=not args.no_flush, timing=2, **kw)
    # nvidia-smi needs ~0.1 s before its first sample and the legs above take milliseconds: keep the same
    # kernel running until rank 0 holds at least 3 samples taken under this load (or 3 s have passed)
    t_load = time.perf_counter()
    while True:
        need = 1.0 if (rank == 0 and len(sampler.lines) < 3 and time.perf_counter() - t_load < 3.0) else 0.0
        if cdist.max_over_ranks(need) == 0.0:
            break
        icp.estimate(max_iter=max(args.steps, 200), flush_l2=False, timing=0, **kw)
    if rank == 0:
        clocks = sampler.stop()
    assert res["iterations"] == args.steps
    ms_total = cdist.max_over_ranks(res["gpu_ms_total"])
    ms_kernel = cdist.max_over_ranks(resk["gpu_ms_search"])
    ms_per_step = ms_total / args.steps
    its = 1e3 / ms_per_step
    value = its * n_src * world
    err = synth.frobenius(res["T"], T_ref)

    # ---- e2e: host buffers -> upload -> index build -> full estimate() -> transform back -----------
    e2e_iters = 15
    e2e_runs = 3
    barrier(world)
    e2e_t = []
    for _ in range(e2e_runs):
        ctx.synchronize()
        barrier(world)
        t0 = time.perf_counter()
        # what the ICP constructor of the shims does: both clouds in one call (the second upload overlaps the
        # first grid build), then the ICP object (means)
        c_dst, c_src = capi.cloud_pair(ctx, dst, nrm, src, None, offset_b=lo)
        t2 = time.perf_counter()
        c_icp = capi.Icp(ctx, c_dst, c_src)
        t3 = time.perf_counter()
        r2 = c_icp.estimate(max_iter=e2e_iters, timing=0, **kw)  # production settings: no event instrumentation
        T_host = np.array(r2["T"])  # result read back on the host
        ctx.synchronize()
        t4 = time.perf_counter()
        barrier(world)
        e2e_t.append(time.perf_counter() - t0)
        if rank == 0:
            print(f"[e2e] uploads + grid builds {1e3 * (t2 - t0):.2f} ms, icp_create {1e3 * (t3 - t2):.2f} ms, "
                  f"estimate({e2e_iters}) {1e3 * (t4 - t3):.2f} ms", file=sys.stderr)
        c_icp.close(); c_src.close(); c_dst.close()
    e2e_s = cdist.max_over_ranks(min(e2e_t))
    e2e_its = e2e_iters / e2e_s
    h2d = (dst.nbytes + src.nbytes + (nrm.nbytes if nrm is not None else 0)) / e2e_iters
    d2h = (16 if w["metric"] == "p2p" else 28) * 8 + 48 / e2e_iters

    if rank != 0:
        return 0
    # ---- roofline of the dominant kernel (fused transform + grid 1-NN + moment accumulation) -------
    peak, peak_src = load_peaks()
    # algorithmic bytes per launch: 16 B query read per source point, every cell-sorted
    # reference point read once (16 B), + one 16 B normal gather per correspondence for the plane term
    # (per-query results are no longer written in the iteration loop)
    algo_bytes = 16 * n_src + 16 * n_dst + (16 * n_src if w["metric"] == "combined" else 0)
    kernel_ms = ms_kernel / args.steps
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": load_traffic(args.workload), "kernel": "icp_pass_kernel<%s,search>" % w["metric"],
                "kernel_ms": kernel_ms, "algorithmic_bytes": algo_bytes, "peak_source": peak_src}

    # ---- CPU baseline on rank 0's host cores, bounded sample ----------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        cb_steps = 2 if w["n"] >= 1_000_000 else 3
        r = cpu_reference_run(w, cb_steps, 0, dst, src, nrm, build_in_timed_region=False)
        cits = r["iters"] / r["seconds"]
        cpu = {"value": cits * n_src, "unit": "correspondences/s", "iterations_per_sec": cits, "cores": r["cores"],
               "kind": r["kind"],
               "sample": (f"{r['iters']} ICP iterations of rank 0's shard ({n_src} queries into {n_dst} reference points), "
                          f"kd-tree prebuilt (build {r['build_s']:.2f} s, 1 thread, not counted); kNN {r['t_knn_s']:.2f} s "
                          f"+ estimate {r['t_est_s']:.2f} s")}
    line = {
        "metric": "icp_correspondences_per_sec", "value": value, "unit": "correspondences/s",
        "iterations_per_sec": its,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, w, world, n_src, n_dst),
        "grid": gi,
        "transform_error_vs_generating_pose": err,
        "wall_ms_per_step_incl_flush": 1e3 * wall / args.steps,
        "clocks": clocks,
        "e2e": {"value": e2e_its * n_src * world, "unit": "correspondences/s", "iterations_per_sec": e2e_its,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "what": (f"cb_cloud_create_pair from pinned host buffers (2 uploads + 2 grid builds) + cb_icp_create + "
                         f"cb_icp_estimate({e2e_iters} iterations) + result on host; best of {e2e_runs}; "
                         f"{e2e_s * 1e3:.2f} ms per call")},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    return 0


def barrier(world):
    if world > 1:
        import torch.distributed as dist

        dist.barrier()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=15)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    from bench_aux import AUX

    ap.add_argument("--workload", default="icp_p2p_1m", choices=sorted(WORKLOADS) + sorted(AUX))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true",
                    help="experiments only: skip the L2 flush between timed iterations (the reported config says so)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.workload in AUX:  # secondary single-GPU workloads (k-means, RANSAC, PCA): bench_aux.py
        if int(os.environ.get("RANK", "0")) == 0:
            AUX[args.workload](args)
        return 0
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference(args, w)
    try:
        return run_ours(args, w)
    finally:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
