#!/usr/bin/env python
"""A/B of the device-resident ICP loop against the host-driven loop on one GPU (experiment harness, not a test).

    python tools/loop_ab.py [n_points] [metric] [iters]
Prints per-iteration CUDA-event times of both loops (L2 flushed / not flushed), the difference of the final
transforms and of the correspondence counts.
"""
import sys
import os
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cilantro_b200 import capi, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    metric = sys.argv[2] if len(sys.argv) > 2 else "p2p"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 15
    survey_pose = os.environ.get("SURVEY_POSE") is not None
    ctx = capi.Context(0)
    dst, src, nrm, T_ref = synth.icp_pair(n, seed=1, noise=0.001, with_normals=(metric == "combined"),
                                          T_ref=synth.t_ref_default() if survey_pose else None)
    d_dst, d_src = capi.Cloud(ctx, dst, nrm), capi.Cloud(ctx, src)
    icp = capi.Icp(ctx, d_dst, d_src)
    max_d2 = np.float32((0.02 if n <= 2_000_000 else 0.01) ** 2)
    kw = dict(metric=metric, tol=0.0, max_d2=max_d2, max_iter=iters)
    if metric == "combined":
        kw.update(w_pt=0.1, w_pl=1.0)
    out = {}
    modes = (("device", False),) if os.environ.get("AB_DEVICE_ONLY") else (("host", True), ("device", False))
    for name, host in modes:
        for flush in (True, False):
            icp.estimate(host_loop=host, flush_l2=flush, **kw)  # warm-up
            ctx.synchronize()
            t0 = time.perf_counter()
            r = icp.estimate(host_loop=host, flush_l2=flush, timing=0, **kw)
            ctx.synchronize()
            wall = time.perf_counter() - t0
            r1 = icp.estimate(host_loop=host, flush_l2=flush, timing=1, **kw)
            ms = r1["iter_ms"]
            print(f"{name:6s} flush={int(flush)}: wall {1e3 * wall:.3f} ms for {r['iterations']} it (timing=0); events: total "
                  f"{ms.sum():.3f} ms, per-iter {np.array2string(ms, precision=3, max_line_width=200)}; launches {r['kernel_launches']}")
            out[(name, flush)] = r
    if os.environ.get("AB_DEVICE_ONLY"):
        print("|T_device - T_ref|_F =", synth.frobenius(out[("device", False)]["T"], T_ref), "num_corr", out[("device", False)]["num_corr"])
        ctx.close()
        return
    a, b = out[("host", False)], out[("device", False)]
    print("num_corr host/device:", a["num_corr"], b["num_corr"])
    print("|T_host - T_device|_F =", float(np.linalg.norm(a["T"].astype(np.float64) - b["T"])))
    print("|T_device - T_ref|_F =", synth.frobenius(b["T"], T_ref))
    # per-iteration agreement: run k iterations with both loops for k = 1..iters
    worst = 0.0
    for k in range(1, iters + 1):
        kk = dict(kw, max_iter=k)
        ra = icp.estimate(host_loop=True, **kk)
        rb = icp.estimate(host_loop=False, **kk)
        worst = max(worst, float(np.linalg.norm(ra["T"].astype(np.float64) - rb["T"])))
        if ra["num_corr"] != rb["num_corr"]:
            print(f"  k={k}: num_corr differs {ra['num_corr']} vs {rb['num_corr']}")
    print("worst |T_host - T_device|_F over k = 1..%d: %.3e" % (iters, worst))
    ctx.close()


if __name__ == "__main__":
    main()
