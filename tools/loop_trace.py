#!/usr/bin/env python
"""One device-resident ICP run with CB_LOOP_TRACE=1 (per-iteration %globaltimer stamps and search counts on stderr)."""
import os
import sys

os.environ["CB_LOOP_TRACE"] = "1"
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cilantro_b200 import capi, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
metric = sys.argv[2] if len(sys.argv) > 2 else "p2p"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 15
ctx = capi.Context(0)
dst, src, nrm, T_ref = synth.icp_pair(n, seed=1, noise=0.001, with_normals=(metric == "combined"),
                                      T_ref=synth.t_ref_default() if os.environ.get("SURVEY_POSE") else None)
icp = capi.Icp(ctx, capi.Cloud(ctx, dst, nrm), capi.Cloud(ctx, src))
kw = dict(metric=metric, tol=0.0, max_d2=np.float32((0.02 if n <= 2_000_000 else 0.01) ** 2), max_iter=iters, timing=0)
if metric == "combined":
    kw.update(w_pt=0.1, w_pl=1.0)
icp.estimate(**kw)
print("---- second run ----", file=sys.stderr)
r = icp.estimate(**kw)
print("err", synth.frobenius(r["T"], T_ref), "corr", r["num_corr"])
