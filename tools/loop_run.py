#!/usr/bin/env python
"""Two device-resident ICP runs on one GPU, nothing else (target of ncu captures)."""
import os
import sys

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
kw = dict(metric=metric, tol=0.0, max_d2=np.float32((0.02 if n <= 2_000_000 else 0.01) ** 2), max_iter=iters, timing=0,
          flush_l2=bool(os.environ.get("FLUSH")))
if metric == "combined":
    kw.update(w_pt=0.1, w_pl=1.0)
for _ in range(2):
    r = icp.estimate(**kw)
print("err", synth.frobenius(r["T"], T_ref), "corr", r["num_corr"])
