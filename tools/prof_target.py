"""Short profiling target (ncu replays every kernel ~40 times): one registration-shaped sequence per workload.

    python tools/prof_target.py c2plane|c4|normals10m [iterations]
"""
import sys
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libpointmatcher_b200 import capi, synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "c2plane"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = capi.Context(0)
if what == "normals10m":
    rf = synth.world_map(10_000_000, 10)
    ctx.set_reference(rf)
    ctx.ref_compute_normals(knn=20)
    ctx.sync()
    print("normals of a 10 M-point map done")
else:
    n = 1_000_000 if what == "c2plane" else 2_000_000
    rd, rf, _ = synth.scan_pair(n)
    mean = rf[:, :3].mean(axis=0).astype(np.float32)
    rf = rf.copy(); rf[:, :3] -= mean
    rd = rd.copy(); rd[:, :3] -= mean
    ctx.set_reference(rf)
    ctx.ref_compute_normals(knn=20)
    ctx.set_reading(rd)
    if what == "c2plane":
        p = capi.make_params(knn=1, filters=[(capi.FILTER_TRIMMEDDIST, 0.75)], minimizer=capi.MIN_P2PLANE, max_iterations=iters)
    else:
        p = capi.make_params(knn=10, max_dist=2.0, filters=[(capi.FILTER_MAXDIST, 1.0), (capi.FILTER_MEDIANDIST, 3.0)], minimizer=capi.MIN_P2PLANE_COV,
                             max_iterations=iters)
    res = ctx.icp_run(p)
    print(what, "iterations", res["iterations"])
ctx.close()
