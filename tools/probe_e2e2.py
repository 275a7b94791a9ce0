"""Where a whole registration's time goes (c2plane shape, pinned host buffers): every phase of pm.ICP bracketed by a device sync."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from libpointmatcher_b200 import capi, pm, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rd, rf, _ = synth.scan_pair(n)


def pin(a):
    t = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
    t.numpy()[...] = a
    return t.numpy(), t


rd_p, _a = pin(rd)
rf_p, _b = pin(rf)
ctx = capi.Context(0)
p = capi.make_params(knn=1, filters=[(capi.FILTER_TRIMMEDDIST, 0.75)], minimizer=capi.MIN_P2PLANE, max_iterations=20)
for rep in range(3):
    t = [time.perf_counter()]
    def mark():
        ctx.sync(); t.append(time.perf_counter())
    ctx.set_reference(rf_p); mark()
    ctx.ref_compute_normals(knn=20); mark()
    ctx.ref_center(rf_p); mark()
    ctx.set_reading(rd_p); mark()
    T = np.eye(4, dtype=np.float32); ctx.reading_apply_transform(T); mark()
    ctx.icp_run(p); mark()
    names = ["set_reference (H2D + build)", "normals knn 20", "centre (host mean + shift)", "set_reading (H2D + Morton)", "apply T", "icp_run 20 it"]
    print("rep", rep, "  ".join("%s %.2f ms" % (nm, 1e3 * (b - a)) for nm, a, b in zip(names, t[:-1], t[1:])), " total %.2f ms" % (1e3 * (t[-1] - t[0])))
ctx.close()
