"""kNN stage 1 / stage 2 split per ICP iteration for several leaf budgets (PMGPU_TIME_STAGE2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
os.environ["PMGPU_TIME_STAGE2"] = "1"
from libpointmatcher_b200 import capi, synth, pm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rd, rf, T_gt = synth.scan_pair(n)
mean = pm.sequential_mean(rf)
rf_c = rf.copy(); rf_c[:, :3] -= mean[:3]
T_in = np.eye(4, dtype=np.float32); T_in[:3, 3] = -mean[:3]
for budget in sys.argv[2:] or ["8", "16", "32", "64", "1000000"]:
    os.environ["PMGPU_KNN_BUDGET"] = budget
    with capi.Context(0) as ctx:
        ctx.set_reference(rf_c); ctx.set_reading(rd); ctx.reading_apply_transform(T_in)
        p = capi.make_params(knn=1, filters=[(capi.FILTER_TRIMMEDDIST, 0.75)], minimizer=capi.MIN_P2POINT, max_iterations=40)
        ctx.icp_reset(None); ctx.icp_enqueue(p, 3); ctx.sync(); ctx.icp_reset(None)
        ctx.timing_enable(True); ctx.timing_collect()
        rows = []
        for it in range(40):
            ctx.icp_enqueue(p, 1); st = ctx.timing_collect()
            rows.append((st["knn"][0], st["covariance"][0]))
        a = np.array(rows)
        print("budget %8s: stage1 %.3f ms  stage2 %.3f ms (mean of 40) | it0 %.3f/%.3f it1 %.3f/%.3f it5 %.3f/%.3f it20 %.3f/%.3f it39 %.3f/%.3f" % (
            budget, a[:, 0].mean(), a[:, 1].mean(), *a[0], *a[1], *a[5], *a[20], *a[39]))
