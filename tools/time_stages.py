"""Stage timings of the wider configs: normals (K8) knn 20, kNN k = 10, covariance, e2e pieces."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from libpointmatcher_b200 import capi, synth, pm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rd, rf, T_gt = synth.scan_pair(n)
def wall(f, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); f(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return best * 1e3
with capi.Context(0) as ctx:
    print("set_reference_centered              %.2f ms" % wall(lambda: ctx.set_reference_centered(rf)))
    print("set_reference (upload + build)      %.2f ms" % wall(lambda: ctx.set_reference(rf)))
    print("set_reading (upload + morton)       %.2f ms" % wall(lambda: ctx.set_reading(rd)))
    print("ref_compute_normals knn=20          %.2f ms" % wall(lambda: ctx.ref_compute_normals(knn=20)))
    print("ref_compute_normals knn=7           %.2f ms" % wall(lambda: ctx.ref_compute_normals(knn=7)))
    ctx.timing_enable(True)
    for k in (1, 5, 10, 20):
        ctx.knn(T_gt.astype(np.float32), k, download=False); ctx.timing_collect()
        ctx.knn(T_gt.astype(np.float32), k, download=False)
        print("knn k=%-2d aligned                    %.3f ms" % (k, ctx.timing_collect()["knn"][0]))
    p = capi.make_params(knn=10, max_dist=2.0, filters=[(0, 1.0), (1, 3.0)], minimizer=capi.MIN_P2PLANE_COV, max_iterations=10)
    ctx.icp_run(p); ctx.timing_collect()
    t = wall(lambda: ctx.icp_run(p), 1)
    st = ctx.timing_collect()
    print("config-4 shape (knn 10, maxDist+Median, P2PlaneWithCov) 10 iterations: %.2f ms total;" % t, {k: round(v[0] / 10, 3) for k, v in st.items()})
    out = ctx.normals(rf, knn=20, keep=("normals", "densities"))
    print("pmgpu_normals (private ctx, upload+build+knn20+eig+download) %.2f ms" % wall(lambda: ctx.normals(rf, knn=20, keep=("normals",)), 2))
print("Context create+destroy                %.2f ms" % wall(lambda: capi.Context(0).close()))
def e2e():
    icp = pm.ICP(0); icp.matcher = pm.KDTreeMatcher(); icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.75"})])
    icp.errorMinimizer = pm.PointToPointErrorMinimizer(); icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "40"})]
    icp(pm.DataPoints(rd), pm.DataPoints(rf)); icp.ctx.close()
print("pm.ICP 40 iterations e2e (pageable)   %.2f ms" % wall(e2e))
