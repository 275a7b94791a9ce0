import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from libpointmatcher_b200 import pm, synth, capi
rd, rf, T_gt = synth.scan_pair(1000000)
def once(normals, minim):
    icp = pm.ICP(0)
    icp.matcher = pm.KDTreeMatcher({"knn": "1"})
    icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.75"})])
    icp.errorMinimizer = minim()
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "40"})]
    icp.referenceDataPointsFilters = [pm.SurfaceNormalDataPointsFilter({"knn": "20"})] if normals else []
    t0 = time.perf_counter()
    T = icp(pm.DataPoints(rd), pm.DataPoints(rf))
    dt = time.perf_counter() - t0
    print("normals", normals, minim.__name__, "%.1f ms" % (dt * 1e3), "iterations", icp.iterationCount, flush=True)
    icp.ctx.close()
once(False, pm.PointToPointErrorMinimizer); once(False, pm.PointToPointErrorMinimizer)
once(True, pm.PointToPlaneErrorMinimizer); once(True, pm.PointToPlaneErrorMinimizer); once(True, pm.PointToPlaneErrorMinimizer)
f = pm.SurfaceNormalDataPointsFilter({"knn": "20"})
for i in range(3):
    t0 = time.perf_counter(); out = f.filter(pm.DataPoints(rf)); print("filter alone %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
