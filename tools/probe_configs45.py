"""BASELINE configs[3] and [4] shapes, timed through the Python mirror (pm.ICP) on one GPU:
  C4: 2 M-point scans, KDTreeMatcher knn 10 maxDist 2.0, MaxDist 1.0 x MedianDist 3, PointToPlaneWithCov, normals knn 20
  C5: independent 200 k x 200 k scan pairs, one after the other on this GPU (the 8-GPU job runs 1024 of them round-robin)
Prints per-registration and per-iteration times; clouds come from pinned host memory, results go back to the host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from libpointmatcher_b200 import pm, synth


def chain_c4():
    icp = pm.ICP(0)
    icp.referenceDataPointsFilters = [pm.SurfaceNormalDataPointsFilter({"knn": "20"})]
    icp.matcher = pm.KDTreeMatcher({"knn": "10", "maxDist": "2.0"})
    icp.outlierFilters = pm.OutlierFilters([pm.MaxDistOutlierFilter({"maxDist": "1.0"}), pm.MedianDistOutlierFilter({"factor": "3"})])
    icp.errorMinimizer = pm.PointToPlaneWithCovErrorMinimizer({"sensorStdDev": "0.01"})
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "40"})]
    return icp


def chain_c5():
    icp = pm.ICP(0)
    icp.referenceDataPointsFilters = [pm.SurfaceNormalDataPointsFilter({"knn": "10"})]
    icp.matcher = pm.KDTreeMatcher({"knn": "1"})
    icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.75"})])
    icp.errorMinimizer = pm.PointToPlaneErrorMinimizer()
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "40"}),
                                  pm.DifferentialTransformationChecker({"minDiffRotErr": "0.001", "minDiffTransErr": "0.01", "smoothLength": "3"})]
    return icp


def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0


n4 = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
rd, rf, T_gt = synth.scan_pair(n4)
icp = chain_c4()
for rep in range(3):
    T, dt = timed(lambda: icp(pm.DataPoints(rd), pm.DataPoints(rf)))
    cov = icp.errorMinimizer.getCovariance()
    print("C4 %d x %d, knn 10, MaxDist x MedianDist, PointToPlaneWithCov: registration %.1f ms (%d iterations, %.3f ms / iteration incl. normals knn 20 + build), "
          "translation error %.4f m, cov trace %.3e" % (n4, n4, dt * 1e3, icp.iterationCount, dt * 1e3 / icp.iterationCount, np.linalg.norm(T[:3, 3] - T_gt[:3, 3]), np.trace(cov)),
          flush=True)
icp.ctx.close()

pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
clouds = [synth.scan_pair(200_000, pair_seed=j + 1) for j in range(pairs)]
icp = chain_c5()
icp(pm.DataPoints(clouds[0][0]), pm.DataPoints(clouds[0][1]))   # warm
its, errs = 0, []
_, dt = timed(lambda: None)
torch.cuda.synchronize(); t0 = time.perf_counter()
for rd5, rf5, T5 in clouds:
    T = icp(pm.DataPoints(rd5), pm.DataPoints(rf5))
    its += icp.iterationCount
    errs.append(np.linalg.norm(T[:3, 3] - T5[:3, 3]))
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("C5 %d independent 200 k x 200 k pairs on one GPU: %.2f ms / pair (%.1f pairs/s, %d iterations in all, %.0f it/s), worst translation error %.4f m"
      % (pairs, dt * 1e3 / pairs, pairs / dt, its, its / dt, max(errs)), flush=True)
icp.ctx.close()
