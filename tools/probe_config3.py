"""BASELINE config 3 shape: a 10 M-point map kept resident (ICPSequence), 1 M-point readings
registered against it with point-to-plane, normals knn 20 computed on the map."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from libpointmatcher_b200 import pm, synth

n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
t0 = time.perf_counter(); world = synth.world_map(n_map); print("synth map %.1f s" % (time.perf_counter() - t0), world.shape, flush=True)
pose = synth.pose_matrix((3.3, 0.2, 0.0), 2.0)
rd = synth.scan(1_000_000, pose, seed=777)      # sensor-frame scan taken inside the mapped corridor
seq = pm.ICPSequence(0)
seq.matcher = pm.KDTreeMatcher({"knn": "1"})
seq.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.75"})])
seq.errorMinimizer = pm.PointToPlaneErrorMinimizer()
seq.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "40"}),
                              pm.DifferentialTransformationChecker({"minDiffRotErr": "0.001", "minDiffTransErr": "0.01", "smoothLength": "3"})]
seq.referenceDataPointsFilters = [pm.SurfaceNormalDataPointsFilter({"knn": "20"})]
torch.cuda.synchronize(); t0 = time.perf_counter()
seq.setMap(pm.DataPoints(world))
torch.cuda.synchronize(); print("setMap (normals knn 20 on %d points + upload + build): %.1f ms" % (n_map, (time.perf_counter() - t0) * 1e3), flush=True)
T_guess = synth.pose_matrix((3.0, 0.0, 0.0), 0.0).astype(np.float32)
for rep in range(4):
    if rep == 3:  # no early stop: how far does it get?
        seq.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "80"})]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    T = seq(pm.DataPoints(rd), T_guess)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    err = np.linalg.norm(T[:3, 3] - pose[:3, 3])
    print("register 1 M reading vs %d-point map: %.1f ms, %d iterations, translation error vs truth %.4f m" % (n_map, dt * 1e3, seq.iterationCount, err), flush=True)
print("device memory in use: %.2f GB" % ((torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9))
