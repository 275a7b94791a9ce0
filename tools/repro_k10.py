import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libpointmatcher_b200 import capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rd, rf, _ = synth.scan_pair(n)
with capi.Context(0) as ctx:
    ctx.set_reference(rf)
    ctx.ref_compute_normals(knn=10)
    ctx.set_reading(rd)
    for mini in (capi.MIN_P2PLANE_COV, capi.MIN_P2POINT):
        p = capi.make_params(knn=10, max_dist=2.0, filters=[(0, 1.0), (1, 3.0)], minimizer=mini, max_iterations=4)
        print(ctx.icp_run(p)["iterations"])
    p = capi.make_params(knn=1, filters=[(2, 0.85)], minimizer=capi.MIN_P2PLANE, max_iterations=40, differential=(0.001, 0.001, 3))
    print(ctx.icp_run(p)["iterations"])
print("ok")
