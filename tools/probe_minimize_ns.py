import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
from libpointmatcher_b200 import capi, synth, pm
rd, rf, T_gt = synth.scan_pair(1000000)
with capi.Context(0) as ctx:
    ctx.set_reference(rf); ctx.set_reading(rd); ctx.ref_compute_normals(knn=10)
    for mn in (capi.MIN_P2POINT, capi.MIN_P2PLANE):
        p = capi.make_params(knn=1, filters=[(capi.FILTER_TRIMMEDDIST, 0.75)], minimizer=mn, max_iterations=6)
        ctx.icp_run(p)
