"""globaltimer breakdown of select_accumulate_kernel (build: PMGPU_VARIANT=prof PMGPU_DEFINES=-DPM_PROFILE_NS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libpointmatcher_b200 import capi, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
rd, rf, T_gt = synth.scan_pair(n)
with capi.Context(0) as ctx:
    ctx.set_reference(rf); ctx.set_reading(rd); ctx.ref_compute_normals(knn=10)
    p = capi.make_params(knn=1, filters=[(capi.FILTER_TRIMMEDDIST, 0.75)], minimizer=capi.MIN_P2PLANE, max_iterations=16)
    ctx.icp_run(p)
