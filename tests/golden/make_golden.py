"""Generates tests/golden/oracle_small.npz: a small seeded input/output vector of the whole path
produced by the oracle, so that later changes to the oracle (or to the GPU path checked against
it) are caught.  The reference itself cannot run in this image (no Eigen/Boost/libnabo), so this
is oracle-generated, not reference-generated; the reference's own known answers are checked
separately in tests/test_oracle_golden.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from libpointmatcher_b200 import synth  # noqa: E402
from oracle import binding as orc  # noqa: E402

reading, reference, T_gt = synth.scan_pair(6400)
k, ratio, nk, iters = 3, 0.75, 10, 12
ids, dists = orc.KdTree(reference).knn(reading, k=k)
w, lim = orc.outlier_weights(dists, [(orc.FILTER_TRIMMEDDIST, ratio)])
nrm = orc.surface_normals(reference, knn=nk)
res = orc.icp(reading, reference, ref_normals=nrm["normals"], filters=[(orc.FILTER_TRIMMEDDIST, ratio)], minimizer=orc.MIN_P2PLANE,
              max_iterations=iters)
np.savez_compressed(os.path.join(os.path.dirname(__file__), "oracle_small.npz"), reading=reading, reference=reference, T_gt=T_gt, k=k,
                    ratio=ratio, ids=ids, dists=dists, weights=w, limit=lim[0], normals_knn=nk, normals=nrm["normals"], gap=nrm["gap"],
                    iterations=iters, T_icp=res["T"])
print("T_icp\n", res["T"], "\nT_gt\n", T_gt)
