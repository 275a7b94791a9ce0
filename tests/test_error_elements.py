"""ErrorElements materialised on request, and getResidualError / getOverlap derived from it, against the
plain-loop restatement in oracle/error_elements.py.  The CPU tests feed host arrays; the GPU test goes through pm.ICP."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import error_elements as orc  # noqa: E402


@pytest.fixture(scope="module")
def pm():
    from libpointmatcher_b200 import pm
    return pm


def _case(seed, n=400, m=300, knn=3, noise=True, dens=True):
    rng = np.random.default_rng(seed)
    rd = np.c_[rng.normal(0, 3, (n, 3)), np.ones(n)].astype(np.float32)
    rf = np.c_[rng.normal(0, 3, (m, 3)), np.ones(m)].astype(np.float32)
    ids = rng.integers(0, m, (n, knn)).astype(np.int32)
    dists = rng.uniform(0, 4, (n, knn)).astype(np.float32)
    dists[rng.uniform(size=(n, knn)) < 0.1] = np.inf
    w = (rng.uniform(size=(n, knn)) < 0.7).astype(np.float32) * rng.uniform(0.2, 1, (n, knn)).astype(np.float32)
    w[dists == np.inf] = 0
    w[7] = 0                                                    # a point with no kept match
    rdd = {"simpleSensorNoise": rng.uniform(0.5, 3, (n, 1)).astype(np.float32)} if noise else {}
    rfd = {"normals": rng.normal(0, 1, (m, 3)).astype(np.float32)}
    if noise:
        rfd["simpleSensorNoise"] = rng.uniform(0.5, 3, (m, 1)).astype(np.float32)
    if dens:
        rfd["densities"] = rng.uniform(0.1, 5, (m, 1)).astype(np.float32)
    return rd, rdd, rf, rfd, w, ids, dists


@pytest.mark.parametrize("seed,noise,dens", [(0, True, True), (1, True, False), (2, False, False)])
def test_error_elements_and_derived_numbers_match_oracle(pm, seed, noise, dens):
    rd, rdd, rf, rfd, w, ids, dists = _case(seed, noise=noise, dens=dens)
    ee = pm.ErrorElements(pm.DataPoints(rd, rdd), pm.DataPoints(rf, rfd), w, pm.Matches(dists, ids))
    eo = orc.error_elements(rd, rdd, rf, rfd, w, ids, dists)
    assert (ee.reading.features == eo["reading"]).all() and (ee.reference.features == eo["reference"]).all()
    assert (ee.weights == eo["weights"]).all() and (ee.matches.ids[:, 0] == eo["ids"]).all() and (ee.matches.dists[:, 0] == eo["dists"]).all()
    assert (ee.reference.descriptors["normals"] == eo["reference_desc"]["normals"]).all()
    assert (ee.nbRejectedMatches, ee.nbRejectedPoints) == (eo["nbRejectedMatches"], eo["nbRejectedPoints"]) and ee.nbRejectedPoints >= 1
    assert ee.pointUsedRatio == eo["pointUsedRatio"] and abs(ee.weightedPointUsedRatio - eo["weightedPointUsedRatio"]) < 1e-6
    assert abs(pm.point_to_point_residual(ee) - orc.point_to_point_residual(eo)) < 1e-6 * orc.point_to_point_residual(eo)
    for f2d in (False, True):
        assert abs(pm.point_to_plane_residual(ee, f2d) - orc.point_to_plane_residual(eo, f2d)) < 1e-6 * orc.point_to_plane_residual(eo, f2d)
    assert pm.point_to_point_overlap(ee) == orc.point_to_point_overlap(eo) and (pm.point_to_point_overlap(ee) is None) == (not noise)
    assert pm.point_to_plane_overlap(ee) == orc.point_to_plane_overlap(eo) and (pm.point_to_plane_overlap(ee) is None) == (not noise)


def test_error_elements_without_a_kept_match_throws(pm):
    rd, rdd, rf, rfd, w, ids, dists = _case(3)
    with pytest.raises(pm.ConvergenceError):
        pm.ErrorElements(pm.DataPoints(rd, rdd), pm.DataPoints(rf, rfd), np.zeros_like(w), pm.Matches(dists, ids))


def test_rigid_apply_turns_normals_and_keeps_other_descriptors(pm):
    rng = np.random.default_rng(4)
    c, s = np.cos(0.3), np.sin(0.3)
    T = np.array([[c, -s, 0, 1], [s, c, 0, 2], [0, 0, 1, 3], [0, 0, 0, 1]], np.float32)
    cloud = pm.DataPoints(np.c_[rng.normal(0, 1, (50, 3)), np.ones(50)].astype(np.float32),
                          {"normals": rng.normal(0, 1, (50, 3)).astype(np.float32), "intensity": rng.uniform(size=(50, 1)).astype(np.float32)})
    out = pm.rigid_apply(T, cloud)
    assert np.abs(out.features - cloud.features @ T.T).max() < 1e-5 and (out.features[:, 3] == 1).all()
    assert np.abs(out.descriptors["normals"] - cloud.descriptors["normals"] @ T[:3, :3].T).max() < 1e-6
    assert (out.descriptors["intensity"] == cloud.descriptors["intensity"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("plane", [False, True])
def test_icp_error_elements_residual_and_overlap_on_gpu(pm, plane):
    """after a fused (capped) loop: ErrorElements from the resident matches reproduce the device's own statistics, and the residual /
    overlap of the mirror equal the oracle's on the same pairs"""
    from libpointmatcher_b200 import synth
    rd, rf, _ = synth.scan_pair(30000)
    icp = pm.ICP()
    icp.readingDataPointsFilters = [pm.SimpleSensorNoiseDataPointsFilter()]
    icp.referenceDataPointsFilters = [pm.SimpleSensorNoiseDataPointsFilter(), pm.SurfaceNormalDataPointsFilter({"knn": "10", "keepDensities": "1"})]
    icp.matcher = pm.KDTreeMatcher({"knn": "2"})
    icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.8"})])
    icp.errorMinimizer = pm.PointToPlaneErrorMinimizer() if plane else pm.PointToPointErrorMinimizer()
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "8"})]
    icp(pm.DataPoints(rd), pm.DataPoints(rf))
    em = icp.errorMinimizer
    ee = em.getErrorElements()
    stats = em._stats
    assert abs(ee.pointUsedRatio - stats["pointUsedRatio"]) < 1e-6 and abs(ee.weightedPointUsedRatio - stats["weightedPointUsedRatio"]) < 1e-6
    assert ee.nbRejectedMatches == stats["nbRejectedMatches"] and ee.nbRejectedPoints == stats["nbRejectedPoints"]
    assert set(ee.reference.descriptors) >= {"normals", "densities", "simpleSensorNoise"} and "simpleSensorNoise" in ee.reading.descriptors
    # kept pairs are what the matcher reported: |p - q|^2 is the match distance, bit for bit
    d = (ee.reading.features[:, :3] - ee.reference.features[:, :3]).astype(np.float32)
    d2 = ((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]).astype(np.float32) + d[:, 2] * d[:, 2]).astype(np.float32)
    assert (d2 == ee.matches.dists[:, 0]).all()
    eo = dict(reading=ee.reading.features, reading_desc=ee.reading.descriptors, reference=ee.reference.features,
              reference_desc=ee.reference.descriptors, weights=ee.weights, nbRejectedPoints=ee.nbRejectedPoints)
    res = em.getResidualError()
    ref = orc.point_to_plane_residual(eo) if plane else orc.point_to_point_residual(eo)
    assert res > 0 and abs(res - ref) < 1e-6 * ref
    ov = em.getOverlap()
    assert ov == (orc.point_to_plane_overlap(eo) if plane else orc.point_to_point_overlap(eo)) and 0 < ov <= 1
    icp.ctx.close()
