"""The reference's OWN golden vectors for this path (tests/golden/reference_fixture.npz, packed from
/root/reference/examples/data by tests/golden/make_reference_fixture.py):

  * examples/data/icp_data/*.ref_trans — `TEST(icpTest, icpTest)` (utest/utest.cpp:81-160): ICP of
    cloud.00001 onto cloud.00000 must agree with the stored transform to < 3 % median relative point
    displacement.  The goldens were produced with the CPU-only SamplingSurfaceNormal pre-filter; here the
    reference normals come from SurfaceNormalDataPointsFilter (knn 10) — the pass criterion is the
    reference's own.
  * validT3d — `IcpHelper::validate3dTransformation` (utest/utest.h:66-84): car_cloud401 onto
    car_cloud400, |t| within 0.1 and quaternion angular distance within 0.1 rad.

The oracle is checked on the CPU; the GPU path (through the C ABI) is checked with `-m gpu`.
"""
import os

import numpy as np
import pytest

FIXTURE = os.path.join(os.path.dirname(__file__), "golden", "reference_fixture.npz")
# golden name -> (minimizer id, max iterations): the matcher / filter / checker settings of the YAMLs
CASES = {
    "defaultIdentityDataPointsFilter": (1, 40),
    "defaultPointToPlaneMinDistDataPointsFilter": (1, 40),
    "defaultPointToPlaneWithCovErrorMinimizer": (3, 40),
    "defaultPointToPointWithCovErrorMinimizer": (2, 40),
    "defaultPointToPointMinDistDataPointsFilter": (0, 150),
}


def homog(p):
    return np.ascontiguousarray(np.c_[p[:, :3], np.ones(len(p))].astype(np.float32))


def rel_err(curT, refT, data):
    """utest/utest.cpp:146-158"""
    cur = curT.astype(np.float64) @ data.T.astype(np.float64)
    ref = refT.astype(np.float64) @ data.T.astype(np.float64)
    return np.median(np.abs(cur - ref)) / np.median(np.abs(cur))


@pytest.fixture(scope="module")
def fx():
    return np.load(FIXTURE)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden_ref_trans(oracle, fx, name):
    mini, iters = CASES[name]
    ref, data = homog(fx["cloud0"]), homog(fx["cloud1"])
    nrm = oracle.surface_normals(ref, knn=10, nthreads=4)["normals"]
    r = oracle.icp(data, ref, ref_normals=nrm, filters=[(oracle.FILTER_TRIMMEDDIST, 0.75)], minimizer=mini, max_iterations=iters,
                   differential=(0.001, 0.01, 4), nthreads=4)
    assert rel_err(r["T"], fx["golden_" + name], data) < 0.03


@pytest.mark.parametrize("mini", [0, 1])
def test_oracle_reproduces_validT3d(oracle, fx, mini):
    ref, data = homog(fx["car400"]), homog(fx["car401"])
    r = oracle.icp(data, ref, ref_normals=np.ascontiguousarray(fx["car400"][:, 3:6]), filters=[(oracle.FILTER_TRIMMEDDIST, 0.85)],
                   minimizer=mini, max_iterations=40, differential=(0.001, 0.001, 3), nthreads=4)
    valid = fx["validT3d"]
    assert abs(np.linalg.norm(r["T"][:3, 3]) - np.linalg.norm(valid[:3, 3])) < 0.1
    assert oracle.angular_distance(r["T"], valid.astype(np.float32)) < 0.1


def _gpu_icp(mini, iters, ratio, diff, data, ref, normals=None, normals_knn=None):
    from libpointmatcher_b200 import pm
    icp = pm.ICP()
    if normals_knn:
        icp.referenceDataPointsFilters = [pm.SurfaceNormalDataPointsFilter({"knn": str(normals_knn)})]
    icp.matcher = pm.KDTreeMatcher({"knn": "1", "epsilon": "0"})
    icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": repr(ratio)})])
    icp.errorMinimizer = [pm.PointToPointErrorMinimizer, pm.PointToPlaneErrorMinimizer, pm.PointToPointWithCovErrorMinimizer,
                          pm.PointToPlaneWithCovErrorMinimizer][mini]()
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": str(iters)}),
                                  pm.DifferentialTransformationChecker({"minDiffRotErr": repr(diff[0]), "minDiffTransErr": repr(diff[1]),
                                                                        "smoothLength": str(diff[2])})]
    desc = {} if normals is None else {"normals": np.ascontiguousarray(normals, np.float32)}
    T = icp(pm.DataPoints(data), pm.DataPoints(ref, desc))
    icp.ctx.close()
    return T


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_reproduces_golden_ref_trans(fx, name):
    mini, iters = CASES[name]
    ref, data = homog(fx["cloud0"]), homog(fx["cloud1"])
    T = _gpu_icp(mini, iters, 0.75, (0.001, 0.01, 4), data, ref, normals_knn=10)
    assert rel_err(T, fx["golden_" + name], data) < 0.03


@pytest.mark.gpu
@pytest.mark.parametrize("mini", [0, 1])
def test_gpu_reproduces_validT3d(oracle, fx, mini):
    ref, data = homog(fx["car400"]), homog(fx["car401"])
    T = _gpu_icp(mini, 40, 0.85, (0.001, 0.001, 3), data, ref, normals=fx["car400"][:, 3:6])
    valid = fx["validT3d"]
    assert abs(np.linalg.norm(T[:3, 3]) - np.linalg.norm(valid[:3, 3])) < 0.1
    assert oracle.angular_distance(T, valid.astype(np.float32)) < 0.1
