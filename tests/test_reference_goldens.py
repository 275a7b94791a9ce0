"""The reference's OWN golden vectors for this path (tests/golden/reference_fixture.npz, packed from
/root/reference/examples/data by tests/golden/make_reference_fixture.py):

  * examples/data/icp_data/*.ref_trans — `TEST(icpTest, icpTest)` (utest/utest.cpp:81-160): ICP of
    cloud.00001 onto cloud.00000 must agree with the stored transform to < 3 % median relative point
    displacement.  The goldens were produced with the CPU-only SamplingSurfaceNormal pre-filter; here the
    reference normals come from SurfaceNormalDataPointsFilter (knn 10) — the pass criterion is the
    reference's own.
  * validT3d — `IcpHelper::validate3dTransformation` (utest/utest.h:66-84): car_cloud401 onto
    car_cloud400, |t| within 0.1 and quaternion angular distance within 0.1 rad.

  * the same goldens through the reference's own chain files (YAML text in the fixture) now that the
    CPU pre-filters they name are built: `icp.loadFromYaml(...)`, `icp(data, ref)`,
    exactly what `TEST(icpTest, icpTest)` does.

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


# ---- the reference's chain files, end to end (pre-filters on the host, loop on the GPU) -----------
YAML_CHAINS = ["defaultIdentityDataPointsFilter", "defaultPointToPlaneMinDistDataPointsFilter", "defaultPointToPointMinDistDataPointsFilter",
               "defaultMaxDistDataPointsFilter", "SamplingSurfaceNormalDataPointsFilter1", "SamplingSurfaceNormalDataPointsFilter2",
               "SamplingSurfaceNormalDataPointsFilter3", "defaultRobustOutlierFilter"]


@pytest.mark.parametrize("name", ["defaultIdentityDataPointsFilter", "defaultPointToPlaneMinDistDataPointsFilter"])
def test_oracle_chain_with_prefilters_reproduces_golden(oracle, fx, name):
    """oracle pre-filters (bin sampling, one point per bin) + oracle ICP = the chain of the YAML file"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import prefilters as pre
    ref, data = homog(fx["cloud0"]), homog(fx["cloud1"])
    if "MinDist" in name:
        data_f = data[pre.min_dist(data, -1, 1.0)]
    else:
        data_f = data
    ref_f, normals, _, _, _ = pre.sampling_surface_normal_method1(ref, 10)
    r = oracle.icp(np.ascontiguousarray(data_f), np.ascontiguousarray(ref_f), ref_normals=np.ascontiguousarray(normals),
                   filters=[(oracle.FILTER_TRIMMEDDIST, 0.75)], minimizer=1, max_iterations=40, differential=(0.001, 0.01, 4), nthreads=4)
    assert rel_err(r["T"], fx["golden_" + name], data) < 0.03


def test_oracle_robust_chain_reproduces_golden(oracle, fx):
    """defaultRobustOutlierFilter.yaml: knn 10, RobustOutlierFilter(cauchy, mad, tuning 1), PointToPoint"""
    ref, data = homog(fx["cloud0"]), homog(fx["cloud1"])
    r = oracle.icp(data, ref, knn=10, filters=[(oracle.robust_word("cauchy", oracle.SCALE_MAD), 1.0)], minimizer=0, max_iterations=40,
                   differential=(0.001, 0.01, 4), nthreads=4)
    assert rel_err(r["T"], fx["golden_defaultRobustOutlierFilter"], data) < 0.03


def test_oracle_similarity_chain_reproduces_golden(oracle, fx):
    """defaultSimilarityPointToPointMinDistDataPointsFilter.yaml: MinDist 1 on both clouds, TrimmedDist 0.75,
    PointToPointSimilarityErrorMinimizer, 150 iterations"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import prefilters as pre
    ref, data = homog(fx["cloud0"]), homog(fx["cloud1"])
    ref_f, data_f = ref[pre.min_dist(ref, -1, 1.0)], data[pre.min_dist(data, -1, 1.0)]
    r = oracle.icp(np.ascontiguousarray(data_f), np.ascontiguousarray(ref_f), filters=[(oracle.FILTER_TRIMMEDDIST, 0.75)],
                   minimizer=oracle.MIN_P2POINT_SIM, max_iterations=150, differential=(0.001, 0.01, 4), nthreads=4)
    assert rel_err(r["T"], fx["golden_defaultSimilarityPointToPointMinDistDataPointsFilter"], data) < 0.03


@pytest.mark.parametrize("name", ["defaultBoundingBoxDataPointsFilter", "defaultDistanceLimitDataPointsFilter", "defaultMaxQuantileOnAxisDataPointsFilter",
                                  "defaultRemoveNaNDataPointsFilter", "defaultMaxPointCountDataPointsFilter"])
def test_oracle_chain_with_more_prefilters_reproduces_golden(oracle, fx, name):
    """oracle restatement of the reading filter (oracle/prefilters.py, parameters of the YAML file) + bin sampling of the
    reference + oracle ICP (knn 1, TrimmedDist 0.75, PointToPlane) against the reference's golden transform"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import prefilters as pre
    ref, data = homog(fx["cloud0"]), homog(fx["cloud1"])
    keep = {"defaultBoundingBoxDataPointsFilter": lambda: pre.bounding_box(data, (0.2, -1, -1), (1, 1, 1), True),
            "defaultDistanceLimitDataPointsFilter": lambda: pre.distance_limit(data, -1, 200, 0),
            "defaultMaxQuantileOnAxisDataPointsFilter": lambda: pre.max_quantile_on_axis(data, 0, 0.72),
            "defaultRemoveNaNDataPointsFilter": lambda: pre.remove_nan(data),
            "defaultMaxPointCountDataPointsFilter": lambda: pre.max_point_count(len(data), 1, 500)}[name]()
    ref_f, normals, _, _, _ = pre.sampling_surface_normal_method1(ref, 10)
    r = oracle.icp(np.ascontiguousarray(data[keep]), np.ascontiguousarray(ref_f), ref_normals=np.ascontiguousarray(normals),
                   filters=[(oracle.FILTER_TRIMMEDDIST, 0.75)], minimizer=1, max_iterations=40, differential=(0.001, 0.01, 4), nthreads=4)
    assert rel_err(r["T"], fx["golden_" + name], data) < 0.03


# the per-cloud host filters of the remaining chain files (BoundingBox, DistanceLimit, FixStepSampling, MaxDensity, MaxPointCount,
# MaxQuantileOnAxis, RemoveNaN, Shadow, SimpleSensorNoise, ObservationDirection, OrientNormals): all 21 files of utest icpTest
MORE_CHAINS = ["defaultBoundingBoxDataPointsFilter", "defaultDistanceLimitDataPointsFilter", "defaultFixStepSamplingDataPointsFilter",
               "defaultMaxDensityDataPointsFilter", "defaultMaxPointCountDataPointsFilter", "defaultMaxQuantileOnAxisDataPointsFilter",
               "defaultObservationDirectionDataPointsFilter", "defaultOrientNormalsDataPointsFilter", "defaultRemoveNaNDataPointsFilter",
               "defaultShadowDataPointsFilter", "defaultSimpleSensorNoiseDataPointsFilter"]


def test_fixture_holds_all_21_chain_files(fx):
    names = sorted(k[5:] for k in fx.files if k.startswith("yaml_") and k != "yaml_default")
    assert len(names) == 21 and all("golden_" + n in fx.files for n in names)
    assert sorted(YAML_CHAINS + MORE_CHAINS + ["defaultSimilarityPointToPointMinDistDataPointsFilter", "force4DOFForPointToPlaneMinimizer"]) == names


@pytest.mark.gpu
@pytest.mark.parametrize("name", YAML_CHAINS + MORE_CHAINS + ["defaultSimilarityPointToPointMinDistDataPointsFilter", "force4DOFForPointToPlaneMinimizer"])
def test_gpu_runs_reference_yaml_chain_to_golden(fx, name):
    from libpointmatcher_b200 import capi, pm
    ref, data = homog(fx["cloud0"]), homog(fx["cloud1"])
    capi.lib.pmgpu_host_srand(1)
    icp = pm.ICP()
    icp.loadFromYaml(str(fx["yaml_" + name]))
    T = icp(pm.DataPoints(data), pm.DataPoints(ref))
    icp.ctx.close()
    assert rel_err(T, fx["golden_" + name], data) < 0.03
    assert 1 <= icp.iterationCount <= 150


@pytest.mark.gpu
def test_gpu_default_chain_config1(fx):
    """BASELINE config 1: examples/icp_simple = ICP::setDefault() on cloud.00001 -> cloud.00000; default.yaml is the
    same chain with knn 10 and TrimmedDist 0.75 (its VTK inspector replaced by the Null one)."""
    from libpointmatcher_b200 import capi, pm
    ref, data = homog(fx["cloud0"]), homog(fx["cloud1"])
    golden = fx["golden_SamplingSurfaceNormalDataPointsFilter1"]   # same clouds, same kind of chain
    capi.lib.pmgpu_host_srand(1)
    icp = pm.ICP()
    icp.setDefault()
    T = icp(pm.DataPoints(data), pm.DataPoints(ref))
    assert rel_err(T, golden, data) < 0.03
    text = str(fx["yaml_default"])
    text = text[:text.index("inspector:\n VTKFileInspector")] + "inspector:\n  NullInspector\n\nlogger:\n  NullLogger\n"
    icp.loadFromYaml(text)
    T2 = icp(pm.DataPoints(data), pm.DataPoints(ref))
    icp.ctx.close()
    assert rel_err(T2, golden, data) < 0.03
