"""the CPU pre-filters of the default chain (RandomSampling, SamplingSurfaceNormal,
MinDist / MaxDist, Identity) and the YAML chain loader, against the oracle restatement in
oracle/prefilters.py.  These run on the host in the reference too, so everything here is a CPU test;
the chains that use them are run end to end against the reference's golden transforms with `-m gpu`
(tests/test_reference_goldens.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import prefilters as orc  # noqa: E402


@pytest.fixture(scope="module")
def pm():
    from libpointmatcher_b200 import pm
    return pm


@pytest.fixture(scope="module")
def capi():
    from libpointmatcher_b200 import capi
    return capi


def surface_cloud(n, seed=0):
    """a noisy curved sheet plus a wall: well-defined local normals, no repeated coordinates"""
    rng = np.random.default_rng(seed)
    xy = rng.uniform(-10, 10, (n, 2))
    z = 0.05 * xy[:, 0] ** 2 * 0.1 + rng.normal(0, 0.002, n)
    pts = np.c_[xy, z]
    wall = n // 4
    pts[:wall] = np.c_[rng.uniform(-10, 10, wall), np.full(wall, 10.0) + rng.normal(0, 0.002, wall), rng.uniform(0, 4, wall)]
    return np.ascontiguousarray(np.c_[pts, np.ones(n)].astype(np.float32))


@pytest.mark.parametrize("prob,seed", [(0.5, 1), (0.75, 7), (0.05, 123), (1.0, 3)])
def test_random_sampling_is_the_reference_rand_stream(pm, capi, prob, seed):
    cloud = surface_cloud(5000)
    desc = {"intensity": np.arange(5000, dtype=np.float32)[:, None]}
    capi.lib.pmgpu_host_srand(seed)
    out = pm.RandomSamplingDataPointsFilter({"prob": repr(prob)}).filter(pm.DataPoints(cloud, desc))
    orc.srand(seed)
    keep = orc.random_sampling(5000, prob)
    assert (out.descriptors["intensity"][:, 0].astype(np.int64) == keep).all()
    assert (out.features == cloud[keep]).all()
    assert abs(len(keep) / 5000 - min(prob, 1.0)) < 0.03


def _match_rows(a, b, tol):
    """rows of a and b as the same multiset (lexicographic order), within tol"""
    ia, ib = np.lexsort(a[:, ::-1].T), np.lexsort(b[:, ::-1].T)
    assert a.shape == b.shape
    assert np.abs(a[ia] - b[ib]).max() < tol
    return ia, ib


def test_sampling_surface_normal_one_point_per_bin(pm):
    cloud = surface_cloud(6000, seed=2)
    colour = np.random.default_rng(5).uniform(0, 1, (6000, 2)).astype(np.float32)
    f = pm.SamplingSurfaceNormalDataPointsFilter({"knn": "10", "samplingMethod": "1", "keepDensities": "1", "averageExistingDescriptors": "1"})
    out = f.filter(pm.DataPoints(cloud, {"colour": colour}))
    feats, normals, dens, descs, unfit = orc.sampling_surface_normal_method1(cloud, 10, descriptors=colour)
    assert unfit == f.unfitPointsCount == 0
    ia, ib = _match_rows(out.features, feats, 1e-4)                       # bins as sets: the same bin means
    dots = np.abs((out.descriptors["normals"][ia] * normals[ib]).sum(axis=1))
    assert (dots > 0.999).mean() > 0.99 and np.median(dots) > 0.999999    # one normal per bin, sign free
    assert np.allclose(np.linalg.norm(out.descriptors["normals"], axis=1), 1, atol=1e-5)
    assert np.allclose(out.descriptors["densities"][ia, 0], dens[ib], rtol=1e-3)
    assert np.allclose(out.descriptors["colour"][ia], descs[ib], atol=1e-5)
    assert (out.features[:, 3] == 1).all()
    assert 6000 / 10 <= len(out.features) <= 6000 / 5                     # bins hold between knn/2 and knn points


def test_sampling_surface_normal_random_subsampling(pm, capi):
    cloud = surface_cloud(6000, seed=3)
    tag = np.arange(6000, dtype=np.float32)[:, None]
    capi.lib.pmgpu_host_srand(1)
    f = pm.SamplingSurfaceNormalDataPointsFilter({"knn": "10", "samplingMethod": "0", "ratio": "0.7"})
    out = f.filter(pm.DataPoints(cloud, {"tag": tag}))
    kept = out.descriptors["tag"][:, 0].astype(np.int64)
    assert (np.diff(kept) > 0).all()                                      # sorted by original index (SamplingSurfaceNormal.cpp:146)
    assert (out.features == cloud[kept]).all()                            # points are kept as they are
    assert abs(len(kept) / 6000 - 0.7) < 0.03
    owner = np.empty(6000, np.int64)
    fused = {}
    for b, idx in enumerate(orc.bins(cloud, 10)):
        owner[idx] = b
        fused[b] = orc.fuse(cloud, idx)
    dots = np.array([abs(float(out.descriptors["normals"][i] @ fused[owner[k]][1])) for i, k in enumerate(kept)])
    assert (dots > 0.999).mean() > 0.99                                   # every kept point carries its bin's normal


def test_sampling_surface_normal_drops_large_and_degenerate_bins(pm):
    cloud = surface_cloud(3000, seed=4)
    f = pm.SamplingSurfaceNormalDataPointsFilter({"knn": "12", "samplingMethod": "1", "maxBoxDim": "0.9"})
    out = f.filter(pm.DataPoints(cloud))
    feats, _, _, _, unfit = orc.sampling_surface_normal_method1(cloud, 12, max_box_dim=0.9)
    assert f.unfitPointsCount == unfit > 0 and len(out.features) == len(feats) > 0
    # collinear points: the covariance has rank 1 -> no normal, every point is dropped
    line = np.zeros((64, 4), np.float32)
    line[:, 0] = np.linspace(0, 1, 64, dtype=np.float32) ** 2
    line[:, 3] = 1
    f = pm.SamplingSurfaceNormalDataPointsFilter({"knn": "8", "samplingMethod": "1"})
    out = f.filter(pm.DataPoints(line))
    assert len(out.features) == 0 and f.unfitPointsCount == 64


def test_min_max_dist_filters(pm):
    cloud = surface_cloud(4000, seed=6)
    for dim, val in ((-1, 6.5), (-1, -6.5), (0, 1.25), (2, 0.3)):
        out = pm.MinDistDataPointsFilter({"dim": str(dim), "minDist": repr(val)}).filter(pm.DataPoints(cloud))
        assert (out.features == cloud[orc.min_dist(cloud, dim, val)]).all()
        out = pm.MaxDistDataPointsFilter({"dim": str(dim), "maxDist": repr(val)}).filter(pm.DataPoints(cloud))
        assert (out.features == cloud[orc.max_dist(cloud, dim, val)]).all()
    out = pm.IdentityDataPointsFilter().filter(pm.DataPoints(cloud))
    assert (out.features == cloud).all()
    with pytest.raises(pm.InvalidParameter):
        pm.MinDistDataPointsFilter({"dim": "3"})


def test_parameter_tables_of_the_prefilters(pm):
    """names, defaults and bounds as in RandomSampling.h:58-63, SamplingSurfaceNormal.h:60-74, MinDist.h:56-62"""
    assert pm.RandomSamplingDataPointsFilter().prob == 0.75
    f = pm.SamplingSurfaceNormalDataPointsFilter()
    assert (f.ratio, f.knn, f.samplingMethod, f.maxBoxDim) == (0.5, 7, 0, np.inf)
    assert (f.averageExistingDescriptors, f.keepNormals, f.keepDensities, f.keepEigenValues, f.keepEigenVectors) == (True, True, False, False, False)
    for bad in ({"knn": "2"}, {"ratio": "1.5"}, {"samplingMethod": "2"}, {"nope": "1"}):
        with pytest.raises(pm.InvalidParameter):
            pm.SamplingSurfaceNormalDataPointsFilter(bad)
    with pytest.raises(pm.InvalidParameter):
        pm.RandomSamplingDataPointsFilter({"prob": "1.5"})
    m = pm.MinDistDataPointsFilter()
    assert (m.dim, m.minDist) == (-1, 1.0)
    for name in ("RandomSamplingDataPointsFilter", "SamplingSurfaceNormalDataPointsFilter", "MinDistDataPointsFilter", "MaxDistDataPointsFilter",
                 "IdentityDataPointsFilter", "SurfaceNormalDataPointsFilter"):
        assert name in pm.DataPointsFilterRegistrar


def test_yaml_chains_of_the_reference_load(pm):
    """ICPChainBase::loadFromYaml on the reference's own chain files (text packed into the fixture)"""
    fx = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixture.npz"))
    icp = pm.ICP()
    icp.loadFromYaml(str(fx["yaml_defaultPointToPlaneMinDistDataPointsFilter"]))
    assert [type(f).__name__ for f in icp.readingDataPointsFilters] == ["MinDistDataPointsFilter"]
    ssn = icp.referenceDataPointsFilters[0]
    assert type(ssn).__name__ == "SamplingSurfaceNormalDataPointsFilter" and (ssn.knn, ssn.samplingMethod, ssn.averageExistingDescriptors) == (10, 1, False)
    assert abs(ssn.ratio - 0.666666) < 1e-7
    assert type(icp.errorMinimizer).__name__ == "PointToPlaneErrorMinimizer" and icp.matcher.knn == 1
    assert abs(icp.outlierFilters[0].get("ratio") - 0.75) < 1e-7
    assert [type(c).__name__ for c in icp.transformationCheckers] == ["CounterTransformationChecker", "DifferentialTransformationChecker"]
    assert icp.transformationCheckers[1].smoothLength == 4
    icp.loadFromYaml(str(fx["yaml_defaultPointToPointMinDistDataPointsFilter"]))
    assert type(icp.errorMinimizer).__name__ == "PointToPointErrorMinimizer"
    assert [type(f).__name__ for f in icp.readingDataPointsFilters] == ["MinDistDataPointsFilter", "RandomSamplingDataPointsFilter"]
    with pytest.raises(pm.InvalidElement):      # default.yaml asks for the VTKFileInspector: C++ host layer only
        icp.loadFromYaml(str(fx["yaml_default"]))
    with pytest.raises(pm.InvalidElement):
        icp.loadFromYaml("matcher:\n  NoSuchMatcher:\n    knn: 1\n")
    icp.setDefault()                            # ICP.cpp:100-113
    assert type(icp.readingDataPointsFilters[0]).__name__ == "RandomSamplingDataPointsFilter"
    assert type(icp.referenceDataPointsFilters[0]).__name__ == "SamplingSurfaceNormalDataPointsFilter"
    assert abs(icp.outlierFilters[0].get("ratio") - 0.85) < 1e-7


def test_loaders_read_the_reference_data_layouts(pm, tmp_path):
    """DataPoints::load (IO.cpp:376-392): CSV with the header of examples/data/car_cloud400.csv, ASCII legacy VTK with the
    layout of examples/data/cloud.00000.vtk; written from the packed fixture and read back bit for bit."""
    fx = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixture.npz"))
    car = fx["car400"][:500]
    with open(tmp_path / "car.csv", "w") as f:
        f.write("x,y,z,nx,ny,nz\n")
        for r in car:
            f.write("%r , %r , %r , %r,%r,%r\n" % tuple(float(v) for v in r))
    c = pm.DataPoints.load(str(tmp_path / "car.csv"))
    assert (c.features[:, :3] == car[:, :3]).all() and (c.features[:, 3] == 1).all()
    assert (c.descriptors["normals"] == car[:, 3:6]).all()
    pts = fx["cloud0"][:700]
    with open(tmp_path / "cloud.vtk", "w") as f:
        f.write("# vtk DataFile Version 3.0\ndata\nASCII\nDATASET POLYDATA\nPOINTS %d float\n" % len(pts))
        for p in pts:
            f.write("%r %r %r \n" % tuple(float(v) for v in p))
        f.write("VERTICES %d %d\n" % (len(pts), 2 * len(pts)))
        for i in range(len(pts)):
            f.write("1 %d\n" % i)
        f.write("POINT_DATA %d\nNORMALS normals float\n" % len(pts))
        for i in range(len(pts)):
            f.write("0 0 1\n")
        f.write("SCALARS densities float 1\nLOOKUP_TABLE default\n" + "\n".join(str(float(i)) for i in range(len(pts))) + "\n")
    v = pm.DataPoints.load(str(tmp_path / "cloud.vtk"))
    assert (v.features[:, :3] == pts).all() and (v.features[:, 3] == 1).all()
    assert (v.descriptors["normals"] == [0, 0, 1]).all() and (v.descriptors["densities"][:, 0] == np.arange(len(pts))).all()
    (tmp_path / "two.csv").write_text("0.5 1.5\n2.5 3.5\n")
    two = pm.DataPoints.load(str(tmp_path / "two.csv"))
    assert two.features.tolist() == [[0.5, 1.5, 1.0], [2.5, 3.5, 1.0]]
    (tmp_path / "five.csv").write_text("1 2 3 4 5\n")
    with pytest.raises(RuntimeError):
        pm.DataPoints.load(str(tmp_path / "five.csv"))           # the reference asks on stdin here
    with pytest.raises(RuntimeError):
        pm.DataPoints.load(str(tmp_path / "missing.vtk"))
    with pytest.raises(RuntimeError):
        pm.DataPoints.load(str(tmp_path / "cloud.xyz"))
    ref_dir = "/root/reference/examples/data"                    # only in the build container
    if os.path.isdir(ref_dir):
        assert (pm.DataPoints.load(ref_dir + "/cloud.00000.vtk").features[:, :3] == fx["cloud0"]).all()
        car_full = pm.DataPoints.load(ref_dir + "/car_cloud400.csv")
        assert (car_full.features[:, :3] == fx["car400"][:, :3]).all() and (car_full.descriptors["normals"] == fx["car400"][:, 3:6]).all()


# ---- the other per-cloud host filters of the reference's golden chain files ------------------------
def chain_cloud(n=3000, seed=3):
    rng = np.random.default_rng(seed)
    f = np.c_[rng.normal(0, 2.5, (n, 3)), np.ones(n)].astype(np.float32)
    normals = rng.normal(0, 1, (n, 3)).astype(np.float32)
    normals[::97] = 0                                  # zero normals: normalized() leaves them alone
    dens = rng.uniform(0, 1, (n, 1)).astype(np.float32)
    dens[::50] = dens.max()                            # saturated densities
    return f, normals, dens


def _run(pm, name, params, f, normals, dens):
    cloud = pm.DataPoints(f, {"normals": normals, "densities": dens, "index": np.arange(len(f), dtype=np.float32)[:, None]})
    return pm.DataPointsFilterRegistrar.create(name, params).filter(cloud)


@pytest.mark.parametrize("params,lo,hi,inside", [({"xMin": "0.2"}, (0.2, -1, -1), (1, 1, 1), True),
                                                  ({"xMin": "-3", "xMax": "2", "yMin": "-1", "yMax": "4", "zMin": "-2", "zMax": "2", "removeInside": "0"},
                                                   (-3, -1, -2), (2, 4, 2), False)])
def test_bounding_box_filter(pm, params, lo, hi, inside):
    f, nrm, dens = chain_cloud()
    out = _run(pm, "BoundingBoxDataPointsFilter", params, f, nrm, dens)
    keep = orc.bounding_box(f, lo, hi, inside)
    assert 0 < len(keep) < len(f)
    assert (out.descriptors["index"][:, 0] == keep).all() and (out.features == f[keep]).all() and (out.descriptors["normals"] == nrm[keep]).all()


@pytest.mark.parametrize("dim,dist,inside", [(-1, 3.0, 0), (-1, -3.0, 1), (1, -0.5, 1), (2, 0.7, 0)])
def test_distance_limit_filter(pm, dim, dist, inside):
    f, nrm, dens = chain_cloud()
    out = _run(pm, "DistanceLimitDataPointsFilter", {"dim": str(dim), "dist": repr(dist), "removeInside": str(inside)}, f, nrm, dens)
    keep = orc.distance_limit(f, dim, dist, inside)
    assert 0 < len(keep) < len(f) and (out.descriptors["index"][:, 0] == keep).all()
    with pytest.raises(pm.InvalidParameter):
        pm.DistanceLimitDataPointsFilter({"dim": "2"}).filter(pm.DataPoints(np.ones((5, 3), np.float32)))


def test_fix_step_sampling_filter_walks_its_step(pm, capi):
    f, nrm, dens = chain_cloud()
    flt = pm.FixStepSamplingDataPointsFilter({"startStep": "7", "endStep": "3", "stepMult": "0.7"})
    capi.lib.pmgpu_host_srand(11)
    got = []
    for _ in range(5):
        got.append(flt.filter(pm.DataPoints(f, {"index": np.arange(len(f), dtype=np.float32)[:, None]})).descriptors["index"][:, 0])
    orc.srand(11)
    step = 7.0
    for g in got:
        assert (g == orc.fix_step(len(f), step)).all()
        step = orc.fix_step_next(step, 7, 3, 0.7)
    assert step == 3.0 and len(got[0]) in (428, 429) and len(got[-1]) == 1000
    flt.init()
    assert flt.step == 7.0


@pytest.mark.parametrize("count,seed", [(500, 1), (1, 9), (2999, 4), (3000, 1), (100000, 5)])
def test_max_point_count_filter(pm, count, seed):
    f, nrm, dens = chain_cloud()
    out = _run(pm, "MaxPointCountDataPointsFilter", {"maxCount": str(count), "seed": str(seed)}, f, nrm, dens)
    keep = orc.max_point_count(len(f), seed, count)
    assert len(keep) == min(count, len(f))
    assert (out.descriptors["index"][:, 0] == keep).all() and (out.features == f[keep]).all()


@pytest.mark.parametrize("dim,ratio", [(0, 0.72), (2, 0.333), (1, 0.9999999), (0, 0.0000001)])
def test_max_quantile_on_axis_filter(pm, dim, ratio):
    f, nrm, dens = chain_cloud()
    f[5:25, dim] = f[100, dim]                         # ties at one value
    out = _run(pm, "MaxQuantileOnAxisDataPointsFilter", {"dim": str(dim), "ratio": repr(ratio)}, f, nrm, dens)
    keep = orc.max_quantile_on_axis(f, dim, ratio)
    assert (out.descriptors["index"][:, 0] == keep).all()
    assert len(keep) <= int(np.float32(len(f)) * np.float32(ratio))


def test_remove_nan_filter(pm):
    f, nrm, dens = chain_cloud()
    f[3, 0] = f[77, 2] = f[2999, 1] = np.nan
    out = _run(pm, "RemoveNaNDataPointsFilter", {}, f, nrm, dens)
    keep = orc.remove_nan(f)
    assert len(keep) == len(f) - 3 and (out.descriptors["index"][:, 0] == keep).all()


@pytest.mark.parametrize("limit,seed", [(0.3, 1), (0.05, 8), (2.0, 1)])
def test_max_density_filter(pm, capi, limit, seed):
    f, nrm, dens = chain_cloud()
    capi.lib.pmgpu_host_srand(seed)
    out = _run(pm, "MaxDensityDataPointsFilter", {"maxDensity": repr(limit)}, f, nrm, dens)
    orc.srand(seed)
    keep = orc.max_density(dens[:, 0], limit)
    assert (out.descriptors["index"][:, 0] == keep).all()
    assert (len(keep) == len(f)) == (limit > 1)
    with pytest.raises(pm.InvalidField):
        pm.MaxDensityDataPointsFilter().filter(pm.DataPoints(f))


@pytest.mark.parametrize("eps", [0.00001, 0.3, 1.2])
def test_shadow_filter(pm, eps):
    f, nrm, dens = chain_cloud()
    out = _run(pm, "ShadowDataPointsFilter", {"eps": repr(eps)}, f, nrm, dens)
    keep = orc.shadow(f, nrm, eps)
    assert 0 < len(keep) < len(f) and (out.descriptors["index"][:, 0] == keep).all()
    with pytest.raises(pm.InvalidField):
        pm.ShadowDataPointsFilter().filter(pm.DataPoints(f))


@pytest.mark.parametrize("sensor", [0, 1, 2, 3, 4])
def test_simple_sensor_noise_filter(pm, sensor):
    f, nrm, dens = chain_cloud(500)
    out = _run(pm, "SimpleSensorNoiseDataPointsFilter", {"sensorType": str(sensor), "gain": "2"}, f, nrm, dens)
    assert (out.descriptors["simpleSensorNoise"][:, 0] == orc.simple_sensor_noise(f, sensor)).all() and len(out.features) == 500
    with pytest.raises(pm.InvalidParameter):
        pm.SimpleSensorNoiseDataPointsFilter({"sensorType": "5"})


def test_all_21_reference_chain_files_load(pm):
    """every chain file of utest icpTest parses into modules registered under the reference's names"""
    fx = np.load(os.path.join(ROOT, "tests", "golden", "reference_fixture.npz"))
    names = [k for k in fx.files if k.startswith("yaml_") and k != "yaml_default"]
    assert len(names) == 21
    for k in names:
        icp = pm.ICP()
        icp.loadFromYaml(str(fx[k]))
        assert icp.matcher is not None and icp.errorMinimizer is not None and len(icp.transformationCheckers) >= 1
        # the inspector they name (Null or Performance) collects statistics at most: the loop stays fused
        assert type(icp.inspector).__name__ in ("NullInspector", "PerformanceInspector") and not icp.inspector.needsIterationData()


def test_datapoints_concatenate(pm):
    """DataPoints::concatenate (DataPoints.cpp:225-330): only the descriptors both clouds carry survive"""
    a = pm.DataPoints(np.ones((2, 4), np.float32), {"normals": np.full((2, 3), 5, np.float32), "densities": np.full((2, 1), 6, np.float32)})
    b = pm.DataPoints(np.full((3, 4), 2, np.float32), {"densities": np.full((3, 1), 7, np.float32), "color": np.full((3, 4), 8, np.float32)})
    a.concatenate(b)
    assert a.features.shape == (5, 4) and a.features[1, 0] == 1 and a.features[4, 3] == 2
    assert list(a.descriptors) == ["densities"] and a.descriptors["densities"][:, 0].tolist() == [6, 6, 7, 7, 7]
    with pytest.raises(pm.InvalidField):
        a.concatenate(pm.DataPoints(np.zeros((1, 3), np.float32)))
    c = pm.DataPoints(np.zeros((1, 4), np.float32), {"normals": np.zeros((1, 3), np.float32)})
    with pytest.raises(pm.InvalidField):
        c.concatenate(pm.DataPoints(np.zeros((1, 4), np.float32), {"normals": np.zeros((1, 2), np.float32)}))


def test_null_outlier_filter_contributes_no_filter_word(pm):
    """NullOutlierFilter (OutlierFiltersImpl.cpp:45-58) is a factor of one: registered, loadable from YAML, absent from the words the
    device is given"""
    chain = pm.OutlierFilters([pm.OutlierFilterRegistrar.create("NullOutlierFilter"), pm.TrimmedDistOutlierFilter({"ratio": "0.7"})])
    assert [t for t, _ in chain.specs()] == [pm.TrimmedDistOutlierFilter.TYPE] and abs(chain.specs()[0][1] - 0.7) < 1e-7
    icp = pm.ICP()
    icp.loadFromYaml("outlierFilters:\n  - NullOutlierFilter\nmatcher:\n  KDTreeMatcher\n")
    assert [type(f).__name__ for f in icp.outlierFilters] == ["NullOutlierFilter"] and icp.outlierFilters.specs() == []
    with pytest.raises(pm.InvalidParameter):
        pm.NullOutlierFilter({"ratio": "1"})


def test_min_dist_outlier_filter_module(pm):
    f = pm.OutlierFilterRegistrar.create("MinDistOutlierFilter", {"minDist": "0.2"})
    assert f.spec()[0] == 6 and abs(f.spec()[1] - 0.2) < 1e-7
    with pytest.raises(pm.InvalidParameter):
        pm.MinDistOutlierFilter({"minDist": "0"})
