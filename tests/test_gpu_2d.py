"""2-D clouds (features.rows() == 3) on the GPU path against the 2-D oracle (oracle/oracle2d.py) and the reference's own 2-D
fixture (2D_twoBoxes -> 2D_oneBox, validT2d; utest/utest.h:44-60).  The device holds a 2-D point as (x, y, 0, w): every z term
is an exact zero, so the kernels compute what the reference computes on 3-row matrices."""
import os

import numpy as np
import pytest

from helpers import cloud

pytestmark = pytest.mark.gpu
FIXTURE = os.path.join(os.path.dirname(__file__), "golden", "reference_fixture.npz")


@pytest.fixture(scope="module")
def boxes():
    z = np.load(FIXTURE)
    one, two = z["box2d_one"], z["box2d_two"]
    return (np.c_[two, np.ones(len(two))].astype(np.float32), np.c_[one, np.ones(len(one))].astype(np.float32), z["validT2d"])


def close_to_valid(T, valid):
    return (abs(np.linalg.norm(T[:2, 2]) - np.linalg.norm(valid[:2, 2])) < 0.05 and abs(np.arccos(T[0, 0]) - np.arccos(valid[0, 0])) < 0.05)


def rot2_err(Ta, Tb):
    a = np.arctan2(Ta[1, 0], Ta[0, 0]) - np.arctan2(Tb[1, 0], Tb[0, 0])
    return abs(float(np.arctan2(np.sin(a), np.cos(a))))


def test_knn_2d_bit_exact_vs_bruteforce(gpu_ctx, oracle):
    rng = np.random.default_rng(31)
    for n in (5, 300, 20000):
        ref = cloud(rng, n, "uniform")[:, [0, 1, 3]].copy()
        q = cloud(rng, 500, "uniform")[:, [0, 1, 3]].copy()
        gpu_ctx.set_reference(ref)
        gpu_ctx.set_reading(q)
        T = np.array([[np.cos(0.1), -np.sin(0.1), 0.3], [np.sin(0.1), np.cos(0.1), -0.2], [0, 0, 1]], np.float32)
        from oracle import oracle2d as o2
        for k, md in ((1, np.inf), (4, np.inf), (3, 1.5)):
            if k > n:
                continue
            ib, db = oracle.bruteforce_knn(ref, q, k, md)
            ig, dg, _ = gpu_ctx.knn(None, k, 0.0, md)
            assert (ib == ig).all() and (db.view(np.uint32) == dg.view(np.uint32)).all(), (n, k, md)
            ib, db = oracle.bruteforce_knn(ref, o2.transform(T, q), k, md)
            ig, dg, _ = gpu_ctx.knn(T, k, 0.0, md)
            assert (ib == ig).all() and (db.view(np.uint32) == dg.view(np.uint32)).all(), ("T", n, k, md)
        assert gpu_ctx.get_reading().shape == (500, 3) and np.array_equal(gpu_ctx.get_reading(), q)


def test_minimizers_2d_match_oracle(gpu_ctx, oracle, boxes):
    from libpointmatcher_b200 import capi
    from oracle import oracle2d as o2
    rd, rf, _ = boxes
    nrm = o2.surface_normals(rf, 7)["normals"]
    gpu_ctx.set_reference(rf, nrm)
    gpu_ctx.set_reading(rd)
    ids, dists, _ = gpu_ctx.knn(None, 2)
    w, lim = gpu_ctx.weights([(capi.FILTER_TRIMMEDDIST, 0.85)])
    wo, lo = oracle.outlier_weights(dists, [(oracle.FILTER_TRIMMEDDIST, 0.85)])
    assert lim[0] == lo[0] and (w == wo).all()
    for kind, name in ((capi.MIN_P2POINT, "point"), (capi.MIN_P2PLANE, "plane")):
        T, _, stats = gpu_ctx.minimize(kind)
        To = o2.point_to_point(rd, rf, ids, dists, w) if name == "point" else o2.point_to_plane(rd, rf, nrm, ids, dists, w)
        assert T.shape == (3, 3) and T[2, 0] == 0 and T[2, 1] == 0 and T[2, 2] == 1
        assert rot2_err(T, To) <= 1e-5 and np.linalg.norm(T[:2, 2].astype(np.float64) - To[:2, 2]) <= 1e-5, (name, T, To)
        assert stats["nbKept"] == int((w != 0).sum())
    with pytest.raises(capi.PmGpuError) as e:
        gpu_ctx.minimize(capi.MIN_P2PLANE_COV)
    assert e.value.code == capi.ERR_UNSUPPORTED


@pytest.mark.parametrize("minimizer", ["point", "plane"])
def test_icp_2d_matches_oracle_and_validT2d(oracle, boxes, minimizer):
    """whole registrations through the mirror: iteration count (Counter + Differential with its 3x3 quirk) and transform
    equal the oracle's, and the result is the reference's validT2d within its own 0.05"""
    from libpointmatcher_b200 import pm
    from oracle import oracle2d as o2
    rd, rf, valid = boxes
    nrm = o2.surface_normals(rf, 7)["normals"]
    ro = o2.icp(rd, rf, normals=nrm, filters=[(2, 0.85)], minimizer=minimizer, max_iterations=40, differential=(1e-3, 1e-3, 3))
    icp = pm.ICP()
    icp.matcher = pm.KDTreeMatcher()
    icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.85"})])
    icp.errorMinimizer = pm.PointToPointErrorMinimizer() if minimizer == "point" else pm.PointToPlaneErrorMinimizer()
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "40"}),
                                  pm.DifferentialTransformationChecker({"minDiffRotErr": "0.001", "minDiffTransErr": "0.001", "smoothLength": "3"})]
    T = icp(pm.DataPoints(rd), pm.DataPoints(rf, {"normals": nrm}))
    assert T.shape == (3, 3)
    assert icp.iterationCount == ro["iterations"]
    assert rot2_err(T, ro["T"]) <= 1e-5 and np.linalg.norm(T[:2, 2].astype(np.float64) - ro["T"][:2, 2]) <= 1e-5, (T, ro["T"])
    assert close_to_valid(T, valid)
    # with an initial guess (3 x 3)
    T0 = np.array([[np.cos(0.05), -np.sin(0.05), 0.02], [np.sin(0.05), np.cos(0.05), 0.03], [0, 0, 1]], np.float32)
    r1 = o2.icp(rd, rf, normals=nrm, T_init=T0, filters=[(2, 0.85)], minimizer=minimizer, max_iterations=12)
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "12"})]
    T1 = icp(pm.DataPoints(rd), pm.DataPoints(rf, {"normals": nrm}), T0)
    assert rot2_err(T1, r1["T"]) <= 1e-5 and np.linalg.norm(T1[:2, 2].astype(np.float64) - r1["T"][:2, 2]) <= 1e-5
    icp.ctx.close()


def test_surface_normals_2d_match_oracle(gpu_ctx, boxes):
    from oracle import oracle2d as o2
    _, rf, _ = boxes
    o = o2.surface_normals(rf, 7)
    g = gpu_ctx.normals(rf, knn=7, keep=("normals", "densities", "eigValues", "matchedIds"))
    assert g["normals"].shape == (len(rf), 2) and g["eigValues"].shape == (len(rf), 2)
    same = (g["matchedIds"].astype(np.int32) == o["ids"]).all(axis=1)
    assert same.mean() > 0.98
    gap = (o["eigValues"][:, 1] - o["eigValues"][:, 0]) / np.maximum(o["eigValues"][:, 1], 1e-30)
    well = same & (gap > 1e-3)
    dots = np.abs((g["normals"] * o["normals"]).sum(1))
    assert (1.0 - dots[well]).max() <= 1e-5
    assert np.allclose(g["densities"][same, 0], o["densities"][same], rtol=1e-4)
    assert np.allclose(g["eigValues"][same], o["eigValues"][same], rtol=1e-3, atol=1e-7)


@pytest.mark.parametrize("minimizer", ["point", "plane"])
def test_reference_default_chain_on_2d_clouds(boxes, minimizer):
    """utest/ui/ErrorMinimizers.cpp:34-46: setDefault() (RandomSampling on the reading, SamplingSurfaceNormal on the reference, both
    on the host with libc's rand stream) with either minimiser registers 2D_twoBoxes on 2D_oneBox within 0.05 of validT2d"""
    from libpointmatcher_b200 import capi, pm
    rd, rf, valid = boxes
    capi.lib.pmgpu_host_srand(1)
    icp = pm.ICP()
    icp.setDefault()
    if minimizer == "point":
        icp.errorMinimizer = pm.PointToPointErrorMinimizer()
    T = icp(pm.DataPoints(rd), pm.DataPoints(rf))
    icp.ctx.close()
    assert T.shape == (3, 3) and close_to_valid(T, valid), T
