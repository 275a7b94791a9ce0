"""Parity at BASELINE.json's full sizes (1 M x 1 M): against the oracle where it finishes in
seconds on the box's host cores, and through size-independent properties."""
import numpy as np
import pytest

from helpers import assert_transform_close, classify_id_mismatches, small_pose

pytestmark = pytest.mark.gpu
N = 1_000_000


@pytest.fixture(scope="module")
def pair(synth):
    return synth.scan_pair(N)


def test_config2_knn_quantile_and_transform_at_1m(pair, oracle):
    """BASELINE configs[1]: 1 M reading vs 1 M reference, knn 1, TrimmedDist 0.75, PointToPoint"""
    from libpointmatcher_b200 import capi
    rd, rf, T_gt = pair
    threads = oracle.num_threads()
    with capi.Context(0) as ctx:
        ctx.set_reference(rf)
        ctx.set_reading(rd)
        ids, d, visits = ctx.knn(None, 1)
        io, do = oracle.KdTree(rf).knn(rd, 1, nthreads=threads)
        assert (d.view(np.uint32) == do.view(np.uint32)).all()          # distances bit-exact
        ndiff, nties = classify_id_mismatches(io, do, ids, d)
        assert ndiff == nties                                            # ids bit-exact, exact ties excepted
        w, lim = ctx.weights([(capi.FILTER_TRIMMEDDIST, 0.75)])
        wo, lo = oracle.outlier_weights(do, [(oracle.FILTER_TRIMMEDDIST, 0.75)])
        assert lim[0] == lo[0] and (w == wo).all()                       # quantile bit-exact, weights identical
        assert abs(w.mean() - 0.75) < 1e-3
        T, _, st = ctx.minimize(capi.MIN_P2POINT)
        To, _, so = oracle.minimize(oracle.MIN_P2POINT, rd, rf, None, ids, d, w, acc_double=True)
        assert_transform_close(T, To, 1e-5, 1e-5)
        assert st["nbKept"] == so["nbKept"]
        # properties: sorted / idempotent self-match
        ctx.set_reading(rf[:200000])
        ids2, d2, _ = ctx.knn(None, 3)
        assert (d2[:, 0] == 0).all() and (ids2[:, 0] == np.arange(200000)).all()   # a reference point's nearest neighbour is itself
        assert (np.diff(d2, axis=1) >= 0).all()                                   # ascending distances


def test_point_to_plane_icp_converges_at_1m(pair, oracle):
    """the north-star target config: 1 M x 1 M point-to-plane, normals from K8 (knn 20)"""
    from libpointmatcher_b200 import capi, pm
    rd, rf, T_gt = pair
    icp = pm.ICP()
    icp.referenceDataPointsFilters = [pm.SurfaceNormalDataPointsFilter({"knn": "20"})]
    icp.matcher = pm.KDTreeMatcher()
    icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.75"})])
    icp.errorMinimizer = pm.PointToPlaneErrorMinimizer()
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "40"})]
    T = icp(pm.DataPoints(rd), pm.DataPoints(rf))
    assert icp.iterationCount == 40
    assert np.linalg.norm(T[:3, 3].astype(np.float64) - T_gt[:3, 3]) < 2e-3      # the true pose of the re-scan
    R = T[:3, :3].astype(np.float64)
    assert np.allclose(R.T @ R, np.eye(3), atol=1e-5)
    icp.ctx.close()


def test_config4_shape_knn10_filters_cov(oracle, synth):
    """BASELINE configs[3] shape (knn 10, maxDist + MedianDist, PointToPlaneWithCov) on 500 k-point scans,
    the largest size the oracle finishes in seconds"""
    from libpointmatcher_b200 import capi
    rd, rf, _ = synth.scan_pair(500_000)
    threads = oracle.num_threads()
    with capi.Context(0) as ctx:
        ctx.set_reference(rf)
        ctx.ref_compute_normals(knn=20)
        ctx.set_reading(rd)
        ids, d, _ = ctx.knn(None, 10, 0.0, 2.0)
        io, do = oracle.KdTree(rf).knn(rd, 10, max_dist=2.0, nthreads=threads)
        assert (d.view(np.uint32) == do.view(np.uint32)).all()
        ndiff, nties = classify_id_mismatches(io, do, ids, d)
        assert ndiff == nties
        chain = [(capi.FILTER_MAXDIST, 1.0), (capi.FILTER_MEDIANDIST, 3.0)]
        w, lim = ctx.weights(chain)
        wo, lo = oracle.outlier_weights(do, chain)
        assert (lim.view(np.uint32) == lo.view(np.uint32)).all() and (w == wo).all()
        # a second search of the same reading under a new transform starts from the radius the first one's matches give (knn.cu,
        # k > 1): still the exact answer
        T = small_pose(np.random.default_rng(7), trans=0.2, ang=0.02)
        ids2, d2, _ = ctx.knn(T, 10, 0.0, 2.0)
        io2, do2 = oracle.KdTree(rf).knn(oracle.rigid_transform(T, rd), 10, max_dist=2.0, nthreads=threads)
        assert (d2.view(np.uint32) == do2.view(np.uint32)).all()
        ndiff, nties = classify_id_mismatches(io2, do2, ids2, d2)
        assert ndiff == nties


def test_surface_normals_at_1m(pair, oracle):
    """K8 at 1 M points, knn 20, every point against the oracle's SurfaceNormalDataPointsFilter (its own kd-tree over the
    full cloud, float eigen-solver) and against float64 eigenvectors of the same neighbourhoods"""
    from libpointmatcher_b200 import capi
    rd, rf, _ = pair
    with capi.Context(0) as ctx:
        g = ctx.normals(rf, knn=20, sort_eigen=True, keep=("normals", "matchedIds", "densities", "eigValues"))
    n = g["normals"]
    o = oracle.surface_normals(rf, knn=20, nthreads=oracle.num_threads(), sort_eigen=True)
    assert g["degenerate"] == o["degenerate"] == 0
    assert np.abs(np.linalg.norm(n, axis=1) - 1.0).max() < 1e-5
    assert (g["matchedIds"][:, 0] == np.arange(len(rf))).all()       # every point is its own first neighbour
    same = (g["matchedIds"].astype(np.int32) == o["ids"]).all(axis=1)  # exact distance ties may reorder a few neighbourhoods
    assert same.mean() > 0.999
    well = same & (o["gap"] > 1e-3)
    dots = np.abs((n * o["normals"]).sum(1))
    assert (1.0 - dots[well]).max() <= 1e-5
    assert np.allclose(g["densities"][same, 0], o["densities"][same], rtol=1e-5)
    scale = o["eigValues"][same].max(axis=1, keepdims=True)
    assert (np.abs(g["eigValues"][same] - o["eigValues"][same]) <= 2e-5 * scale + 1e-9).all()
    # and against float64 arithmetic on a sample of the agreed neighbourhoods
    sample = np.flatnonzero(same)[:: max(1, int(same.sum()) // 20000)]
    P = rf[o["ids"][sample]][:, :, :3].astype(np.float64)
    Pc = P - P.mean(1, keepdims=True)
    wv, V = np.linalg.eigh(np.einsum("nki,nkj->nij", Pc, Pc))
    gap = (wv[:, 1] - wv[:, 0]) / wv.sum(1)
    d64 = np.abs((V[:, :, 0] * n[sample]).sum(1))
    assert (1.0 - d64[gap > 1e-3]).max() < 1e-5


def _run_host_icp(rd, rf, nrm, matcher, filters, minimizer, iters):
    from libpointmatcher_b200 import pm
    icp = pm.ICP()
    icp.matcher = pm.KDTreeMatcher(matcher)
    icp.outlierFilters = pm.OutlierFilters([pm.OutlierFilterRegistrar.create(n, p) for n, p in filters])
    icp.errorMinimizer = getattr(pm, minimizer)()
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": str(iters)})]
    return icp(pm.DataPoints(rd), pm.DataPoints(rf, {"normals": nrm})), icp


def test_config3_10m_map_1m_reading(oracle, synth):
    """BASELINE configs[2] at its full size: a 10 M-point map (10 scans in the world frame), 1 M-point reading, knn 1, TrimmedDist
    0.75, PointToPlane, SurfaceNormal knn 20 on the map.  The oracle finishes this in seconds on the box's host threads (kd-tree
    build ~2 s, 1 M queries ~0.1 s), so matching is compared in full; the map's normals on a random sample."""
    from libpointmatcher_b200 import capi
    rf = synth.world_map(10_000_000, 10)
    rd = synth.scan(1_000_000, synth.READING_POSE, seed=synth.SEED + 1)
    threads = oracle.num_threads()
    chain = [(capi.FILTER_TRIMMEDDIST, 0.75)]
    with capi.Context(0) as ctx:
        ctx.set_reference(rf)
        ctx.ref_compute_normals(knn=20)
        nrm = ctx.ref_normals()
        ctx.set_reading(rd)
        ids, d, _ = ctx.knn(None, 1)
        tree = oracle.KdTree(rf)
        io, do = tree.knn(rd, 1, nthreads=threads)
        assert (d.view(np.uint32) == do.view(np.uint32)).all()
        ndiff, nties = classify_id_mismatches(io, do, ids, d)
        assert ndiff == nties
        w, lim = ctx.weights(chain)
        wo, lo = oracle.outlier_weights(do, chain)
        assert lim[0] == lo[0] and (w == wo).all()
        # map normals: unit length everywhere, and the oracle's neighbourhoods / float64 eigenvectors on a sample
        assert np.abs(np.linalg.norm(nrm, axis=1) - 1.0).max() < 1e-5
        rng = np.random.default_rng(11)
        sample = np.sort(rng.choice(len(rf), 20000, replace=False))
        isamp, _ = tree.knn(rf[sample], 20, nthreads=threads)
        del tree
        P = rf[isamp][:, :, :3].astype(np.float64)
        Pc = P - P.mean(1, keepdims=True)
        wv, V = np.linalg.eigh(np.einsum("nki,nkj->nij", Pc, Pc))
        gap = (wv[:, 1] - wv[:, 0]) / wv.sum(1)
        dots = np.abs((V[:, :, 0] * nrm[sample]).sum(1))
        # a neighbourhood with an exact distance tie at rank 20 may hold a different 20th point: allow 0.1 % of the sample
        assert ((1.0 - dots[gap > 1e-3]) > 1e-5).mean() < 1e-3
    # the whole registration through the host driver (which centres both clouds on the map's mean like ICP.cpp:291-299, as the
    # oracle's loop does) against the oracle's loop over the same normals
    res_o = oracle.icp(rd, rf, ref_normals=nrm, knn=1, filters=chain, minimizer=capi.MIN_P2PLANE, max_iterations=6,
                       nthreads=threads, acc_double=True)
    T, icp = _run_host_icp(rd, rf, nrm, {"knn": "1"}, [("TrimmedDistOutlierFilter", {"ratio": "0.75"})], "PointToPlaneErrorMinimizer", 6)
    assert icp.iterationCount == res_o["iterations"] == 6
    assert_transform_close(T, res_o["T"], 1e-5, 1e-5)
    icp.ctx.close()


def test_config4_at_2m(oracle, synth):
    """BASELINE configs[3] at its full size: 2 M x 2 M, knn 10, maxDist 2, MaxDist 1 x MedianDist 3, PointToPlaneWithCov"""
    from libpointmatcher_b200 import capi
    rd, rf, _ = synth.scan_pair(2_000_000)
    threads = oracle.num_threads()
    chain = [(capi.FILTER_MAXDIST, 1.0), (capi.FILTER_MEDIANDIST, 3.0)]
    with capi.Context(0) as ctx:
        ctx.set_reference(rf)
        ctx.ref_compute_normals(knn=20)
        nrm = ctx.ref_normals()
        ctx.set_reading(rd)
        ids, d, _ = ctx.knn(None, 10, 0.0, 2.0)
        io, do = oracle.KdTree(rf).knn(rd, 10, max_dist=2.0, nthreads=threads)
        assert (d.view(np.uint32) == do.view(np.uint32)).all()
        ndiff, nties = classify_id_mismatches(io, do, ids, d)
        assert ndiff == nties
        w, lim = ctx.weights(chain)
        wo, lo = oracle.outlier_weights(do, chain)
        assert (lim.view(np.uint32) == lo.view(np.uint32)).all() and (w == wo).all()
        res_o = oracle.icp(rd, rf, ref_normals=nrm, knn=10, max_dist=2.0, filters=chain, minimizer=capi.MIN_P2PLANE_COV,
                           max_iterations=4, nthreads=threads, acc_double=True)
    T, icp = _run_host_icp(rd, rf, nrm, {"knn": "10", "maxDist": "2.0"},
                           [("MaxDistOutlierFilter", {"maxDist": "1.0"}), ("MedianDistOutlierFilter", {"factor": "3.0"})],
                           "PointToPlaneWithCovErrorMinimizer", 4)
    assert icp.iterationCount == res_o["iterations"] == 4
    assert_transform_close(T, res_o["T"], 1e-5, 1e-5)
    c, co = np.asarray(icp.errorMinimizer.getCovariance(), np.float64), res_o["cov"].astype(np.float64)
    assert np.abs(c - co).max() <= 2e-3 * np.abs(co).max()
    icp.ctx.close()
