"""GPU parity tests: the CUDA path, called through the C ABI (capi), against the oracle on the
same seeded inputs.  Bars (BASELINE.json north_star): match indices bit-exact (exact ties
excepted when the oracle is the kd-tree; none excepted against the brute-force oracle), squared
distances bit-exact, quantile limits bit-exact, weights identical, transforms within
1e-5 rad / 1e-5 m.
"""
import numpy as np
import pytest

from helpers import assert_transform_close, classify_id_mismatches, cloud, small_pose

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


# ---------------------------------------------------------------------------------- K1 + K2
@pytest.mark.parametrize("kind", ["uniform", "grid", "plane", "cluster"])
def test_knn_bit_exact_vs_bruteforce(gpu_ctx, oracle, kind):
    rng = np.random.default_rng(100)
    for n in (1, 2, 8, 9, 100, 3000):
        ref, q = cloud(rng, n, kind), cloud(rng, 257, kind)
        if kind != "grid":
            q[:, :3] += rng.normal(0, 0.5, (257, 3)).astype(np.float32)
        gpu_ctx.set_reference(ref)
        gpu_ctx.set_reading(q)
        for k in (1, 2, 5, 10, 20, 40):
            if k > n:
                continue
            for md in (np.inf, 1.0):
                ib, db = oracle.bruteforce_knn(ref, q, k, md, nthreads=4)
                ig, dg, visits = gpu_ctx.knn(None, k, 0.0, md)
                assert (ib == ig).all() and (bits(db) == bits(dg)).all(), (kind, n, k, md)
                assert ((dg == np.inf) == (ig == -1)).all()


@pytest.mark.parametrize("kind", ["uniform", "grid", "plane"])
def test_knn_bit_exact_with_upper_build_levels(gpu_ctx, oracle, kind):
    """clouds large enough for the level-by-level part of the structure build (segments > 4096 points: radix select + count +
    scatter, tree_build.cu) at sizes where segments straddle the 4096-position chunks unevenly, and with many EQUAL coordinates
    (grid / plane: the median key has ties that must fill the left half exactly)"""
    rng = np.random.default_rng(104)
    for n in (4097, 8191, 8193, 12289, 40961, 100003):
        ref, q = cloud(rng, n, kind), cloud(rng, 300, kind)
        if kind == "plane":
            ref[:, 0] = np.round(ref[:, 0], 1)   # ~100 distinct x values: long runs of equal keys whenever x is the split axis
        if kind != "grid":
            q[:, :3] += rng.normal(0, 0.3, (300, 3)).astype(np.float32)
        gpu_ctx.set_reference(ref)
        gpu_ctx.set_reading(q)
        for k in (1, 7):
            ib, db = oracle.bruteforce_knn(ref, q, k, np.inf, nthreads=8)
            ig, dg, _ = gpu_ctx.knn(None, k, 0.0, np.inf)
            assert (ib == ig).all() and (bits(db) == bits(dg)).all(), (kind, n, k)
    # the same cloud builds the same structure: two contexts agree on the leaf order (sharded map normals rely on it)
    from libpointmatcher_b200 import capi
    ref = cloud(rng, 50000, kind)
    with capi.Context(0) as other:
        other.set_reference(ref)
        other.ref_compute_normals(knn=8)
        gpu_ctx.set_reference(ref)
        gpu_ctx.ref_compute_normals(knn=8)
        assert (bits(other.ref_normals()) == bits(gpu_ctx.ref_normals())).all()


def test_knn_fused_transform(gpu_ctx, oracle):
    rng = np.random.default_rng(101)
    ref, q = cloud(rng, 50000, "uniform"), cloud(rng, 4000, "uniform")
    gpu_ctx.set_reference(ref)
    gpu_ctx.set_reading(q)
    for _ in range(3):
        T = small_pose(rng)
        qt = oracle.rigid_transform(T, q)
        ib, db = oracle.bruteforce_knn(ref, qt, 3, nthreads=8)
        ig, dg, _ = gpu_ctx.knn(T, 3)
        assert (ib == ig).all() and (bits(db) == bits(dg)).all()


def test_knn_lidar_vs_kdtree_oracle(gpu_ctx, oracle, synth):
    rd, rf, _ = synth.scan_pair(200000)
    tree = oracle.KdTree(rf)
    gpu_ctx.set_reference(rf)
    gpu_ctx.set_reading(rd)
    for k, md in ((1, np.inf), (10, 2.0)):
        ik, dk = tree.knn(rd, k, max_dist=md, nthreads=8)
        ig, dg, visits = gpu_ctx.knn(None, k, 0.0, md)
        assert (bits(dk) == bits(dg)).all()
        ndiff, nties = classify_id_mismatches(ik, dk, ig, dg)
        assert ndiff == nties, "%d id mismatches that are not exact ties" % (ndiff - nties)
        assert visits > 0


def test_knn_errors(gpu_ctx):
    from libpointmatcher_b200 import capi
    rng = np.random.default_rng(102)
    ref = cloud(rng, 10)
    gpu_ctx.set_reference(ref)
    gpu_ctx.set_reading(ref)
    with pytest.raises(capi.PmGpuError) as e:
        gpu_ctx.knn(None, 11)
    assert e.value.code == capi.ERR_KNN_TOO_LARGE
    T = np.eye(4, dtype=np.float32)
    T[0, 0] = 1.01
    with pytest.raises(capi.PmGpuError) as e:
        gpu_ctx.knn(T, 1)
    assert e.value.code == capi.ERR_NOT_ORTHOGONAL
    with pytest.raises(capi.PmGpuError) as e:
        gpu_ctx.set_reference(np.ones((5, 5), np.float32))  # neither 2-D nor 3-D: explicit error, never silently wrong
    assert e.value.code == capi.ERR_UNSUPPORTED
    # empty reading is fine
    gpu_ctx.set_reference(ref)
    gpu_ctx.set_reading(np.zeros((0, 4), np.float32))
    ids, d, _ = gpu_ctx.knn(None, 1)
    assert ids.shape == (0, 1)


# ---------------------------------------------------------------------------------- K3
def test_quantile_and_weights_bit_exact(gpu_ctx, oracle, synth):
    rd, rf, _ = synth.scan_pair(100000)
    gpu_ctx.set_reference(rf)
    gpu_ctx.set_reading(rd)
    for k, md in ((1, np.inf), (3, 0.7)):
        ids, d, _ = gpu_ctx.knn(None, k, 0.0, md)
        chains = [
            [],
            [(0, 0.5)],
            [(2, 0.75)], [(2, 0.85)], [(2, 1.0)], [(2, 1e-7)], [(2, 0.3333333)],
            [(1, 3.0)], [(1, 0.5)],
            [(0, 1.0), (1, 3.0)], [(2, 0.9), (0, 0.3), (1, 2.0)],
        ]
        for chain in chains:
            wo, lo = oracle.outlier_weights(d, chain)
            wg, lg = gpu_ctx.weights(chain)
            assert (bits(lo) == bits(lg)).all(), (chain, lo, lg)
            assert (wo == wg).all(), chain


def test_quantile_ties_and_errors(gpu_ctx, oracle):
    from libpointmatcher_b200 import capi
    # the {4,5,5,5,5} vector of utest/ui/Outliers.cpp:126-152, via a 5-point cloud on a line
    ref = np.array([[0, 0, 0, 1]], np.float32)
    rd = np.array([[2, 0, 0, 1], [0, np.sqrt(5), 0, 1], [0, 0, np.sqrt(5), 1], [np.sqrt(5), 0, 0, 1], [0, -np.sqrt(5), 0, 1]], np.float32)
    gpu_ctx.set_reference(ref)
    gpu_ctx.set_reading(rd)
    ids, d, _ = gpu_ctx.knn(None, 1)
    for ratio in (0.19, 0.2, 0.5, 1.0):
        wo, lo = oracle.outlier_weights(d, [(2, ratio)])
        wg, lg = gpu_ctx.weights([(2, ratio)])
        assert (bits(lo) == bits(lg)).all() and (wo == wg).all()
    # all matches beyond maxDist -> "no outlier to filter" (Matches.cpp:76-77)
    gpu_ctx.knn(None, 1, 0.0, 0.5)
    with pytest.raises(capi.PmGpuError) as e:
        gpu_ctx.weights([(2, 0.85)])
    assert e.value.code == capi.ERR_NO_OUTLIER_TO_FILTER
    # ... and the empty chain then leaves nothing to minimise (ErrorMinimizer.cpp:76-77)
    gpu_ctx.weights([])
    with pytest.raises(capi.PmGpuError) as e:
        gpu_ctx.minimize(0)
    assert e.value.code == capi.ERR_NO_POINT_TO_MINIMIZE


# ---------------------------------------------------------------------------------- K4-K7
@pytest.mark.parametrize("k", [1, 4])
def test_minimizers_match_oracle(gpu_ctx, oracle, synth, k):
    rd, rf, _ = synth.scan_pair(100000)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    gpu_ctx.set_reference(rf, nrm)
    gpu_ctx.set_reading(rd)
    T0 = small_pose(np.random.default_rng(103), 0.05, 0.01)
    ids, d, _ = gpu_ctx.knn(T0, k)
    rdt = oracle.rigid_transform(T0, rd)
    chain = [(2, 0.8), (0, 1.5)]
    w, _ = gpu_ctx.weights(chain)
    for mini in (0, 1, 2, 3):
        Tg, covg, sg = gpu_ctx.minimize(mini, 0.02)
        # The bar (1e-5 rad / 1e-5 m) is checked against the oracle with fp64 sums.  The
        # float-faithful oracle accumulates ~1e5..1e6 terms *sequentially* in float, which alone
        # costs up to ~1e-3 m on un-centred clouds (Eigen's blocked reductions lose less), so it
        # can only confirm that the GPU result lies within that oracle's own rounding noise.
        To, covo, so = oracle.minimize(mini, rdt, rf, nrm, ids, d, w, 0.02, acc_double=True)
        Tf, _, sf = oracle.minimize(mini, rdt, rf, nrm, ids, d, w, 0.02, acc_double=False)
        assert_transform_close(Tg, To, 1e-5, 1e-5)
        noise_t = float(np.linalg.norm(To[:3, 3].astype(np.float64) - Tf[:3, 3]))
        assert float(np.linalg.norm(Tg[:3, 3].astype(np.float64) - Tf[:3, 3])) <= noise_t + 1e-5
        for s_ in (so, sf):
            assert sg["nbKept"] == s_["nbKept"] and sg["nbRejectedMatches"] == s_["nbRejectedMatches"]
            assert sg["nbRejectedPoints"] == s_["nbRejectedPoints"]
            assert abs(sg["pointUsedRatio"] - s_["pointUsedRatio"]) < 1e-6
            assert abs(sg["weightedPointUsedRatio"] - s_["weightedPointUsedRatio"]) < 1e-5
        if mini == 3:
            # H^-1 is ill-conditioned in float -> relative tolerance
            assert np.allclose(covg, covo, rtol=2e-3, atol=1e-3 * np.abs(covo).max()), (covg, covo)
        if mini == 2:
            # PointToPointWithCov uses the pseudo-normal (1,1,1) (PointToPointWithCov.cpp:77): the
            # first three rows of J_hessian are identical, the matrix is singular and its inverse
            # (hence the reference's covariance) is not finite.  Same on the GPU.
            assert not np.isfinite(covo).all() and not np.isfinite(covg).all()


def test_point_to_plane_without_normals_is_invalid_field(gpu_ctx, oracle):
    from libpointmatcher_b200 import capi
    rng = np.random.default_rng(104)
    ref = cloud(rng, 1000)
    gpu_ctx.set_reference(ref)
    gpu_ctx.set_reading(ref)
    gpu_ctx.knn(None, 1)
    with pytest.raises(capi.PmGpuError) as e:
        gpu_ctx.minimize(1)
    assert e.value.code == capi.ERR_NO_NORMALS


def test_singular_plane_minimum_norm_solution(gpu_ctx):
    """icpSingular (utest/utest.cpp:162-198) on the GPU modules: rank-deficient normal matrix"""
    pts = np.array([[i, j, 0, 1] for i in range(-5, 5) for j in range(-5, 5)], np.float32)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (100, 1))
    shifted = pts.copy()
    shifted[:, 2] += 1.0
    gpu_ctx.set_reference(shifted, nrm)
    gpu_ctx.set_reading(pts)
    gpu_ctx.knn(None, 1)
    gpu_ctx.weights([(2, 0.85)])
    T, _, _ = gpu_ctx.minimize(1)
    expect = np.eye(4, dtype=np.float32)
    expect[2, 3] = 1.0
    assert np.allclose(T, expect, atol=1e-6)


# ---------------------------------------------------------------------------------- K8
def test_surface_normals_match_oracle(gpu_ctx, oracle, synth):
    rf = synth.scan(60000, cache=False)
    for knn in (5, 20):
        o = oracle.surface_normals(rf, knn=knn, nthreads=8, sort_eigen=True)
        g = gpu_ctx.normals(rf, knn=knn, sort_eigen=True, keep=("normals", "densities", "eigValues", "eigVectors", "matchedIds", "meanDists"))
        mism = g["matchedIds"].astype(np.int32) != o["ids"]
        same = ~mism.any(axis=1)  # points whose neighbourhood is identical (ties may reorder a few)
        assert same.mean() > 0.999
        well = same & (o["gap"] > 1e-3)
        dots = np.abs((g["normals"] * o["normals"]).sum(1))
        assert (1.0 - dots[well]).max() <= 1e-5, (1.0 - dots[well]).max()
        assert np.allclose(g["densities"][same, 0], o["densities"][same], rtol=1e-5)
        assert np.allclose(g["meanDists"][same, 0], o["meanDists"][same], rtol=1e-4, atol=1e-6)
        scale = o["eigValues"][same].max(axis=1, keepdims=True)
        assert (np.abs(g["eigValues"][same] - o["eigValues"][same]) <= 2e-5 * scale + 1e-9).all()
        assert g["degenerate"] == o["degenerate"]


def test_surface_normals_degenerate_points(gpu_ctx, oracle):
    # an exact line: scatter matrix of rank 1 -> degenerate -> zero normal, zero density
    line = np.array([[i, 0, 0, 1] for i in range(50)], np.float32)
    o = oracle.surface_normals(line, knn=5)
    g = gpu_ctx.normals(line, knn=5, keep=("normals", "densities"))
    assert g["degenerate"] == o["degenerate"] == 50
    assert (g["normals"] == 0).all() and (g["densities"] == 0).all()


# ---------------------------------------------------------------------------------- ICP loop
def _icp_both(oracle, synth, n, minimizer, filters, k=1, max_dist=np.inf, iters=15, differential=None, T_init=None):
    from libpointmatcher_b200 import pm
    rd, rf, T_gt = synth.scan_pair(n)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    res_o = oracle.icp(rd, rf, ref_normals=nrm, T_init=T_init, knn=k, max_dist=max_dist, filters=filters, minimizer=minimizer,
                       max_iterations=iters, differential=differential, nthreads=8, acc_double=True)
    icp = pm.ICP()
    icp.matcher = pm.KDTreeMatcher({"knn": str(k), "maxDist": str(max_dist)})
    names = {0: ("MaxDistOutlierFilter", "maxDist"), 1: ("MedianDistOutlierFilter", "factor"), 2: ("TrimmedDistOutlierFilter", "ratio")}
    icp.outlierFilters = pm.OutlierFilters([pm.OutlierFilterRegistrar.create(names[t][0], {names[t][1]: repr(float(v))}) for t, v in filters])
    icp.errorMinimizer = [pm.PointToPointErrorMinimizer, pm.PointToPlaneErrorMinimizer, pm.PointToPointWithCovErrorMinimizer,
                          pm.PointToPlaneWithCovErrorMinimizer][minimizer]()
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": str(iters)})]
    if differential:
        icp.transformationCheckers.append(pm.DifferentialTransformationChecker(
            {"minDiffRotErr": repr(differential[0]), "minDiffTransErr": repr(differential[1]), "smoothLength": str(differential[2])}))
    T = icp(pm.DataPoints(rd), pm.DataPoints(rf, {"normals": nrm}), T_init)
    return T, icp, res_o, T_gt


@pytest.mark.parametrize("minimizer", [0, 1])
def test_icp_fixed_iterations_matches_oracle(oracle, synth, minimizer):
    T, icp, res_o, T_gt = _icp_both(oracle, synth, 100000, minimizer, [(2, 0.75)], iters=15)
    assert icp.iterationCount == res_o["iterations"] == 15
    assert_transform_close(T, res_o["T"], 1e-5, 1e-5)


def test_icp_knn10_median_maxdist_cov(oracle, synth):
    """config 4 shape: knn 10, maxDist + MedianDist, PointToPlaneWithCov"""
    T, icp, res_o, _ = _icp_both(oracle, synth, 50000, 3, [(0, 1.0), (1, 3.0)], k=10, max_dist=2.0, iters=8)
    assert_transform_close(T, res_o["T"], 1e-5, 1e-5)
    co, cg = res_o["cov"], icp.errorMinimizer.getCovariance()
    assert np.allclose(cg, co, rtol=5e-3, atol=1e-3 * np.abs(co).max())


def test_icp_differential_checker_stops_like_oracle(oracle, synth):
    T, icp, res_o, T_gt = _icp_both(oracle, synth, 100000, 1, [(2, 0.85)], iters=40, differential=(0.001, 0.001, 3))
    assert icp.iterationCount == res_o["iterations"] < 40
    assert_transform_close(T, res_o["T"], 1e-5, 1e-5)
    assert_transform_close(T, T_gt.astype(np.float32), 5e-3, 2e-2)  # and it is the right registration


def test_icp_with_initial_guess(oracle, synth):
    Ti = synth.pose_matrix((0.5, -0.3, 0.0), 3.0).astype(np.float32)
    T, icp, res_o, _ = _icp_both(oracle, synth, 50000, 1, [(2, 0.8)], iters=10, T_init=Ti)
    assert_transform_close(T, res_o["T"], 1e-5, 1e-5)


def test_icp_sequence_resident_map(oracle, synth):
    """ICPSequence (ICP.cpp:455-609): the map is indexed once; registering against it gives exactly
    what ICP::operator() gives, before and after other registrations"""
    from libpointmatcher_b200 import pm
    rd, rf, _ = synth.scan_pair(50000)
    rd2, _, _ = synth.scan_pair(50000, pair_seed=3)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]

    def chain(obj):
        obj.matcher = pm.KDTreeMatcher()
        obj.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.8"})])
        obj.errorMinimizer = pm.PointToPlaneErrorMinimizer()
        obj.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "10"})]
        return obj

    ref = pm.DataPoints(rf, {"normals": nrm})
    plain = chain(pm.ICP())
    T_plain = plain(pm.DataPoints(rd), ref)
    seq = chain(pm.ICPSequence())
    assert not seq.hasMap()
    assert (seq(pm.DataPoints(rd)) == np.eye(4)).all()      # no map yet: identity (ICP.cpp:598-604)
    assert seq.setMap(ref) and seq.hasMap()
    T1 = seq(pm.DataPoints(rd))
    seq(pm.DataPoints(rd2))                                   # another reading in between
    T3 = seq(pm.DataPoints(rd))
    assert (T1 == T_plain).all() and (T3 == T_plain).all()
    plain.ctx.close()
    seq.ctx.close()


# ---------------------------------------------------------------------------------- capped matching
def _fused_run(monkeypatch, env, rd, rf, nrm, params_kw, T0=None):
    """One fused registration through the C ABI in a fresh context created under `env`."""
    from libpointmatcher_b200 import capi
    for k in ("PMGPU_NO_CAP", "PMGPU_CAP_MARGIN"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    with capi.Context(0) as ctx:
        ctx.set_reference(rf, normals=nrm)
        ctx.set_reading(rd)
        return ctx.icp_run(capi.make_params(**params_kw), T0)


@pytest.mark.parametrize("chain", [
    dict(knn=1, filters=[(2, 0.75)], minimizer=0, max_iterations=12),
    dict(knn=1, filters=[(2, 0.9)], minimizer=1, max_iterations=12),
    dict(knn=3, filters=[(1, 2.5)], minimizer=1, max_iterations=8),
    dict(knn=1, filters=[(0, 0.4), (1, 3.0)], minimizer=3, max_iterations=8),
    dict(knn=1, filters=[(0, 0.05)], minimizer=0, max_iterations=6),
    dict(knn=1, filters=[(6, 0.02), (2, 0.8)], minimizer=1, max_iterations=8),      # MinDist x TrimmedDist
])
def test_capped_loop_is_bit_identical_to_uncapped(monkeypatch, oracle, synth, chain):
    """The adaptive search radius of the fused loop (pmgpu.h "capped matching") must not change a
    single bit of what the loop returns: T, iteration count, statistics, covariance."""
    rd, rf, _ = synth.scan_pair(60000)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    plain = _fused_run(monkeypatch, {"PMGPU_NO_CAP": "1"}, rd, rf, nrm, chain)
    capped = _fused_run(monkeypatch, {}, rd, rf, nrm, chain)
    assert plain["cap_redos"] == 0
    assert capped["iterations"] == plain["iterations"] == chain["max_iterations"]
    assert (bits(capped["T_iter"]) == bits(plain["T_iter"])).all()
    assert capped["stats"] == plain["stats"]
    assert (bits(capped["cov"]) == bits(plain["cov"])).all()


def test_cap_violation_voids_the_slot_and_is_repaired(monkeypatch, oracle, synth):
    """A cap that is always too small (margin < 1) is detected by the select kernels every time it
    is used: the slot is void, the next one runs uncapped, and the result is still exact."""
    rd, rf, _ = synth.scan_pair(60000)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    chain = dict(knn=1, filters=[(2, 0.8)], minimizer=1, max_iterations=9)
    plain = _fused_run(monkeypatch, {"PMGPU_NO_CAP": "1"}, rd, rf, nrm, chain)
    tight = _fused_run(monkeypatch, {"PMGPU_CAP_MARGIN": "0.5"}, rd, rf, nrm, chain)
    assert tight["cap_redos"] >= 4
    assert tight["iterations"] == plain["iterations"] == 9
    assert (bits(tight["T_iter"]) == bits(plain["T_iter"])).all()
    assert tight["stats"] == plain["stats"]


def test_cap_not_used_with_finite_maxdist_or_staged_calls(monkeypatch, oracle, synth):
    """A finite matcher maxDist decides which matches count as missing (infinite distance, left out
    of the quantile population), so the cap must stay off; and staged calls return exact matches."""
    from libpointmatcher_b200 import capi
    rd, rf, _ = synth.scan_pair(30000)
    chain = dict(knn=2, max_dist=0.3, filters=[(2, 0.7)], minimizer=0, max_iterations=6)
    plain = _fused_run(monkeypatch, {"PMGPU_NO_CAP": "1"}, rd, rf, None, chain)
    capped = _fused_run(monkeypatch, {"PMGPU_CAP_MARGIN": "0.01"}, rd, rf, None, chain)
    assert capped["cap_redos"] == 0
    assert (bits(capped["T_iter"]) == bits(plain["T_iter"])).all() and capped["stats"] == plain["stats"]
    monkeypatch.delenv("PMGPU_CAP_MARGIN", raising=False)
    with capi.Context(0) as ctx:      # after a fused run, staged matching is still exact and uncapped
        ctx.set_reference(rf)
        ctx.set_reading(rd)
        ctx.icp_run(capi.make_params(knn=1, filters=[(2, 0.5)], minimizer=0, max_iterations=5))
        ids, dists, _ = ctx.knn(None, 1, 0.0, np.inf)
        ib, db = oracle.bruteforce_knn(rf, rd, 1, np.inf, nthreads=8)
        assert (bits(dists) == bits(db)).all() and (ids >= 0).all()
        w, lim = ctx.weights([(2, 0.5)])
        T, _, _ = ctx.minimize(0)
        assert np.isfinite(T).all()


# ---------------------------------------------------------------------------------- RobustOutlierFilter (8f row 3)
@pytest.mark.parametrize("fct", ["cauchy", "welsch", "sc", "gm", "tukey", "huber", "L1", "student"])
def test_robust_filter_weights_match_oracle(gpu_ctx, oracle, synth, fct):
    """scale = sqrt(MAD) bit-exact (two exact selects), weights in float like Eigen's arrays"""
    rd, rf, _ = synth.scan_pair(40000)
    gpu_ctx.set_reference(rf)
    gpu_ctx.set_reading(rd)
    ids, dists, _ = gpu_ctx.knn(None, 3, 0.0, np.inf)
    for scale in (oracle.SCALE_MAD, oracle.SCALE_NONE):
        word, tuning = oracle.robust_word(fct, scale), 1.3
        wo, so = oracle.outlier_weights(dists, [(word, tuning)])
        wg, sg = gpu_ctx.weights([(word, tuning)])
        assert bits(sg)[0] == bits(so)[0], (fct, scale, sg, so)
        if fct in ("cauchy", "sc", "gm", "tukey"):
            assert (bits(wg) == bits(wo)).all()
        else:                       # exp / pow / sqrt differ from libm in the last places
            assert np.allclose(wg, wo, rtol=2e-6, atol=1e-30)
    # in a chain the weights multiply (OutlierFilter.cpp:96)
    chain = [(2, 0.8), (oracle.robust_word(fct, oracle.SCALE_MAD), 0.7)]
    wo, _ = oracle.outlier_weights(dists, chain)
    wg, _ = gpu_ctx.weights(chain)
    assert np.allclose(wg, wo, rtol=2e-6, atol=1e-30) and ((wg == 0) == (wo == 0)).all()


@pytest.mark.parametrize("minimizer,k", [(0, 1), (0, 4), (1, 2)])
def test_icp_with_robust_filter_matches_oracle(oracle, synth, minimizer, k):
    from libpointmatcher_b200 import capi
    rd, rf, _ = synth.scan_pair(50000)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    word = oracle.robust_word("cauchy", oracle.SCALE_MAD)
    res_o = oracle.icp(rd, rf, ref_normals=nrm, knn=k, filters=[(word, 1.0)], minimizer=minimizer, max_iterations=12, nthreads=8, acc_double=True)
    with capi.Context(0) as ctx:
        ctx.set_reference(rf, normals=nrm)
        ctx.set_reading(rd)
        res = ctx.icp_run(capi.make_params(knn=k, filters=[(word, 1.0)], minimizer=minimizer, max_iterations=12))
    assert res["iterations"] == res_o["iterations"] == 12 and res["cap_redos"] == 0
    assert_transform_close(res["T_iter"], res_o["T"], 1e-5, 1e-5)
    assert abs(res["stats"]["weightedPointUsedRatio"] - float(res_o["stats"][1])) < 1e-4


def test_robust_filter_scale_frozen_after_n_iterations(oracle, synth):
    """nbIterationForScale: the scale is re-estimated in the first n calls only (OutlierFiltersImpl.cpp:510-514)"""
    from libpointmatcher_b200 import capi
    rd, rf, _ = synth.scan_pair(30000)
    word = oracle.robust_word("huber", oracle.SCALE_MAD, 2)
    res_o = oracle.icp(rd, rf, knn=1, filters=[(word, 1.0)], minimizer=0, max_iterations=8, nthreads=8, acc_double=True)
    with capi.Context(0) as ctx:
        ctx.set_reference(rf)
        ctx.set_reading(rd)
        res = ctx.icp_run(capi.make_params(knn=1, filters=[(word, 1.0)], minimizer=0, max_iterations=8))
    assert_transform_close(res["T_iter"], res_o["T"], 1e-5, 1e-5)


@pytest.mark.parametrize("est", ["berg", "std"])
@pytest.mark.parametrize("fct", ["cauchy", "tukey", "welsch"])
def test_robust_filter_berg_and_std_scale_estimators(oracle, synth, est, fct):
    """scaleEstimator "berg" (1.9 sqrt(median) at the first call, then 0.85 (scale - target) + target, Bergstrom's tuning constants)
    and "std" (sqrt of the standard deviation of all distances), OutlierFiltersImpl.cpp:420-432, 516-537; with `approximation`"""
    from libpointmatcher_b200 import capi
    rd, rf, _ = synth.scan_pair(40000)
    scale = oracle.SCALE_BERG if est == "berg" else oracle.SCALE_STD
    word, tuning = oracle.robust_word(fct, scale), 0.6
    for approx in (float("inf"), 0.9):
        oracle.set_robust_approximation(approx)
        try:
            with capi.Context(0) as ctx:   # a fresh context = a fresh filter object (iteration 1)
                ctx.set_reference(rf)
                ctx.set_reading(rd)
                ids, dists, _ = ctx.knn(None, 2, 0.0, np.inf)
                ctx.set_robust_approximation(approx)
                wo, so = oracle.outlier_weights(dists, [(word, tuning)])
                wg, sg = ctx.weights([(word, tuning)])
        finally:
            oracle.set_robust_approximation(float("inf"))
        if est == "berg":
            assert bits(sg)[0] == bits(so)[0], (fct, sg, so)   # an exact select and one rounding
        else:
            assert abs(sg[0] - so[0]) <= 2e-6 * abs(so[0])      # two fp64 sums against the restatement's
        assert np.allclose(wg, wo, rtol=2e-5, atol=1e-30) and ((wg == 0) == (wo == 0)).mean() > 0.9999
        if approx != float("inf"):
            assert (wg == 0).any() and (wg != 0).any()


def test_robust_filter_point_to_plane_distance(oracle, synth):
    """distanceType point2plane (computePointToPlaneDistance, OutlierFiltersImpl.cpp:468-500): the weight function sees
    dot(n / |n|, p - q)^2, the scale estimator still the match distances; one evaluation and a fused ICP run against the oracle"""
    from libpointmatcher_b200 import capi
    rd, rf, _ = synth.scan_pair(40000)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    word = oracle.robust_word("cauchy", oracle.SCALE_MAD) | oracle.ROBUST_P2PLANE
    with capi.Context(0) as ctx:
        ctx.set_reference(rf, normals=nrm)
        ctx.set_reading(rd)
        ids, dists, _ = ctx.knn(None, 2, 0.0, np.inf)
        wo, so = oracle.outlier_weights_geom(dists, ids, [(word, 1.0)], rd, rf, nrm)
        wg, sg = ctx.weights([(word, 1.0)])
        assert bits(sg)[0] == bits(so)[0]
        assert np.allclose(wg, wo, rtol=2e-5, atol=1e-30)
        w_p2p, _ = ctx.weights([(oracle.robust_word("cauchy", oracle.SCALE_MAD), 1.0)])
        assert np.abs(wg - w_p2p).max() > 0.1                       # it is a different distance
        for minimizer in (0, 1):
            res_o = oracle.icp(rd, rf, ref_normals=nrm, knn=1, filters=[(word, 1.0)], minimizer=minimizer, max_iterations=8, nthreads=8, acc_double=True)
            ctx.set_reading(rd)
            res = ctx.icp_run(capi.make_params(knn=1, filters=[(word, 1.0)], minimizer=minimizer, max_iterations=8))
            assert res["iterations"] == res_o["iterations"] == 8
            assert_transform_close(res["T_iter"], res_o["T"], 1e-5, 1e-5)
    with capi.Context(0) as ctx:                                    # without reference normals: the reference's InvalidField
        ctx.set_reference(rf)
        ctx.set_reading(rd)
        ctx.knn(None, 1, 0.0, np.inf)
        with pytest.raises(capi.PmGpuError) as e:
            ctx.weights([(word, 1.0)])
        assert e.value.code == capi.ERR_NO_NORMALS


def test_icp_with_berg_scale_matches_oracle(oracle, synth):
    """the berg scale decays from iteration to iteration inside the fused loop exactly as in the oracle's filter object"""
    from libpointmatcher_b200 import capi
    rd, rf, _ = synth.scan_pair(40000)
    word = oracle.robust_word("cauchy", oracle.SCALE_BERG, 0)
    res_o = oracle.icp(rd, rf, knn=1, filters=[(word, 0.1)], minimizer=0, max_iterations=8, nthreads=8, acc_double=True)
    with capi.Context(0) as ctx:
        ctx.set_reference(rf)
        ctx.set_reading(rd)
        res = ctx.icp_run(capi.make_params(knn=1, filters=[(word, 0.1)], minimizer=0, max_iterations=8))
    assert res["iterations"] == res_o["iterations"] == 8
    assert_transform_close(res["T_iter"], res_o["T"], 1e-5, 1e-5)


# ---------------------------------------------------------------------------------- force4DOF (8f row 3)
def test_point_to_plane_force4dof_matches_oracle(gpu_ctx, oracle, synth):
    """PointToPlaneErrorMinimizer force4DOF (PointToPlane.cpp:203-214, 266-281): rotation about z + translation,
    one minimiser call and a whole ICP run against the oracle"""
    from libpointmatcher_b200 import capi, pm
    rd, rf, _ = synth.scan_pair(50000)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    gpu_ctx.set_reference(rf, normals=nrm)
    gpu_ctx.set_reading(rd)
    ids, dists, _ = gpu_ctx.knn(None, 1, 0.0, np.inf)
    w, _ = gpu_ctx.weights([(2, 0.8)])
    Tg, _, _ = gpu_ctx.minimize(capi.MIN_P2PLANE | capi.MIN_FORCE4DOF)
    To, _, _ = oracle.minimize(oracle.MIN_P2PLANE | oracle.MIN_FORCE4DOF, rd, rf, nrm, ids, dists, w, acc_double=True)
    assert_transform_close(Tg, To, 1e-5, 1e-5)
    assert Tg[2, 0] == 0 and Tg[2, 1] == 0 and Tg[0, 2] == 0 and Tg[1, 2] == 0 and abs(Tg[2, 2] - 1) < 1e-6   # a yaw-only rotation
    T6, _, _ = gpu_ctx.minimize(capi.MIN_P2PLANE)
    assert np.abs(T6[:3, :3] - Tg[:3, :3]).max() > 1e-7                                                       # and not the 6-DOF answer
    res_o = oracle.icp(rd, rf, ref_normals=nrm, filters=[(2, 0.8)], minimizer=oracle.MIN_P2PLANE | oracle.MIN_FORCE4DOF, max_iterations=10,
                       nthreads=8, acc_double=True)
    icp = pm.ICP()
    icp.matcher = pm.KDTreeMatcher()
    icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.8"})])
    icp.errorMinimizer = pm.PointToPlaneErrorMinimizer({"force4DOF": "1"})
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "10"})]
    T = icp(pm.DataPoints(rd), pm.DataPoints(rf, {"normals": nrm}))
    icp.ctx.close()
    assert_transform_close(T, res_o["T"], 1e-5, 1e-5)
    with pytest.raises(pm.ConfigurationError):
        pm.PointToPlaneErrorMinimizer({"force2D": "1", "force4DOF": "1"})
    with pytest.raises(capi.PmGpuError):
        gpu_ctx.minimize(capi.MIN_P2POINT | capi.MIN_FORCE4DOF)


# ---------------------------------------------------------------------------------- PointToPointSimilarity (8f row 3)
def test_similarity_minimizer_matches_oracle(gpu_ctx, oracle, synth):
    from libpointmatcher_b200 import capi
    rd, rf, _ = synth.scan_pair(50000)
    rd = rd.copy()
    rd[:, :3] *= np.float32(1.03)                       # a reading that really needs a scale
    gpu_ctx.set_reference(rf)
    gpu_ctx.set_reading(rd)
    ids, dists, _ = gpu_ctx.knn(None, 2, 0.0, np.inf)
    w, _ = gpu_ctx.weights([(2, 0.8)])
    Tg, _, _ = gpu_ctx.minimize(capi.MIN_P2POINT_SIM)
    To, _, _ = oracle.minimize(oracle.MIN_P2POINT_SIM, rd, rf, None, ids, dists, w, acc_double=True)
    assert np.abs(Tg - To).max() < 2e-5
    s_g = np.cbrt(np.linalg.det(Tg[:3, :3].astype(np.float64)))
    assert 0.9 < s_g < 1.0                              # it shrinks the inflated reading
    res_o = oracle.icp(rd, rf, filters=[(2, 0.8)], minimizer=oracle.MIN_P2POINT_SIM, max_iterations=15, nthreads=8, acc_double=True)
    with capi.Context(0) as ctx:
        ctx.set_reference(rf)
        ctx.set_reading(rd)
        res = ctx.icp_run(capi.make_params(knn=1, filters=[(2, 0.8)], minimizer=capi.MIN_P2POINT_SIM, max_iterations=15))
    assert res["iterations"] == res_o["iterations"] == 15
    assert np.abs(res["T_iter"] - res_o["T"]).max() < 5e-5
    assert abs(np.cbrt(np.linalg.det(res["T_iter"][:3, :3].astype(np.float64))) - 1 / 1.03) < 1.5e-2   # on its way to 1 / 1.03


# ---------------------------------------------------------------------------------- SurfaceNormalOutlierFilter (8f row 3)
def test_surface_normal_outlier_filter_matches_oracle(gpu_ctx, oracle, synth):
    from libpointmatcher_b200 import capi, pm
    rd, rf, _ = synth.scan_pair(40000)
    nq = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    nr = oracle.surface_normals(rd, knn=10, nthreads=8)["normals"]
    gpu_ctx.set_reference(rf, normals=nq)
    gpu_ctx.set_reading(rd)
    gpu_ctx.set_reading_normals(nr)
    T = synth.pose_matrix((0.05, -0.02, 0.01), 2.0).astype(np.float32)
    ids, dists, _ = gpu_ctx.knn(T, 3, 0.0, np.inf)
    nr_rot = (nr @ T[:3, :3].T).astype(np.float32)
    for chain in ([(4, 0.3)], [(2, 0.8), (4, 0.5)], [(4, 1.57)]):
        wo, lo = oracle.outlier_weights_sn(dists, ids, chain, nr_rot, nq)
        wg, lg = gpu_ctx.weights(chain)
        assert (wg != wo).mean() < 2e-4, chain      # |dot| within an ulp of cos(maxAngle) may fall either side (rotation rounding)
        assert 0.02 < (wg == 0).mean() < 0.98 or chain == [(4, 1.57)]
    # without reading normals the filter does nothing (OutlierFiltersImpl.cpp:268-277)
    gpu_ctx.set_reading_normals(None)
    gpu_ctx.knn(T, 3, 0.0, np.inf)
    wg, _ = gpu_ctx.weights([(4, 0.3)])
    assert (wg == 1).all()
    # whole loop: point-to-plane with Trimmed + SurfaceNormal filters, reading normals turning with the reading
    chain = [(2, 0.8), (4, 0.4)]
    res_o = oracle.icp(rd, rf, ref_normals=nq, reading_normals=nr, knn=2, filters=chain, minimizer=1, max_iterations=10, nthreads=8, acc_double=True)
    icp = pm.ICP()
    icp.matcher = pm.KDTreeMatcher({"knn": "2"})
    icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.8"}), pm.SurfaceNormalOutlierFilter({"maxAngle": "0.4"})])
    icp.errorMinimizer = pm.PointToPlaneErrorMinimizer()
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "10"})]
    Tg = icp(pm.DataPoints(rd, {"normals": nr}), pm.DataPoints(rf, {"normals": nq}))
    icp.ctx.close()
    assert_transform_close(Tg, res_o["T"], 2e-5, 2e-5)


# ---------------------------------------------------------------------------------- KDTreeVarDistMatcher (8f row 3)
def test_var_dist_matcher_bit_exact_vs_bruteforce(gpu_ctx, oracle, synth):
    """one maximum search distance per reading point (MatchersImpl.cpp:132-150): ids and dists as libnabo's brute force"""
    from libpointmatcher_b200 import capi, pm
    rd, rf, _ = synth.scan_pair(20000)
    rng = np.random.default_rng(7)
    radii = rng.uniform(0.0, 0.3, len(rd)).astype(np.float32)
    radii[::17] = np.inf
    radii[::23] = 0.0
    gpu_ctx.set_reference(rf)
    gpu_ctx.set_reading(rd)
    gpu_ctx.set_reading_max_dists(radii)
    for k in (1, 5):
        ib, db = oracle.bruteforce_knn_var(rf, rd, k, radii, nthreads=8)
        ig, dg, _ = gpu_ctx.knn(None, k, 0.0, -1.0)
        assert (ib == ig).all() and (bits(db) == bits(dg)).all()
        assert 0.05 < (ig[:, 0] == -1).mean() < 0.95
        ig2, dg2, _ = gpu_ctx.knn(None, k, 0.0, -1.0)          # seeded second call (k = 1): same answer
        assert (ig2 == ig).all() and (bits(dg2) == bits(dg)).all()
    gpu_ctx.set_reading(rd)                                    # a new reading forgets the distances
    with pytest.raises(capi.PmGpuError):
        gpu_ctx.knn(None, 1, 0.0, -1.0)
    # the module: fused loop == staged calls, and a missing descriptor is an InvalidField
    m = pm.KDTreeVarDistMatcher({"knn": "1"})
    m.bind(gpu_ctx)
    with pytest.raises(pm.InvalidField):
        m.findClosests(pm.DataPoints(rd))
    reading = pm.DataPoints(rd, {"maxSearchDist": radii[:, None]})
    mt = m.findClosests(reading)
    ib, db = oracle.bruteforce_knn_var(rf, rd, 1, radii, nthreads=8)
    assert (mt.ids == ib).all()
    icp = pm.ICP()
    icp.matcher = pm.KDTreeVarDistMatcher()
    icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.9"})])
    icp.errorMinimizer = pm.PointToPointErrorMinimizer()
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "6"})]
    wide = pm.DataPoints(rd, {"maxSearchDist": np.full((len(rd), 1), 1e3, np.float32)})
    T_var = icp(wide, pm.DataPoints(rf))
    icp.matcher = pm.KDTreeMatcher()
    T_plain = icp(pm.DataPoints(rd), pm.DataPoints(rf))
    icp.ctx.close()
    assert (bits(T_var) == bits(T_plain)).all()                # radii that never bind: the plain matcher's answer


def test_trailing_normals_filter_runs_on_the_matchers_structure(synth):
    """a chain whose last reference filter is SurfaceNormalDataPointsFilter computes the normals on the structure the matcher needs
    anyway (pmgpu_ref_set -> pmgpu_ref_compute_normals -> pmgpu_ref_center): the same transform, bit for bit, as filtering first"""
    from libpointmatcher_b200 import pm
    rd, rf, _ = synth.scan_pair(50000)

    def chain(filters):
        icp = pm.ICP()
        icp.referenceDataPointsFilters = filters
        icp.matcher = pm.KDTreeMatcher()
        icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.8"})])
        icp.errorMinimizer = pm.PointToPlaneErrorMinimizer()
        icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "8"})]
        return icp

    fused = chain([pm.SurfaceNormalDataPointsFilter({"knn": "12"})])
    T_fused = fused(pm.DataPoints(rd), pm.DataPoints(rf))
    fused.ctx.close()
    filtered = pm.SurfaceNormalDataPointsFilter({"knn": "12"}).filter(pm.DataPoints(rf))
    plain = chain([])
    T_plain = plain(pm.DataPoints(rd), filtered)
    plain.ctx.close()
    assert (bits(T_fused) == bits(T_plain)).all()
    # earlier filters stay on the host: the same chain with the pre-filter applied by hand
    cut = pm.MinDistDataPointsFilter({"minDist": "8"}).filter(pm.DataPoints(rf))
    assert 0 < len(cut.features) < len(rf)
    both = chain([pm.MinDistDataPointsFilter({"minDist": "8"}), pm.SurfaceNormalDataPointsFilter({"knn": "12"})])
    T_both = both(pm.DataPoints(rd), pm.DataPoints(rf))
    both.ctx.close()
    by_hand = chain([pm.SurfaceNormalDataPointsFilter({"knn": "12"})])
    T_hand = by_hand(pm.DataPoints(rd), cut)
    by_hand.ctx.close()
    assert (bits(T_both) == bits(T_hand)).all() and not (bits(T_both) == bits(T_fused)).all()


def test_new_entry_points_reject_misuse(gpu_ctx, synth):
    """argument checking of the row-8f entry points: loud errors, never a silent answer"""
    from libpointmatcher_b200 import capi
    rd, rf, _ = synth.scan_pair(5000)
    with capi.Context(0) as fresh:
        with pytest.raises(capi.PmGpuError) as e:       # no reading yet
            fresh.set_reading_normals(np.zeros((10, 3), np.float32))
        assert e.value.code == capi.ERR_NO_READING
        with pytest.raises(capi.PmGpuError) as e:       # no reference yet
            fresh.ref_center(rf)
        assert e.value.code == capi.ERR_NO_REFERENCE
    gpu_ctx.set_reference(rf)
    gpu_ctx.set_reading(rd)
    with pytest.raises(capi.PmGpuError):                # a different cloud than the resident one
        gpu_ctx.ref_center(rf[:-1])
    gpu_ctx.knn(None, 1, 0.0, np.inf)
    for bad in ([(capi.FILTER_ROBUST | (99 << 8), 1.0)],                       # unknown robust function
                [(capi.FILTER_ROBUST | (5 << 16), 1.0)],                       # scale estimator that is not built
                [(capi.FILTER_ROBUST, 1.0), (capi.FILTER_ROBUST, 2.0)],        # two robust filters
                [(7, 1.0)]):                                                   # unknown filter
        with pytest.raises(capi.PmGpuError):
            gpu_ctx.weights(bad)
    with pytest.raises(capi.PmGpuError):
        gpu_ctx.minimize(capi.MIN_P2POINT_SIM + 1)
    w, _ = gpu_ctx.weights([(2, 0.9)])                  # and the context is still usable afterwards
    assert w.shape == (len(rd), 1) and 0.85 < w.mean() < 0.95


# ---------------------------------------------------------------------------------- force2D (8a row a11)
@pytest.mark.gpu
def test_point_to_plane_force2d_matches_oracle(gpu_ctx, oracle, synth):
    """PointToPlaneErrorMinimizer force2D on 3-D clouds (PointToPlane.cpp:177-186, 294-310): rotation about z and x/y
    translation from the 3x3 system without z; one minimiser call and a whole ICP run against the oracle"""
    from libpointmatcher_b200 import capi, pm
    rd, rf, _ = synth.scan_pair(50000)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    gpu_ctx.set_reference(rf, normals=nrm)
    gpu_ctx.set_reading(rd)
    ids, dists, _ = gpu_ctx.knn(None, 1, 0.0, np.inf)
    w, _ = gpu_ctx.weights([(2, 0.8)])
    Tg, _, _ = gpu_ctx.minimize(capi.MIN_P2PLANE | capi.MIN_FORCE2D)
    To, _, _ = oracle.minimize(oracle.MIN_P2PLANE | oracle.MIN_FORCE2D, rd, rf, nrm, ids, dists, w, acc_double=True)
    assert_transform_close(Tg, To, 1e-5, 1e-5)
    assert (Tg[2] == [0, 0, 1, 0]).all() and (Tg[:, 2] == [0, 0, 1, 0]).all() and (Tg[3] == [0, 0, 0, 1]).all()   # z is left alone
    T4, _, _ = gpu_ctx.minimize(capi.MIN_P2PLANE | capi.MIN_FORCE4DOF)
    assert np.abs(T4[:2, :] - Tg[:2, :]).max() > 1e-7                                                             # not the 4-DOF answer
    res_o = oracle.icp(rd, rf, ref_normals=nrm, filters=[(2, 0.8)], minimizer=oracle.MIN_P2PLANE | oracle.MIN_FORCE2D, max_iterations=10,
                       nthreads=8, acc_double=True)
    icp = pm.ICP()
    icp.matcher = pm.KDTreeMatcher()
    icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.8"})])
    icp.errorMinimizer = pm.PointToPlaneErrorMinimizer({"force2D": "1"})
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "10"})]
    T = icp(pm.DataPoints(rd), pm.DataPoints(rf, {"normals": nrm}))
    icp.ctx.close()
    assert_transform_close(T, res_o["T"], 1e-5, 1e-5)
    assert T[2, 3] == 0 and T[2, 2] == 1
    with pytest.raises(pm.ConfigurationError):
        pm.PointToPlaneWithCovErrorMinimizer({"force2D": "1"})
    for bad in (capi.MIN_P2POINT | capi.MIN_FORCE2D, capi.MIN_P2PLANE_COV | capi.MIN_FORCE2D, capi.MIN_P2PLANE | capi.MIN_FORCE2D | capi.MIN_FORCE4DOF):
        with pytest.raises(capi.PmGpuError):
            gpu_ctx.minimize(bad)


# ---------------------------------------------------------------------------------- VarTrimmedDist (8f row 3)
def test_var_trimmed_reference_known_answer_on_gpu(gpu_ctx):
    """utest/ui/Outliers.cpp:126-152 on the device: squared distances {4,5,5,5,5}"""
    from libpointmatcher_b200 import capi
    ref = np.array([[20.0 * i, 0, 0, 1] for i in range(5)], np.float32)
    rd = ref.copy()
    rd[0, 0] += 2
    rd[1:, 0] += 1
    rd[1:, 1] += 2
    gpu_ctx.set_reference(ref)
    gpu_ctx.set_reading(rd)
    _, dists, _ = gpu_ctx.knn(None, 1, 0.0, np.inf)
    assert dists[:, 0].tolist() == [4, 5, 5, 5, 5]
    gpu_ctx.set_var_trimmed_ratios(0.0000001, 1.0)
    w, lim = gpu_ctx.weights([(capi.FILTER_VARTRIMMEDDIST, 0.0)])
    assert w[:, 0].tolist() == [1, 0, 0, 0, 0] and lim[0] == 4 and gpu_ctx.var_trimmed_ratio() == 0
    w, lim = gpu_ctx.weights([(capi.FILTER_VARTRIMMEDDIST, 1.0)])
    assert w[:, 0].tolist() == [1, 1, 1, 1, 1] and lim[0] == 5 and gpu_ctx.var_trimmed_ratio() == np.float32(0.8)
    with pytest.raises(capi.PmGpuError):
        gpu_ctx.set_var_trimmed_ratios(0.5, 0.5)
    with pytest.raises(capi.PmGpuError):
        gpu_ctx.weights([(capi.FILTER_VARTRIMMEDDIST, 1.0), (capi.FILTER_VARTRIMMEDDIST, 2.0)])
    gpu_ctx.set_var_trimmed_ratios()


@pytest.mark.parametrize("n,k,max_dist,ratios,lam", [(50000, 1, np.inf, (0.05, 0.99), 2.35), (200000, 1, np.inf, (0.3, 0.95), 1.0),
                                                      (30000, 5, 0.8, (0.05, 0.99), 2.35), (1000000, 1, np.inf, (0.05, 0.99), 2.35)])
def test_var_trimmed_ratio_limit_and_weights_match_oracle(gpu_ctx, oracle, synth, n, k, max_dist, ratios, lam):
    """optimizeInlierRatio (OutlierFiltersImpl.cpp:177-218): ratio, limit and weights against the oracle — bit for bit; with a finite
    maxDist some matches are +inf and (k > 1, exact duplicates) some zero, which both sides leave out of the sorted list"""
    from libpointmatcher_b200 import capi
    rd, rf, _ = synth.scan_pair(n)
    rd = rd.copy()
    rd[:50] = rf[:50]                                     # zero distances
    gpu_ctx.set_reference(rf)
    gpu_ctx.set_reading(rd)
    _, dists, _ = gpu_ctx.knn(None, k, 0.0, max_dist)
    gpu_ctx.set_var_trimmed_ratios(*ratios)
    oracle.set_var_trimmed_ratios(*ratios)
    try:
        w, lim = gpu_ctx.weights([(capi.FILTER_VARTRIMMEDDIST, lam), (capi.FILTER_MAXDIST, 3.0)])
        wo, limo = oracle.outlier_weights(dists, [(oracle.FILTER_VARTRIMMEDDIST, lam), (oracle.FILTER_MAXDIST, 3.0)])
        ro = oracle.var_trimmed_ratio(dists, ratios[0], ratios[1], lam)
    finally:
        oracle.set_var_trimmed_ratios()
        rg = gpu_ctx.var_trimmed_ratio()
        gpu_ctx.set_var_trimmed_ratios()
    assert ratios[0] <= rg <= ratios[1]
    assert rg == ro and lim[0] == limo[0] and lim[1] == limo[1]
    assert (w == wo).all() and 0 < w.sum() < w.size


@pytest.mark.parametrize("minimizer", [0, 1])
def test_icp_with_var_trimmed_filter_matches_oracle(oracle, synth, minimizer):
    from libpointmatcher_b200 import pm
    rd, rf, _ = synth.scan_pair(60000)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    oracle.set_var_trimmed_ratios(0.05, 0.99)
    res_o = oracle.icp(rd, rf, ref_normals=nrm, filters=[(oracle.FILTER_VARTRIMMEDDIST, 2.35)], minimizer=minimizer, max_iterations=12, nthreads=8,
                       acc_double=True)
    icp = pm.ICP()
    icp.matcher = pm.KDTreeMatcher()
    icp.outlierFilters = pm.OutlierFilters([pm.VarTrimmedDistOutlierFilter()])
    icp.errorMinimizer = pm.PointToPlaneErrorMinimizer() if minimizer else pm.PointToPointErrorMinimizer()
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "12"})]
    T = icp(pm.DataPoints(rd), pm.DataPoints(rf, {"normals": nrm}))
    ratio = icp.ctx.var_trimmed_ratio()
    icp.ctx.close()
    assert_transform_close(T, res_o["T"], 1e-5, 1e-5)
    assert 0.5 < ratio < 0.99
    with pytest.raises(pm.InvalidParameter):
        pm.VarTrimmedDistOutlierFilter({"minRatio": "0.9", "maxRatio": "0.5"})


# ---------------------------------------------------------------------------------- resident matches on request (8f row 4)
def test_matches_get_after_staged_calls_and_after_the_fused_loop(gpu_ctx, oracle, synth):
    """pmgpu_matches_get: what getErrorElements / the inspectors read after the fact — identical to the staged outputs, and after a
    fused (capped) loop the kept pairs are those of an exact search at T_match with the chain's weights"""
    from libpointmatcher_b200 import capi, pm
    rd, rf, _ = synth.scan_pair(40000)
    gpu_ctx.set_reference(rf)
    gpu_ctx.set_reading(rd)
    ids, dists, _ = gpu_ctx.knn(None, 2, 0.0, np.inf)
    w, _ = gpu_ctx.weights([(capi.FILTER_TRIMMEDDIST, 0.7)])
    ids2, dists2, w2, T = gpu_ctx.matches()
    assert (ids2 == ids).all() and (dists2 == dists).all() and (w2 == w).all() and (T == np.eye(4, dtype=np.float32)).all()
    icp = pm.ICP()
    icp.matcher = pm.KDTreeMatcher()
    icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.7"})])
    icp.errorMinimizer = pm.PointToPointErrorMinimizer()
    icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "6"})]
    icp(pm.DataPoints(rd), pm.DataPoints(rf))
    m, wf, Tm = icp.getMatches()
    stats = icp.errorMinimizer._stats
    assert abs(wf.sum() / wf.size - stats["weightedPointUsedRatio"]) < 1e-6
    # the same search, staged and uncapped, at the same transform: kept pairs agree exactly
    ids_e, dists_e, _ = icp.ctx.knn(Tm, 1, 0.0, np.inf)
    we, _ = icp.ctx.weights([(capi.FILTER_TRIMMEDDIST, 0.7)])
    icp.ctx.close()
    kept = wf[:, 0] != 0
    assert (we == wf).all() and (m.ids[kept] == ids_e[kept]).all() and (m.dists[kept] == dists_e[kept]).all()
    assert ((m.ids[~kept] == ids_e[~kept]) | (m.ids[~kept] == -2)).all()


# ---------------------------------------------------------------------------------- k > 1 starting radius from the previous matches
def test_knn_k_gt_1_seeded_radius_is_exact(gpu_ctx, oracle):
    """consecutive searches of the same reading with the same k: the previous k matches, re-measured, set the starting radius (knn.cu).
    Whatever the change of transform in between, and with empty slots under a finite maxDist, the answer is the brute-force one."""
    rng = np.random.default_rng(202)
    ref, q = cloud(rng, 60000, "uniform"), cloud(rng, 5000, "uniform")   # mean spacing ~0.5: both maxDist values leave some slots empty
    gpu_ctx.set_reference(ref)
    gpu_ctx.set_reading(q)
    poses = [np.eye(4, dtype=np.float32), small_pose(rng), small_pose(rng)]
    big = np.eye(4, dtype=np.float32)
    big[:3, 3] = (3.0, -2.0, 1.0)                                  # far from the previous answer: the radius is loose but valid
    poses.append(big)
    for k, md in ((10, np.inf), (10, 0.6), (4, 0.3)):
        for T in poses:
            qt = oracle.rigid_transform(T, q)
            ib, db = oracle.bruteforce_knn(ref, qt, k, md, nthreads=8)
            ig, dg, _ = gpu_ctx.knn(T, k, 0.0, md)
            assert (ib == ig).all() and (bits(db) == bits(dg)).all(), (k, md)
        if md != np.inf:
            assert (ig == -1).any() and (ig[:, 0] >= 0).any()      # some slots empty: those queries search unseeded next time


# ---------------------------------------------------------------------------------- round-2 review items
@pytest.mark.gpu
def test_differential_only_chain_stops_without_a_counter(oracle, synth):
    """A chain with no CounterTransformationChecker is valid (max_iterations = INT_MAX): the fused loop is enqueued in
    bounded chunks and stops when the Differential checker says so, with the oracle's iteration count."""
    from libpointmatcher_b200 import pm
    rd, rf, _ = synth.scan_pair(30000)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    res_o = oracle.icp(rd, rf, ref_normals=nrm, filters=[(2, 0.8)], minimizer=1, max_iterations=1000, differential=(1e-3, 1e-3, 3), nthreads=8, acc_double=True)
    icp = pm.ICP()
    icp.matcher = pm.KDTreeMatcher()
    icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.8"})])
    icp.errorMinimizer = pm.PointToPlaneErrorMinimizer()
    icp.transformationCheckers = [pm.DifferentialTransformationChecker({"minDiffRotErr": "0.001", "minDiffTransErr": "0.001", "smoothLength": "3"})]
    launches0 = icp.ctx.launch_count
    T = icp(pm.DataPoints(rd), pm.DataPoints(rf, {"normals": nrm}))
    assert icp.iterationCount == res_o["iterations"] < 200
    assert icp.ctx.launch_count - launches0 < 8 * 64 * (icp.iterationCount // 64 + 2)   # bounded: chunks of 64 slots
    icp.ctx.close()
    assert_transform_close(T, res_o["T"], 2e-5, 2e-5)


@pytest.mark.gpu
def test_materialised_weights_after_fused_loop_use_the_match_transform(oracle, synth):
    """After a fused iteration T_iter is already the composed transform; the SurfaceNormalOutlierFilter weights that
    pmgpu_matches_get materialises must turn the reading normals with T_match, the transform the matches (and the
    minimiser) saw."""
    from libpointmatcher_b200 import capi
    rd, rf, _ = synth.scan_pair(30000)
    nq = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    nr = oracle.surface_normals(rd, knn=10, nthreads=8)["normals"]
    chain = [(capi.FILTER_TRIMMEDDIST, 0.8), (capi.FILTER_SURFACENORMAL, 0.4)]
    with capi.Context(0) as ctx:
        ctx.set_reference(rf, nq)
        ctx.set_reading(rd)
        ctx.set_reading_normals(nr)
        p = capi.make_params(knn=2, filters=chain, minimizer=capi.MIN_P2PLANE, max_iterations=3)
        res = ctx.icp_run(p)
        ids, dists, w, T_match = ctx.matches()
        assert not np.array_equal(T_match, res["T_iter"])     # the loop has composed one more increment
        nr_rot = (nr @ T_match[:3, :3].T).astype(np.float32)
        real = ids >= 0
        # capped far matches (id -2, dist FLT_MAX) stay in the quantile population as finite, rejected matches
        wo, _ = oracle.outlier_weights_sn(dists, np.where(real, ids, 0).astype(np.int32), chain, nr_rot, nq)
        # |dot| within an ulp of the limit may fall either side
        assert ((w != wo) & real).mean() < 2e-4
        assert res["stats"]["nbKept"] == int((w != 0).sum())   # what the minimiser used is what is materialised


@pytest.mark.gpu
def test_knn_larger_than_64_is_exact(gpu_ctx, oracle):
    """knn > 64 (MatchersImpl.h:80 allows any unsigned): the warp-per-query kernel with the list distributed over the lanes;
    ids and dists bit-exact against libnabo's brute-force semantics, misses included"""
    rng = np.random.default_rng(21)
    ref, q = cloud(rng, 3000, "uniform"), cloud(rng, 500, "uniform")
    gpu_ctx.set_reference(ref)
    gpu_ctx.set_reading(q)
    for k, md in ((65, np.inf), (100, np.inf), (200, 3.0), (700, np.inf), (1024, 8.0)):
        ib, db = oracle.bruteforce_knn(ref, q, k, md)
        ig, dg, _ = gpu_ctx.knn(None, k, 0.0, md)
        assert (ib == ig).all() and (db.view(np.uint32) == dg.view(np.uint32)).all(), (k, md)
    with pytest.raises(Exception):
        gpu_ctx.knn(None, 1025)


@pytest.mark.gpu
def test_smooth_normals_match_oracle(gpu_ctx, oracle, synth):
    """smoothNormals (SurfaceNormal.cpp:259-283): the in-place, point-after-point average of the neighbours' normals — a serial
    recurrence, run on the host over the device's normals and neighbour ids; equal to the oracle's up to each normal's sign"""
    rf = synth.scan(30000, cache=False)
    o = oracle.surface_normals(rf, knn=8, nthreads=8, smooth_normals=True)
    g = gpu_ctx.normals(rf, knn=8, keep=("normals",), smooth=True)
    no, ng = np.linalg.norm(o["normals"], axis=1), np.linalg.norm(g["normals"], axis=1)
    ok = np.abs(no - ng) <= 1e-4
    cos = np.abs((o["normals"] * g["normals"]).sum(1)) / np.maximum(no * ng, 1e-12)
    ok &= (cos >= 1 - 1e-4) | (no < 1e-3)
    assert ok.mean() > 0.995, ok.mean()     # a tie-reordered neighbourhood or an ill-conditioned normal spreads to its neighbours
    # and the resident-reference variant takes the same path
    gpu_ctx.set_reference(rf)
    gpu_ctx.ref_compute_normals(knn=8, smooth=True)
    assert np.array_equal(gpu_ctx.ref_normals(), g["normals"])


def test_reading_set_sharded_uploads_the_ranks_chunks(gpu_ctx, synth):
    """pmgpu_reading_set_sharded: one strided copy straight from the whole reading = pmgpu_reading_set on dist.shard_take's slice"""
    from libpointmatcher_b200 import dist as pmdist
    rd, rf, _ = synth.scan_pair(30000)
    gpu_ctx.set_reference(rf)
    for n in (len(rd), 4096 * 3, 4096 * 3 + 5, 100):
        for world in (2, 3):
            for rank in range(world):
                gpu_ctx.set_reading_sharded(rd[:n], rank, world, pmdist.SHARD_CHUNK)
                want = pmdist.shard_take(rd[:n], rank, world)
                assert gpu_ctx.nq == len(want)
                if len(want):
                    assert (gpu_ctx.get_reading().view(np.uint32) == np.ascontiguousarray(want).view(np.uint32)).all()


# ---------------------------------------------------------------------------------- Inspector / host checkers of the Python mirror
def test_python_inspector_is_shown_every_iteration(oracle, synth):
    """Inspector::dumpIteration (ICP.cpp:403-405) through the Python mirror: an inspector that overrides it takes the loop off the
    fused path (pmgpu_icp_step: one exactly-matched slot at a time) and is shown, per iteration, T_iter, the centred reference, the
    reading as the iteration sees it, every match and weight; the registration itself is the NullInspector's, bit for bit.  A
    BoundTransformationChecker rides on the same path, now with a WithCov minimiser too."""
    from libpointmatcher_b200 import pm
    rd, rf, _ = synth.scan_pair(40000)
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]

    def chain(inspector=None, extra=()):
        icp = pm.ICP()
        icp.matcher = pm.KDTreeMatcher({"knn": "3"})
        icp.outlierFilters = pm.OutlierFilters([pm.TrimmedDistOutlierFilter({"ratio": "0.8"})])
        icp.errorMinimizer = pm.PointToPlaneWithCovErrorMinimizer()
        icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "12"}),
                                      pm.DifferentialTransformationChecker({"minDiffRotErr": "0.001", "minDiffTransErr": "0.001", "smoothLength": "3"})] + list(extra)
        if inspector is not None:
            icp.inspector = inspector
        return icp

    class Recorder(pm.Inspector):
        def __init__(self):
            pm.Inspector.__init__(self)
            self.calls, self.stats, self.finished, self.inits = [], {}, None, 0

        def init(self):
            self.inits += 1

        def dumpIteration(self, iterationNumber, parameters, filteredReference, reading, matches, outlierWeights, transformationCheckers):
            self.calls.append((iterationNumber, np.array(parameters), filteredReference, reading, matches, np.array(outlierWeights), len(transformationCheckers)))

        def addStat(self, name, data):
            self.stats[name] = data

        def finish(self, iterationCount):
            self.finished = iterationCount

    reading, reference = pm.DataPoints(rd), pm.DataPoints(rf, {"normals": nrm})
    plain = chain()
    T0 = plain(reading, reference)
    cov0 = np.asarray(plain.errorMinimizer.getCovariance())
    plain.ctx.close()
    rec = Recorder()
    assert rec.needsIterationData() and not pm.NullInspector().needsIterationData() and not pm.PerformanceInspector().needsIterationData()
    icp = chain(rec)
    T1 = icp(reading, reference)
    cov1 = np.asarray(icp.errorMinimizer.getCovariance())
    icp.ctx.close()
    assert (T0.view(np.uint32) == T1.view(np.uint32)).all() and plain.iterationCount == icp.iterationCount
    assert np.allclose(cov0, cov1, rtol=1e-5, atol=1e-12)
    assert rec.inits == 1 and rec.finished == icp.iterationCount == len(rec.calls) < 12
    assert rec.stats["IterationsCount"] == icp.iterationCount and abs(rec.stats["OverlapRatio"] - 0.8) < 1e-3
    assert [c[0] for c in rec.calls] == list(range(len(rec.calls))) and (rec.calls[0][1] == np.eye(4, dtype=np.float32)).all()
    tree = None
    for n, T_iter, ref, seen, matches, w, ncheck in rec.calls[:1] + rec.calls[-2:]:
        assert ncheck == 2 and ref.descriptorExists("normals") and w.shape == matches.ids.shape == (len(rd), 3)
        if tree is None:
            tree = oracle.KdTree(ref.features)            # the reference minus its mean: what the matcher was built on
        io, do = tree.knn(seen.features, 3, nthreads=8)
        assert (matches.dists.view(np.uint32) == do.view(np.uint32)).all()      # every match, none cut off by a cap
        ndiff, nties = classify_id_mismatches(io, do, matches.ids, matches.dists)
        assert ndiff == nties
        wo, _ = oracle.outlier_weights(do, [(oracle.FILTER_TRIMMEDDIST, 0.8)])
        assert (w == wo).all()
    # host-side Bound checker: generous bounds change nothing, tight ones stop the registration with the reference's exception
    loose = chain(extra=[pm.BoundTransformationChecker({"maxRotationNorm": "1.0", "maxTranslationNorm": "5.0"})])
    T2 = loose(reading, reference)
    loose.ctx.close()
    assert (T0.view(np.uint32) == T2.view(np.uint32)).all()
    tight = chain(extra=[pm.BoundTransformationChecker({"maxRotationNorm": "0.001", "maxTranslationNorm": "0.001"})])
    with pytest.raises(pm.ConvergenceError):
        tight(reading, reference)
    tight.ctx.close()


def test_null_outlier_filter_is_a_factor_of_one(synth):
    """NullOutlierFilter (OutlierFiltersImpl.cpp:45-58) in a chain: the same registration as without it, bit for bit; alone: the
    empty chain"""
    from libpointmatcher_b200 import pm
    rd, rf, _ = synth.scan_pair(30000)

    def run(filters):
        icp = pm.ICP()
        icp.matcher = pm.KDTreeMatcher()
        icp.outlierFilters = pm.OutlierFilters(filters)
        icp.errorMinimizer = pm.PointToPointErrorMinimizer()
        icp.transformationCheckers = [pm.CounterTransformationChecker({"maxIterationCount": "8"})]
        T = icp(pm.DataPoints(rd), pm.DataPoints(rf))
        ratio = icp.errorMinimizer.getWeightedPointUsedRatio()
        icp.ctx.close()
        return T, ratio

    null = pm.OutlierFilterRegistrar.create("NullOutlierFilter")
    T_a, r_a = run([pm.TrimmedDistOutlierFilter({"ratio": "0.7"})])
    T_b, r_b = run([null, pm.TrimmedDistOutlierFilter({"ratio": "0.7"}), pm.NullOutlierFilter()])
    assert (T_a.view(np.uint32) == T_b.view(np.uint32)).all() and r_a == r_b
    T_c, r_c = run([])
    T_d, r_d = run([pm.NullOutlierFilter()])
    assert (T_c.view(np.uint32) == T_d.view(np.uint32)).all() and r_c == r_d == 1.0


def test_min_dist_outlier_filter_matches_oracle(gpu_ctx, oracle, synth):
    """MinDistOutlierFilter (OutlierFiltersImpl.cpp:87-101): weight 0 below minDist^2, alone and in chains; the whole loop with it.
    A match without a neighbour (dist = +inf) reads 0 on the device and 1 in the reference's matrix — ErrorElements drops it either
    way (ErrorMinimizer.cpp:103-106)"""
    from libpointmatcher_b200 import capi
    rd, rf, _ = synth.scan_pair(60000)
    gpu_ctx.set_reference(rf)
    gpu_ctx.set_reading(rd)
    for k, md in ((1, np.inf), (3, 0.6)):
        ids, d, _ = gpu_ctx.knn(None, k, 0.0, md)
        finite = np.isfinite(d)
        for chain in ([(6, 0.05)], [(6, 0.3), (6, 0.1)], [(6, 0.05), (2, 0.8)], [(0, 0.5), (6, 0.1), (1, 2.0)], [(6, 100.0)]):
            wo, lo = oracle.outlier_weights(d, chain)
            wg, lg = gpu_ctx.weights(chain)
            assert (bits(lo) == bits(lg)).all(), (chain, lo, lg)
            assert (wo[finite] == wg[finite]).all() and (wg[~finite] == 0).all(), chain
            assert 0 < wg.sum() < wg.size or chain == [(6, 100.0)]
    nrm = oracle.surface_normals(rf, knn=10, nthreads=8)["normals"]
    chain = [(6, 0.03), (2, 0.8)]
    for minimizer in (0, 1):
        res_o = oracle.icp(rd, rf, ref_normals=nrm, knn=1, filters=chain, minimizer=minimizer, max_iterations=10, nthreads=8, acc_double=True)
        with capi.Context(0) as ctx:
            ctx.set_reference(rf, normals=nrm)
            ctx.set_reading(rd)
            res = ctx.icp_run(capi.make_params(knn=1, filters=chain, minimizer=minimizer, max_iterations=10))
        assert res["iterations"] == res_o["iterations"] == 10
        assert_transform_close(res["T_iter"], res_o["T"], 1e-5, 1e-5)
        assert abs(res["stats"]["weightedPointUsedRatio"] - float(res_o["stats"][1])) < 1e-4
