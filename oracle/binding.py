"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package.

Conventions: clouds are numpy float32 arrays of shape (N, 4) (C-contiguous), which is the same
memory as the reference's 4 x N column-major Eigen matrix (PointMatcher.h:169,331).  ids / dists
are (N, k) — the same memory as the reference's k x N column-major Matches (PointMatcher.h:373).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

FILTER_MAXDIST, FILTER_MEDIANDIST, FILTER_TRIMMEDDIST, FILTER_ROBUST, FILTER_SURFACENORMAL, FILTER_VARTRIMMEDDIST, FILTER_MINDIST = 0, 1, 2, 3, 4, 5, 6
ROBUST_FCTS = dict(cauchy=0, welsch=1, sc=2, gm=3, tukey=4, huber=5, L1=6, student=7)
SCALE_NONE, SCALE_MAD, SCALE_BERG, SCALE_STD = 0, 1, 2, 3


ROBUST_P2PLANE = 1 << 28   # distanceType point2plane, or-ed into the robust filter word


def robust_word(fct="cauchy", scale=SCALE_MAD, nb_iteration_for_scale=0):
    """filter word of a RobustOutlierFilter (oracle.h): use as the filter type, with the tuning as the parameter"""
    return FILTER_ROBUST | (ROBUST_FCTS[fct] << 8) | (scale << 16) | (nb_iteration_for_scale << 20)
MIN_P2POINT, MIN_P2PLANE, MIN_P2POINT_COV, MIN_P2PLANE_COV, MIN_P2POINT_SIM = 0, 1, 2, 3, 4
MIN_FORCE4DOF = 0x100  # or-ed into a point-to-plane minimizer id
MIN_FORCE2D = 0x200
ERRORS = {
    1: "ConvergenceError: no outlier to filter",
    2: "ConvergenceError: ErrorMnimizer: no point to minimize",
    3: "TransformationError: rotation matrix is not orthogonal",
    4: "ConvergenceError: quantile must be between 0 and 1",
    5: "ConvergenceError: not a number",
    6: "bad argument",
    7: "knn larger than the number of reference points",
}


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(ERRORS.get(code, "oracle error %d" % code))
        self.code = code


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("oracle.cpp", "oracle.h"))
    ):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class IcpConfig(C.Structure):
    _fields_ = [
        ("knn", C.c_int), ("epsilon", C.c_float), ("max_dist", C.c_float), ("search_type", C.c_int),
        ("nfilters", C.c_int), ("filter_type", C.c_int * 8), ("filter_param", C.c_float * 8),
        ("minimizer", C.c_int), ("sensor_std_dev", C.c_float), ("max_iterations", C.c_int),
        ("use_differential", C.c_int), ("min_diff_rot_err", C.c_float), ("min_diff_trans_err", C.c_float),
        ("smooth_length", C.c_int), ("acc_double", C.c_int), ("nthreads", C.c_int),
    ]


_lib = None
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_kdtree_create.restype = C.c_void_p
        _lib.orc_kdtree_create.argtypes = [_fp, C.c_int, C.c_int]
        _lib.orc_kdtree_destroy.argtypes = [C.c_void_p]
        _lib.orc_kdtree_knn.restype = C.c_long
        _lib.orc_kdtree_knn.argtypes = [C.c_void_p, _fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _ip, _fp, C.c_int]
        _lib.orc_bruteforce_knn.restype = C.c_long
        _lib.orc_bruteforce_knn.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int, C.c_float, _ip, _fp, C.c_int]
        _lib.orc_rigid_transform.argtypes = [_fp, _fp, C.c_int, _fp]
        _lib.orc_rotate_normals.argtypes = [_fp, _fp, C.c_int, _fp]
        _lib.orc_dists_quantile.argtypes = [_fp, C.c_long, C.c_float, _fp]
        _lib.orc_set_robust_approximation.argtypes = [C.c_float]
        _lib.orc_set_robust_approximation.restype = None
        _lib.orc_set_var_trimmed_ratios.argtypes = [C.c_float, C.c_float]
        _lib.orc_set_var_trimmed_ratios.restype = None
        _lib.orc_var_trimmed_ratio.argtypes = [_fp, C.c_long, C.c_float, C.c_float, C.c_float, _fp]
        _lib.orc_outlier_weights.argtypes = [_fp, C.c_int, C.c_int, C.c_int, _ip, _fp, _fp, _fp]
        _lib.orc_minimize.argtypes = [C.c_int, _fp, C.c_int, _fp, C.c_int, _fp, _ip, _fp, _fp, C.c_int, C.c_float, C.c_int, _fp, _fp, _fp]
        _lib.orc_surface_normals.argtypes = [_fp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int,
                                             _fp, _fp, _fp, _fp, _fp, _fp, _ip, _fp, _fp, C.POINTER(C.c_int)]
        _lib.orc_icp.argtypes = [_fp, C.c_int, _fp, C.c_int, _fp, _fp, C.POINTER(IcpConfig), _fp, _fp, C.POINTER(C.c_int), _fp, _fp]
        _lib.orc_quaternion_angular_distance.argtypes = [_fp, _fp, _fp]
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_last_timings.argtypes = [C.POINTER(C.c_double)]
    return _lib


def _f(a):
    return a.ctypes.data_as(_fp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _cloud(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4, "clouds are (N, 4) float32"
    return a


def _check(rc):
    if rc != 0:
        raise OracleError(rc)


def num_threads():
    return lib().orc_num_threads()


class KdTree:
    """libnabo-style bucketed kd-tree over a (N, 4) cloud (MatchersImpl.cpp:77-83)."""

    def __init__(self, reference):
        self.ref = _cloud(reference)
        self.handle = lib().orc_kdtree_create(_f(self.ref), 4, self.ref.shape[0])

    def knn(self, query, k=1, eps=0.0, max_dist=np.inf, nthreads=1):
        q = _cloud(query)
        ids = np.empty((q.shape[0], k), np.int32)
        dists = np.empty((q.shape[0], k), np.float32)
        visits = lib().orc_kdtree_knn(self.handle, _f(q), 4, q.shape[0], k, eps, max_dist, _i(ids), _f(dists), nthreads)
        if visits < 0:
            raise OracleError(-visits)
        self.last_visits = visits
        return ids, dists

    def __del__(self):
        if getattr(self, "handle", None):
            lib().orc_kdtree_destroy(self.handle)
            self.handle = None


def bruteforce_knn(reference, query, k=1, max_dist=np.inf, nthreads=1):
    """clouds (N, 4), or (N, 3) for 2-D clouds: the search runs over the first rows - 1 coordinates (MatchersImpl.cpp:82)"""
    r, q = np.ascontiguousarray(reference, np.float32), np.ascontiguousarray(query, np.float32)
    assert r.ndim == 2 and r.shape[1] in (3, 4) and q.shape[1] == r.shape[1], "clouds are (N, 4) or (N, 3) float32"
    ids = np.empty((q.shape[0], k), np.int32)
    dists = np.empty((q.shape[0], k), np.float32)
    rc = lib().orc_bruteforce_knn(_f(r), r.shape[1], r.shape[0], _f(q), q.shape[0], k, max_dist, _i(ids), _f(dists), nthreads)
    if rc < 0:
        raise OracleError(-rc)
    return ids, dists


def rigid_transform(T, cloud):
    """T: (4, 4) numpy in the usual row/col sense.  Returns T applied to every point."""
    c = _cloud(cloud)
    Tc = np.asfortranarray(np.asarray(T, np.float32))
    out = np.empty_like(c)
    _check(lib().orc_rigid_transform(_f(Tc), _f(c), c.shape[0], _f(out)))
    return out


def dists_quantile(dists, quantile):
    d = np.ascontiguousarray(dists, np.float32)
    out = np.zeros(1, np.float32)
    _check(lib().orc_dists_quantile(_f(d), d.size, quantile, _f(out)))
    return out[0]


def set_robust_approximation(approximation=float("inf")):
    """`approximation` of the RobustOutlierFilters evaluated from now on (metres; inf: none)"""
    lib().orc_set_robust_approximation(approximation)


def set_var_trimmed_ratios(min_ratio=0.05, max_ratio=0.99):
    """minRatio / maxRatio of the VarTrimmedDist filters (type FILTER_VARTRIMMEDDIST, param lambda) evaluated from now on"""
    lib().orc_set_var_trimmed_ratios(min_ratio, max_ratio)


def var_trimmed_ratio(dists, min_ratio=0.05, max_ratio=0.99, lam=2.35):
    """VarTrimmedDistOutlierFilter::optimizeInlierRatio (OutlierFiltersImpl.cpp:177-218)"""
    d = np.ascontiguousarray(dists, np.float32)
    out = np.zeros(1, np.float32)
    _check(lib().orc_var_trimmed_ratio(_f(d), d.size, min_ratio, max_ratio, lam, _f(out)))
    return out[0]


def outlier_weights(dists, filters):
    """filters: list of (type, param).  Returns (weights (N, k), limits)."""
    d = np.ascontiguousarray(dists, np.float32)
    n, k = d.shape
    types = np.array([f[0] for f in filters], np.int32)
    params = np.array([f[1] for f in filters], np.float32)
    w = np.empty_like(d)
    limits = np.zeros(max(1, len(filters)), np.float32)
    _check(lib().orc_outlier_weights(_f(d), k, n, len(filters), _i(types), _f(params), _f(w), _f(limits)))
    return w, limits[: len(filters)]


def bruteforce_knn_var(reference, query, k, max_radii, nthreads=1):
    """KDTreeVarDistMatcher semantics: per-query maximum radius.  Returns (ids (N, k), dists (N, k))."""
    rf, q = _cloud(reference), _cloud(query)
    mr = np.ascontiguousarray(max_radii, np.float32).ravel()
    ids = np.empty((q.shape[0], k), np.int32)
    dists = np.empty((q.shape[0], k), np.float32)
    L = lib()
    L.orc_bruteforce_knn_var.restype = C.c_long
    L.orc_bruteforce_knn_var.argtypes = [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int, _fp, _ip, _fp, C.c_int]
    rc = L.orc_bruteforce_knn_var(_f(rf), 4, rf.shape[0], _f(q), q.shape[0], k, _f(mr), _i(ids), _f(dists), nthreads)
    if rc < 0:
        _check(-rc)
    return ids, dists


def outlier_weights_geom(dists, ids, filters, reading, reference, ref_normals):
    """a chain that may contain a RobustOutlierFilter with distanceType point2plane (robust_word(...) | ROBUST_P2PLANE): reading
    (N, 4) as the iteration sees it, reference (Nr, 4), ref_normals (Nr, 3)"""
    d = np.ascontiguousarray(dists, np.float32)
    i = np.ascontiguousarray(ids, np.int32)
    n, k = d.shape
    types = np.array([f[0] for f in filters], np.int32)
    params = np.array([f[1] for f in filters], np.float32)
    rd, rf = _cloud(reading), _cloud(reference)
    qn = np.ascontiguousarray(ref_normals, np.float32)
    w = np.empty_like(d)
    limits = np.zeros(max(1, len(filters)), np.float32)
    L = lib()
    L.orc_outlier_weights_geom.argtypes = [_fp, _ip, C.c_int, C.c_int, C.c_int, _ip, _fp, _fp, _fp, _fp, _fp, _fp]
    _check(L.orc_outlier_weights_geom(_f(d), _i(i), k, n, len(filters), _i(types), _f(params), _f(rd), _f(rf), _f(qn), _f(w), _f(limits)))
    return w, limits[: len(filters)]


def outlier_weights_sn(dists, ids, filters, reading_normals, ref_normals):
    """a chain that may contain SurfaceNormalOutlierFilter (type FILTER_SURFACENORMAL, param maxAngle); reading_normals
    (N, 3) already rotated like the reading, ref_normals (Nr, 3)"""
    d = np.ascontiguousarray(dists, np.float32)
    i = np.ascontiguousarray(ids, np.int32)
    n, k = d.shape
    types = np.array([f[0] for f in filters], np.int32)
    params = np.array([f[1] for f in filters], np.float32)
    rn = None if reading_normals is None else np.ascontiguousarray(reading_normals, np.float32)
    qn = None if ref_normals is None else np.ascontiguousarray(ref_normals, np.float32)
    w = np.empty_like(d)
    limits = np.zeros(max(1, len(filters)), np.float32)
    L = lib()
    L.orc_outlier_weights_sn.argtypes = [_fp, _ip, C.c_int, C.c_int, C.c_int, _ip, _fp, _fp, _fp, _fp, _fp]
    _check(L.orc_outlier_weights_sn(_f(d), _i(i), k, n, len(filters), _i(types), _f(params), _f(rn), _f(qn), _f(w), _f(limits)))
    return w, limits[: len(filters)]


def minimize(minimizer, reading, reference, ref_normals, ids, dists, weights, sensor_std_dev=0.01, acc_double=False):
    """Returns (T (4,4), cov (6,6) or None, stats dict)."""
    rd, rf = _cloud(reading), _cloud(reference)
    nr = None if ref_normals is None else np.ascontiguousarray(ref_normals, np.float32)
    ids = np.ascontiguousarray(ids, np.int32)
    dists = np.ascontiguousarray(dists, np.float32)
    weights = np.ascontiguousarray(weights, np.float32)
    knn = ids.shape[1]
    T = np.zeros((4, 4), np.float32, order="F")
    cov = np.zeros((6, 6), np.float32, order="F")
    stats = np.zeros(5, np.float32)
    _check(lib().orc_minimize(minimizer, _f(rd), rd.shape[0], _f(rf), rf.shape[0], _f(nr), _i(ids), _f(dists), _f(weights), knn,
                              sensor_std_dev, int(acc_double), _f(T), _f(cov), _f(stats)))
    st = dict(pointUsedRatio=stats[0], weightedPointUsedRatio=stats[1], nbRejectedMatches=int(stats[2]),
              nbRejectedPoints=int(stats[3]), nbKept=int(stats[4]))
    return np.array(T), (np.array(cov) if (minimizer & 0xff) in (2, 3) else None), st


def surface_normals(cloud, knn=5, eps=0.0, max_dist=np.inf, sort_eigen=False, smooth_normals=False, nthreads=1):
    c = _cloud(cloud)
    n = c.shape[0]
    out = dict(
        normals=np.zeros((n, 3), np.float32), densities=np.zeros(n, np.float32), eigValues=np.zeros((n, 3), np.float32),
        eigVectors=np.zeros((n, 9), np.float32), matchedIds=np.zeros((n, knn), np.float32), meanDists=np.zeros(n, np.float32),
        ids=np.zeros((n, knn), np.int32), dists=np.zeros((n, knn), np.float32), gap=np.zeros(n, np.float32),
    )
    deg = C.c_int(0)
    _check(lib().orc_surface_normals(_f(c), 4, n, knn, eps, max_dist, int(sort_eigen), int(smooth_normals), nthreads,
                                     _f(out["normals"]), _f(out["densities"]), _f(out["eigValues"]), _f(out["eigVectors"]),
                                     _f(out["matchedIds"]), _f(out["meanDists"]), _i(out["ids"]), _f(out["dists"]), _f(out["gap"]),
                                     C.byref(deg)))
    out["degenerate"] = deg.value
    return out


def make_config(knn=1, epsilon=0.0, max_dist=np.inf, search_type=1, filters=(), minimizer=MIN_P2POINT, sensor_std_dev=0.01,
                max_iterations=40, differential=None, acc_double=False, nthreads=1):
    cfg = IcpConfig()
    cfg.knn, cfg.epsilon, cfg.max_dist, cfg.search_type = knn, epsilon, max_dist, search_type
    cfg.nfilters = len(filters)
    for i, (t, p) in enumerate(filters):
        cfg.filter_type[i] = t
        cfg.filter_param[i] = p
    cfg.minimizer, cfg.sensor_std_dev, cfg.max_iterations = minimizer, sensor_std_dev, max_iterations
    if differential is not None:
        cfg.use_differential = 1
        cfg.min_diff_rot_err, cfg.min_diff_trans_err, cfg.smooth_length = differential
    cfg.acc_double, cfg.nthreads = int(acc_double), nthreads
    return cfg


def icp(reading, reference, ref_normals=None, T_init=None, reading_normals=None, **kw):
    """ICP::operator() (ICP.cpp:243-449).  Returns dict(T, iterations, T_iters, cov, stats).  reading_normals: the
    reading's "normals" descriptor, used by SurfaceNormalOutlierFilter only."""
    cfg = kw.pop("config", None) or make_config(**kw)
    rn = None if reading_normals is None else np.ascontiguousarray(reading_normals, np.float32)
    lib().orc_set_reading_normals.argtypes = [_fp]
    lib().orc_set_reading_normals.restype = None
    lib().orc_set_reading_normals(_f(rn))
    rd, rf = _cloud(reading), _cloud(reference)
    nr = None if ref_normals is None else np.ascontiguousarray(ref_normals, np.float32)
    Ti = np.asfortranarray(np.eye(4, dtype=np.float32) if T_init is None else np.asarray(T_init, np.float32))
    T = np.zeros((4, 4), np.float32, order="F")
    T_iters = np.zeros((max(1, cfg.max_iterations), 16), np.float32)
    cov = np.zeros((6, 6), np.float32, order="F")
    stats = np.zeros(5, np.float32)
    it = C.c_int(0)
    _check(lib().orc_icp(_f(rd), rd.shape[0], _f(rf), rf.shape[0], _f(nr), _f(Ti), C.byref(cfg), _f(T), _f(T_iters), C.byref(it),
                         _f(cov), _f(stats)))
    iters = [np.array(T_iters[i].reshape(4, 4, order="F")) for i in range(it.value)]
    return dict(T=np.array(T), iterations=it.value, T_iters=iters, cov=np.array(cov), stats=stats)


def last_timings():
    """dict(build_s, loop_s, match_s, iterations) of the last icp() call"""
    t = (C.c_double * 4)()
    lib().orc_last_timings(t)
    return dict(build_s=t[0], loop_s=t[1], match_s=t[2], iterations=int(t[3]))


def angular_distance(Ta, Tb):
    a = np.asfortranarray(np.asarray(Ta, np.float32))
    b = np.asfortranarray(np.asarray(Tb, np.float32))
    out = np.zeros(1, np.float32)
    lib().orc_quaternion_angular_distance(_f(a), _f(b), _f(out))
    return float(out[0])
