"""ctypes binding of the C ABI in include/pmgpu.h (libpmgpu.so, CUDA sm_100a).

There is no CPU fallback: importing this module fails loudly when the extension has not been
built, and creating a Context fails when no CUDA device is present.

Array conventions (zero-copy views of the reference's column-major Eigen matrices):
  clouds   (N, 4) float32 C-contiguous  == 4 x N column-major `DataPoints::features`
  normals  (N, 3) float32               == the 3 `normals` rows of `DataPoints::descriptors`
  ids/dists/weights (N, k)              == k x N column-major `Matches` / `OutlierWeights`
  transforms (4, 4) numpy arrays in the mathematical (row, col) sense
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PMGPU_VARIANT selects a tuning build (libpmgpu_<name>.so, see build.py); default: libpmgpu.so
LIB_PATH = os.path.join(_HERE, "libpmgpu%s.so" % ("_" + os.environ["PMGPU_VARIANT"] if os.environ.get("PMGPU_VARIANT") else ""))

FILTER_MAXDIST, FILTER_MEDIANDIST, FILTER_TRIMMEDDIST, FILTER_ROBUST, FILTER_SURFACENORMAL, FILTER_VARTRIMMEDDIST, FILTER_MINDIST = 0, 1, 2, 3, 4, 5, 6
ROBUST_P2PLANE = 1 << 28   # RobustOutlierFilter distanceType point2plane, or-ed into its filter word
MIN_P2POINT, MIN_P2PLANE, MIN_P2POINT_COV, MIN_P2PLANE_COV, MIN_P2POINT_SIM = 0, 1, 2, 3, 4
MIN_FORCE4DOF = 0x100  # or-ed into a point-to-plane minimiser id
MIN_FORCE2D = 0x200    # likewise (PointToPlaneErrorMinimizer only)
NORMALS_SORT_EIGEN, NORMALS_SMOOTH = 1, 2

OK = 0
ERR_CUDA, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_NO_REFERENCE, ERR_NO_READING, ERR_NO_MATCHES = 1, 2, 3, 4, 5, 6
ERR_NO_OUTLIER_TO_FILTER, ERR_BAD_QUANTILE, ERR_NO_POINT_TO_MINIMIZE, ERR_NO_NORMALS = 7, 8, 9, 10
ERR_NOT_ORTHOGONAL, ERR_KNN_TOO_LARGE, ERR_NAN, ERR_COMM = 11, 12, 13, 14

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libpointmatcher_b200: %s is missing — build it with `python -m libpointmatcher_b200.build` "
        "(there is no CPU fallback)" % LIB_PATH
    )

_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int32)


class IcpParams(C.Structure):
    """pmgpu_icp_params"""
    _fields_ = [
        ("knn", C.c_int), ("epsilon", C.c_float), ("max_dist", C.c_float), ("nfilters", C.c_int),
        ("filter_type", C.c_int * 8), ("filter_param", C.c_float * 8), ("minimizer", C.c_int),
        ("sensor_std_dev", C.c_float), ("max_iterations", C.c_int), ("use_differential", C.c_int),
        ("min_diff_rot_err", C.c_float), ("min_diff_trans_err", C.c_float), ("smooth_length", C.c_int),
    ]


class NormalsOut(C.Structure):
    """pmgpu_normals_out"""
    _fields_ = [
        ("normals", _fp), ("normals_ld", C.c_int), ("densities", _fp), ("densities_ld", C.c_int),
        ("eig_values", _fp), ("eig_values_ld", C.c_int), ("eig_vectors", _fp), ("eig_vectors_ld", C.c_int),
        ("matched_ids", _fp), ("matched_ids_ld", C.c_int), ("mean_dists", _fp), ("mean_dists_ld", C.c_int),
    ]


# every symbol include/pmgpu.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "pmgpu_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "pmgpu_ctx_destroy": (None, [C.c_void_p]),
    "pmgpu_last_error": (C.c_char_p, [C.c_void_p]),
    "pmgpu_status_string": (C.c_char_p, [C.c_int]),
    "pmgpu_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "pmgpu_sync": (C.c_int, [C.c_void_p]),
    "pmgpu_launch_count": (C.c_uint64, [C.c_void_p]),
    "pmgpu_timing_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "pmgpu_timing_collect": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "pmgpu_ref_set": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "pmgpu_ref_set_centered": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, _fp]),
    "pmgpu_ref_set_normals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "pmgpu_reading_set": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "pmgpu_reading_set_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "pmgpu_reading_set_normals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "pmgpu_reading_set_max_dists": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "pmgpu_reading_apply_transform": (C.c_int, [C.c_void_p, _fp]),
    "pmgpu_reading_get": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pmgpu_ref_get_normals": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pmgpu_knn": (C.c_int, [C.c_void_p, _fp, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "pmgpu_weights": (C.c_int, [C.c_void_p, C.c_int, _ip, _fp, C.c_void_p, _fp]),
    "pmgpu_minimize": (C.c_int, [C.c_void_p, C.c_int, C.c_float, _fp, _fp, _fp]),
    "pmgpu_normals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int,
                                C.POINTER(NormalsOut), C.POINTER(C.c_int)]),
    "pmgpu_ref_compute_normals": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_int]),
    "pmgpu_ref_center": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, _fp]),
    "pmgpu_icp_run": (C.c_int, [C.c_void_p, C.POINTER(IcpParams), _fp, _fp, C.POINTER(C.c_int), _fp, _fp]),
    "pmgpu_icp_enqueue": (C.c_int, [C.c_void_p, C.POINTER(IcpParams), C.c_int]),
    "pmgpu_icp_reset": (C.c_int, [C.c_void_p, _fp]),
    "pmgpu_icp_result": (C.c_int, [C.c_void_p, _fp, C.POINTER(C.c_int), _fp, _fp]),
    "pmgpu_icp_step": (C.c_int, [C.c_void_p, C.POINTER(IcpParams), _fp, C.POINTER(C.c_int), _fp, _fp]),
    "pmgpu_icp_cap_redos": (C.c_int, [C.c_void_p]),
    "pmgpu_matches_get": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pmgpu_set_var_trimmed_ratios": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "pmgpu_set_robust_approximation": (C.c_int, [C.c_void_p, C.c_float]),
    "pmgpu_host_pin": (C.c_int, [C.c_void_p, C.c_size_t]),
    "pmgpu_host_unpin": (C.c_int, [C.c_void_p]),
    "pmgpu_var_trimmed_ratio": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pmgpu_host_srand": (None, [C.c_uint]),
    "pmgpu_host_random_sampling": (C.c_int, [C.c_int, C.c_float, C.c_void_p]),
    "pmgpu_host_rand": (C.c_int, []),
    "pmgpu_host_max_point_count": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]),
    "pmgpu_host_max_density": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "pmgpu_host_sampling_surface_normal": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float, C.c_int,
                                                    C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]),
    "pmgpu_comm_unique_id": (C.c_int, [C.c_void_p]),
    "pmgpu_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "pmgpu_comm_peer_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "pmgpu_comm_peer_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "pmgpu_comm_destroy": (C.c_int, [C.c_void_p]),
}

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError if the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args


class PmGpuError(RuntimeError):
    """A non-zero pmgpu status; `.code` is the PMGPU_* value (see pmgpu.h for the reference
    exception each one maps to)."""

    def __init__(self, code, message):
        super().__init__("%s (pmgpu status %d)" % (message, code))
        self.code = code


def _f(a):
    return None if a is None else a.ctypes.data_as(_fp)


def _ptr(a):
    """numpy array -> void*, int -> raw (device) pointer, None -> NULL"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.ctypes.data)


def _cloud(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2:
        raise ValueError("clouds are 2-D arrays (N, rows)")
    return a


def _T(T):
    return None if T is None else np.asfortranarray(np.asarray(T, dtype=np.float32))


def make_params(knn=1, epsilon=0.0, max_dist=np.inf, filters=(), minimizer=MIN_P2POINT, sensor_std_dev=0.01, max_iterations=40,
                differential=None):
    p = IcpParams()
    p.knn, p.epsilon, p.max_dist = knn, epsilon, max_dist
    p.nfilters = len(filters)
    if len(filters) > 8:
        raise ValueError("at most 8 outlier filters")
    for i, (t, v) in enumerate(filters):
        p.filter_type[i] = t
        p.filter_param[i] = v
    p.minimizer, p.sensor_std_dev, p.max_iterations = minimizer, sensor_std_dev, max_iterations
    if differential is not None:
        p.use_differential = 1
        p.min_diff_rot_err, p.min_diff_trans_err, p.smooth_length = differential
    return p


class Context:
    """One pmgpu context: a CUDA stream plus the resident reference / reading / matches."""

    def __init__(self, device=0):
        h = C.c_void_p()
        rc = lib.pmgpu_ctx_create(device, C.byref(h))
        if rc != OK:
            raise PmGpuError(rc, "cannot create a pmgpu context on CUDA device %d: %s (no CPU fallback exists)"
                             % (device, lib.pmgpu_status_string(rc).decode()))
        self.h = h
        self.device = device
        self.nq = self.nr = self.k = 0
        self.dimh = 4   # features.rows() of the resident clouds: 4, or 3 for 2-D clouds (transforms are then 3 x 3)

    def close(self):
        if getattr(self, "h", None):
            lib.pmgpu_ctx_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != OK:
            raise PmGpuError(rc, lib.pmgpu_last_error(self.h).decode() or lib.pmgpu_status_string(rc).decode())

    # ---- plumbing
    @property
    def stream(self):
        return lib.pmgpu_ctx_stream(self.h)

    def sync(self):
        self._check(lib.pmgpu_sync(self.h))

    @property
    def launch_count(self):
        return int(lib.pmgpu_launch_count(self.h))

    def timing_enable(self, on=True):
        self._check(lib.pmgpu_timing_enable(self.h, int(on)))

    def timing_collect(self):
        """{stage: (total ms, launches-of-that-stage)} since the last collect; synchronises."""
        ms = (C.c_double * 4)()
        cnt = (C.c_int * 4)()
        self._check(lib.pmgpu_timing_collect(self.h, ms, cnt))
        names = ("knn", "select", "minimize", "covariance")
        return {n: (ms[i], cnt[i]) for i, n in enumerate(names)}

    # ---- K1
    def set_reference(self, features, normals=None):
        """KDTreeMatcher::init.  `features`: (N, rows) array, or (device_ptr, n) tuple."""
        if isinstance(features, tuple):
            ptr, n = features
            rows = 4
            nrm = normals
            self._check(lib.pmgpu_ref_set(self.h, _ptr(ptr), rows, n, _ptr(nrm), 3 if nrm is not None else 0))
        else:
            f = _cloud(features)
            n, rows = f.shape
            nrm = None if normals is None else np.ascontiguousarray(normals, np.float32)
            ld = 0 if nrm is None else nrm.shape[1]
            self._check(lib.pmgpu_ref_set(self.h, _ptr(f), rows, n, _ptr(nrm), ld))
            self.dimh = rows
        self.nr = n

    def set_reference_centered(self, features, normals=None):
        """ICP::compute's preamble + KDTreeMatcher::init: centre the reference on its mean (float row
        sums / N, ICP.cpp:291-299) and build the structure.  Returns the mean (3,)."""
        f = _cloud(features)
        n, rows = f.shape
        nrm = None if normals is None else np.ascontiguousarray(normals, np.float32)
        mean = np.zeros(4, np.float32)
        self._check(lib.pmgpu_ref_set_centered(self.h, _ptr(f), rows, n, _ptr(nrm), 0 if nrm is None else nrm.shape[1], _f(mean)))
        self.nr, self.dimh = n, rows
        return mean[:rows - 1].copy()

    def set_reference_normals(self, normals):
        nrm = None if normals is None else np.ascontiguousarray(normals, np.float32)
        self._check(lib.pmgpu_ref_set_normals(self.h, _ptr(nrm), 0 if nrm is None else nrm.shape[1]))

    def set_reading(self, features):
        if isinstance(features, tuple):
            ptr, n = features
            self._check(lib.pmgpu_reading_set(self.h, _ptr(ptr), 4, n))
        else:
            f = _cloud(features)
            n, rows = f.shape
            self._check(lib.pmgpu_reading_set(self.h, _ptr(f), rows, n))
            if self.nr == 0:
                self.dimh = rows
        self.nq = n

    def set_reading_sharded(self, features, rank, world, chunk):
        """this rank's chunks of the whole reading (N, 4), uploaded straight from `features` (dist.shard_columns' columns)"""
        f = _cloud(features)
        n, rows = f.shape
        self._check(lib.pmgpu_reading_set_sharded(self.h, _ptr(f), rows, n, rank, world, chunk))
        nchunks = n // chunk
        mine = (nchunks - rank + world - 1) // world if nchunks > rank else 0
        self.nq = mine * chunk + (n - nchunks * chunk if nchunks % world == rank else 0)

    def reading_apply_transform(self, T):
        self._check(lib.pmgpu_reading_apply_transform(self.h, _f(_T(T))))

    def set_reading_normals(self, normals):
        """the reading's "normals" descriptor (N, 3), after set_reading; used by SurfaceNormalOutlierFilter"""
        nrm = None if normals is None else np.ascontiguousarray(normals, np.float32)
        self._check(lib.pmgpu_reading_set_normals(self.h, _ptr(nrm), 0 if nrm is None else nrm.shape[1]))

    def set_reading_max_dists(self, max_dists):
        """KDTreeVarDistMatcher: one maximum search distance per reading point (N,) — then pass max_dist < 0 to knn / make_params"""
        md = None if max_dists is None else np.ascontiguousarray(np.asarray(max_dists, np.float32).reshape(-1))
        self._check(lib.pmgpu_reading_set_max_dists(self.h, _ptr(md), 1))

    def get_reading(self):
        out = np.empty((self.nq, self.dimh), np.float32)
        self._check(lib.pmgpu_reading_get(self.h, _ptr(out)))
        return out

    # ---- K2
    def knn(self, T=None, k=1, epsilon=0.0, max_dist=np.inf, download=True):
        """KDTreeMatcher::findClosests on T * reading.  Returns (ids, dists, visits)."""
        ids = np.empty((self.nq, k), np.int32) if download else None
        dists = np.empty((self.nq, k), np.float32) if download else None
        visits = C.c_uint64(0)
        self._check(lib.pmgpu_knn(self.h, _f(_T(T)), k, epsilon, max_dist, _ptr(ids), _ptr(dists), C.byref(visits)))
        self.k = k
        return ids, dists, int(visits.value)

    # ---- K3
    def weights(self, filters=(), download=True):
        """OutlierFilters::compute.  filters: [(type, param)].  Returns (weights, limits)."""
        types = np.array([f[0] for f in filters], np.int32)
        params = np.array([f[1] for f in filters], np.float32)
        w = np.empty((self.nq, self.k), np.float32) if download else None
        limits = np.zeros(max(1, len(filters)), np.float32)
        self._check(lib.pmgpu_weights(self.h, len(filters), types.ctypes.data_as(_ip), _f(params), _ptr(w), _f(limits)))
        return w, limits[: len(filters)]

    def ref_normals(self):
        """the reference's normals as resident on the device (nr, 3)"""
        out = np.empty((self.nr, self.dimh - 1), np.float32)
        self._check(lib.pmgpu_ref_get_normals(self.h, _ptr(out)))
        return out

    def matches(self, weights=True):
        """the resident matches of the last evaluation: (ids (nq, k), dists (nq, k), weights (nq, k) or None, T_match (4, 4))"""
        ids = np.empty((self.nq, self.k), np.int32)
        dists = np.empty((self.nq, self.k), np.float32)
        w = np.empty((self.nq, self.k), np.float32) if weights else None
        T = np.zeros((self.dimh, self.dimh), np.float32, order="F")
        self._check(lib.pmgpu_matches_get(self.h, _ptr(ids), _ptr(dists), _ptr(w), _f(T)))
        return ids, dists, w, np.array(T)

    def set_robust_approximation(self, approximation=float("inf")):
        """RobustOutlierFilter `approximation` (metres; inf: none) of the chain evaluated from here on"""
        self._check(lib.pmgpu_set_robust_approximation(self.h, approximation))

    def set_var_trimmed_ratios(self, min_ratio=0.05, max_ratio=0.99):
        """minRatio / maxRatio of the chain's VarTrimmedDistOutlierFilter (type FILTER_VARTRIMMEDDIST, param lambda)"""
        self._check(lib.pmgpu_set_var_trimmed_ratios(self.h, min_ratio, max_ratio))

    def var_trimmed_ratio(self):
        """the inlier ratio the last evaluation of a VarTrimmedDist filter chose"""
        out = np.zeros(1, np.float32)
        self._check(lib.pmgpu_var_trimmed_ratio(self.h, _f(out)))
        return out[0]

    # ---- K4-K7
    def minimize(self, minimizer, sensor_std_dev=0.01):
        """ErrorMinimizer::compute.  Returns (T (4,4), cov (6,6) or None, stats dict)."""
        T = np.zeros((self.dimh, self.dimh), np.float32, order="F")
        cov = np.zeros((6, 6), np.float32, order="F")
        stats = np.zeros(5, np.float32)
        self._check(lib.pmgpu_minimize(self.h, minimizer, sensor_std_dev, _f(T), _f(cov), _f(stats)))
        return np.array(T), (np.array(cov) if (minimizer & 0xff) in (2, 3) else None), _stats(stats)

    # ---- K8
    def normals(self, features, knn=5, epsilon=0.0, max_dist=np.inf, sort_eigen=False, keep=("normals",), smooth=False):
        """SurfaceNormalDataPointsFilter on a cloud.  keep: subset of normals, densities, eigValues,
        eigVectors, matchedIds, meanDists.  Returns dict of arrays (+ 'degenerate')."""
        f = _cloud(features)
        n, dn = f.shape[0], f.shape[1] - 1
        spans = dict(normals=dn, densities=1, eigValues=dn, eigVectors=dn * dn, matchedIds=knn, meanDists=1)
        arrays = {name: np.zeros((n, spans[name]), np.float32) for name in keep}
        o = NormalsOut()
        for name, field in (("normals", "normals"), ("densities", "densities"), ("eigValues", "eig_values"),
                            ("eigVectors", "eig_vectors"), ("matchedIds", "matched_ids"), ("meanDists", "mean_dists")):
            if name in arrays:
                setattr(o, field, _f(arrays[name]))
                setattr(o, field + "_ld", spans[name])
        deg = C.c_int(0)
        flags = (NORMALS_SORT_EIGEN if sort_eigen else 0) | (NORMALS_SMOOTH if smooth else 0)
        self._check(lib.pmgpu_normals(self.h, _ptr(f), f.shape[1], n, knn, epsilon, max_dist, flags, C.byref(o), C.byref(deg)))
        arrays["degenerate"] = deg.value
        return arrays

    def ref_compute_normals(self, knn=5, epsilon=0.0, max_dist=np.inf, smooth=False):
        self._check(lib.pmgpu_ref_compute_normals(self.h, knn, epsilon, max_dist, NORMALS_SMOOTH if smooth else 0))

    def ref_center(self, features):
        """centre the resident reference on the mean of the host cloud it was set from; returns the mean (4,)"""
        f = _cloud(features)
        mean = np.zeros(4, np.float32)
        self._check(lib.pmgpu_ref_center(self.h, _ptr(f), f.shape[1], f.shape[0], _f(mean)))
        return mean[:f.shape[1]]

    # ---- fused loop
    def icp_run(self, params, T_iter_init=None):
        """ICP::computeWithTransformedReference loop.  Returns dict(T_iter, iterations, cov, stats)."""
        T = np.zeros((self.dimh, self.dimh), np.float32, order="F")
        cov = np.zeros((6, 6), np.float32, order="F")
        stats = np.zeros(5, np.float32)
        it = C.c_int(0)
        self._check(lib.pmgpu_icp_run(self.h, C.byref(params), _f(_T(T_iter_init)), _f(T), C.byref(it), _f(cov), _f(stats)))
        self.k = params.knn
        return dict(T_iter=np.array(T), iterations=it.value, cov=np.array(cov), stats=_stats(stats), cap_redos=int(lib.pmgpu_icp_cap_redos(self.h)))

    def icp_reset(self, T_iter_init=None):
        self._check(lib.pmgpu_icp_reset(self.h, _f(_T(T_iter_init))))

    def icp_enqueue(self, params, n_iterations):
        self._check(lib.pmgpu_icp_enqueue(self.h, C.byref(params), n_iterations))
        self.k = params.knn

    def icp_step(self, params):
        """one exactly-matched iteration slot (after icp_reset), then the state: dict as icp_result"""
        T = np.zeros((self.dimh, self.dimh), np.float32, order="F")
        cov = np.zeros((6, 6), np.float32, order="F")
        stats = np.zeros(5, np.float32)
        it = C.c_int(0)
        self._check(lib.pmgpu_icp_step(self.h, C.byref(params), _f(T), C.byref(it), _f(cov), _f(stats)))
        self.k = params.knn
        return dict(T_iter=np.array(T), iterations=it.value, cov=np.array(cov), stats=_stats(stats), cap_redos=int(lib.pmgpu_icp_cap_redos(self.h)))

    def icp_result(self):
        T = np.zeros((self.dimh, self.dimh), np.float32, order="F")
        cov = np.zeros((6, 6), np.float32, order="F")
        stats = np.zeros(5, np.float32)
        it = C.c_int(0)
        self._check(lib.pmgpu_icp_result(self.h, _f(T), C.byref(it), _f(cov), _f(stats)))
        return dict(T_iter=np.array(T), iterations=it.value, cov=np.array(cov), stats=_stats(stats), cap_redos=int(lib.pmgpu_icp_cap_redos(self.h)))

    # ---- multi-GPU
    def comm_init(self, unique_id, rank, nranks):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(lib.pmgpu_comm_init(self.h, buf, rank, nranks))

    def comm_peer_handle(self):
        """allocate this context's peer mailbox; returns its 128-byte handle (to be gathered over all ranks)"""
        buf = C.create_string_buffer(128)
        self._check(lib.pmgpu_comm_peer_handle(self.h, buf))
        return bytes(buf.raw)

    def comm_peer_init(self, handles, rank):
        """handles: the ranks' mailbox handles in rank order"""
        blob = b"".join(bytes(h) for h in handles)
        buf = C.create_string_buffer(blob, len(blob))
        self._check(lib.pmgpu_comm_peer_init(self.h, buf, rank, len(handles)))

    def comm_destroy(self):
        self._check(lib.pmgpu_comm_destroy(self.h))


def comm_unique_id():
    buf = C.create_string_buffer(128)
    rc = lib.pmgpu_comm_unique_id(buf)
    if rc != OK:
        raise PmGpuError(rc, "cannot create an NCCL unique id")
    return bytes(buf.raw)


def _stats(s):
    return dict(pointUsedRatio=float(s[0]), weightedPointUsedRatio=float(s[1]), nbRejectedMatches=int(s[2]),
                nbRejectedPoints=int(s[3]), nbKept=int(s[4]))
