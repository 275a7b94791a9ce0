"""Oracle restatement of the CPU pre-filters of the default chain.  TEST INFRASTRUCTURE
ONLY — nothing in the product imports this file.

  RandomSamplingDataPointsFilter          pointmatcher/DataPointsFilters/RandomSampling.cpp:58-75
  SamplingSurfaceNormalDataPointsFilter   pointmatcher/DataPointsFilters/SamplingSurfaceNormal.cpp:80-342,
                                          utils/utils.h:105-139
  MinDist / MaxDistDataPointsFilter       MinDist.cpp:60-100, MaxDist.cpp:60-100

Parity status: the reference draws from the C library's rand() in point order, which this file
reproduces through ctypes (same libc), so RandomSampling is pinned bit for bit.  The bins of
SamplingSurfaceNormal are defined by recursive median cuts of the CELL's widest dimension; as SETS they
are unique whenever no two points share the cut coordinate, and that is what is restated here (numpy
argpartition).  The ORDER inside a bin — which decides which points the random subsampling keeps and
which descriptor "the first one" is — is whatever std::nth_element leaves behind in the reference's
C++ runtime: parity unpinned, and not restated.
"""
import ctypes
import ctypes.util

import numpy as np

_libc = ctypes.CDLL(ctypes.util.find_library("c"))
_libc.rand.restype = ctypes.c_int
RAND_MAX = 2147483647


def srand(seed):
    _libc.srand(ctypes.c_uint(seed))


def random_sampling(n, prob):
    """RandomSampling.cpp:63-72: r = (float)rand() / (float)RAND_MAX; keep if r < prob (prob held in a double)."""
    keep = []
    p = float(np.float32(prob))
    for i in range(n):
        r = np.float32(_libc.rand()) / np.float32(RAND_MAX)
        if float(r) < p:
            keep.append(i)
    return np.array(keep, np.int64)


def bins(features, knn):
    """SamplingSurfaceNormal.cpp:177-230.  features: (N, 4).  Returns a list of index arrays."""
    f = np.asarray(features, np.float32)
    out = []

    def rec(idx, lo, hi):
        count = len(idx)
        if count <= knn:
            out.append(idx)
            return
        cut_dim = int(np.argmax(hi - lo))  # first maximum, homogeneous row included (extent 0)
        right = count // 2
        left = count - right
        order = np.argpartition(f[idx, cut_dim], left)
        idx = idx[order]
        cut_val = f[idx[left], cut_dim]
        left_hi = hi.copy()
        left_hi[cut_dim] = cut_val
        right_lo = lo.copy()
        right_lo[cut_dim] = cut_val
        rec(idx[:left], lo, left_hi)
        rec(idx[left:], right_lo, hi)

    if len(f):
        rec(np.arange(len(f)), f.min(axis=0), f.max(axis=0))
    return out


def fuse(features, idx, max_box_dim=np.inf):
    """SamplingSurfaceNormal.cpp:232-270 for one bin: (mean, normal, density) or None when the bin is
    dropped (box too large, or the covariance has rank < 2)."""
    d = np.asarray(features, np.float32)[idx, :3]
    if (d.max(axis=0) - d.min(axis=0)).max() > max_box_dim:
        return None
    mean = (np.cumsum(d, axis=0, dtype=np.float32)[-1] / np.float32(len(d))).astype(np.float32)
    nn = (d - mean).astype(np.float32)
    c = (nn.T.astype(np.float64) @ nn.astype(np.float64))
    if np.linalg.matrix_rank(c.astype(np.float32), tol=3 * np.finfo(np.float32).eps * np.abs(c).max()) + 1 < 3:
        return None
    w, v = np.linalg.eigh(c)
    normal = v[:, int(np.argmin(w))].astype(np.float32)
    r = np.sqrt((nn.astype(np.float64) ** 2).sum(axis=1)).max()
    density = np.float32(len(d) / ((4.0 / 3.0) * np.pi * r ** 3))
    return mean, normal, density


def sampling_surface_normal_method1(features, knn, max_box_dim=np.inf, descriptors=None, average=True):
    """samplingMethod 1 (SamplingSurfaceNormal.cpp:296-331): one point per bin at the bin's mean.  Returns
    (features (M, 4), normals (M, 3), densities (M,), averaged descriptors or None, unfit count); row order
    is the bins' traversal order (the reference sorts by the kept index, which is not restated)."""
    feats, normals, dens, descs, unfit = [], [], [], [], 0
    for idx in bins(features, knn):
        r = fuse(features, idx, max_box_dim)
        if r is None:
            unfit += len(idx)
            continue
        mean, normal, density = r
        feats.append(np.r_[mean, np.float32(1)])
        normals.append(normal)
        dens.append(density)
        if descriptors is not None and average:
            acc = np.zeros(descriptors.shape[1], np.float32)
            for i in idx:
                acc = (acc + descriptors[i]).astype(np.float32)
            descs.append(acc / np.float32(len(idx)))
    return (np.array(feats, np.float32).reshape(-1, 4), np.array(normals, np.float32).reshape(-1, 3), np.array(dens, np.float32),
            np.array(descs, np.float32) if descs else None, unfit)


def min_dist(features, dim, min_dist_value):
    f = np.asarray(features, np.float32)
    keep = []
    for i in range(len(f)):
        if dim == -1:
            v = np.sqrt(np.float32(np.float32(f[i, 0] * f[i, 0] + f[i, 1] * f[i, 1]) + f[i, 2] * f[i, 2]))
            if v > np.float32(abs(min_dist_value)):
                keep.append(i)
        elif f[i, dim] > np.float32(min_dist_value):
            keep.append(i)
    return np.array(keep, np.int64)


def max_dist(features, dim, max_dist_value):
    f = np.asarray(features, np.float32)
    keep = []
    for i in range(len(f)):
        if dim == -1:
            v = np.sqrt(np.float32(np.float32(f[i, 0] * f[i, 0] + f[i, 1] * f[i, 1]) + f[i, 2] * f[i, 2]))
            if v < np.float32(abs(max_dist_value)):
                keep.append(i)
        elif f[i, dim] < np.float32(max_dist_value):
            keep.append(i)
    return np.array(keep, np.int64)


# ---- the other per-cloud filters of the reference's golden chain files (plain loops, one point at a time) ----
#   BoundingBox.cpp:71-103, DistanceLimit.cpp:66-127, FixStepSampling.cpp:76-110, MaxPointCount.cpp:71-110,
#   MaxQuantileOnAxis.cpp:65-103, RemoveNaN.cpp:52-72, MaxDensity.cpp:60-105, Shadow.cpp:62-90,
#   SimpleSensorNoise.cpp:75-140.  The rand()-driven ones are pinned through libc like random_sampling above.
def _f32(x):
    return np.float32(x)


def _norm3(v):
    acc = _f32(0)
    for a in v:
        acc = _f32(acc + _f32(a) * _f32(a))
    return _f32(np.sqrt(acc))


def bounding_box(features, lo, hi, remove_inside):
    keep = []
    for i, p in enumerate(np.asarray(features, np.float32)):
        inside = all(p[a] > _f32(lo[a]) and p[a] < _f32(hi[a]) for a in range(len(p) - 1))
        if inside != bool(remove_inside):
            keep.append(i)
    return np.array(keep, np.int64)


def distance_limit(features, dim, dist, remove_inside):
    keep = []
    for i, p in enumerate(np.asarray(features, np.float32)):
        v, lim = (_norm3(p[:-1]), _f32(abs(dist))) if dim == -1 else (p[dim], _f32(dist))
        if (v > lim) if remove_inside else (v < lim):
            keep.append(i)
    return np.array(keep, np.int64)


def fix_step(n, step):
    """one call: phase = rand() % int(step), then every int(step)-th column"""
    i_step = int(step)
    phase = _libc.rand() % i_step
    return np.arange(phase, n, i_step)


def fix_step_next(step, start, end, mult):
    delta = start * mult - start
    step *= mult
    if delta < 0 and step < end:
        step = end
    if delta > 0 and step > end:
        step = end
    return step


def max_point_count(n, seed, max_count):
    """the columns the reference ends up with: column j takes column idx (the Eigen-view 'swap' never writes idx back)"""
    N = n - 1
    if not max_count <= N:
        return np.arange(n)
    srand(seed)
    cur = list(range(n))
    for j in range(max_count):
        r = np.float32(np.float32(_libc.rand()) / np.float32(RAND_MAX))
        idx = j + int(np.float32(N - j) * r)
        view_j = j                      # `const auto feat = col(j)`: a view of column j, not its value
        cur[j] = cur[idx]
        cur[idx] = cur[view_j]          # reads the NEW column j: no-op
    return np.array(cur[:max_count], np.int64)


def max_quantile_on_axis(features, dim, ratio):
    v = np.asarray(features, np.float32)[:, dim]
    rank = int(np.float32(len(v)) * np.float32(ratio))
    limit = np.sort(v)[rank]
    return np.array([i for i in range(len(v)) if v[i] < limit], np.int64)


def remove_nan(features):
    return np.array([i for i, p in enumerate(np.asarray(features, np.float32)) if not any(x != x for x in p)], np.int64)


def max_density(densities, max_density_):
    d = np.asarray(densities, np.float32)
    last = d.max()
    saturated = int((d == last).sum())
    keep = []
    for i in range(len(d)):
        if d[i] > _f32(max_density_):
            r = np.float32(_libc.rand()) / np.float32(RAND_MAX)
            accept = _f32(_f32(max_density_) / d[i])
            if d[i] == last:
                accept = _f32(accept * _f32(1 - saturated // len(d)))
            if r < accept:
                keep.append(i)
        else:
            keep.append(i)
    return np.array(keep, np.int64)


def shadow(features, normals, eps):
    lim = np.sin(np.float32(eps))
    keep = []
    for i, (p, nrm) in enumerate(zip(np.asarray(features, np.float32), np.asarray(normals, np.float32))):
        nn, pn = _norm3(nrm), _norm3(p[:-1])
        nu = [_f32(x / nn) if nn > 0 else x for x in nrm]
        pu = [_f32(x / pn) if pn > 0 else x for x in p[:-1]]
        dot = _f32(0)
        for a, b in zip(nu, pu):
            dot = _f32(dot + _f32(a * b))
        if abs(dot) > lim:
            keep.append(i)
    return np.array(keep, np.int64)


_LASERS = {0: (0.012, 0.0068, 0.0008), 1: (0.028, 0.0013, 0.0001), 2: (0.018, 0.0006, 0.0015), 4: (0.004, 0.0053, -0.0092)}


def simple_sensor_noise(features, sensor_type):
    out = []
    for p in np.asarray(features, np.float32):
        norm = _norm3(p[:-1])
        if sensor_type == 3:
            out.append(_f32(_f32(norm * norm) * _f32(0.5 * 0.00285)))
        else:
            mn, ang, cst = (_f32(v) for v in _LASERS[sensor_type])
            out.append(max(_f32(_f32(ang * norm) + cst), mn))
    return np.array(out, np.float32)
