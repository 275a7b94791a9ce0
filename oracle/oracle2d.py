"""oracle/oracle2d.py — TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU restatement, in numpy, of the reference's path for 2-D clouds (features.rows() == 3: x, y, w), which every module of
the reference branches on.  Each function cites the lines it follows.  The nearest-neighbour search is the dimension-
generic brute force of oracle.cpp (libnabo's brute-force semantics), the distance filters are the dimension-independent
ones of oracle.cpp; what is restated here is what differs in 2-D: the 3x3 rigid transform, the 2x2 point-to-point
rotation, the 2-D point-to-plane system, the Differential checker's quaternion built from the whole 3x3 homogeneous
matrix, and the 2x2 surface-normal eigen-problem.
Pinned by the reference's own 2-D fixture (tests/test_oracle_golden.py): `2D_twoBoxes -> 2D_oneBox` must land within
0.05 of validT2d in translation norm and rotation angle (utest/utest.h:44-60, utest/utest.cpp:347-350) with both
error minimisers; anything finer (1e-5 agreement with the GPU path) is pinned by this restatement only.
Long sums are carried in float64 (the "truth" variant the GPU path is compared with, like oracle.cpp's acc_double).
"""
import numpy as np

from . import binding as orc

F = np.float32


def transform(T, cloud):
    """RigidTransformation::compute on a 3 x N cloud (TransformationsImpl.cpp:49-87): features' = T * features, float,
    the depth-3 sums left to right"""
    T = np.asarray(T, F)
    c = np.asarray(cloud, F)
    out = np.empty_like(c)
    for r in range(3):
        acc = (T[r, 0] * c[:, 0]).astype(F)
        acc = (acc + (T[r, 1] * c[:, 1]).astype(F)).astype(F)
        acc = (acc + (T[r, 2] * c[:, 2]).astype(F)).astype(F)
        out[:, r] = acc
    return out


def rotate_normals(T, normals):
    """`R * inputDesc` for the "normals" descriptor of a 2-D cloud (TransformationsImpl.cpp:71-84)"""
    T = np.asarray(T, F)
    n = np.asarray(normals, F)
    out = np.empty_like(n)
    for r in range(2):
        out[:, r] = ((T[r, 0] * n[:, 0]).astype(F) + (T[r, 1] * n[:, 1]).astype(F)).astype(F)
    return out


def is_rigid(T):
    """RigidTransformation::checkParameters (TransformationsImpl.cpp:90-105) on the 2x2 block"""
    T = np.asarray(T, F)
    det = F(T[0, 0] * T[1, 1]) - F(T[0, 1] * T[1, 0])
    return not (abs(F(1) - det) > F(0.001))


def knn(reference, query, k=1, max_dist=np.inf):
    """KDTreeMatcher::findClosests over the first 2 rows (MatchersImpl.cpp:85-101)"""
    return orc.bruteforce_knn(reference, query, k, max_dist)


def kept_pairs(reading, ids, dists, weights):
    """ErrorElements (ErrorMinimizer.cpp:98-135): reading point index, matched id, weight of every kept pair, point-major"""
    n, k = ids.shape
    keep = (weights != 0) & np.isfinite(dists)
    qi = np.repeat(np.arange(n), k).reshape(n, k)[keep]
    return qi, ids[keep], weights[keep].astype(np.float64)


def point_to_point(reading_t, reference, ids, dists, weights):
    """PointToPointErrorMinimizer::compute_in_place for dimCount == 3 (PointToPoint.cpp:61-101): weighted centroids,
    m = ref' diag(w) read'^T (2x2), R = U V^T, a reflection fixed by negating row dimCount - 2 of V^T"""
    qi, ri, w = kept_pairs(reading_t, ids, dists, weights)
    if len(w) == 0:
        raise ValueError("no point to minimize")
    p = reading_t[qi, :2].astype(np.float64)
    q = reference[ri, :2].astype(np.float64)
    W = w.sum()
    mp = F((p * w[:, None]).sum(0) / W)   # the reference keeps the centroids in float
    mq = F((q * w[:, None]).sum(0) / W)
    pd, qd = p - mp.astype(np.float64), q - mq.astype(np.float64)
    m = (qd * w[:, None]).T @ pd
    U, _, Vt = np.linalg.svd(m)
    R = U @ Vt
    if np.linalg.det(R) < 0:
        Vt = Vt.copy()
        Vt[1, :] *= -1          # row dimCount - 2 = 1
        R = U @ Vt
    R = R.astype(F)
    t = (mq.astype(np.float64) - R.astype(np.float64) @ mp.astype(np.float64)).astype(F)
    T = np.eye(3, dtype=F)
    T[:2, :2] = R
    T[:2, 2] = t
    return T


def point_to_plane(reading_t, reference, normals, ids, dists, weights):
    """PointToPlaneErrorMinimizer::compute_in_place for dim == 3 (PointToPlane.cpp:171-312): cross = x n_y - y n_x
    (ErrorMinimizer.cpp:308-313), F = [cross; n], A = wF F^T (3x3), b = -wF (n . (p - q)), T = Rotation2D(x0) + (x1, x2)"""
    qi, ri, w = kept_pairs(reading_t, ids, dists, weights)
    if len(w) == 0:
        raise ValueError("no point to minimize")
    p, q, n = reading_t[qi, :2], reference[ri, :2], np.asarray(normals, F)[ri, :2]
    cross = (F(1) * (p[:, 0] * n[:, 1]).astype(F) - (p[:, 1] * n[:, 0]).astype(F)).astype(F)
    Fm = np.stack([cross, n[:, 0], n[:, 1]], axis=0).astype(F)           # 3 x M
    wF = (Fm * w.astype(F)[None, :]).astype(F)
    dot = ((p[:, 0] - q[:, 0]).astype(F) * n[:, 0]).astype(F)
    dot = (dot + ((p[:, 1] - q[:, 1]).astype(F) * n[:, 1]).astype(F)).astype(F)
    A = wF.astype(np.float64) @ Fm.astype(np.float64).T
    b = -(wF.astype(np.float64) @ dot.astype(np.float64))
    x = np.linalg.lstsq(A, b, rcond=3 * np.finfo(np.float32).eps)[0]        # minimum norm when rank-deficient (:108-161)
    ang = F(x[0])
    T = np.eye(3, dtype=F)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1] = np.cos(ang), -np.sin(ang), np.sin(ang), np.cos(ang)
    T[0, 2], T[1, 2] = F(x[1]), F(x[2])
    return T


def quat_from_matrix3(m):
    """Eigen's Matrix3 -> Quaternion conversion (no normalisation), float"""
    m = np.asarray(m, F)
    t = F(m[0, 0] + m[1, 1] + m[2, 2])
    if t > 0:
        t = np.sqrt(F(t + F(1)))
        w = F(0.5) * t
        t = F(0.5) / t
        return np.array([w, (m[2, 1] - m[1, 2]) * t, (m[0, 2] - m[2, 0]) * t, (m[1, 0] - m[0, 1]) * t], F)
    i = 0
    if m[1, 1] > m[0, 0]:
        i = 1
    if m[2, 2] > m[i, i]:
        i = 2
    j, k = (i + 1) % 3, (i + 2) % 3
    t = np.sqrt(F(m[i, i] - m[j, j] - m[k, k] + F(1)))
    v = np.zeros(3, F)
    v[i] = F(0.5) * t
    t = F(0.5) / t
    w = (m[k, j] - m[j, k]) * t
    v[j] = (m[j, i] + m[i, j]) * t
    v[k] = (m[k, i] + m[i, k]) * t
    return np.array([w, v[0], v[1], v[2]], F)


def angular_distance(a, b):
    """QuaternionBase::angularDistance (Eigen 3.3): d = a * conj(b); 2 atan2(|d.vec|, |d.w|)"""
    dw = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]
    dx = -a[0] * b[1] + a[1] * b[0] - a[2] * b[3] + a[3] * b[2]
    dy = -a[0] * b[2] + a[2] * b[0] - a[3] * b[1] + a[1] * b[3]
    dz = -a[0] * b[3] + a[3] * b[0] - a[1] * b[2] + a[2] * b[1]
    return F(2) * np.arctan2(np.sqrt(F(dx * dx + dy * dy + dz * dz)), abs(F(dw)))


class Differential:
    """DifferentialTransformationChecker on 3x3 parameters (TransformationCheckersImpl.cpp:103-158): init() embeds the 2x2
    rotation in an identity (:116-121); check() builds its quaternion from topLeftCorner(3, 3) of the HOMOGENEOUS matrix,
    translation column included (:131) — restated as it is"""

    def __init__(self, min_rot, min_trans, smooth):
        self.limits, self.smooth = (F(min_rot), F(min_trans)), int(smooth)

    def init(self, T):
        m = np.eye(3, dtype=F)
        m[:2, :2] = np.asarray(T, F)[:2, :2]
        self.rot = [quat_from_matrix3(m)]
        self.tr = [np.asarray(T, F)[:2, 2].copy()]

    def check(self, T):
        T = np.asarray(T, F)
        self.rot.append(quat_from_matrix3(T))
        self.tr.append(T[:2, 2].copy())
        if len(self.rot) > self.smooth:
            c0 = c1 = F(0)
            for i in range(len(self.rot) - 1, len(self.rot) - 1 - self.smooth, -1):
                c0 += abs(angular_distance(self.rot[i], self.rot[i - 1]))
                c1 += abs(F(np.linalg.norm((self.tr[i] - self.tr[i - 1]).astype(F))))
            c0, c1 = c0 / F(self.smooth), c1 / F(self.smooth)
            if c0 < self.limits[0] and c1 < self.limits[1]:
                return False
        return True


def surface_normals(cloud, knn_=5, max_dist=np.inf):
    """SurfaceNormalDataPointsFilter on a 2-D cloud (SurfaceNormal.cpp:166-252 with featDim - 1 == 2): mean and 2x2
    scatter matrix of the valid neighbours (the point itself included), normal = eigenvector of the smallest eigenvalue,
    density = k / (4/3 pi r_max^3) (utils.h:105-120).  Returns dict(normals (N, 2), densities (N,), eigValues (N, 2))."""
    c = np.asarray(cloud, F)
    ids, dists = orc.bruteforce_knn(c, c, knn_, max_dist)
    n = len(c)
    normals, dens, eva = np.zeros((n, 2), F), np.zeros(n, F), np.zeros((n, 2), F)
    for i in range(n):
        nb = c[ids[i][np.isfinite(dists[i])], :2]
        mean = np.zeros(2, F)
        for p in nb:
            mean = (mean + p).astype(F)
        mean = mean / F(len(nb))
        NN = (nb - mean).astype(F)
        C = np.zeros((2, 2), F)
        for v in NN:
            C = (C + np.outer(v, v).astype(F)).astype(F)
        if np.linalg.matrix_rank(C.astype(np.float64), tol=2 * np.finfo(np.float32).eps * np.abs(C).max()) + 1 >= 2:
            w, V = np.linalg.eigh(C.astype(np.float64))
            eva[i] = w.astype(F)
            normals[i] = np.clip(V[:, 0].astype(F), -1, 1)
            r = float(np.sqrt((NN.astype(F) ** 2).sum(1)).max())
            dens[i] = F(len(nb)) / F((4.0 / 3.0) * np.pi * r ** 3)
    return dict(normals=normals, densities=dens, eigValues=eva, ids=ids)


def icp(reading, reference, normals=None, T_init=None, knn_=1, max_dist=np.inf, filters=(), minimizer="point", max_iterations=40,
        differential=None):
    """ICP::compute + computeWithTransformedReference for 3-row clouds (ICP.cpp:264-449): the reference is centred on its
    float mean (:291-299), the reading moved into that frame (:345-347), T_iter <- dT * T_iter (:411-412), Counter and
    Differential decide.  Returns dict(T (3, 3), iterations)."""
    rf, rd = np.asarray(reference, F).copy(), np.asarray(reading, F)
    n = len(rf)
    s = np.zeros(2, F)
    for i in range(n):           # rowwise().sum() in float, column after column (:292)
        s = (s + rf[i, :2]).astype(F)
    mean = (s / F(n)).astype(F)
    rf[:, :2] = (rf[:, :2] - mean).astype(F)
    T_ref_mean = np.eye(3, dtype=F)
    T_ref_mean[:2, 2] = mean
    T_mean_ref = np.eye(3, dtype=F)
    T_mean_ref[:2, 2] = -mean
    T0 = np.eye(3, dtype=F) if T_init is None else np.asarray(T_init, F)
    T_mean_data = mul3(T_mean_ref, T0)
    rd0 = transform(T_mean_data, rd)
    T_iter = np.eye(3, dtype=F)
    diff = Differential(*differential) if differential else None
    if diff:
        diff.init(T_iter)
    it, iterate = 0, True
    while iterate:
        step = transform(T_iter, rd0)
        ids, dists = knn(rf, step, knn_, max_dist)
        w, _ = orc.outlier_weights(dists, list(filters))
        dT = point_to_point(step, rf, ids, dists, w) if minimizer == "point" else point_to_plane(step, rf, normals, ids, dists, w)
        T_iter = mul3(dT, T_iter)
        it += 1
        if it >= max_iterations:      # Counter first, like every reference chain
            break
        if diff and not diff.check(T_iter):
            break
        if not is_rigid(T_iter):
            raise ValueError("rotation matrix is not orthogonal")
    return dict(T=mul3(mul3(T_ref_mean, T_iter), T_mean_data), iterations=it, T_iter=T_iter)


def mul3(A, B):
    A, B = np.asarray(A, F), np.asarray(B, F)
    out = np.zeros((3, 3), F)
    for i in range(3):
        for j in range(3):
            acc = F(A[i, 0] * B[0, j])
            acc = F(acc + F(A[i, 1] * B[1, j]))
            acc = F(acc + F(A[i, 2] * B[2, j]))
            out[i, j] = acc
    return out
