"""Shared helpers of the test-suite."""
import numpy as np


def rot_angle(Ra, Rb):
    """geodesic distance between two rotation matrices (radians)"""
    # |Ra - Rb|_F = 2 sqrt(2) sin(theta / 2): well conditioned near theta = 0, unlike acos(trace)
    d = np.linalg.norm(np.asarray(Ra, np.float64)[:3, :3] - np.asarray(Rb, np.float64)[:3, :3])
    return float(2.0 * np.arcsin(min(1.0, d / (2.0 * np.sqrt(2.0)))))


def assert_transform_close(Ta, Tb, rot_tol=1e-5, trans_tol=1e-5):
    """the north-star tolerance: 1e-5 rad / 1e-5 m"""
    ra = rot_angle(Ta, Tb)
    dt = float(np.linalg.norm(np.asarray(Ta, np.float64)[:3, 3] - np.asarray(Tb, np.float64)[:3, 3]))
    assert ra <= rot_tol and dt <= trans_tol, "rotation differs by %.3g rad, translation by %.3g m\n%s\n%s" % (ra, dt, Ta, Tb)


def cloud(rng, n, kind="uniform"):
    if kind == "uniform":
        p = rng.uniform(-10, 10, (n, 3))
    elif kind == "grid":  # many exact ties and duplicates
        p = rng.integers(-3, 4, (n, 3)).astype(float)
    elif kind == "plane":
        p = np.c_[rng.uniform(-5, 5, (n, 2)), np.zeros(n)]
    elif kind == "cluster":
        p = rng.normal(0, 0.01, (n, 3)) + rng.integers(0, 3, (n, 1)) * 50
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(np.c_[p, np.ones(n)].astype(np.float32))


def small_pose(rng, trans=0.3, ang=0.05):
    a = rng.uniform(-ang, ang, 3)
    cx, sx, cy, sy, cz, sz = np.cos(a[0]), np.sin(a[0]), np.cos(a[1]), np.sin(a[1]), np.cos(a[2]), np.sin(a[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = rng.uniform(-trans, trans, 3)
    return T.astype(np.float32)


def classify_id_mismatches(ids_a, dists_a, ids_b, dists_b):
    """(number of differing ids, how many of them sit on an exact distance tie)"""
    diff = ids_a != ids_b
    ties = diff & (dists_a.view(np.uint32) == dists_b.view(np.uint32))
    return int(diff.sum()), int(ties.sum())
