"""CPU tests of the exact per-thread device algorithms (csrc/core/*.h) through the host harness
tests/emu/emu.cpp, checked against the oracle.  The kernels themselves are tested on the GPU
(tests/test_gpu_*.py); this file keeps the logic they share honest without one.
"""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_transform_close, cloud, small_pose

FP = C.POINTER(C.c_float)
IP = C.POINTER(C.c_int32)
DP = C.POINTER(C.c_double)


def f(a):
    return a.ctypes.data_as(FP)


def d(a):
    return a.ctypes.data_as(DP)


def emu_knn(emu, ref, q, k, max_dist=np.inf, T=None):
    t = emu.emu_tree_build(f(ref), len(ref))
    try:
        assert emu.emu_tree_check(t) == 0
        ids = np.empty((len(q), k), np.int32)
        dist = np.empty((len(q), k), np.float32)
        Tc = None if T is None else np.asfortranarray(np.asarray(T, np.float32))
        visits = emu.emu_knn(t, None if T is None else f(Tc), f(q), len(q), k, max_dist, ids.ctypes.data_as(IP), f(dist))
        return ids, dist, visits
    finally:
        emu.emu_tree_free(t)


def test_segment_arithmetic(emu):
    for n in (1, 2, 7, 8, 9, 1000, 1000003):
        for level in (0, 1, 3, 7):
            if (1 << level) > n:
                continue
            segs = 1 << level
            begins = [emu.emu_seg_begin(level, s, n) for s in range(segs + 1)]
            assert begins[0] == 0 and begins[-1] == n and all(b1 > b0 for b0, b1 in zip(begins, begins[1:]))
            for s in range(segs):
                assert emu.emu_seg_of(begins[s], level, n) == s
                assert emu.emu_seg_of(begins[s + 1] - 1, level, n) == s
                # children are exact sub-ranges of the parent
                assert emu.emu_seg_begin(level + 1, 2 * s, n) == begins[s]


def test_float_ord_is_monotone(emu):
    v = np.array([-np.inf, -1e30, -1.5, -1e-40, -0.0, 0.0, 1e-40, 1.5, 3e38, np.inf], np.float32)
    o = [emu.emu_float_ord(float(x)) for x in v]
    assert all(a <= b for a, b in zip(o, o[1:]))
    for x in v:
        assert emu.emu_ord_float(emu.emu_float_ord(float(x))) == x


@pytest.mark.parametrize("kind", ["uniform", "grid", "plane", "cluster"])
def test_knn_bit_exact_vs_bruteforce(emu, oracle, kind):
    """ids AND dists bit-exact against libnabo's brute-force semantics, exact ties included."""
    rng = np.random.default_rng(10)
    for n in (1, 2, 8, 9, 100, 3000):
        ref, q = cloud(rng, n, kind), cloud(rng, 200, kind)
        if kind != "grid":
            q[:, :3] += rng.normal(0, 0.5, (200, 3)).astype(np.float32)
        for k in (1, 2, 5, 10, 20, 40):
            if k > n:
                continue
            for md in (np.inf, 1.0):
                ib, db = oracle.bruteforce_knn(ref, q, k, md)
                ie, de, _ = emu_knn(emu, ref, q, k, md)
                assert (ib == ie).all() and (db.view(np.uint32) == de.view(np.uint32)).all(), (kind, n, k, md)


def test_knn_with_transform_matches_oracle_transform(emu, oracle):
    rng = np.random.default_rng(11)
    ref, q = cloud(rng, 20000, "uniform"), cloud(rng, 1000, "uniform")
    T = small_pose(rng)
    qt = oracle.rigid_transform(T, q)
    ib, db = oracle.bruteforce_knn(ref, qt, 4)
    ie, de, _ = emu_knn(emu, ref, q, 4, T=T)
    assert (ib == ie).all() and (db.view(np.uint32) == de.view(np.uint32)).all()


def test_knn_vs_kdtree_on_lidar(emu, oracle, synth):
    rd, rf, _ = synth.scan_pair(30000)
    ik, dk = oracle.KdTree(rf).knn(rd, 1)
    ie, de, visits = emu_knn(emu, rf, rd, 1)
    assert (dk.view(np.uint32) == de.view(np.uint32)).all()
    assert ((ik == ie) | (dk == de)).all()
    assert visits / len(rd) < 40  # box pruning keeps the leaf visits low on surface data


def test_knn_seed_candidate_keeps_the_answer_exact(emu, oracle):
    """k = 1: the previous match, re-measured, may seed the search (it only tightens the bounds);
    any seed — right, wrong or missing — gives the same answer"""
    rng = np.random.default_rng(16)
    ref, q = cloud(rng, 5000, "grid"), cloud(rng, 600, "grid")  # grid: exact ties with the seed
    ref2, q2 = cloud(rng, 5000, "uniform"), cloud(rng, 600, "uniform")
    for r_, q_ in ((ref, q), (ref2, q2)):
        t = emu.emu_tree_build(f(r_), len(r_))
        try:
            for md in (np.inf, 0.3):
                ib, db = oracle.bruteforce_knn(r_, q_, 1, md)
                for seed in (ib[:, 0].copy(), rng.integers(-1, len(r_), len(q_)).astype(np.int32)):
                    emu.emu_set_seed(seed.ctypes.data_as(C.c_void_p))
                    ids = np.empty((len(q_), 1), np.int32)
                    dist = np.empty((len(q_), 1), np.float32)
                    emu.emu_knn(t, None, f(q_), len(q_), 1, md, ids.ctypes.data_as(IP), f(dist))
                    assert (ib == ids).all() and (db.view(np.uint32) == dist.view(np.uint32)).all()
        finally:
            emu.emu_set_seed(None)
            emu.emu_tree_free(t)


def test_solve_psd6(emu, oracle):
    rng = np.random.default_rng(12)
    x = np.zeros(6)
    for _ in range(20):
        J = rng.normal(size=(50, 6)) * np.array([30, 30, 30, 1, 1, 1])
        A = np.asfortranarray(J.T @ J)
        xt = rng.normal(size=6)
        b = A @ xt
        rank = emu.emu_solve_psd6(d(A), d(b), d(x))
        assert rank == 6 and np.allclose(x, xt, rtol=1e-8, atol=1e-10)
    # rank-deficient (planar scene, icpSingular): minimum-norm solution == pinv
    for r in (1, 3, 5):
        J = rng.normal(size=(40, r)) @ rng.normal(size=(r, 6))
        A = np.asfortranarray(J.T @ J)
        b = A @ rng.normal(size=6)
        rank = emu.emu_solve_psd6(d(A), d(b), d(x))
        assert rank == r
        assert np.allclose(x, np.linalg.pinv(A, rcond=1e-9) @ b, rtol=1e-6, atol=1e-9)
    A = np.zeros((6, 6), order="F")
    assert emu.emu_solve_psd6(d(A), d(np.zeros(6)), d(x)) == 0 and (x == 0).all()


def test_rotation_from_crosscov(emu):
    rng = np.random.default_rng(13)
    R = np.zeros((3, 3), order="F")
    for trial in range(30):
        m = rng.normal(size=(3, 3))
        if trial % 3 == 1:
            m = m @ np.diag([1.0, 0.5, 0.0])  # rank 2
        if trial % 3 == 2:
            m[:, 0] *= -1
        m = np.asfortranarray(m)
        emu.emu_rotation_from_crosscov(d(m), d(R))
        U, s, Vt = np.linalg.svd(m)
        Rn = U @ Vt
        if np.linalg.det(Rn) < 0:
            Vt[2] *= -1
            Rn = U @ Vt
        assert np.allclose(R.T @ R, np.eye(3), atol=1e-12) and np.linalg.det(R) > 0
        if s[1] > 1e-9 * s[0] and (s[2] > 1e-6 * s[0] or trial % 3 == 1):
            assert np.allclose(R, Rn, atol=1e-7), (trial, s)


def test_jacobi_eig3_and_rank3(emu):
    rng = np.random.default_rng(14)
    w, V = np.zeros(3), np.zeros((3, 3), order="F")
    for _ in range(30):
        B = rng.normal(size=(3, 6))
        A = np.asfortranarray(B @ B.T)
        emu.emu_jacobi_eig3(d(A), d(w), d(V))
        wn = np.linalg.eigvalsh(A)
        assert np.allclose(np.sort(w), wn, rtol=1e-10)
        assert np.allclose(A @ V, V * w, atol=1e-9 * wn[-1])
    assert emu.emu_rank3(f(np.eye(3, dtype=np.float32))) == 3
    plane = np.diag([2.0, 1.0, 0.0]).astype(np.float32)
    assert emu.emu_rank3(f(np.asfortranarray(plane))) == 2
    line = np.outer([1, 2, 3], [1, 2, 3]).astype(np.float32)
    assert emu.emu_rank3(f(np.asfortranarray(line))) == 1
    assert emu.emu_rank3(f(np.zeros((3, 3), np.float32))) == 0


def test_angle_axis_and_checkers_match_oracle(emu, oracle):
    rng = np.random.default_rng(15)
    T = np.zeros((4, 4), np.float32, order="F")
    for _ in range(20):
        x = (rng.normal(size=6) * [0.1, 0.1, 0.1, 1, 1, 1]).astype(np.float32)
        emu.emu_angle_axis(f(x), f(T))
        R = T[:3, :3].astype(np.float64)
        assert np.allclose(R.T @ R, np.eye(3), atol=1e-6) and np.allclose(T[:3, 3], x[3:])
        ang = np.linalg.norm(x[:3].astype(np.float64))
        assert abs(np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1)) - ang) < 1e-3
        T2 = small_pose(rng)
        a = emu.emu_angular_distance(f(np.asfortranarray(T)), f(np.asfortranarray(T2)))
        assert abs(a - oracle.angular_distance(T, T2)) < 1e-6
    emu.emu_angle_axis(f(np.zeros(6, np.float32)), f(T))  # zero motion -> identity (PointToPlane.cpp:286-292)
    assert (T == np.eye(4)).all()


