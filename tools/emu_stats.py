"""Algorithmic work of the kNN search (core/tree.h) counted on the CPU harness: descent steps, plane
tests, box tests, re-descents and leaves per query, for an unseeded and a seeded (aligned) pass.

    python tools/emu_stats.py [points]
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libpointmatcher_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "_build", "libemu.so"))
fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int32)
lib.emu_tree_build.restype = C.c_void_p
lib.emu_tree_build.argtypes = [fp, C.c_int]
lib.emu_tree_depth.argtypes = [C.c_void_p]
lib.emu_set_seed.argtypes = [C.c_void_p]
lib.emu_knn.restype = C.c_long
lib.emu_knn.argtypes = [C.c_void_p, fp, fp, C.c_int, C.c_int, C.c_float, ip, fp]
lib.emu_stats.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]

rd, rf, T_gt = synth.scan_pair(n)
rf = np.ascontiguousarray(rf, dtype=np.float32)
rd = np.ascontiguousarray(rd, dtype=np.float32)
tree = lib.emu_tree_build(rf.ctypes.data_as(fp), len(rf))
print("points", n, "depth", lib.emu_tree_depth(tree))


def run(T, seed, label):
    ids = np.empty(len(rd), np.int32)
    d = np.empty(len(rd), np.float32)
    lib.emu_set_seed(seed.ctypes.data_as(C.c_void_p) if seed is not None else None)
    out = (C.c_ulonglong * 5)()
    lib.emu_stats(out, 1)
    T = np.ascontiguousarray(T.T, dtype=np.float32)  # column-major
    lib.emu_knn(tree, T.ctypes.data_as(fp), rd.ctypes.data_as(fp), len(rd), 1, np.float32(np.inf), ids.ctypes.data_as(ip), d.ctypes.data_as(fp))
    lib.emu_stats(out, 1)
    names = ["descent_steps", "plane_tests", "box_tests", "redescents", "leaves"]
    print("%-28s" % label, "  ".join("%s %.2f" % (k, v / len(rd)) for k, v in zip(names, out)))
    return ids


T_gt = np.asarray(T_gt, dtype=np.float64)
Tn = T_gt.copy()
Tn[:3, 3] += [0.01, -0.008, 0.004]  # 1 cm off: a late ICP iteration
ids0 = run(np.eye(4), None, "identity, unseeded")
ids1 = run(Tn, None, "near ground truth, unseeded")
run(T_gt, ids1, "ground truth, seeded")
run(Tn, ids0, "near gt, seeded from identity")
