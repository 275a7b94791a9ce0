import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure, oracle/)."""
    from oracle import binding
    binding.build()
    binding.lib()
    return binding


@pytest.fixture(scope="session")
def emu():
    """Host harness around the PM_HD device functions (tests/emu/emu.cpp)."""
    import ctypes as C
    here = os.path.join(ROOT, "tests", "emu")
    out = os.path.join(here, "_build", "libemu.so")
    src = os.path.join(here, "emu.cpp")
    core = os.path.join(ROOT, "libpointmatcher_b200", "csrc", "core")
    deps = [src] + [os.path.join(core, f) for f in os.listdir(core)]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wno-unknown-pragmas",
                               "-I", os.path.join(ROOT, "libpointmatcher_b200", "csrc"), "-o", out, src])
    lib = C.CDLL(out)
    fp, ip, dp = C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_double)
    lib.emu_tree_build.restype = C.c_void_p
    lib.emu_tree_build.argtypes = [fp, C.c_int]
    lib.emu_tree_check.argtypes = [C.c_void_p]
    lib.emu_tree_free.argtypes = [C.c_void_p]
    lib.emu_tree_depth.argtypes = [C.c_void_p]
    lib.emu_set_seed.argtypes = [C.c_void_p]
    lib.emu_knn.restype = C.c_long
    lib.emu_knn.argtypes = [C.c_void_p, fp, fp, C.c_int, C.c_int, C.c_float, ip, fp]
    lib.emu_solve_psd6.argtypes = [dp, dp, dp]
    lib.emu_rotation_from_crosscov.argtypes = [dp, dp]
    lib.emu_jacobi_eig3.argtypes = [dp, dp, dp]
    lib.emu_rank3.argtypes = [fp]
    lib.emu_angle_axis.argtypes = [fp, fp]
    lib.emu_angular_distance.restype = C.c_float
    lib.emu_angular_distance.argtypes = [fp, fp]
    lib.emu_mat4_mul.argtypes = [fp, fp, fp]
    lib.emu_seg_begin.restype = C.c_uint32
    lib.emu_seg_begin.argtypes = [C.c_int, C.c_uint32, C.c_uint32]
    lib.emu_seg_of.restype = C.c_uint32
    lib.emu_seg_of.argtypes = [C.c_uint32, C.c_int, C.c_uint32]
    lib.emu_float_ord.restype = C.c_uint32
    lib.emu_float_ord.argtypes = [C.c_float]
    lib.emu_ord_float.restype = C.c_float
    lib.emu_ord_float.argtypes = [C.c_uint32]
    return lib


@pytest.fixture(scope="session")
def synth():
    from libpointmatcher_b200 import synth as s
    return s


@pytest.fixture(scope="session")
def gpu_ctx():
    from libpointmatcher_b200 import capi
    ctx = capi.Context(0)
    yield ctx
    ctx.close()
