"""Builds libpmgpu.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m libpointmatcher_b200.build [--force]

The shared library is written next to this file so it travels to the GPU box with the repo
snapshot (it is git-ignored).  nvcc cross-compiles without a GPU.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# tuning variants: PMGPU_VARIANT=<name> PMGPU_DEFINES="-DPM_LEAF_MAX=16 ..." builds libpmgpu_<name>.so
VARIANT = os.environ.get("PMGPU_VARIANT", "")
DEFINES = os.environ.get("PMGPU_DEFINES", "").split()
OBJ = os.path.join(HERE, "_build" + ("_" + VARIANT if VARIANT else ""))
LIB = os.path.join(HERE, "libpmgpu%s.so" % ("_" + VARIANT if VARIANT else ""))
SOURCES = ["api.cu", "tree_build.cu", "knn.cu", "select.cu", "minimize.cu", "comm.cu", "host_filters.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--expt-relaxed-constexpr",
    "-I", os.path.join(HERE, "..", "include"),
]


def _headers():
    out = [os.path.join(HERE, "..", "include", "pmgpu.h")]
    for root, _, files in os.walk(CSRC):
        out += [os.path.join(root, f) for f in files if f.endswith((".h", ".cuh"))]
    return out


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not _stale(obj, [path] + _headers()):
        return obj
    cmd = [NVCC] + FLAGS + DEFINES + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout))
    if verbose:
        sys.stderr.write(r.stdout)
    return obj


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    if force or _stale(LIB, objs):
        tmp = LIB + ".tmp%d" % os.getpid()   # link next to the target, then rename: a reader never sees a partial library
        cmd = [NVCC, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
        os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
