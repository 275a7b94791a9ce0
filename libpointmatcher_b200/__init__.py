"""libpointmatcher_b200 — B200-native ICP hot path behind libpointmatcher's plugin surface.

    capi   ctypes binding of include/pmgpu.h (libpmgpu.so: hand-written sm_100a kernels)
    pm     Python mirror of the reference's module classes / registrars / ICP driver
    synth  deterministic Velodyne-like synthetic clouds (bench + tests)
    build  nvcc build of libpmgpu.so

The CUDA extension is mandatory: `capi` raises ImportError when libpmgpu.so has not been built,
and every module raises when no CUDA device is present.  There is no CPU fallback.
"""
__version__ = "0.1.0"
