// core/common.h — host/device building blocks shared by the CUDA kernels.
//
// Everything here is PM_HD (host + device) so the per-thread algorithms (tree addressing,
// traversal, top-k insertion, small solves) can also be unit-tested by a plain g++ harness
// (tests/emu/) without a GPU.  The product only ever runs them inside kernels.
#pragma once

#include <stdint.h>
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define PM_HD __host__ __device__ __forceinline__
#define PM_D __device__ __forceinline__
typedef float4 f4;
typedef float2 f2;
#else
#define PM_HD inline
struct alignas(16) f4 { float x, y, z, w; };
struct alignas(8) f2 { float x, y; };
#endif

namespace pm {

#define PM_INF_BITS 0x7f800000u

PM_HD f4 make_f4(float x, float y, float z, float w) {
    f4 r;
    r.x = x; r.y = y; r.z = z; r.w = w;
    return r;
}

PM_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c;
    c.f = f;
    return c.u;
#endif
}
PM_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c;
    c.u = u;
    return c.f;
#endif
}
PM_HD float pm_inf() { return u2f(PM_INF_BITS); }

// Single-rounded float ops.  The reference (and libnabo) are built without FMA contraction
// (CMakeLists.txt:69-71: plain -O3, SSE2), so distances and transformed coordinates must never
// be fused on the device either; the host build of the harness uses -ffp-contract=off.
PM_HD float fmul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
PM_HD float fadd(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
PM_HD float fsub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}

// Squared distance exactly as libnabo accumulates it: ((dx*dx + dy*dy) + dz*dz).
PM_HD float dist2(float qx, float qy, float qz, float px, float py, float pz) {
    const float dx = fsub(qx, px), dy = fsub(qy, py), dz = fsub(qz, pz);
    return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
}

// Order-preserving float <-> uint32 map (for radix sorting and integer atomics).
PM_HD uint32_t float_ord(float f) {
    const uint32_t b = f2u(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
PM_HD float ord_float(uint32_t o) { return u2f((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }

// Rigid transform of one homogeneous point, the depth-4 GEMM order of
// `parameters * input.features` (TransformationsImpl.cpp:69): ((T0 x + T1 y) + T2 z) + T3 w,
// T column-major.
struct Mat4 {
    float m[16];
};
PM_HD f4 transform_point(const Mat4& T, f4 p) {
    f4 r;
    r.x = fadd(fadd(fadd(fmul(T.m[0], p.x), fmul(T.m[4], p.y)), fmul(T.m[8], p.z)), fmul(T.m[12], p.w));
    r.y = fadd(fadd(fadd(fmul(T.m[1], p.x), fmul(T.m[5], p.y)), fmul(T.m[9], p.z)), fmul(T.m[13], p.w));
    r.z = fadd(fadd(fadd(fmul(T.m[2], p.x), fmul(T.m[6], p.y)), fmul(T.m[10], p.z)), fmul(T.m[14], p.w));
    r.w = fadd(fadd(fadd(fmul(T.m[3], p.x), fmul(T.m[7], p.y)), fmul(T.m[11], p.z)), fmul(T.m[15], p.w));
    return r;
}
// C = A * B, same accumulation order (T_iter = dT * T_iter, ICP.cpp:411-412)
PM_HD void mat4_mul(const Mat4& A, const Mat4& B, Mat4& C) {
    Mat4 t;
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            float acc = fmul(A.m[i], B.m[4 * j]);
            acc = fadd(acc, fmul(A.m[i + 4], B.m[1 + 4 * j]));
            acc = fadd(acc, fmul(A.m[i + 8], B.m[2 + 4 * j]));
            acc = fadd(acc, fmul(A.m[i + 12], B.m[3 + 4 * j]));
            t.m[i + 4 * j] = acc;
        }
    C = t;
}
PM_HD void mat4_identity(Mat4& T) {
    for (int i = 0; i < 16; ++i) T.m[i] = (i % 5 == 0) ? 1.f : 0.f;
}
// RigidTransformation::checkParameters (TransformationsImpl.cpp:90-105)
PM_HD bool mat4_is_rigid(const Mat4& T) {
    const float* m = T.m;
    const float det = m[0] * (m[5] * m[10] - m[9] * m[6]) - m[4] * (m[1] * m[10] - m[9] * m[2]) + m[8] * (m[1] * m[6] - m[5] * m[2]);
    return !(fabsf(1.f - det) > 0.001f);
}

}  // namespace pm
