// normals.cuh — K8: the per-point part of SurfaceNormalDataPointsFilter::inPlaceFilter
// (DataPointsFilters/SurfaceNormal.cpp:166-252, utils/utils.h:105-139) as the EPILOGUE of the
// self-kNN search (knn.cu).
//
// When a self-query finishes, its k neighbours sit in the searching thread's registers.  The thread
// gathers their coordinates (the leaves it has just scanned: L1/L2 hits), forms the float mean and
// the un-normalised 3x3 scatter matrix in the reference's accumulation order, runs Eigen's
// FullPivHouseholderQR rank test in float (core/linalg.h) and a cyclic-Jacobi symmetric eigen-solve
// in fp64, and writes one float4 normal (+ the optional descriptors a caller asked for).  The
// k x N ids / dists matrices of the search are never written or read back: algorithmic traffic is
// 16 B read + 16 B written per point (+ 4 k B only when keepMatchedIds is set).
// Normal = eigenvector of the smallest eigenvalue, unit norm, sign arbitrary (as Eigen's general
// EigenSolver leaves it), clamped to [-1, 1].  Eigenvalues / eigenvectors are reported in ascending
// eigenvalue order whether or not sortEigen is set (the reference's unsorted order is whatever
// Eigen's Hessenberg QR produces: unpinned).
// smoothNormals (SurfaceNormal.cpp:259-283) replaces every normal by the sign-aligned mean of its neighbours' normals IN
// PLACE, point after point, so point i sees the already smoothed normals of every neighbour j < i: a serial recurrence
// by definition.  It runs on the host (api.cu smooth_normals_host), once per cloud like the reference's, on the normals
// and exact neighbour ids this epilogue wrote.
#pragma once
#include "core/linalg.h"
#include "pmgpu_internal.cuh"

namespace pm {

// where the epilogue writes; by value in the kernel arguments
struct NormalsSink {
    const f4* pts;       // the cloud in the caller's column order: neighbour ids index it
    f4* normals4;        // (nx, ny, nz, 0) per point, may be null
    float* densities;    // optional, 1 per point
    float* eig_values;   // optional, 3 per point
    float* eig_vectors;  // optional, 9 per point (row-major serialisation, utils.h:89-103)
    float* mean_dists;   // optional, 1 per point
    float* matched_ids;  // optional, k per point, as float (SurfaceNormal.cpp:254-257)
    int32_t* ids_i32;    // optional, k per point, exact (-1 = no neighbour): what smoothNormals walks on the host
    int* degenerate;     // counter
    int dim2;            // the cloud is 2-D (z = 0): 2x2 scatter matrix, eig_values 2 and eig_vectors 4 per point (SurfaceNormal.cpp featDim - 1 = 2)
    int by_position;     // 1: outputs indexed by the query's leaf-order position (sharded K8: contiguous per rank), 0: by original column
    // stage-2 hand-over of the candidates stage 1 found for queries that ran out of budget
    int32_t* scratch_ids;
    float* scratch_d;
    unsigned scratch_cap;  // queries
};

// knn.cu
int launch_knn_normals(pmgpu_ctx* ctx, const TreeView& tree, int pos_lo, int pos_hi, int k, float max_r2, const NormalsSink& sink);

// `best`: the finished search of the point whose outputs go to index `out` (ascending (dist, id); unfilled slots carry
// PM_NO_ID / +inf); (px, py, pz) the point itself.  Returns true when the point is degenerate.
template <int KMAX>
__device__ __forceinline__ bool normals_epilogue(const TopK<KMAX>& best, int k, const NormalsSink& ns, size_t out, float px, float py, float pz) {
    // mean of the valid neighbours (SurfaceNormal.cpp:173-184); the point itself is one of them
    float sx = 0.f, sy = 0.f, sz = 0.f;
    int real_knn = 0;
    static_for<0, KMAX>([&](auto J) {
        if (J < k && best.I(J) != PM_NO_ID && best.D(J) != pm_inf()) {
            const f4 p = __ldg(ns.pts + best.I(J));
            sx = fadd(sx, p.x); sy = fadd(sy, p.y); sz = fadd(sz, p.z);
            ++real_knn;
        }
    });
    const float cnt = (float)real_knn;
    const float mx = sx / cnt, my = sy / cnt, mz = sz / cnt;
    // C = NN * NN^T (un-normalised), and the largest neighbour radius for the density
    float c00 = 0.f, c01 = 0.f, c02 = 0.f, c11 = 0.f, c12 = 0.f, c22 = 0.f, max_norm = 0.f;
    static_for<0, KMAX>([&](auto J) {
        if (J < k && best.I(J) != PM_NO_ID && best.D(J) != pm_inf()) {
            const f4 p = __ldg(ns.pts + best.I(J));
            const float dx = fsub(p.x, mx), dy = fsub(p.y, my), dz = fsub(p.z, mz);
            c00 = fadd(c00, fmul(dx, dx)); c01 = fadd(c01, fmul(dx, dy)); c02 = fadd(c02, fmul(dx, dz));
            c11 = fadd(c11, fmul(dy, dy)); c12 = fadd(c12, fmul(dy, dz)); c22 = fadd(c22, fmul(dz, dz));
            max_norm = fmaxf(max_norm, sqrtf(fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz))));
        }
    });
    if (ns.matched_ids) {
        static_for<0, KMAX>([&](auto J) {
            if (J < k) {
                const bool valid = best.I(J) != PM_NO_ID && best.D(J) != pm_inf();
                ns.matched_ids[out * k + J] = (float)(valid ? best.I(J) : -1);
            }
        });
    }
    if (ns.ids_i32) {
        static_for<0, KMAX>([&](auto J) {
            if (J < k) {
                const bool valid = best.I(J) != PM_NO_ID && best.D(J) != pm_inf();
                ns.ids_i32[out * k + J] = valid ? best.I(J) : -1;
            }
        });
    }
    // the rank test and the eigen-solve only run for the outputs that need them (SurfaceNormal.cpp:190-218): with
    // densities / mean distances alone a point is never marked degenerate
    const bool wants_eigen = ns.normals4 || ns.eig_values || ns.eig_vectors;
    bool is_degenerate = false;
    float va[3] = {0.f, 0.f, 0.f};
    float ve[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // column-major, ascending eigenvalue
    if (wants_eigen) {
        float Cq[9] = {c00, c01, c02, c01, c11, c12, c02, c12, c22};
        if (ns.dim2) {
            // 2-D cloud: C is 2x2, the test is rank + 1 >= 2 with the 2x2 threshold, the eigen-solve is closed-form
            if (real_knn > 0 && fullpiv_qr_rank3(Cq, 2) + 1 >= 2) {
                double w[2], V[4];
                sym_eig2((double)c00, (double)c01, (double)c11, w, V);
                va[0] = (float)w[0]; va[1] = (float)w[1];
                ve[0] = (float)V[0]; ve[1] = (float)V[1];  // column 0 (3-row layout: the third row stays 0)
                ve[3] = (float)V[2]; ve[4] = (float)V[3];
            } else {
                is_degenerate = true;
            }
        } else if (real_knn > 0 && fullpiv_qr_rank3(Cq) + 1 >= 3) {
            double A[9] = {c00, c01, c02, c01, c11, c12, c02, c12, c22}, w[3], V[9];
            jacobi_eig3(A, w, V);
            int o[3] = {0, 1, 2};
            for (int a = 0; a < 2; ++a)
                for (int b = a + 1; b < 3; ++b)
                    if (w[o[b]] < w[o[a]]) { const int t = o[a]; o[a] = o[b]; o[b] = t; }
            for (int c = 0; c < 3; ++c) {
                va[c] = (float)w[o[c]];
                for (int r = 0; r < 3; ++r) ve[r + 3 * c] = (float)V[r + 3 * o[c]];
            }
        } else {
            is_degenerate = true;
        }
    }
    if (ns.normals4) ns.normals4[out] = make_float4(fminf(1.f, fmaxf(-1.f, ve[0])), fminf(1.f, fmaxf(-1.f, ve[1])), fminf(1.f, fmaxf(-1.f, ve[2])), 0.f);
    if (ns.densities) {
        // computeDensity (utils.h:105-120): the volume is evaluated in double, stored as float
        if (is_degenerate) ns.densities[out] = 0.f;
        else {
            const double r = (double)max_norm;
            const float volume = (float)((4. / 3.) * 3.14159265358979323846 * (r * r * r));
            ns.densities[out] = (float)real_knn / volume;
        }
    }
    const int dn = ns.dim2 ? 2 : 3;
    if (ns.eig_values)
        for (int r = 0; r < dn; ++r) ns.eig_values[dn * out + r] = va[r];
    if (ns.eig_vectors)  // serializeEigVec: row-major (utils.h:89-103)
        for (int r = 0; r < dn; ++r)
            for (int c = 0; c < dn; ++c) ns.eig_vectors[dn * dn * out + dn * r + c] = ve[r + 3 * c];
    if (ns.mean_dists) {
        if (is_degenerate) ns.mean_dists[out] = 18446744073709551615.f;  // numeric_limits<size_t>::max() as float, SurfaceNormal.cpp:245
        else {
            const float dx = fsub(px, mx), dy = fsub(py, my), dz = fsub(pz, mz);
            ns.mean_dists[out] = sqrtf(fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)));
        }
    }
    return is_degenerate;
}

}  // namespace pm
