// normals.cu — K8: the per-point part of SurfaceNormalDataPointsFilter::inPlaceFilter
// (DataPointsFilters/SurfaceNormal.cpp:166-252, utils/utils.h:105-139) after the self-kNN of K2.
//
// One thread per point: gather the valid neighbours twice (mean, then scatter matrix; the second
// gather hits L1/L2), float mean and un-normalised 3x3 scatter matrix in the reference's
// accumulation order, Eigen's FullPivHouseholderQR rank test in float (core/linalg.h), then a
// cyclic-Jacobi symmetric eigen-solve in fp64.  Normal = eigenvector of the smallest eigenvalue,
// unit norm, sign arbitrary (as Eigen's general EigenSolver leaves it), clamped to [-1, 1].
// Eigenvalues / eigenvectors are reported in ascending eigenvalue order.
#include "core/linalg.h"
#include "pmgpu_internal.cuh"

namespace pm {

namespace {

__global__ void __launch_bounds__(128) normals_kernel(const f4* __restrict__ pts, int n, const int32_t* __restrict__ ids,
                                                      const float* __restrict__ dists, int knn, f4* __restrict__ normals4,
                                                      float* __restrict__ densities, float* __restrict__ eig_values,
                                                      float* __restrict__ eig_vectors, float* __restrict__ mean_dists, int* degenerate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool is_degenerate = false;
    if (i < n) {
        const int32_t* my_ids = ids + (size_t)i * knn;
        const float* my_d = dists + (size_t)i * knn;
        // mean of the valid neighbours (SurfaceNormal.cpp:173-184); the point itself is one of them
        float sx = 0.f, sy = 0.f, sz = 0.f;
        int real_knn = 0;
        for (int j = 0; j < knn; ++j) {
            if (my_d[j] != pm_inf()) {
                const f4 p = __ldg(pts + my_ids[j]);
                sx = fadd(sx, p.x); sy = fadd(sy, p.y); sz = fadd(sz, p.z);
                ++real_knn;
            }
        }
        const float inv_n = (float)real_knn;
        const float mx = sx / inv_n, my = sy / inv_n, mz = sz / inv_n;
        // C = NN * NN^T (un-normalised), and the largest neighbour radius for the density
        float c00 = 0.f, c01 = 0.f, c02 = 0.f, c11 = 0.f, c12 = 0.f, c22 = 0.f, max_norm = 0.f;
        for (int j = 0; j < knn; ++j) {
            if (my_d[j] != pm_inf()) {
                const f4 p = __ldg(pts + my_ids[j]);
                const float dx = fsub(p.x, mx), dy = fsub(p.y, my), dz = fsub(p.z, mz);
                c00 = fadd(c00, fmul(dx, dx)); c01 = fadd(c01, fmul(dx, dy)); c02 = fadd(c02, fmul(dx, dz));
                c11 = fadd(c11, fmul(dy, dy)); c12 = fadd(c12, fmul(dy, dz)); c22 = fadd(c22, fmul(dz, dz));
                max_norm = fmaxf(max_norm, sqrtf(fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz))));
            }
        }
        float Cq[9] = {c00, c01, c02, c01, c11, c12, c02, c12, c22};
        float va[3] = {0.f, 0.f, 0.f};
        float ve[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // column-major, ascending eigenvalue
        if (real_knn > 0 && fullpiv_qr_rank3(Cq) + 1 >= 3) {
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
        if (normals4) {
            normals4[i] = make_float4(fminf(1.f, fmaxf(-1.f, ve[0])), fminf(1.f, fmaxf(-1.f, ve[1])), fminf(1.f, fmaxf(-1.f, ve[2])), 0.f);
        }
        if (densities) {
            // computeDensity (utils.h:105-120): the volume is evaluated in double, stored as float
            if (is_degenerate) densities[i] = 0.f;
            else {
                const double r = (double)max_norm;
                const float volume = (float)((4. / 3.) * 3.14159265358979323846 * (r * r * r));
                densities[i] = (float)real_knn / volume;
            }
        }
        if (eig_values)
            for (int r = 0; r < 3; ++r) eig_values[3 * (size_t)i + r] = va[r];
        if (eig_vectors)  // serializeEigVec: row-major (utils.h:89-103)
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) eig_vectors[9 * (size_t)i + 3 * r + c] = ve[r + 3 * c];
        if (mean_dists) {
            if (is_degenerate) mean_dists[i] = 18446744073709551615.f;  // numeric_limits<size_t>::max() as float, SurfaceNormal.cpp:245
            else {
                const f4 p = pts[i];
                const float dx = fsub(p.x, mx), dy = fsub(p.y, my), dz = fsub(p.z, mz);
                mean_dists[i] = sqrtf(fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz)));
            }
        }
    }
    const unsigned deg = __ballot_sync(0xffffffffu, is_degenerate);
    if ((threadIdx.x & 31) == 0 && deg) atomicAdd(degenerate, __popc(deg));
}

}  // namespace

int launch_normals(pmgpu_ctx* ctx, const f4* pts, int n, const int32_t* ids, const float* dists, int knn, int flags, f4* normals4, float* densities,
                   float* eig_values, float* eig_vectors, float* mean_dists) {
    (void)flags;
    const int B = 128;
    if (n == 0) return PMGPU_OK;
    normals_kernel<<<(n + B - 1) / B, B, 0, ctx->stream>>>(pts, n, ids, dists, knn, normals4, densities, eig_values, eig_vectors, mean_dists,
                                                          &ctx->state->degenerate);
    ctx->launches += 1;
    PM_CUDA_TRY(ctx, cudaGetLastError());
    return PMGPU_OK;
}

}  // namespace pm
