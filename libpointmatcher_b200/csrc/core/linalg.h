// core/linalg.h — the tiny dense solves that end each ICP iteration (K7) and the per-point 3x3
// eigen-solve of the surface-normal filter (K8).  One thread each; PM_HD so they can be unit
// tested on the host.  Matrices are column-major.
#pragma once
#include "common.h"

namespace pm {

#define PM_FLT_EPS 1.1920929e-07

// ------------------------------------------------------------------------------------------
// A x = b for the symmetric positive semi-definite 6x6 normal matrix of point-to-plane
// (replaces solvePossiblyUnderdeterminedLinearSystem, PointToPlane.cpp:108-161).
// Diagonally pivoted Cholesky P^T A P = L L^T, stopped when the next pivot falls below the
// relative threshold Eigen's FullPivHouseholderQR uses in float (6 * eps_float): full rank ->
// ordinary Cholesky solve (the reference's `A.llt().solve(b)`); rank r < 6 -> the minimum-norm
// solution x = P L (L^T L)^-2 L^T P^T b, which is what the reference's
// R1^T (R1 R1^T)^-1 Q1^T b evaluates to.  Returns the rank.
// ------------------------------------------------------------------------------------------
// (n x n, n <= 6: 6 for the full problem, 4 with force4DOF; the rank threshold is n * eps_float)
PM_HD int solve_psd(const double* A, const double* b, double* x, const int n) {
    double M[36], L[36];
    int perm[6];
    for (int i = 0; i < n * n; ++i) { M[i] = A[i]; L[i] = 0.0; }
    for (int i = 0; i < n; ++i) perm[i] = i;
    const double thr = (double)n * PM_FLT_EPS;
    double d0 = 0.0;
    int rank = 0;
    for (int k = 0; k < n; ++k) {
        int p = k;
        for (int i = k + 1; i < n; ++i)
            if (M[i + n * i] > M[p + n * p]) p = i;
        const double dmax = M[p + n * p];
        if (k == 0) d0 = dmax;
        if (!(dmax > thr * d0) || !(dmax > 0.0)) break;
        if (p != k) {
            for (int j = 0; j < n; ++j) { const double t = M[k + n * j]; M[k + n * j] = M[p + n * j]; M[p + n * j] = t; }
            for (int i = 0; i < n; ++i) { const double t = M[i + n * k]; M[i + n * k] = M[i + n * p]; M[i + n * p] = t; }
            for (int j = 0; j < k; ++j) { const double t = L[k + n * j]; L[k + n * j] = L[p + n * j]; L[p + n * j] = t; }
            const int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
        }
        const double lkk = sqrt(dmax);
        L[k + n * k] = lkk;
        for (int i = k + 1; i < n; ++i) L[i + n * k] = M[i + n * k] / lkk;
        for (int j = k + 1; j < n; ++j)
            for (int i = k + 1; i < n; ++i) M[i + n * j] -= L[i + n * k] * L[j + n * k];
        ++rank;
    }
    double c[6], z[6];
    for (int i = 0; i < n; ++i) { c[i] = b[perm[i]]; z[i] = 0.0; }
    if (rank == n) {
        double y[6];
        for (int i = 0; i < n; ++i) {
            double s = c[i];
            for (int j = 0; j < i; ++j) s -= L[i + n * j] * y[j];
            y[i] = s / L[i + n * i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double s = y[i];
            for (int j = i + 1; j < n; ++j) s -= L[j + n * i] * z[j];
            z[i] = s / L[i + n * i];
        }
    } else if (rank > 0) {
        const int r = rank;
        // G = L^T L (r x r), g = L^T c
        double G[25], g[5], C[25];
        for (int a = 0; a < r; ++a) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) s += L[i + n * a] * c[i];
            g[a] = s;
            for (int bcol = 0; bcol < r; ++bcol) {
                double t = 0.0;
                for (int i = 0; i < n; ++i) t += L[i + n * a] * L[i + n * bcol];
                G[a + 5 * bcol] = t;
            }
        }
        // Cholesky of G
        for (int i = 0; i < 25; ++i) C[i] = 0.0;
        for (int k = 0; k < r; ++k) {
            double v = G[k + 5 * k];
            for (int j = 0; j < k; ++j) v -= C[k + 5 * j] * C[k + 5 * j];
            const double ckk = sqrt(v);
            C[k + 5 * k] = ckk;
            for (int i = k + 1; i < r; ++i) {
                double s = G[i + 5 * k];
                for (int j = 0; j < k; ++j) s -= C[i + 5 * j] * C[k + 5 * j];
                C[i + 5 * k] = s / ckk;
            }
        }
        // two solves with G
        for (int rep = 0; rep < 2; ++rep) {
            double y[5];
            for (int i = 0; i < r; ++i) {
                double s = g[i];
                for (int j = 0; j < i; ++j) s -= C[i + 5 * j] * y[j];
                y[i] = s / C[i + 5 * i];
            }
            for (int i = r - 1; i >= 0; --i) {
                double s = y[i];
                for (int j = i + 1; j < r; ++j) s -= C[j + 5 * i] * g[j];
                g[i] = s / C[i + 5 * i];
            }
        }
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int a = 0; a < r; ++a) s += L[i + n * a] * g[a];
            z[i] = s;
        }
    }
    for (int i = 0; i < n; ++i) x[perm[i]] = z[i];
    return rank;
}
PM_HD int solve_psd6(const double* A, const double* b, double* x) { return solve_psd(A, b, x, 6); }

// ------------------------------------------------------------------------------------------
// Cyclic Jacobi eigen-decomposition of a symmetric 3x3 matrix (double).  A is destroyed;
// w = eigenvalues (unsorted), V = eigenvectors as columns.
// ------------------------------------------------------------------------------------------
PM_HD void jacobi_eig3(double* A, double* w, double* V) {
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 24; ++sweep) {
        const double off = A[3] * A[3] + A[6] * A[6] + A[7] * A[7];
        const double diag = A[0] * A[0] + A[4] * A[4] + A[8] * A[8];
        if (!(off > diag * 1e-30) || off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const double apq = A[p + 3 * q];
                if (apq == 0.0) continue;
                const double theta = (A[q + 3 * q] - A[p + 3 * p]) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[k + 3 * p], akq = A[k + 3 * q];
                    A[k + 3 * p] = c * akp - s * akq;
                    A[k + 3 * q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[p + 3 * k], aqk = A[q + 3 * k];
                    A[p + 3 * k] = c * apk - s * aqk;
                    A[q + 3 * k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k + 3 * p], vkq = V[k + 3 * q];
                    V[k + 3 * p] = c * vkp - s * vkq;
                    V[k + 3 * q] = s * vkp + c * vkq;
                }
            }
    }
    w[0] = A[0]; w[1] = A[4]; w[2] = A[8];
}

// ------------------------------------------------------------------------------------------
// Eigen-decomposition of the symmetric 2x2 [a b; b c] in closed form (2-D clouds): w ascending,
// V = unit eigenvectors as columns (column-major 2x2).
// ------------------------------------------------------------------------------------------
PM_HD void sym_eig2(double a, double b, double c, double* w, double* V) {
    const double half = 0.5 * (a + c), d = 0.5 * (a - c);
    const double r = sqrt(d * d + b * b);
    w[0] = half - r;
    w[1] = half + r;
    // eigenvector of w[1]: (b, w1 - a) or (w1 - c, b), whichever is better conditioned; the other one is orthogonal to it
    double x1 = b, y1 = w[1] - a;
    const double x2 = w[1] - c, y2 = b;
    if (x2 * x2 + y2 * y2 > x1 * x1 + y1 * y1) { x1 = x2; y1 = y2; }
    const double n1 = sqrt(x1 * x1 + y1 * y1);
    if (n1 > 0.0) { x1 /= n1; y1 /= n1; } else { x1 = 0.0; y1 = 1.0; }  // a multiple of the identity: any basis
    V[2] = x1; V[3] = y1;     // column 1: the larger eigenvalue
    V[0] = -y1; V[1] = x1;    // column 0: the smaller one
}

// ------------------------------------------------------------------------------------------
// Rotation of point-to-point: R = U V^T from the SVD of the 3x3 cross-covariance m, with the
// reflection fix of PointToPoint.cpp:82-93 (negate the last row of V^T when det(U V^T) < 0).
// V from the eigen-decomposition of m^T m, U = m V / sigma, rank-deficient columns completed.
// ------------------------------------------------------------------------------------------
// sv (optional): the singular values, descending, the last one negated when the reflection fix was
// applied (what PointToPointSimilarity.cpp:78-89 sums for the scale)
PM_HD void rotation_from_crosscov(const double* m, double* R, double* sv = nullptr) {
    double MtM[9], w[3], V[9];
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += m[k + 3 * i] * m[k + 3 * j];
            MtM[i + 3 * j] = s;
        }
    jacobi_eig3(MtM, w, V);
    // order eigenvalues descending
    int o[3] = {0, 1, 2};
    for (int a = 0; a < 2; ++a)
        for (int bq = a + 1; bq < 3; ++bq)
            if (w[o[bq]] > w[o[a]]) { const int t = o[a]; o[a] = o[bq]; o[bq] = t; }
    double Vs[9], U[9], sig[3];
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) Vs[i + 3 * j] = V[i + 3 * o[j]];
    // make V a proper orthonormal basis (Jacobi keeps it orthonormal; fix handedness is not needed)
    for (int j = 0; j < 3; ++j) {
        double u[3];
        for (int i = 0; i < 3; ++i) u[i] = m[i] * Vs[3 * j] + m[i + 3] * Vs[1 + 3 * j] + m[i + 6] * Vs[2 + 3 * j];
        sig[j] = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        for (int i = 0; i < 3; ++i) U[i + 3 * j] = u[i];
    }
    const double tiny = sig[0] * 1e-12;
    // Gram-Schmidt the columns of U in order of decreasing singular value
    if (sig[0] > 0.0) {
        for (int i = 0; i < 3; ++i) U[i] /= sig[0];
    } else {
        U[0] = 1.0; U[1] = 0.0; U[2] = 0.0;
    }
    {
        double dot = U[0] * U[3] + U[1] * U[4] + U[2] * U[5];
        double v[3] = {U[3] - dot * U[0], U[4] - dot * U[1], U[5] - dot * U[2]};
        double nv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (!(sig[1] > tiny) || !(nv > 0.0)) {
            // any unit vector orthogonal to column 0
            int mi = 0;
            if (fabs(U[1]) < fabs(U[mi])) mi = 1;
            if (fabs(U[2]) < fabs(U[mi])) mi = 2;
            double e[3] = {0.0, 0.0, 0.0};
            e[mi] = 1.0;
            dot = U[mi];
            v[0] = e[0] - dot * U[0]; v[1] = e[1] - dot * U[1]; v[2] = e[2] - dot * U[2];
            nv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        }
        U[3] = v[0] / nv; U[4] = v[1] / nv; U[5] = v[2] / nv;
    }
    {
        // third column: keep its measured direction when sigma_3 is significant, else complete
        const double cx = U[1] * U[5] - U[2] * U[4], cy = U[2] * U[3] - U[0] * U[5], cz = U[0] * U[4] - U[1] * U[3];
        double sgn = 1.0;
        if (sig[2] > tiny) sgn = (cx * U[6] + cy * U[7] + cz * U[8]) < 0.0 ? -1.0 : 1.0;
        U[6] = sgn * cx; U[7] = sgn * cy; U[8] = sgn * cz;
    }
    // R = U V^T ; if det < 0 negate last row of V^T (i.e. last column of V)
    for (int rep = 0; rep < 2; ++rep) {
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) R[i + 3 * j] = U[i] * Vs[j] + U[i + 3] * Vs[j + 3] + U[i + 6] * Vs[j + 6];
        const double det = R[0] * (R[4] * R[8] - R[7] * R[5]) - R[3] * (R[1] * R[8] - R[7] * R[2]) + R[6] * (R[1] * R[5] - R[4] * R[2]);
        if (det < 0.0 && rep == 0) { Vs[6] = -Vs[6]; Vs[7] = -Vs[7]; Vs[8] = -Vs[8]; sig[2] = -sig[2]; }
        else break;
    }
    if (sv) { sv[0] = sig[0]; sv[1] = sig[1]; sv[2] = sig[2]; }
}

// ------------------------------------------------------------------------------------------
// x = [rx ry rz tx ty tz] -> 4x4, as Eigen::AngleAxis(|r|, r/|r|) + translation
// (PointToPlane.cpp:250-292), float like the reference; zero rotation vector -> identity.
// ------------------------------------------------------------------------------------------
PM_HD void angle_axis_to_mat4(const float* x, Mat4& T) {
    const float n2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    const float angle = sqrtf(n2);
    float ax = x[0], ay = x[1], az = x[2];
    if (n2 > 0.f) { ax = x[0] / angle; ay = x[1] / angle; az = x[2] / angle; }
    const float s = sinf(angle), c = cosf(angle);
    const float sx = s * ax, sy = s * ay, sz = s * az;
    const float cx = (1.f - c) * ax, cy = (1.f - c) * ay, cz = (1.f - c) * az;
    mat4_identity(T);
    float tmp;
    tmp = cx * ay; T.m[0 + 4 * 1] = tmp - sz; T.m[1 + 4 * 0] = tmp + sz;
    tmp = cx * az; T.m[0 + 4 * 2] = tmp + sy; T.m[2 + 4 * 0] = tmp - sy;
    tmp = cy * az; T.m[1 + 4 * 2] = tmp - sx; T.m[2 + 4 * 1] = tmp + sx;
    T.m[0] = cx * ax + c; T.m[5] = cy * ay + c; T.m[10] = cz * az + c;
    T.m[12] = x[3]; T.m[13] = x[4]; T.m[14] = x[5];
    bool bad = false;
    for (int i = 0; i < 16; ++i) bad = bad || (T.m[i] != T.m[i]);
    if (bad)
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) T.m[i + 4 * j] = (i == j) ? 1.f : 0.f;
}

// ------------------------------------------------------------------------------------------
// Quaternion helpers of DifferentialTransformationChecker (TransformationCheckersImpl.cpp:
// 107-158): Eigen's matrix->quaternion conversion and QuaternionBase::angularDistance.
// ------------------------------------------------------------------------------------------
struct Quat {
    float w, x, y, z;
};
PM_HD Quat quat_from_mat4(const Mat4& T) {
    const float* m = T.m;
#define PM_M(i, j) m[(i) + 4 * (j)]
    Quat q;
    float t = PM_M(0, 0) + PM_M(1, 1) + PM_M(2, 2);
    if (t > 0.f) {
        t = sqrtf(t + 1.f);
        q.w = 0.5f * t;
        t = 0.5f / t;
        q.x = (PM_M(2, 1) - PM_M(1, 2)) * t;
        q.y = (PM_M(0, 2) - PM_M(2, 0)) * t;
        q.z = (PM_M(1, 0) - PM_M(0, 1)) * t;
    } else {
        int i = 0;
        if (PM_M(1, 1) > PM_M(0, 0)) i = 1;
        if (PM_M(2, 2) > PM_M(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrtf(PM_M(i, i) - PM_M(j, j) - PM_M(k, k) + 1.f);
        float v[3];
        v[i] = 0.5f * t;
        t = 0.5f / t;
        q.w = (PM_M(k, j) - PM_M(j, k)) * t;
        v[j] = (PM_M(j, i) + PM_M(i, j)) * t;
        v[k] = (PM_M(k, i) + PM_M(i, k)) * t;
        q.x = v[0]; q.y = v[1]; q.z = v[2];
    }
#undef PM_M
    return q;
}
PM_HD float quat_angular_distance(const Quat& a, const Quat& b) {
    const float dw = a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z;
    const float dx = -a.w * b.x + a.x * b.w - a.y * b.z + a.z * b.y;
    const float dy = -a.w * b.y + a.y * b.w - a.z * b.x + a.x * b.z;
    const float dz = -a.w * b.z + a.z * b.w - a.x * b.y + a.y * b.x;
    return 2.f * atan2f(sqrtf(dx * dx + dy * dy + dz * dz), fabsf(dw));
}

// ------------------------------------------------------------------------------------------
// Rank of a 3x3 float matrix exactly as Eigen's FullPivHouseholderQR decides it
// (SurfaceNormal.cpp:193 `C.fullPivHouseholderQr().rank()`): full pivoting, Householder
// reflections, pivots compared with 3 * eps * max|pivot|.  A is destroyed.
// ------------------------------------------------------------------------------------------
// `size`: the matrix is size x size in the top-left corner of the 3x3 (the rest zero): a 2-D cloud's 2x2 scatter matrix
// takes the same pivots as its zero-padded 3x3, with Eigen's threshold for a 2x2 (eps * diagonalSize).
PM_HD int fullpiv_qr_rank3(float* A, int size = 3) {
    const int n = 3;
    const float precision = (float)PM_FLT_EPS * (float)size;
    float biggest = 0.f, maxpivot = 0.f;
    float diag[3] = {0.f, 0.f, 0.f};
    int nonzero = n;
    for (int k = 0; k < n; ++k) {
        int br = k, bc = k;
        float big = -1.f;
        for (int j = k; j < n; ++j)
            for (int i = k; i < n; ++i) {
                const float v = fabsf(A[i + n * j]);
                if (v > big) { big = v; br = i; bc = j; }
            }
        if (k == 0) biggest = big;
        if (fabsf(big) <= fabsf(biggest) * precision) { nonzero = k; break; }
        if (br != k)
            for (int j = k; j < n; ++j) { const float t = A[k + n * j]; A[k + n * j] = A[br + n * j]; A[br + n * j] = t; }
        if (bc != k)
            for (int i = 0; i < n; ++i) { const float t = A[i + n * k]; A[i + n * k] = A[i + n * bc]; A[i + n * bc] = t; }
        float tailSq = 0.f;
        for (int i = k + 1; i < n; ++i) tailSq = fadd(tailSq, fmul(A[i + n * k], A[i + n * k]));
        const float c0 = A[k + n * k];
        float beta, tau;
        if (tailSq <= FLT_MIN) {
            tau = 0.f;
            beta = c0;
            for (int i = k + 1; i < n; ++i) A[i + n * k] = 0.f;
        } else {
            beta = sqrtf(fadd(fmul(c0, c0), tailSq));
            if (c0 >= 0.f) beta = -beta;
            for (int i = k + 1; i < n; ++i) A[i + n * k] = A[i + n * k] / fsub(c0, beta);
            tau = fsub(beta, c0) / beta;
        }
        diag[k] = beta;
        if (fabsf(beta) > maxpivot) maxpivot = fabsf(beta);
        for (int j = k + 1; j < n; ++j) {
            float tmp = A[k + n * j];
            for (int i = k + 1; i < n; ++i) tmp = fadd(tmp, fmul(A[i + n * k], A[i + n * j]));
            A[k + n * j] = fsub(A[k + n * j], fmul(tau, tmp));
            for (int i = k + 1; i < n; ++i) A[i + n * j] = fsub(A[i + n * j], fmul(fmul(tau, A[i + n * k]), tmp));
        }
    }
    const float thr = maxpivot * precision;
    int r = 0;
    for (int i = 0; i < nonzero; ++i) r += (fabsf(diag[i]) > thr) ? 1 : 0;
    return r;
}

}  // namespace pm
