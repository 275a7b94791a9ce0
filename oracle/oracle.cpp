/*
 * oracle/oracle.cpp — CPU restatement of the libpointmatcher ICP hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Never linked into, imported by, or called from the
 * product (libpointmatcher_b200/, include/).  It is the parity checker and the timed CPU
 * baseline ("port") because the reference cannot be compiled here (no Eigen/Boost/libnabo).
 *
 * PARITY UNPINNED for kNN indices: libnabo is not under /root/reference; its published
 * algorithm (brute force, and the bucketed kd-tree with implicit bounds + sorted linear heap,
 * libnabo >= 1.0.7 `KDTreeUnbalancedPtInLeavesImplicitBoundsStackOpt`) is restated below.
 * Transforms are pinned against the reference's golden .ref_trans files, validT3d and its
 * known-answer tests (tests/test_reference_goldens.py, tests/test_oracle_golden.py).
 *
 * Float semantics: compiled with -ffp-contract=off so every float op rounds once, like the
 * reference's default -O3 / SSE2 build (CMakeLists.txt:69-71).
 *
 * All matrices are column-major (Eigen default), clouds are rows x N with rows == 4
 * (x, y, z, w) as produced by the reference loaders (IO.cpp:999-1008).
 */
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const float kInf = std::numeric_limits<float>::infinity();
double g_timings[4] = {0, 0, 0, 0};  // last orc_icp: {build s, loop s, match s, iterations}
inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ------------------------------------------------------------------------------------------
// small dense helpers (column-major, runtime n <= 6)
// ------------------------------------------------------------------------------------------
template <typename S>
struct Mat {
    int r, c;
    std::vector<S> d;
    Mat() : r(0), c(0) {}
    Mat(int r_, int c_) : r(r_), c(c_), d(size_t(r_) * c_, S(0)) {}
    S& operator()(int i, int j) { return d[size_t(j) * r + i]; }
    S operator()(int i, int j) const { return d[size_t(j) * r + i]; }
    static Mat identity(int n) {
        Mat m(n, n);
        for (int i = 0; i < n; ++i) m(i, i) = S(1);
        return m;
    }
};

template <typename S>
Mat<S> mul(const Mat<S>& a, const Mat<S>& b) {
    Mat<S> o(a.r, b.c);
    for (int j = 0; j < b.c; ++j)
        for (int i = 0; i < a.r; ++i) {
            S acc = S(0);
            for (int k = 0; k < a.c; ++k) acc += a(i, k) * b(k, j);
            o(i, j) = acc;
        }
    return o;
}

template <typename S>
Mat<S> transpose(const Mat<S>& a) {
    Mat<S> o(a.c, a.r);
    for (int j = 0; j < a.c; ++j)
        for (int i = 0; i < a.r; ++i) o(j, i) = a(i, j);
    return o;
}

// Cholesky solve, the unblocked algorithm Eigen's LLT runs for small matrices
// (used by PointToPlane.cpp:159 `A.llt().solve(b)`).  Returns false if not positive definite;
// Eigen would still "solve" with garbage, the caller decides.
template <typename S>
bool llt_solve(const Mat<S>& A, const std::vector<S>& b, std::vector<S>& x) {
    const int n = A.r;
    Mat<S> L(n, n);
    bool ok = true;
    for (int k = 0; k < n; ++k) {
        S v = A(k, k);
        for (int j = 0; j < k; ++j) v -= L(k, j) * L(k, j);
        if (!(v > S(0))) ok = false;
        const S lkk = std::sqrt(v);
        L(k, k) = lkk;
        for (int i = k + 1; i < n; ++i) {
            S s = A(i, k);
            for (int j = 0; j < k; ++j) s -= L(i, j) * L(k, j);
            L(i, k) = s / lkk;
        }
    }
    std::vector<S> y(n);
    for (int i = 0; i < n; ++i) {
        S s = b[i];
        for (int j = 0; j < i; ++j) s -= L(i, j) * y[j];
        y[i] = s / L(i, i);
    }
    x.assign(n, S(0));
    for (int i = n - 1; i >= 0; --i) {
        S s = y[i];
        for (int j = i + 1; j < n; ++j) s -= L(j, i) * x[j];
        x[i] = s / L(i, i);
    }
    return ok;
}

// Full-pivoting Householder QR as Eigen's FullPivHouseholderQR computes it (the rank decision
// of PointToPlane.cpp:118 and SurfaceNormal.cpp:193 depends on its pivot thresholds).
template <typename S>
struct FullPivQR {
    int n;
    Mat<S> qr;
    std::vector<S> hcoef;
    std::vector<int> rowT, colT;
    int nonzero_pivots;
    S maxpivot;

    explicit FullPivQR(const Mat<S>& A) : n(A.r), qr(A), hcoef(A.r, S(0)), rowT(A.r), colT(A.r) {
        const S eps = std::numeric_limits<S>::epsilon();
        const S precision = eps * S(n);
        nonzero_pivots = n;
        maxpivot = S(0);
        S biggest = S(0);
        for (int k = 0; k < n; ++k) {
            // biggest |coef| in the bottom-right corner; column-major visit, first maximum wins
            int br = k, bc = k;
            S big = S(-1);
            for (int j = k; j < n; ++j)
                for (int i = k; i < n; ++i) {
                    const S v = std::fabs(qr(i, j));
                    if (v > big) { big = v; br = i; bc = j; }
                }
            if (k == 0) biggest = big;
            if (std::fabs(big) <= std::fabs(biggest) * precision) {  // isMuchSmallerThan
                nonzero_pivots = k;
                for (int i = k; i < n; ++i) { rowT[i] = i; colT[i] = i; hcoef[i] = S(0); }
                break;
            }
            rowT[k] = br;
            colT[k] = bc;
            if (k != br)
                for (int j = k; j < n; ++j) std::swap(qr(k, j), qr(br, j));
            if (k != bc)
                for (int i = 0; i < n; ++i) std::swap(qr(i, k), qr(i, bc));
            // makeHouseholderInPlace on qr(k:n, k)
            S tailSq = S(0);
            for (int i = k + 1; i < n; ++i) tailSq += qr(i, k) * qr(i, k);
            const S c0 = qr(k, k);
            S beta, tau;
            if (tailSq <= std::numeric_limits<S>::min()) {
                tau = S(0);
                beta = c0;
                for (int i = k + 1; i < n; ++i) qr(i, k) = S(0);
            } else {
                beta = std::sqrt(c0 * c0 + tailSq);
                if (c0 >= S(0)) beta = -beta;
                for (int i = k + 1; i < n; ++i) qr(i, k) /= (c0 - beta);
                tau = (beta - c0) / beta;
            }
            hcoef[k] = tau;
            qr(k, k) = beta;
            if (std::fabs(beta) > maxpivot) maxpivot = std::fabs(beta);
            // apply H_k to the trailing columns
            for (int j = k + 1; j < n; ++j) {
                S tmp = qr(k, j);
                for (int i = k + 1; i < n; ++i) tmp += qr(i, k) * qr(i, j);
                qr(k, j) -= tau * tmp;
                for (int i = k + 1; i < n; ++i) qr(i, j) -= tau * qr(i, k) * tmp;
            }
        }
    }
    int rank() const {
        const S thr = std::fabs(maxpivot) * (std::numeric_limits<S>::epsilon() * S(n));
        int r = 0;
        for (int i = 0; i < nonzero_pivots; ++i) r += (std::fabs(qr(i, i)) > thr) ? 1 : 0;
        return r;
    }
    bool invertible() const { return rank() == n; }
    Mat<S> matrixQ() const {
        Mat<S> Q = Mat<S>::identity(n);
        for (int k = n - 1; k >= 0; --k) {
            const S tau = hcoef[k];
            for (int j = k; j < n; ++j) {
                S tmp = Q(k, j);
                for (int i = k + 1; i < n; ++i) tmp += qr(i, k) * Q(i, j);
                Q(k, j) -= tau * tmp;
                for (int i = k + 1; i < n; ++i) Q(i, j) -= tau * qr(i, k) * tmp;
            }
            if (rowT[k] != k)
                for (int j = 0; j < n; ++j) std::swap(Q(k, j), Q(rowT[k], j));
        }
        return Q;
    }
    // column permutation matrix P with A * P = Q * R
    Mat<S> colsPermutation() const {
        std::vector<int> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        // Eigen: m_cols_permutation.setIdentity(); for k: applyTranspositionOnTheRight(k, colT[k])
        for (int k = 0; k < n; ++k) std::swap(idx[k], idx[colT[k]]);
        Mat<S> P(n, n);
        for (int j = 0; j < n; ++j) P(idx[j], j) = S(1);
        return P;
    }
};

// cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (values unsorted).
template <typename S>
void jacobi_eig(Mat<S> A, std::vector<S>& w, Mat<S>& V) {
    const int n = A.r;
    V = Mat<S>::identity(n);
    for (int sweep = 0; sweep < 64; ++sweep) {
        S off = S(0), diag = S(0);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) (i == j ? diag : off) += A(i, j) * A(i, j);
        if (off <= std::numeric_limits<S>::min() || off <= diag * std::numeric_limits<S>::epsilon() * std::numeric_limits<S>::epsilon()) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const S apq = A(p, q);
                if (apq == S(0)) continue;
                const S theta = (A(q, q) - A(p, p)) / (S(2) * apq);
                const S t = (theta >= S(0) ? S(1) : S(-1)) / (std::fabs(theta) + std::sqrt(theta * theta + S(1)));
                const S c = S(1) / std::sqrt(t * t + S(1)), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const S akp = A(k, p), akq = A(k, q);
                    A(k, p) = c * akp - s * akq;
                    A(k, q) = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const S apk = A(p, k), aqk = A(q, k);
                    A(p, k) = c * apk - s * aqk;
                    A(q, k) = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const S vkp = V(k, p), vkq = V(k, q);
                    V(k, p) = c * vkp - s * vkq;
                    V(k, q) = s * vkp + c * vkq;
                }
            }
    }
    w.resize(n);
    for (int i = 0; i < n; ++i) w[i] = A(i, i);
}

// 3x3 SVD by one-sided (Hestenes) Jacobi: M = U diag(s) V^T, singular values sorted
// descending as Eigen's JacobiSVD returns them (PointToPoint.cpp:82).
template <typename S>
void svd3(const Mat<S>& M, Mat<S>& U, std::vector<S>& sv, Mat<S>& V) {
    Mat<S> B = M;
    V = Mat<S>::identity(3);
    for (int sweep = 0; sweep < 64; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                S alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; ++i) {
                    alpha += B(i, p) * B(i, p);
                    beta += B(i, q) * B(i, q);
                    gamma += B(i, p) * B(i, q);
                }
                if (std::fabs(gamma) <= std::numeric_limits<S>::epsilon() * std::sqrt(alpha * beta) || gamma == S(0)) continue;
                rotated = true;
                const S zeta = (beta - alpha) / (S(2) * gamma);
                const S t = (zeta >= S(0) ? S(1) : S(-1)) / (std::fabs(zeta) + std::sqrt(S(1) + zeta * zeta));
                const S c = S(1) / std::sqrt(S(1) + t * t), s = c * t;
                for (int i = 0; i < 3; ++i) {
                    const S bp = B(i, p), bq = B(i, q);
                    B(i, p) = c * bp - s * bq;
                    B(i, q) = s * bp + c * bq;
                    const S vp = V(i, p), vq = V(i, q);
                    V(i, p) = c * vp - s * vq;
                    V(i, q) = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    sv.assign(3, S(0));
    U = Mat<S>(3, 3);
    for (int j = 0; j < 3; ++j) {
        S nrm = 0;
        for (int i = 0; i < 3; ++i) nrm += B(i, j) * B(i, j);
        sv[j] = std::sqrt(nrm);
    }
    int order[3] = {0, 1, 2};
    std::sort(order, order + 3, [&](int a, int b) { return sv[a] > sv[b]; });
    Mat<S> Vs(3, 3);
    std::vector<S> ss(3);
    for (int j = 0; j < 3; ++j) {
        const int o = order[j];
        ss[j] = sv[o];
        for (int i = 0; i < 3; ++i) {
            Vs(i, j) = V(i, o);
            U(i, j) = sv[o] > S(0) ? B(i, o) / sv[o] : S(0);
        }
    }
    // complete U to an orthonormal basis when singular values vanish
    auto colnorm = [&](int j) { return std::sqrt(U(0, j) * U(0, j) + U(1, j) * U(1, j) + U(2, j) * U(2, j)); };
    if (colnorm(0) == S(0)) { U(0, 0) = 1; U(1, 0) = 0; U(2, 0) = 0; }
    if (colnorm(1) == S(0)) {
        // any unit vector orthogonal to column 0
        S a[3] = {U(0, 0), U(1, 0), U(2, 0)};
        int m = 0;
        if (std::fabs(a[1]) < std::fabs(a[m])) m = 1;
        if (std::fabs(a[2]) < std::fabs(a[m])) m = 2;
        S e[3] = {0, 0, 0};
        e[m] = 1;
        S dot = a[m];
        S v[3] = {e[0] - dot * a[0], e[1] - dot * a[1], e[2] - dot * a[2]};
        S nv = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        for (int i = 0; i < 3; ++i) U(i, 1) = v[i] / nv;
    }
    if (colnorm(2) == S(0)) {
        U(0, 2) = U(1, 0) * U(2, 1) - U(2, 0) * U(1, 1);
        U(1, 2) = U(2, 0) * U(0, 1) - U(0, 0) * U(2, 1);
        U(2, 2) = U(0, 0) * U(1, 1) - U(1, 0) * U(0, 1);
    }
    sv = ss;
    V = Vs;
}

template <typename S>
S det3(const Mat<S>& R) {
    return R(0, 0) * (R(1, 1) * R(2, 2) - R(1, 2) * R(2, 1)) -
           R(0, 1) * (R(1, 0) * R(2, 2) - R(1, 2) * R(2, 0)) +
           R(0, 2) * (R(1, 0) * R(2, 1) - R(1, 1) * R(2, 0));
}

// general inverse by Gauss-Jordan with partial pivoting (J_hessian.inverse(),
// PointToPlaneWithCov.cpp:155 — Eigen uses PartialPivLU for dynamic sizes).
template <typename S>
Mat<S> inverse(const Mat<S>& A) {
    const int n = A.r;
    Mat<S> M = A, I = Mat<S>::identity(n);
    for (int k = 0; k < n; ++k) {
        int piv = k;
        for (int i = k + 1; i < n; ++i)
            if (std::fabs(M(i, k)) > std::fabs(M(piv, k))) piv = i;
        if (piv != k)
            for (int j = 0; j < n; ++j) { std::swap(M(k, j), M(piv, j)); std::swap(I(k, j), I(piv, j)); }
        const S d = M(k, k);
        for (int j = 0; j < n; ++j) { M(k, j) /= d; I(k, j) /= d; }
        for (int i = 0; i < n; ++i) {
            if (i == k) continue;
            const S f = M(i, k);
            if (f == S(0)) continue;
            for (int j = 0; j < n; ++j) { M(i, j) -= f * M(k, j); I(i, j) -= f * I(k, j); }
        }
    }
    return I;
}

// 4x4 float helpers, column-major
inline void mat4_mul(const float* A, const float* B, float* O) {
    float t[16];
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) {
            // depth-4 GEMM, single accumulator, k ascending (Eigen GEBP without FMA)
            float acc = A[i + 0] * B[0 + 4 * j];
            acc = acc + A[i + 4] * B[1 + 4 * j];
            acc = acc + A[i + 8] * B[2 + 4 * j];
            acc = acc + A[i + 12] * B[3 + 4 * j];
            t[i + 4 * j] = acc;
        }
    std::memcpy(O, t, sizeof(t));
}
inline void mat4_identity(float* T) {
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
}

// ------------------------------------------------------------------------------------------
// kNN — libnabo semantics
// ------------------------------------------------------------------------------------------
// Sorted linear heap (libnabo IndexHeapBruteForceVector): data ascending by value, head =
// largest kept value; replaceHead keeps earlier-inserted entries in front of equal values.
struct LinearHeap {
    struct Entry { int index; float value; };
    std::vector<Entry> data;
    explicit LinearHeap(int k) : data(k) { reset(); }
    void reset() { for (auto& e : data) { e.index = -1; e.value = kInf; } }
    float headValue() const { return data.back().value; }
    void replaceHead(int index, float value) {
        size_t i;
        for (i = data.size() - 1; i > 0; --i) {
            if (data[i - 1].value > value) data[i] = data[i - 1];
            else break;
        }
        data[i].value = value;
        data[i].index = index;
    }
    void get(int32_t* ids, float* dists) const {
        for (size_t i = 0; i < data.size(); ++i) { ids[i] = data[i].index; dists[i] = data[i].value; }
    }
};

struct KdTree {
    static const int kBucketSize = 8;  // libnabo default bucketSize
    int rows = 0, dim = 0, n = 0;
    const float* cloud = nullptr;
    struct Node {
        int cutDim;       // -1: leaf
        float cutVal;
        int rightChild;   // inner: index of right child (left child is this + 1)
        int bucketStart;  // leaf
        int bucketSize;
    };
    struct BucketEntry { float pt[3]; int index; };
    std::vector<Node> nodes;
    std::vector<BucketEntry> buckets;

    float coord(int d, int i) const { return cloud[size_t(i) * rows + d]; }

    int build(std::vector<int>::iterator first, std::vector<int>::iterator last, float* minV, float* maxV) {
        const int count = int(last - first);
        const int pos = int(nodes.size());
        if (count <= kBucketSize) {
            Node leaf{-1, 0.f, -1, int(buckets.size()), count};
            for (auto it = first; it != last; ++it) {
                BucketEntry e;
                for (int d = 0; d < 3; ++d) e.pt[d] = d < dim ? coord(d, *it) : 0.f;
                e.index = *it;
                buckets.push_back(e);
            }
            nodes.push_back(leaf);
            return pos;
        }
        int cutDim = 0;
        for (int d = 1; d < dim; ++d)
            if ((maxV[d] - minV[d]) > (maxV[cutDim] - minV[cutDim])) cutDim = d;
        const int rightCount = count / 2;
        const int leftCount = count - rightCount;
        std::nth_element(first, first + leftCount, last,
                         [&](int a, int b) { return coord(cutDim, a) < coord(cutDim, b); });
        const float cutVal = coord(cutDim, *(first + leftCount));
        float leftMax[3], rightMin[3];
        for (int d = 0; d < 3; ++d) { leftMax[d] = maxV[d]; rightMin[d] = minV[d]; }
        leftMax[cutDim] = cutVal;
        rightMin[cutDim] = cutVal;
        nodes.push_back(Node{cutDim, cutVal, -1, 0, 0});
        build(first, first + leftCount, minV, leftMax);
        const int right = build(first + leftCount, last, rightMin, maxV);
        nodes[pos].rightChild = right;
        return pos;
    }

    KdTree(const float* feat, int rows_, int n_) : rows(rows_), dim(rows_ - 1), n(n_), cloud(feat) {
        std::vector<int> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        float minV[3] = {kInf, kInf, kInf}, maxV[3] = {-kInf, -kInf, -kInf};
        for (int i = 0; i < n; ++i)
            for (int d = 0; d < dim; ++d) {
                minV[d] = std::min(minV[d], coord(d, i));
                maxV[d] = std::max(maxV[d], coord(d, i));
            }
        nodes.reserve(size_t(n) / 2 + 16);
        buckets.reserve(n);
        if (n > 0) build(idx.begin(), idx.end(), minV, maxV);
    }

    // libnabo recurseKnn (allowSelfMatch = true, collectStatistics = true)
    long recurse(const float* q, int ni, float rd, LinearHeap& heap, float* off, float maxError2, float maxRadius2) const {
        const Node& nd = nodes[ni];
        if (nd.cutDim < 0) {
            const BucketEntry* b = &buckets[nd.bucketStart];
            for (int i = 0; i < nd.bucketSize; ++i, ++b) {
                float dist = 0.f;
                for (int d = 0; d < dim; ++d) {
                    const float diff = q[d] - b->pt[d];
                    dist += diff * diff;
                }
                if (dist <= maxRadius2 && dist < heap.headValue()) heap.replaceHead(b->index, dist);
            }
            return nd.bucketSize;
        }
        const int cd = nd.cutDim;
        const float old_off = off[cd];
        const float new_off = q[cd] - nd.cutVal;
        long touched = 0;
        if (new_off > 0) {
            touched += recurse(q, nd.rightChild, rd, heap, off, maxError2, maxRadius2);
            rd += -old_off * old_off + new_off * new_off;
            if (rd <= maxRadius2 && rd * maxError2 < heap.headValue()) {
                off[cd] = new_off;
                touched += recurse(q, ni + 1, rd, heap, off, maxError2, maxRadius2);
                off[cd] = old_off;
            }
        } else {
            touched += recurse(q, ni + 1, rd, heap, off, maxError2, maxRadius2);
            rd += -old_off * old_off + new_off * new_off;
            if (rd <= maxRadius2 && rd * maxError2 < heap.headValue()) {
                off[cd] = new_off;
                touched += recurse(q, nd.rightChild, rd, heap, off, maxError2, maxRadius2);
                off[cd] = old_off;
            }
        }
        return touched;
    }

    long knn(const float* query, int qrows, int nq, int k, float eps, float maxRadius, int32_t* ids, float* dists, int nthreads) const {
        const float maxError2 = (1.f + eps) * (1.f + eps);
        const float maxRadius2 = maxRadius * maxRadius;
        long total = 0;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1) reduction(+ : total)
        {
            LinearHeap heap(k);
#pragma omp for schedule(dynamic, 1024)
            for (int i = 0; i < nq; ++i) {
                heap.reset();
                float off[3] = {0.f, 0.f, 0.f};
                if (n > 0) total += recurse(query + size_t(i) * qrows, 0, 0.f, heap, off, maxError2, maxRadius2);
                heap.get(ids + size_t(i) * k, dists + size_t(i) * k);
            }
        }
        return total;
    }
};

// ------------------------------------------------------------------------------------------
// Matches::getDistsQuantile (Matches.cpp:60-87)
// ------------------------------------------------------------------------------------------
int dists_quantile(const float* dists, long n, float quantile, float* out) {
    std::vector<float> values;
    values.reserve(n);
    for (long i = 0; i < n; ++i)
        if (dists[i] != kInf) values.push_back(dists[i]);
    if (values.empty()) return ORC_ERR_NO_OUTLIER_TO_FILTER;
    if (quantile < 0.0 || quantile > 1.0) return ORC_ERR_BAD_QUANTILE;
    if (quantile == 1.0) {
        *out = *std::max_element(values.begin(), values.end());
        return ORC_OK;
    }
    // `values.size() * quantile` is evaluated in float (size_t -> float), then truncated
    const float pos = float(values.size()) * quantile;
    size_t idx = size_t(pos);
    if (idx >= values.size()) idx = values.size() - 1;  // reference would read out of bounds
    std::nth_element(values.begin(), values.begin() + idx, values.end());
    *out = values[idx];
    return ORC_OK;
}

// VarTrimmedDistOutlierFilter::optimizeInlierRatio (OutlierFiltersImpl.cpp:177-218).  The reference sorts the finite,
// positive distances, takes their running sum in float (std::partial_sum), and minimises
// FRMS(i) = cumsum[i] / i / (i / N)^(2 lambda) over minEl <= i < maxEl, N = rows * cols of the distance matrix, every
// array expression evaluated per coefficient in float.  Where fewer than maxEl distances qualify the reference reads
// past the end of its vector (undefined); the candidates are cut at the number of qualifying distances here.
float g_var_min_ratio = 0.05f, g_var_max_ratio = 0.99f;
int var_trimmed_ratio(const float* dists, long n, float min_ratio, float max_ratio, float lambda, float* ratio_out) {
    std::vector<float> sorted;
    sorted.reserve(n);
    for (long i = 0; i < n; ++i)
        if (dists[i] != kInf && dists[i] > 0) sorted.push_back(dists[i]);
    if (sorted.empty()) return ORC_ERR_NO_OUTLIER_TO_FILTER;
    std::sort(sorted.begin(), sorted.end());
    std::vector<float> cum(sorted.size());
    std::partial_sum(sorted.begin(), sorted.end(), cum.begin());  // running sum in float, one element after the other
    const int points_nbr = (int)n;
    const int minEl = (int)std::floor(min_ratio * points_nbr);
    const int maxEl = (int)std::floor(max_ratio * points_nbr);
    const long end = std::min<long>(maxEl, (long)cum.size());
    int minIndex = 0;
    float best = 0.f;
    bool have = false;
    for (long e = minEl; e < end; ++e) {
        const float id = float(minEl + 1) + float(e - minEl) * 1.f;  // LinSpaced(maxEl - minEl, minEl + 1, maxEl): step 1
        const float ratio = id / float(points_nbr);
        const float deno = std::pow(ratio, lambda);
        const float inv = 1.f / deno;
        const float frms = (cum[e] * (1.f / id)) * (inv * inv);
        if (!have || frms < best) { best = frms; minIndex = int(e - minEl); have = true; }  // minCoeff: the first minimum
    }
    *ratio_out = (float)(minIndex + minEl) / (float)points_nbr;
    return ORC_OK;
}

// Matches::getMedianAbsDeviation (Matches.cpp:88-122): median(|x - median(x)|) over the finite
// squared distances, both medians at position size/2.
int median_abs_deviation(const float* dists, long n, float* out) {
    std::vector<float> values;
    values.reserve(n);
    for (long i = 0; i < n; ++i)
        if (dists[i] != kInf) values.push_back(dists[i]);
    if (values.empty()) return ORC_ERR_NO_OUTLIER_TO_FILTER;
    std::nth_element(values.begin(), values.begin() + values.size() / 2, values.end());
    const float median = values[values.size() / 2];
    for (float& v : values) v = std::fabs(v - median);
    std::nth_element(values.begin(), values.begin() + values.size() / 2, values.end());
    *out = values[values.size() / 2];
    return ORC_OK;
}

// state a RobustOutlierFilter object carries from one compute() to the next (OutlierFiltersImpl.cpp:420-438)
struct RobustState {
    int iteration = 1;
    float scale = 0.f;
};

// `approximation` of the RobustOutlierFilter under test, squared like the constructor squares it (OutlierFiltersImpl.cpp:400);
// set through orc_set_robust_approximation (the filter-chain arrays carry one float per filter)
static float g_robust_approx2 = std::numeric_limits<float>::infinity();

// Matches::getStandardDeviation (Matches.cpp:124-129): sqrt(sum (d - mean)^2 / (size - 1)) over all entries.  Eigen sums a float
// array packet-wise; the restatement sums in double and rounds once (the difference is the reference's own summation noise).
static float standard_deviation(const float* dists, long total) {
    double s = 0.0;
    for (long i = 0; i < total; ++i) s += dists[i];
    const float mean = (float)(s / (double)total);
    double v = 0.0;
    for (long i = 0; i < total; ++i) { const float t = dists[i] - mean; v += (double)(t * t); }
    return std::sqrt((float)(v / (double)(total - 1)));
}

// RobustOutlierFilter::robustFiltering (OutlierFiltersImpl.cpp:503-598), distanceType point2point.
// All array arithmetic in float, as Eigen's Array<T> evaluates it.
int robust_weights(const float* dists, long total, int word, float tuning, RobustState& st, float* w_out, float* scale_out,
                   const float* weight_dists = nullptr) {
    const int fct = (word >> 8) & 0xff, scale_est = (word >> 16) & 0xf, nb_iter = (word >> 20) & 0xff;
    const bool re = st.iteration <= nb_iter || nb_iter == 0;
    if (scale_est == ORC_SCALE_MAD) {
        if (re) {
            float mad;
            const int rc = median_abs_deviation(dists, total, &mad);
            if (rc) return rc;
            st.scale = std::sqrt(mad);
        }
    } else if (scale_est == ORC_SCALE_STD) {
        if (re) st.scale = std::sqrt(standard_deviation(dists, total));
    } else if (scale_est == ORC_SCALE_BERG) {
        // constructor (OutlierFiltersImpl.cpp:420-432): the tuning given is the target scale, the tuning constant Bergstrom's
        const float target = tuning;
        if (fct == ORC_ROBUST_CAUCHY) tuning = 4.3040f;
        else if (fct == ORC_ROBUST_TUKEY) tuning = 7.0589f;
        else if (fct == ORC_ROBUST_HUBER) tuning = 2.0138f;
        if (re) {
            if (st.iteration == 1) {
                float median;
                const int rc = dists_quantile(dists, total, 0.5f, &median);
                if (rc) return rc;
                st.scale = (float)(1.9 * (double)std::sqrt(median));
            } else {
                const float rate = 0.85f;
                st.scale = rate * (st.scale - target) + target;
            }
        }
    } else if (scale_est == ORC_SCALE_NONE) {
        st.scale = 1.f;
    } else {
        return ORC_ERR_BAD_ARG;
    }
    st.iteration++;
    if (scale_out) *scale_out = st.scale;
    const float k = tuning, k2 = k * k, s2 = st.scale * st.scale;
    const float* wd = weight_dists ? weight_dists : dists;
    for (long i = 0; i < total; ++i) {
        const float e2 = wd[i] / s2;
        float w;
        switch (fct) {
            case ORC_ROBUST_CAUCHY: w = 1.f / (1.f + e2 / k2); break;
            case ORC_ROBUST_WELSCH: w = std::exp(-e2 / k2); break;
            case ORC_ROBUST_SC: { const float a = k + e2; w = (e2 >= k) ? 4.0f * k2 * (1.f / (a * a)) : 1.f; break; }
            case ORC_ROBUST_GM: { const float a = k + e2; w = k2 * (1.f / (a * a)); break; }
            case ORC_ROBUST_TUKEY: { const float a = 1.f - e2 / k2; w = (e2 >= k2) ? 0.f : a * a; break; }
            case ORC_ROBUST_HUBER: w = (e2 >= k2) ? k * (1.f / std::sqrt(e2)) : 1.f; break;
            case ORC_ROBUST_L1: w = 1.f / std::sqrt(e2); break;
            case ORC_ROBUST_STUDENT: { const float d = 3.f; w = std::pow(1.f + e2 / k, -(k + d) / 2.f) * (k + d) * (1.f / (k + e2)); break; }
            default: return ORC_ERR_BAD_ARG;
        }
        // `w <= 1e-50` -> 1e-50, which is 0 once stored in a float array; an infinite distance gives weight 0
        // for every function but is dropped by ErrorElements anyway
        if (w <= 1e-50f) w = (float)1e-50;
        if (g_robust_approx2 != std::numeric_limits<float>::infinity() && e2 >= g_robust_approx2) w = 0.f;  // OutlierFiltersImpl.cpp:591-595
        w_out[i] = w;
    }
    return ORC_OK;
}

// OutlierFilters::compute (OutlierFilter.cpp:63-103) over the in-scope filters
// (OutlierFiltersImpl.cpp:66-81, 109-147, 420-598).
// what SurfaceNormalOutlierFilter::compute reads besides the distances (OutlierFiltersImpl.cpp:236-285)
struct SnContext {
    const int32_t* ids = nullptr;
    const float* reading_normals = nullptr;  // 3 x n, rotated like the reading
    const float* ref_normals = nullptr;      // 3 x nr
    // RobustOutlierFilter distanceType point2plane (OutlierFiltersImpl.cpp:468-500) also reads the clouds themselves
    const float* reading = nullptr;          // 4 x n, the reading as this iteration sees it
    const float* reference = nullptr;        // 4 x nr
};

// `.normalized()`: v / |v| when |v|^2 > 0 (Eigen >= 3.3), in float
inline void normalized3(const float* v, float* out) {
    const float n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    if (n2 > 0.f) {
        const float n = std::sqrt(n2);
        out[0] = v[0] / n; out[1] = v[1] / n; out[2] = v[2] / n;
    } else {
        out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
    }
}

int outlier_weights(const float* dists, int knn, int n, int nfilters, const int* types, const float* params, float* w, float* limits_out,
                    RobustState* robust = nullptr, const SnContext* sn = nullptr) {
    const long total = long(knn) * n;
    if (nfilters == 0) {
        for (long i = 0; i < total; ++i) w[i] = (dists[i] == kInf) ? 0.f : 1.f;
        return ORC_OK;
    }
    for (int f = 0; f < nfilters; ++f) {
        float limit = 0.f;
        if (types[f] == ORC_FILTER_MAXDIST) {
            // maxDist(pow(get<T>("maxDist"), 2)): pow(float, int) promotes to double
            limit = float(std::pow(double(params[f]), 2));
        } else if (types[f] == ORC_FILTER_MINDIST) {
            // minDist(pow(get<T>("minDist"), 2)); weights = dists >= minDist (OutlierFiltersImpl.cpp:87-101): an infinite
            // distance passes, and is dropped later by ErrorElements (ErrorMinimizer.cpp:103-106)
            limit = float(std::pow(double(params[f]), 2));
            if (limits_out) limits_out[f] = limit;
            for (long i = 0; i < total; ++i) {
                const float wf = (dists[i] >= limit) ? 1.f : 0.f;
                w[i] = (f == 0) ? wf : w[i] * wf;
            }
            continue;
        } else if (types[f] == ORC_FILTER_MEDIANDIST) {
            float median;
            const int rc = dists_quantile(dists, total, 0.5f, &median);
            if (rc) return rc;
            limit = params[f] * median;
        } else if (types[f] == ORC_FILTER_TRIMMEDDIST) {
            const int rc = dists_quantile(dists, total, params[f], &limit);
            if (rc) return rc;
        } else if (types[f] == ORC_FILTER_VARTRIMMEDDIST) {
            float tuned;
            int rc = var_trimmed_ratio(dists, total, g_var_min_ratio, g_var_max_ratio, params[f], &tuned);
            if (rc) return rc;
            rc = dists_quantile(dists, total, tuned, &limit);
            if (rc) return rc;
        } else if (types[f] == ORC_FILTER_SURFACENORMAL) {
            const float eps = std::cos(params[f]);  // eps(cos(maxAngle)), OutlierFiltersImpl.cpp:227
            if (limits_out) limits_out[f] = eps;
            const bool have = sn && sn->ids && sn->reading_normals && sn->ref_normals;
            for (int x = 0; x < n; ++x) {
                float nr[3] = {0.f, 0.f, 0.f};
                if (have) normalized3(sn->reading_normals + 3 * size_t(x), nr);
                for (int y = 0; y < knn; ++y) {
                    const long i = long(x) * knn + y;
                    float wf = 1.f;  // no normals: "Skipping filtering", all ones
                    if (have) {
                        const int id = sn->ids[i];
                        if (id < 0) wf = 0.f;
                        else {
                            float nq[3];
                            normalized3(sn->ref_normals + 3 * size_t(id), nq);
                            const float value = std::fabs(nr[0] * nq[0] + nr[1] * nq[1] + nr[2] * nq[2]);
                            wf = (value < eps) ? 0.f : 1.f;
                        }
                    }
                    w[i] = (f == 0) ? wf : w[i] * wf;
                }
            }
            continue;
        } else if ((types[f] & 0xff) == ORC_FILTER_ROBUST) {
            RobustState fresh;
            std::vector<float> wr(total);
            float scale = 0.f;
            // the scale estimators always read the match distances; distanceType point2plane (bit 28) replaces the distances the
            // weight function sees by dot(n / |n|, p - q)^2 (computePointToPlaneDistance, OutlierFiltersImpl.cpp:468-500)
            std::vector<float> pp;
            if (types[f] & ORC_ROBUST_P2PLANE) {
                if (!sn || !sn->ids || !sn->ref_normals || !sn->reading || !sn->reference) return ORC_ERR_BAD_ARG;  // the reference throws InvalidField("Field normals not found")
                pp.assign(total, 0.f);
                for (int x = 0; x < n; ++x)
                    for (int y = 0; y < knn; ++y) {
                        const long i = long(x) * knn + y;
                        const int id = sn->ids[i];
                        if (id < 0) continue;
                        float nq[3];
                        normalized3(sn->ref_normals + 3 * size_t(id), nq);
                        const float* pr = sn->reading + 4 * size_t(x);
                        const float* qr = sn->reference + 4 * size_t(id);
                        const float dot = (nq[0] * (pr[0] - qr[0]) + nq[1] * (pr[1] - qr[1])) + nq[2] * (pr[2] - qr[2]);
                        pp[i] = dot * dot;
                    }
            }
            const int rc = robust_weights(dists, total, types[f], params[f], robust ? robust[f] : fresh, wr.data(), &scale, pp.empty() ? nullptr : pp.data());
            if (rc) return rc;
            if (limits_out) limits_out[f] = scale;  // the scale, for the tests
            for (long i = 0; i < total; ++i) w[i] = (f == 0) ? wr[i] : w[i] * wr[i];
            continue;
        } else {
            return ORC_ERR_BAD_ARG;
        }
        if (limits_out) limits_out[f] = limit;
        for (long i = 0; i < total; ++i) {
            const float wf = (dists[i] <= limit) ? 1.f : 0.f;
            w[i] = (f == 0) ? wf : w[i] * wf;
        }
    }
    return ORC_OK;
}

// ------------------------------------------------------------------------------------------
// RigidTransformation (TransformationsImpl.cpp:49-105)
// ------------------------------------------------------------------------------------------
bool check_rigid(const float* T) {
    Mat<float> R(3, 3);
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) R(i, j) = T[i + 4 * j];
    return !(std::fabs(1.f - det3(R)) > 0.001f);
}

void rigid_apply(const float* T, const float* in, int n, float* out) {
    for (int p = 0; p < n; ++p) {
        const float x = in[4 * p], y = in[4 * p + 1], z = in[4 * p + 2], w = in[4 * p + 3];
        for (int r = 0; r < 4; ++r) {
            float acc = T[r] * x;
            acc = acc + T[r + 4] * y;
            acc = acc + T[r + 8] * z;
            acc = acc + T[r + 12] * w;
            out[4 * p + r] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------
// ErrorElements (ErrorMinimizer.cpp:58-193)
// ------------------------------------------------------------------------------------------
struct ErrorElements {
    std::vector<float> reading;    // 4 x M
    std::vector<float> reference;  // 4 x M
    std::vector<float> normals;    // 3 x M (if available)
    std::vector<float> weights;    // M
    std::vector<int> ids;
    std::vector<float> dists;
    int M = 0;
    int nbRejectedMatches = 0, nbRejectedPoints = 0;
    float pointUsedRatio = 0.f, weightedPointUsedRatio = 0.f;
};

int build_error_elements(const float* reading, int nq, const float* reference, const float* ref_normals, const int32_t* ids, const float* dists, const float* w, int knn, ErrorElements& e) {
    long pointsCount = 0;
    for (long i = 0; i < long(knn) * nq; ++i) pointsCount += (w[i] != 0.0f) ? 1 : 0;
    if (pointsCount == 0) return ORC_ERR_NO_POINT_TO_MINIMIZE;
    e.reading.reserve(4 * pointsCount);
    e.weightedPointUsedRatio = 0.f;
    int j = 0;
    for (int i = 0; i < nq; ++i) {
        bool matchExist = false;
        for (int k = 0; k < knn; ++k) {
            const float matchDist = dists[size_t(i) * knn + k];
            if (matchDist == kInf) continue;
            const float wk = w[size_t(i) * knn + k];
            if (wk != 0.0f) {
                for (int r = 0; r < 4; ++r) e.reading.push_back(reading[4 * size_t(i) + r]);
                e.ids.push_back(ids[size_t(i) * knn + k]);
                e.dists.push_back(matchDist);
                e.weights.push_back(wk);
                ++j;
                e.weightedPointUsedRatio += wk;
                matchExist = true;
            } else {
                e.nbRejectedMatches++;
            }
        }
        if (!matchExist) e.nbRejectedPoints++;
    }
    e.M = j;
    if (j == 0) return ORC_ERR_NO_POINT_TO_MINIMIZE;
    e.pointUsedRatio = float(j) / float(knn * nq);
    e.weightedPointUsedRatio /= float(knn * nq);
    e.reference.resize(4 * size_t(j));
    if (ref_normals) e.normals.resize(3 * size_t(j));
    for (int i = 0; i < j; ++i) {
        const int ri = e.ids[i];
        for (int r = 0; r < 4; ++r) e.reference[4 * size_t(i) + r] = reference[4 * size_t(ri) + r];
        if (ref_normals)
            for (int r = 0; r < 3; ++r) e.normals[3 * size_t(i) + r] = ref_normals[3 * size_t(ri) + r];
    }
    return ORC_OK;
}

// ------------------------------------------------------------------------------------------
// PointToPlane (PointToPlane.cpp:108-312)
// ------------------------------------------------------------------------------------------
// solvePossiblyUnderdeterminedLinearSystem (PointToPlane.cpp:108-161).  S = float restates the
// reference; S = double is the "truth" variant.
template <typename S>
void solve_possibly_underdetermined(const Mat<S>& A, const std::vector<S>& b, std::vector<S>& x) {
    const int n = A.r;
    FullPivQR<S> qr(A);
    if (!qr.invertible()) {
        const int rank = qr.rank();
        if (rank == 0) { x.assign(n, S(0)); return; }
        const Mat<S> Qt = transpose(qr.matrixQ());
        Mat<S> Q1t(rank, n);
        for (int i = 0; i < rank; ++i)
            for (int j = 0; j < n; ++j) Q1t(i, j) = Qt(i, j);
        const Mat<S> P = qr.colsPermutation();
        const Mat<S> full = mul(mul(Q1t, A), P);
        Mat<S> R1(rank, n);
        for (int i = 0; i < rank; ++i)
            for (int j = 0; j < n; ++j) R1(i, j) = full(i, j);
        Mat<S> R1R1t = mul(R1, transpose(R1));
        std::vector<S> q1tb(rank, S(0));
        for (int i = 0; i < rank; ++i)
            for (int j = 0; j < n; ++j) q1tb[i] += Q1t(i, j) * b[j];
        std::vector<S> y;
        llt_solve(R1R1t, q1tb, y);
        // x = R1.triangularView<Upper>().transpose() * y
        std::vector<S> xp(n, S(0));
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < rank; ++i)
                if (j >= i) xp[j] += R1(i, j) * y[i];
        x.assign(n, S(0));
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) x[i] += P(i, j) * xp[j];
        // accuracy check b.isApprox(A x, 1e-5)
        std::vector<S> ax(n, S(0));
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) ax[i] += A(i, j) * x[j];
        S diff2 = 0, nb2 = 0, nax2 = 0;
        for (int i = 0; i < n; ++i) { diff2 += (b[i] - ax[i]) * (b[i] - ax[i]); nb2 += b[i] * b[i]; nax2 += ax[i] * ax[i]; }
        if (!(diff2 <= S(1e-5) * S(1e-5) * std::min(nb2, nax2))) {
            // double-precision pseudo-inverse solve (the reference's jacobiSvd fallback)
            Mat<double> Ad(n, n);
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) Ad(i, j) = double(A(i, j));
            std::vector<double> w;
            Mat<double> V;
            jacobi_eig(Ad, w, V);
            double wmax = 0;
            for (double v : w) wmax = std::max(wmax, std::fabs(v));
            std::vector<double> xd(n, 0.0);
            for (int e = 0; e < n; ++e) {
                if (std::fabs(w[e]) <= wmax * n * std::numeric_limits<double>::epsilon()) continue;
                double proj = 0;
                for (int i = 0; i < n; ++i) proj += V(i, e) * double(b[i]);
                for (int i = 0; i < n; ++i) xd[i] += V(i, e) * proj / w[e];
            }
            for (int i = 0; i < n; ++i) x[i] = S(xd[i]);
        }
    } else {
        llt_solve(A, b, x);
    }
}

// Eigen::AngleAxis(angle, axis).toRotationMatrix() then Transform::matrix()
// (PointToPlane.cpp:250-292)
void angle_axis_to_T(const float x[6], float* T) {
    const float n2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
    const float angle = std::sqrt(n2);
    float ax[3] = {x[0], x[1], x[2]};
    if (n2 > 0.f) {  // Eigen >= 3.3 normalized(): divides only when the squared norm is > 0
        const float nrm = std::sqrt(n2);
        ax[0] = x[0] / nrm; ax[1] = x[1] / nrm; ax[2] = x[2] / nrm;
    }
    const float s = std::sin(angle), c = std::cos(angle);
    const float sx = s * ax[0], sy = s * ax[1], sz = s * ax[2];
    const float cx = (1.f - c) * ax[0], cy = (1.f - c) * ax[1], cz = (1.f - c) * ax[2];
    float R[9];
    float tmp;
    tmp = cx * ax[1]; R[0 + 3 * 1] = tmp - sz; R[1 + 3 * 0] = tmp + sz;
    tmp = cx * ax[2]; R[0 + 3 * 2] = tmp + sy; R[2 + 3 * 0] = tmp - sy;
    tmp = cy * ax[2]; R[1 + 3 * 2] = tmp - sx; R[2 + 3 * 1] = tmp + sx;
    R[0] = cx * ax[0] + c; R[4] = cy * ax[1] + c; R[8] = cz * ax[2] + c;
    mat4_identity(T);
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) T[i + 4 * j] = R[i + 3 * j];
    T[12] = x[3]; T[13] = x[4]; T[14] = x[5];
    bool nan = false;
    for (int i = 0; i < 16; ++i) nan |= (T[i] != T[i]);
    if (nan)
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) T[i + 4 * j] = (i == j) ? 1.f : 0.f;
}

// force4DOF (PointToPlane.cpp:203-214, 266-281): the unknowns are the rotation about z and the translation;
// `cross` collapses to its z row, (matrixGamma * p)^T n = -p_y n_x + p_x n_y.
template <typename S>
int minimize_p2plane_4dof(const ErrorElements& e, float* T_out) {
    const int M = e.M;
    Mat<S> A(4, 4);
    std::vector<S> b(4, S(0));
    for (int p = 0; p < M; ++p) {
        const float* r = &e.reading[4 * size_t(p)];
        const float* q = &e.reference[4 * size_t(p)];
        const float* nr = &e.normals[3 * size_t(p)];
        const float w = e.weights[p];
        float F[4], wF[4];
        F[0] = (-r[1]) * nr[0] + r[0] * nr[1];
        F[1] = nr[0]; F[2] = nr[1]; F[3] = nr[2];
        for (int i = 0; i < 4; ++i) wF[i] = w * F[i];
        float dot = 0.f;
        for (int i = 0; i < 3; ++i) dot += (r[i] - q[i]) * nr[i];
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 4; ++i) A(i, j) += S(wF[i]) * S(F[j]);
        for (int i = 0; i < 4; ++i) b[i] += S(wF[i]) * S(dot);
    }
    for (int i = 0; i < 4; ++i) b[i] = -b[i];
    float x4[4];
    if (sizeof(S) == sizeof(float)) {
        Mat<float> Af(4, 4);
        std::vector<float> bf(4), xf;
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < 4; ++i) Af(i, j) = float(A(i, j));
        for (int i = 0; i < 4; ++i) bf[i] = float(b[i]);
        solve_possibly_underdetermined<float>(Af, bf, xf);
        for (int i = 0; i < 4; ++i) x4[i] = xf[i];
    } else {
        std::vector<S> xs;
        solve_possibly_underdetermined<S>(A, b, xs);
        for (int i = 0; i < 4; ++i) x4[i] = float(xs[i]);
    }
    const float x[6] = {0.f, 0.f, x4[0], x4[1], x4[2], x4[3]};  // AngleAxis(x(0), unitZ), translation x(1..3)
    angle_axis_to_T(x, T_out);
    return ORC_OK;
}

// force2D on 3-D clouds (PointToPlane.cpp:177-186, 294-310): the clouds become [x, y, 1], the normals (nx, ny);
// `cross` is the pseudo cross product x*ny - y*nx (ErrorMinimizer.cpp:308-313), the residual has no z term, the
// unknowns are (angle, tx, ty), and the result is the identity with Rotation2D / translation in its xy block.
template <typename S>
int minimize_p2plane_2d(const ErrorElements& e, float* T_out) {
    const int M = e.M;
    Mat<S> A(3, 3);
    std::vector<S> b(3, S(0));
    for (int p = 0; p < M; ++p) {
        const float* r = &e.reading[4 * size_t(p)];
        const float* q = &e.reference[4 * size_t(p)];
        const float* nr = &e.normals[3 * size_t(p)];
        const float w = e.weights[p];
        float F[3], wF[3];
        F[0] = r[0] * nr[1] - r[1] * nr[0];
        F[1] = nr[0]; F[2] = nr[1];
        for (int i = 0; i < 3; ++i) wF[i] = w * F[i];
        float dot = 0.f;
        for (int i = 0; i < 2; ++i) dot += (r[i] - q[i]) * nr[i];
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) A(i, j) += S(wF[i]) * S(F[j]);
        for (int i = 0; i < 3; ++i) b[i] += S(wF[i]) * S(dot);
    }
    for (int i = 0; i < 3; ++i) b[i] = -b[i];
    float x[3];
    if (sizeof(S) == sizeof(float)) {
        Mat<float> Af(3, 3);
        std::vector<float> bf(3), xf;
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) Af(i, j) = float(A(i, j));
        for (int i = 0; i < 3; ++i) bf[i] = float(b[i]);
        solve_possibly_underdetermined<float>(Af, bf, xf);
        for (int i = 0; i < 3; ++i) x[i] = xf[i];
    } else {
        std::vector<S> xs;
        solve_possibly_underdetermined<S>(A, b, xs);
        for (int i = 0; i < 3; ++i) x[i] = float(xs[i]);
    }
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i) T_out[i + 4 * j] = (i == j) ? 1.f : 0.f;
    const float s = std::sin(x[0]), c = std::cos(x[0]);  // Eigen::Rotation2D<float>::toRotationMatrix
    T_out[0] = c; T_out[4] = -s;
    T_out[1] = s; T_out[5] = c;
    T_out[12] = x[1]; T_out[13] = x[2];
    return ORC_OK;
}

template <typename S>
int minimize_p2plane(const ErrorElements& e, float* T_out) {
    const int M = e.M;
    Mat<S> A(6, 6);
    std::vector<S> b(6, S(0));
    for (int p = 0; p < M; ++p) {
        const float* r = &e.reading[4 * size_t(p)];
        const float* q = &e.reference[4 * size_t(p)];
        const float* nr = &e.normals[3 * size_t(p)];
        const float w = e.weights[p];
        float F[6], wF[6];
        // crossProduct (ErrorMinimizer.cpp:304-306)
        F[0] = r[1] * nr[2] - r[2] * nr[1];
        F[1] = r[2] * nr[0] - r[0] * nr[2];
        F[2] = r[0] * nr[1] - r[1] * nr[0];
        F[3] = nr[0]; F[4] = nr[1]; F[5] = nr[2];
        for (int i = 0; i < 6; ++i) wF[i] = w * F[i];
        // dotProd = sum_i deltas.row(i) * normalRef.row(i), accumulated from zero
        float dot = 0.f;
        for (int i = 0; i < 3; ++i) dot += (r[i] - q[i]) * nr[i];
        for (int j = 0; j < 6; ++j)
            for (int i = 0; i < 6; ++i) A(i, j) += S(wF[i]) * S(F[j]);
        for (int i = 0; i < 6; ++i) b[i] += S(wF[i]) * S(dot);
    }
    for (int i = 0; i < 6; ++i) b[i] = -b[i];
    // the reference solves in T = float on the float A, b
    Mat<float> Af(6, 6);
    std::vector<float> bf(6), xf;
    for (int j = 0; j < 6; ++j)
        for (int i = 0; i < 6; ++i) Af(i, j) = float(A(i, j));
    for (int i = 0; i < 6; ++i) bf[i] = float(b[i]);
    float x[6];
    if (sizeof(S) == sizeof(float)) {
        solve_possibly_underdetermined<float>(Af, bf, xf);
        for (int i = 0; i < 6; ++i) x[i] = xf[i];
    } else {
        std::vector<S> xs;
        solve_possibly_underdetermined<S>(A, b, xs);
        for (int i = 0; i < 6; ++i) x[i] = float(xs[i]);
    }
    angle_axis_to_T(x, T_out);
    return ORC_OK;
}

// Censi covariance (PointToPlaneWithCov.cpp:71-162, PointToPointWithCov.cpp:61-145).
// `pts_reading` / `pts_reference` are 4 x M (for point-to-point: the de-meaned clouds),
// `normals` 3 x M or nullptr for the (1,1,1) pseudo-normal of the point-to-point variant.
template <typename S>
void estimate_covariance(const float* pts_reading, const float* pts_reference, const float* normals, int M, const float* T, float sensorStdDev, float* cov_out) {
    Mat<S> J(6, 6), DDt(6, 6);
    const float beta = -std::asin(T[2 + 4 * 0]);
    const float alpha = std::atan2(T[2 + 4 * 1], T[2 + 4 * 2]);
    const float gamma = std::atan2(T[1 + 4 * 0] / std::cos(beta), T[0 + 4 * 0] / std::cos(beta));
    const float t_x = T[12], t_y = T[13], t_z = T[14];
    for (int i = 0; i < M; ++i) {
        const float* rp = pts_reading + 4 * size_t(i);
        const float* fp = pts_reference + 4 * size_t(i);
        float nrm[3] = {1.f, 1.f, 1.f};
        if (normals) { nrm[0] = normals[3 * size_t(i)]; nrm[1] = normals[3 * size_t(i) + 1]; nrm[2] = normals[3 * size_t(i) + 2]; }
        const float reading_range = std::sqrt(rp[0] * rp[0] + rp[1] * rp[1] + rp[2] * rp[2]);
        const float rd[3] = {rp[0] / reading_range, rp[1] / reading_range, rp[2] / reading_range};
        const float reference_range = std::sqrt(fp[0] * fp[0] + fp[1] * fp[1] + fp[2] * fp[2]);
        const float fd[3] = {fp[0] / reference_range, fp[1] / reference_range, fp[2] / reference_range};
        const float n_alpha = nrm[2] * rd[1] - nrm[1] * rd[2];
        const float n_beta = nrm[0] * rd[2] - nrm[2] * rd[0];
        const float n_gamma = nrm[1] * rd[0] - nrm[0] * rd[1];
        float E = nrm[0] * (rp[0] - gamma * rp[1] + beta * rp[2] + t_x - fp[0]);
        E += nrm[1] * (gamma * rp[0] + rp[1] - alpha * rp[2] + t_y - fp[1]);
        E += nrm[2] * (-beta * rp[0] + alpha * rp[1] + rp[2] + t_z - fp[2]);
        float N_reading = nrm[0] * (rd[0] - gamma * rd[1] + beta * rd[2]);
        N_reading += nrm[1] * (gamma * rd[0] + rd[1] - alpha * rd[2]);
        N_reading += nrm[2] * (-beta * rd[0] + alpha * rd[1] + rd[2]);
        const float N_reference = -(nrm[0] * fd[0] + nrm[1] * fd[1] + nrm[2] * fd[2]);
        const float v[6] = {nrm[0], nrm[1], nrm[2], reading_range * n_alpha, reading_range * n_beta, reading_range * n_gamma};
        const float er = E + reading_range * N_reading;
        const float d1[6] = {nrm[0] * N_reading, nrm[1] * N_reading, nrm[2] * N_reading, n_alpha * er, n_beta * er, n_gamma * er};
        const float d2[6] = {nrm[0] * N_reference, nrm[1] * N_reference, nrm[2] * N_reference,
                             reference_range * n_alpha * N_reference, reference_range * n_beta * N_reference, reference_range * n_gamma * N_reference};
        for (int c = 0; c < 6; ++c)
            for (int r = 0; r < 6; ++r) {
                J(r, c) += S(v[r]) * S(v[c]);
                DDt(r, c) += S(d1[r]) * S(d1[c]) + S(d2[r]) * S(d2[c]);
            }
    }
    const Mat<S> Jinv = inverse(J);
    const Mat<S> cov = mul(mul(Jinv, DDt), Jinv);
    const S s2 = S(sensorStdDev * sensorStdDev);
    for (int c = 0; c < 6; ++c)
        for (int r = 0; r < 6; ++r) cov_out[r + 6 * c] = float(s2 * cov(r, c));
}

// PointToPoint (PointToPoint.cpp:61-101).  De-means e.reading / e.reference in place, as the
// reference does (the WithCov variant depends on it).
// `similarity`: PointToPointSimilarityErrorMinimizer (PointToPointSimilarity.cpp:55-101) — the same solve plus
// the scale  sum(singular values, sign-fixed) / sum_p w |reading_p - mean|^2  (1 when that spread is < 1e-4).
template <typename S>
int minimize_p2point(ErrorElements& e, float* T_out, bool similarity = false) {
    const int M = e.M;
    S wsum = 0;
    for (int p = 0; p < M; ++p) wsum += S(e.weights[p]);
    const S w_sum_inv = S(1) / wsum;
    S mr[3] = {0, 0, 0}, mf[3] = {0, 0, 0};
    for (int p = 0; p < M; ++p)
        for (int d = 0; d < 3; ++d) {
            mr[d] += S(e.reading[4 * size_t(p) + d] * e.weights[p]);
            mf[d] += S(e.reference[4 * size_t(p) + d] * e.weights[p]);
        }
    float meanReading[3], meanReference[3];
    for (int d = 0; d < 3; ++d) { meanReading[d] = float(mr[d] * w_sum_inv); meanReference[d] = float(mf[d] * w_sum_inv); }
    for (int p = 0; p < M; ++p)
        for (int d = 0; d < 3; ++d) {
            e.reading[4 * size_t(p) + d] -= meanReading[d];
            e.reference[4 * size_t(p) + d] -= meanReference[d];
        }
    Mat<S> m(3, 3);
    for (int p = 0; p < M; ++p)
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r)
                m(r, c) += S(e.reference[4 * size_t(p) + r] * e.weights[p]) * S(e.reading[4 * size_t(p) + c]);
    Mat<S> U, V;
    std::vector<S> sv;
    svd3(m, U, sv, V);
    Mat<S> Vt = transpose(V);
    Mat<S> R = mul(U, Vt);
    if (det3(R) < S(0)) {
        for (int j = 0; j < 3; ++j) Vt(2, j) = -Vt(2, j);
        R = mul(U, Vt);
        sv[2] = -sv[2];
    }
    S scale = S(1);
    if (similarity) {
        S sigma = 0;
        for (int p = 0; p < M; ++p) {
            const float* r = &e.reading[4 * size_t(p)];
            sigma += S(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]) * S(e.weights[p]);
        }
        scale = (sv[0] + sv[1] + sv[2]) / sigma;
        if (sigma < S(0.0001)) scale = S(1);
    }
    mat4_identity(T_out);
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) T_out[i + 4 * j] = float(scale * R(i, j));
    for (int i = 0; i < 3; ++i) {
        S acc = 0;
        for (int j = 0; j < 3; ++j) acc += S(T_out[i + 4 * j]) * S(meanReading[j]);
        T_out[12 + i] = float(S(meanReference[i]) - acc);
    }
    return ORC_OK;
}

template <typename S>
int minimize_impl(int minimizer_word, ErrorElements& e, float sensorStdDev, float* T_out, float* cov_out) {
    const int minimizer = minimizer_word & 0xff;
    const bool force4dof = (minimizer_word & ORC_MIN_FORCE4DOF) != 0;
    if (minimizer == ORC_MIN_P2PLANE || minimizer == ORC_MIN_P2PLANE_COV) {
        if (e.normals.empty()) return ORC_ERR_BAD_ARG;
        const bool force2d = (minimizer_word & ORC_MIN_FORCE2D) != 0;
        if (force2d && (force4dof || minimizer == ORC_MIN_P2PLANE_COV)) return ORC_ERR_BAD_ARG;
        const int rc = force2d ? minimize_p2plane_2d<S>(e, T_out) : force4dof ? minimize_p2plane_4dof<S>(e, T_out) : minimize_p2plane<S>(e, T_out);
        if (rc) return rc;
        if (minimizer == ORC_MIN_P2PLANE_COV && cov_out)
            estimate_covariance<S>(e.reading.data(), e.reference.data(), e.normals.data(), e.M, T_out, sensorStdDev, cov_out);
        return ORC_OK;
    }
    const int rc = minimize_p2point<S>(e, T_out, minimizer == ORC_MIN_P2POINT_SIM);
    if (rc) return rc;
    if (minimizer == ORC_MIN_P2POINT_COV && cov_out)
        estimate_covariance<S>(e.reading.data(), e.reference.data(), nullptr, e.M, T_out, sensorStdDev, cov_out);
    return ORC_OK;
}

// ------------------------------------------------------------------------------------------
// TransformationCheckers (TransformationCheckersImpl.cpp:45-158)
// ------------------------------------------------------------------------------------------
struct Quat { float w, x, y, z; };
Quat quat_from_T(const float* T) {
    auto m = [&](int i, int j) { return T[i + 4 * j]; };
    Quat q;
    float t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0.f) {
        t = std::sqrt(t + 1.f);
        q.w = 0.5f * t;
        t = 0.5f / t;
        q.x = (m(2, 1) - m(1, 2)) * t;
        q.y = (m(0, 2) - m(2, 0)) * t;
        q.z = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.f);
        float v[3];
        v[i] = 0.5f * t;
        t = 0.5f / t;
        q.w = (m(k, j) - m(j, k)) * t;
        v[j] = (m(j, i) + m(i, j)) * t;
        v[k] = (m(k, i) + m(i, k)) * t;
        q.x = v[0]; q.y = v[1]; q.z = v[2];
    }
    return q;
}
// Eigen 3.3 QuaternionBase::angularDistance: d = a * conj(b); 2 atan2(|d.vec|, |d.w|)
float quat_angular_distance(const Quat& a, const Quat& b) {
    const Quat c{b.w, -b.x, -b.y, -b.z};
    const float dw = a.w * c.w - a.x * c.x - a.y * c.y - a.z * c.z;
    const float dx = a.w * c.x + a.x * c.w + a.y * c.z - a.z * c.y;
    const float dy = a.w * c.y + a.y * c.w + a.z * c.x - a.x * c.z;
    const float dz = a.w * c.z + a.z * c.w + a.x * c.y - a.y * c.x;
    return 2.f * std::atan2(std::sqrt(dx * dx + dy * dy + dz * dz), std::fabs(dw));
}

struct Checkers {
    int maxIter;
    bool useDiff;
    float minRot, minTrans;
    unsigned smooth;
    float counter = 0;
    std::vector<Quat> rotations;
    std::vector<std::vector<float>> translations;
    void init(const float* T) {
        counter = 0;
        rotations.clear();
        translations.clear();
        if (useDiff) {
            rotations.push_back(quat_from_T(T));
            translations.push_back({T[12], T[13], T[14]});
        }
    }
    // returns status; sets iterate / maxReached like ICP.cpp:419-427
    int check(const float* T, bool& iterate) {
        // chain order used by the reference configs: Counter first, then Differential
        counter += 1.f;
        if (counter >= float(maxIter)) { iterate = false; return ORC_OK; }  // MaxNumIterationsReached
        if (useDiff) {
            rotations.push_back(quat_from_T(T));
            translations.push_back({T[12], T[13], T[14]});
            float c0 = 0.f, c1 = 0.f;
            if (rotations.size() > smooth) {
                for (size_t i = rotations.size() - 1; i >= rotations.size() - smooth; --i) {
                    c0 += std::fabs(quat_angular_distance(rotations[i], rotations[i - 1]));
                    const float dx = translations[i][0] - translations[i - 1][0];
                    const float dy = translations[i][1] - translations[i - 1][1];
                    const float dz = translations[i][2] - translations[i - 1][2];
                    c1 += std::fabs(std::sqrt(dx * dx + dy * dy + dz * dz));
                    if (i == 0) break;
                }
                c0 /= float(smooth);
                c1 /= float(smooth);
                if (c0 < minRot && c1 < minTrans) iterate = false;
            }
            if (c0 != c0 || c1 != c1) return ORC_ERR_NAN;
        }
        return ORC_OK;
    }
};

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void* orc_kdtree_create(const float* feat, int rows, int n) { return new KdTree(feat, rows, n); }
void orc_kdtree_destroy(void* tree) { delete static_cast<KdTree*>(tree); }

long orc_kdtree_knn(void* tree, const float* query, int rows, int nq, int k, float eps, float max_radius, int32_t* ids, float* dists, int nthreads) {
    const KdTree* t = static_cast<const KdTree*>(tree);
    if (k > t->n) return -ORC_ERR_KNN_TOO_LARGE;  // libnabo throws when k > number of points
    return t->knn(query, rows, nq, k, eps, max_radius, ids, dists, nthreads);
}

// libnabo BruteForceSearch: all points in index order, strict `<` against the heap head, so
// the result is the top-k in lexicographic (dist, index) order.
long orc_bruteforce_knn(const float* ref, int rows, int nr, const float* query, int nq, int k, float max_radius, int32_t* ids, float* dists, int nthreads) {
    if (k > nr) return -ORC_ERR_KNN_TOO_LARGE;
    const int dim = rows - 1;
    const float maxRadius2 = max_radius * max_radius;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        LinearHeap heap(k);
#pragma omp for schedule(static)
        for (int i = 0; i < nq; ++i) {
            heap.reset();
            const float* q = query + size_t(i) * rows;
            for (int j = 0; j < nr; ++j) {
                const float* p = ref + size_t(j) * rows;
                float dist = 0.f;
                for (int d = 0; d < dim; ++d) {
                    const float diff = q[d] - p[d];
                    dist += diff * diff;
                }
                if (dist <= maxRadius2 && dist < heap.headValue()) heap.replaceHead(j, dist);
            }
            heap.get(ids + size_t(i) * k, dists + size_t(i) * k);
        }
    }
    return long(nq) * nr;
}

// KDTreeVarDistMatcher (MatchersImpl.cpp:132-150): libnabo's knn with one maximum radius per query
long orc_bruteforce_knn_var(const float* ref, int rows, int nr, const float* query, int nq, int k, const float* max_radii, int32_t* ids, float* dists,
                            int nthreads) {
    if (k > nr) return -ORC_ERR_KNN_TOO_LARGE;
    const int dim = rows - 1;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        LinearHeap heap(k);
#pragma omp for schedule(static)
        for (int i = 0; i < nq; ++i) {
            heap.reset();
            const float maxRadius2 = max_radii[i] * max_radii[i];
            const float* q = query + size_t(i) * rows;
            for (int j = 0; j < nr; ++j) {
                const float* p = ref + size_t(j) * rows;
                float dist = 0.f;
                for (int d = 0; d < dim; ++d) {
                    const float diff = q[d] - p[d];
                    dist += diff * diff;
                }
                if (dist <= maxRadius2 && dist < heap.headValue()) heap.replaceHead(j, dist);
            }
            heap.get(ids + size_t(i) * k, dists + size_t(i) * k);
        }
    }
    return long(nq) * nr;
}

int orc_rigid_transform(const float* T16, const float* in, int n, float* out) {
    if (!check_rigid(T16)) return ORC_ERR_NOT_ORTHOGONAL;
    rigid_apply(T16, in, n, out);
    return ORC_OK;
}

int orc_rotate_normals(const float* T, const float* in3, int n, float* out3) {
    for (int p = 0; p < n; ++p) {
        const float x = in3[3 * p], y = in3[3 * p + 1], z = in3[3 * p + 2];
        for (int r = 0; r < 3; ++r) {
            float acc = T[r] * x;
            acc = acc + T[r + 4] * y;
            acc = acc + T[r + 8] * z;
            out3[3 * p + r] = acc;
        }
    }
    return ORC_OK;
}

static const float* g_reading_normals = nullptr;
void orc_set_reading_normals(const float* normals3xn) { g_reading_normals = normals3xn; }

int orc_outlier_weights_sn(const float* dists, const int32_t* ids, int knn, int n, int nfilters, const int* types, const float* params,
                           const float* reading_normals, const float* ref_normals, float* weights, float* limits_out) {
    SnContext sn;
    sn.ids = ids; sn.reading_normals = reading_normals; sn.ref_normals = ref_normals;
    return outlier_weights(dists, knn, n, nfilters, types, params, weights, limits_out, nullptr, &sn);
}

int orc_outlier_weights_geom(const float* dists, const int32_t* ids, int knn, int n, int nfilters, const int* types, const float* params,
                             const float* reading4xn, const float* reference4xnr, const float* ref_normals, float* weights, float* limits_out) {
    SnContext sn;
    sn.ids = ids; sn.ref_normals = ref_normals; sn.reading = reading4xn; sn.reference = reference4xnr;
    return outlier_weights(dists, knn, n, nfilters, types, params, weights, limits_out, nullptr, &sn);
}

int orc_dists_quantile(const float* dists, long n, float quantile, float* out) { return dists_quantile(dists, n, quantile, out); }
void orc_set_robust_approximation(float approximation) { g_robust_approx2 = (float)((double)approximation * (double)approximation); }
void orc_set_var_trimmed_ratios(float min_ratio, float max_ratio) { g_var_min_ratio = min_ratio; g_var_max_ratio = max_ratio; }
int orc_var_trimmed_ratio(const float* dists, long n, float min_ratio, float max_ratio, float lambda, float* ratio_out) {
    return var_trimmed_ratio(dists, n, min_ratio, max_ratio, lambda, ratio_out);
}

int orc_outlier_weights(const float* dists, int knn, int n, int nfilters, const int* types, const float* params, float* weights, float* limits_out) {
    return outlier_weights(dists, knn, n, nfilters, types, params, weights, limits_out);
}

int orc_minimize(int minimizer, const float* reading, int nq, const float* reference, int nr, const float* ref_normals, const int32_t* ids, const float* dists, const float* weights, int knn, float sensor_std_dev, int acc_double, float* T_out, float* cov_out, float* stats_out) {
    (void)nr;
    ErrorElements e;
    const int rc = build_error_elements(reading, nq, reference, ref_normals, ids, dists, weights, knn, e);
    if (rc) return rc;
    if (stats_out) {
        stats_out[0] = e.pointUsedRatio;
        stats_out[1] = e.weightedPointUsedRatio;
        stats_out[2] = float(e.nbRejectedMatches);
        stats_out[3] = float(e.nbRejectedPoints);
        stats_out[4] = float(e.M);
    }
    return acc_double ? minimize_impl<double>(minimizer, e, sensor_std_dev, T_out, cov_out)
                      : minimize_impl<float>(minimizer, e, sensor_std_dev, T_out, cov_out);
}

int orc_surface_normals(const float* feat, int rows, int n, int knn, float eps, float max_dist, int sort_eigen, int smooth_normals, int nthreads, float* normals, float* densities, float* eig_values, float* eig_vectors, float* matched_ids, float* mean_dists, int32_t* ids_out, float* dists_out, float* gap_out, int* degenerate_out) {
    if (rows != 4) return ORC_ERR_BAD_ARG;
    if (knn > n) return ORC_ERR_KNN_TOO_LARGE;
    std::vector<int32_t> ids(size_t(knn) * n);
    std::vector<float> dists(size_t(knn) * n);
    {
        KdTree tree(feat, rows, n);
        tree.knn(feat, rows, n, knn, eps, max_dist, ids.data(), dists.data(), nthreads);
    }
    std::vector<float> nrm_local;
    float* nrm = normals;
    if (!nrm && smooth_normals) return ORC_ERR_BAD_ARG;
    int degenerateCount = 0;
#pragma omp parallel for num_threads(nthreads > 0 ? nthreads : 1) reduction(+ : degenerateCount) schedule(static)
    for (int i = 0; i < n; ++i) {
        bool isDegenerate = false;
        std::vector<float> d;  // 3 x realKnn
        int realKnn = 0;
        for (int j = 0; j < knn; ++j) {
            if (dists[size_t(i) * knn + j] != kInf) {
                const int ri = ids[size_t(i) * knn + j];
                for (int r = 0; r < 3; ++r) d.push_back(feat[4 * size_t(ri) + r]);
                ++realKnn;
            }
        }
        float mean[3] = {0.f, 0.f, 0.f};
        for (int j = 0; j < realKnn; ++j)
            for (int r = 0; r < 3; ++r) mean[r] += d[3 * j + r];
        for (int r = 0; r < 3; ++r) mean[r] /= float(realKnn);
        std::vector<float> NN(d.size());
        for (int j = 0; j < realKnn; ++j)
            for (int r = 0; r < 3; ++r) NN[3 * j + r] = d[3 * j + r] - mean[r];
        Mat<float> C(3, 3);
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) {
                float acc = 0.f;
                for (int j = 0; j < realKnn; ++j) acc += NN[3 * j + r] * NN[3 * j + c];
                C(r, c) = acc;
            }
        float eigenVa[3] = {0.f, 0.f, 0.f};
        float eigenVe[9] = {0.f};  // column-major
        float gap = 0.f;
        FullPivQR<float> qr(C);
        if (realKnn > 0 && qr.rank() + 1 >= 3) {
            // EigenSolver(C): C is symmetric, so its real eigen-pairs are those of a symmetric
            // solver, eigenvectors unit-norm with arbitrary sign.
            Mat<double> Cd(3, 3);
            for (int c = 0; c < 3; ++c)
                for (int r = 0; r < 3; ++r) Cd(r, c) = double(C(r, c));
            std::vector<double> w;
            Mat<double> V;
            jacobi_eig(Cd, w, V);
            int order[3] = {0, 1, 2};
            if (sort_eigen) std::sort(order, order + 3, [&](int a, int b) { return w[a] < w[b]; });
            for (int c = 0; c < 3; ++c) {
                eigenVa[c] = float(w[order[c]]);
                for (int r = 0; r < 3; ++r) eigenVe[r + 3 * c] = float(V(r, order[c]));
            }
            double ws[3] = {w[0], w[1], w[2]};
            std::sort(ws, ws + 3);
            const double tr = ws[0] + ws[1] + ws[2];
            gap = tr > 0 ? float((ws[1] - ws[0]) / tr) : 0.f;
        } else {
            ++degenerateCount;
            isDegenerate = true;
        }
        if (gap_out) gap_out[i] = gap;
        if (nrm) {
            int smallestId = 0;
            if (!sort_eigen) {
                float smallestValue = std::numeric_limits<float>::max();
                for (int j = 0; j < 3; ++j)
                    if (eigenVa[j] < smallestValue) { smallestId = j; smallestValue = eigenVa[j]; }
            }
            for (int r = 0; r < 3; ++r) nrm[3 * size_t(i) + r] = std::min(1.f, std::max(-1.f, eigenVe[r + 3 * smallestId]));
        }
        if (densities) {
            if (isDegenerate) densities[i] = 0.f;
            else {
                float maxn = 0.f;
                for (int j = 0; j < realKnn; ++j) {
                    const float nn = std::sqrt(NN[3 * j] * NN[3 * j] + NN[3 * j + 1] * NN[3 * j + 1] + NN[3 * j + 2] * NN[3 * j + 2]);
                    maxn = std::max(maxn, nn);
                }
                const float volume = float((4. / 3.) * M_PI * std::pow(double(maxn), 3));
                densities[i] = float(realKnn) / volume;
            }
        }
        if (eig_values)
            for (int r = 0; r < 3; ++r) eig_values[3 * size_t(i) + r] = eigenVa[r];
        if (eig_vectors)  // serializeEigVec: row-major
            for (int k = 0; k < 3; ++k)
                for (int c = 0; c < 3; ++c) eig_vectors[9 * size_t(i) + 3 * k + c] = eigenVe[k + 3 * c];
        if (mean_dists) {
            if (isDegenerate) mean_dists[i] = float(std::numeric_limits<std::size_t>::max());
            else {
                const float dx = feat[4 * size_t(i)] - mean[0], dy = feat[4 * size_t(i) + 1] - mean[1], dz = feat[4 * size_t(i) + 2] - mean[2];
                mean_dists[i] = std::sqrt(dx * dx + dy * dy + dz * dz);
            }
        }
    }
    if (matched_ids)
        for (size_t i = 0; i < ids.size(); ++i) matched_ids[i] = float(ids[i]);
    if (smooth_normals) {
        std::vector<float> orig(nrm, nrm + 3 * size_t(n));
        // the reference smooths in place, point after point: later points see already
        // smoothed neighbours (SurfaceNormal.cpp:259-283)
        for (int i = 0; i < n; ++i) {
            const float cur[3] = {nrm[3 * size_t(i)], nrm[3 * size_t(i) + 1], nrm[3 * size_t(i) + 2]};
            float mean[3] = {0.f, 0.f, 0.f};
            int cnt = 0;
            for (int j = 0; j < knn; ++j) {
                if (dists[size_t(i) * knn + j] != kInf) {
                    const int ri = ids[size_t(i) * knn + j];
                    const float* nn = nrm + 3 * size_t(ri);
                    const float dot = cur[0] * nn[0] + cur[1] * nn[1] + cur[2] * nn[2];
                    if (dot > 0.f) for (int r = 0; r < 3; ++r) mean[r] += nn[r];
                    else for (int r = 0; r < 3; ++r) mean[r] -= nn[r];
                    ++cnt;
                }
            }
            for (int r = 0; r < 3; ++r) nrm[3 * size_t(i) + r] = mean[r] / float(cnt);
        }
    }
    if (ids_out) std::memcpy(ids_out, ids.data(), ids.size() * sizeof(int32_t));
    if (dists_out) std::memcpy(dists_out, dists.data(), dists.size() * sizeof(float));
    if (degenerate_out) *degenerate_out = degenerateCount;
    return ORC_OK;
}

int orc_icp(const float* readingIn, int nq, const float* referenceIn, int nr, const float* ref_normals, const float* T_init, const orc_icp_config* cfg, float* T_out, float* T_iters_out, int* iterations_out, float* cov_out, float* stats_out) {
    // ICP::compute (ICP.cpp:264-313): centre the reference on its mean
    std::vector<float> reference(referenceIn, referenceIn + 4 * size_t(nr));
    float sum[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p = 0; p < nr; ++p)
        for (int r = 0; r < 4; ++r) sum[r] += reference[4 * size_t(p) + r];
    float meanRef[4];
    for (int r = 0; r < 4; ++r) meanRef[r] = sum[r] / float(nr);
    float T_refIn_refMean[16];
    mat4_identity(T_refIn_refMean);
    for (int r = 0; r < 3; ++r) T_refIn_refMean[12 + r] = meanRef[r];
    for (int p = 0; p < nr; ++p)
        for (int r = 0; r < 3; ++r) reference[4 * size_t(p) + r] -= meanRef[r];

    const double t_build0 = now_s();
    KdTree* tree = nullptr;
    if (cfg->search_type != 0) tree = new KdTree(reference.data(), 4, nr);
    g_timings[0] = now_s() - t_build0;
    g_timings[2] = 0;
    if (cfg->knn > nr) { delete tree; return ORC_ERR_KNN_TOO_LARGE; }

    // computeWithTransformedReference (ICP.cpp:316-449)
    float T_refMean_refIn[16];
    mat4_identity(T_refMean_refIn);
    for (int r = 0; r < 3; ++r) T_refMean_refIn[12 + r] = -meanRef[r];
    float T_refMean_dataIn[16];
    mat4_mul(T_refMean_refIn, T_init, T_refMean_dataIn);
    std::vector<float> reading(4 * size_t(nq));
    if (!check_rigid(T_refMean_dataIn)) { delete tree; return ORC_ERR_NOT_ORTHOGONAL; }
    rigid_apply(T_refMean_dataIn, readingIn, nq, reading.data());

    float T_iter[16];
    mat4_identity(T_iter);
    Checkers checkers{cfg->max_iterations, cfg->use_differential != 0, cfg->min_diff_rot_err, cfg->min_diff_trans_err, unsigned(cfg->smooth_length)};
    checkers.init(T_iter);
    bool iterate = true;
    int iterationCount = 0;
    std::vector<float> stepReading(4 * size_t(nq));
    const int knn = cfg->knn;
    std::vector<int32_t> ids(size_t(knn) * nq);
    std::vector<float> dists(size_t(knn) * nq), w(size_t(knn) * nq);
    int rc = ORC_OK;
    RobustState robustState[8];  // one per filter slot: the filter objects live as long as the ICP object
    std::vector<float> stepNormals, baseNormals;
    const double t_loop0 = now_s();
    while (iterate) {
        // SimilarityTransformation::checkParameters accepts anything (TransformationsImpl.cpp:199-204)
        if ((cfg->minimizer & 0xff) != ORC_MIN_P2POINT_SIM && !check_rigid(T_iter)) { rc = ORC_ERR_NOT_ORTHOGONAL; break; }
        rigid_apply(T_iter, reading.data(), nq, stepReading.data());
        const double t_m0 = now_s();
        if (tree) tree->knn(stepReading.data(), 4, nq, knn, cfg->epsilon, cfg->max_dist, ids.data(), dists.data(), cfg->nthreads);
        else orc_bruteforce_knn(reference.data(), 4, nr, stepReading.data(), nq, knn, cfg->max_dist, ids.data(), dists.data(), cfg->nthreads);
        g_timings[2] += now_s() - t_m0;
        SnContext sn;
        sn.ids = ids.data(); sn.ref_normals = ref_normals; sn.reading = stepReading.data(); sn.reference = reference.data();
        if (g_reading_normals && ref_normals) {
            // the reading's normals turn with it (RigidTransformation::compute, TransformationsImpl.cpp:71-84):
            // T_iter * (T_refMean_dataIn * n), rotation blocks only
            if (stepNormals.empty()) {
                stepNormals.resize(3 * size_t(nq));
                baseNormals.resize(3 * size_t(nq));
                orc_rotate_normals(T_refMean_dataIn, g_reading_normals, nq, baseNormals.data());
            }
            orc_rotate_normals(T_iter, baseNormals.data(), nq, stepNormals.data());
            sn.ids = ids.data(); sn.reading_normals = stepNormals.data(); sn.ref_normals = ref_normals;
        }
        rc = outlier_weights(dists.data(), knn, nq, cfg->nfilters, cfg->filter_type, cfg->filter_param, w.data(), nullptr, robustState, &sn);
        if (rc) break;
        float dT[16];
        rc = orc_minimize(cfg->minimizer, stepReading.data(), nq, reference.data(), nr, ref_normals, ids.data(), dists.data(), w.data(), knn, cfg->sensor_std_dev, cfg->acc_double, dT, cov_out, stats_out);
        if (rc) break;
        mat4_mul(dT, T_iter, T_iter);
        if (T_iters_out) std::memcpy(T_iters_out + 16 * size_t(iterationCount), T_iter, sizeof(T_iter));
        rc = checkers.check(T_iter, iterate);
        ++iterationCount;
        if (rc) break;
    }
    g_timings[1] = now_s() - t_loop0;
    g_timings[3] = iterationCount;
    delete tree;
    if (iterations_out) *iterations_out = iterationCount;
    if (rc) return rc;
    float tmp[16];
    mat4_mul(T_refIn_refMean, T_iter, tmp);
    mat4_mul(tmp, T_refMean_dataIn, T_out);
    return ORC_OK;
}

void orc_last_timings(double* out4) {
    for (int i = 0; i < 4; ++i) out4[i] = g_timings[i];
}

void orc_quaternion_angular_distance(const float* Ta, const float* Tb, float* out) {
    *out = quat_angular_distance(quat_from_T(Ta), quat_from_T(Tb));
}

}  // extern "C"
