// tests/emu/emu.cpp — host harness for the PM_HD per-thread algorithms of
// libpointmatcher_b200/csrc/core/*.h (tree addressing, traversal, top-k insertion, small solves).
//
// TEST INFRASTRUCTURE ONLY: it lets the CPU test-suite (pytest -m "not gpu") exercise the exact
// device functions without a GPU.  It is compiled by tests/ into tests/emu/_build/libemu.so and is
// never part of the product library; the product has no CPU path.
// Built with -ffp-contract=off so that fmul/fadd/fsub round once, as the device intrinsics do.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

#define PM_EMU_STATS 1
#include "core/common.h"
#include "core/linalg.h"
#include "core/tree.h"

using namespace pm;

namespace {

struct EmuTree {
    std::vector<f2> splits;
    std::vector<f4> boxes, pts;
    const float* orig = nullptr;  // the caller's cloud (4 x n), original order
    TreeView view;
};

// the level-by-level build of tree_build.cu, sequentially
EmuTree* build(const float* feat, int n) {
    EmuTree* t = new EmuTree();
    t->orig = feat;
    const uint32_t N = (uint32_t)n;
    const int D = tree_depth_for(N);
    const uint32_t nnodes = 2u << D;
    std::vector<uint32_t> box(6 * (size_t)nnodes);
    for (uint32_t i = 0; i < nnodes; ++i) {
        for (int a = 0; a < 3; ++a) { box[6 * i + a] = 0xffffffffu; box[6 * i + 3 + a] = 0u; }
    }
    std::vector<uint32_t> perm(N);
    std::iota(perm.begin(), perm.end(), 0u);
    auto coord = [&](uint32_t i, int d) { return feat[4 * (size_t)i + d]; };
    auto widest = [&](uint32_t node) {
        const uint32_t* b = &box[6 * (size_t)node];
        const float ex = fsub(ord_float(b[3]), ord_float(b[0]));
        const float ey = fsub(ord_float(b[4]), ord_float(b[1]));
        const float ez = fsub(ord_float(b[5]), ord_float(b[2]));
        int dim = 0;
        float best = ex;
        if (ey > best) { dim = 1; best = ey; }
        if (ez > best) { dim = 2; }
        return dim;
    };
    t->splits.assign((size_t)1 << D, f2{0.f, 0.f});
    for (int l = 0; l <= D; ++l) {
        for (uint32_t p = 0; p < N; ++p) {
            const uint32_t node = (1u << l) + seg_of(p, l, N);
            for (int a = 0; a < 3; ++a) {
                const uint32_t o = float_ord(coord(perm[p], a));
                box[6 * node + a] = std::min(box[6 * node + a], o);
                box[6 * node + 3 + a] = std::max(box[6 * node + 3 + a], o);
            }
        }
        if (l == D) break;
        std::vector<std::pair<uint64_t, uint32_t>> kv(N);
        for (uint32_t p = 0; p < N; ++p) {
            const uint32_t seg = seg_of(p, l, N);
            const int dim = widest((1u << l) + seg);
            kv[p] = {((uint64_t)seg << 32) | float_ord(coord(perm[p], dim)), perm[p]};
        }
        std::stable_sort(kv.begin(), kv.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
        for (uint32_t p = 0; p < N; ++p) perm[p] = kv[p].second;
        for (uint32_t s = 0; s < (1u << l); ++s) {
            const uint32_t node = (1u << l) + s;
            const int dim = widest(node);
            const uint32_t mid = seg_begin(l + 1, 2 * s + 1, N);
            t->splits[node] = f2{coord(perm[mid], dim), u2f((uint32_t)dim)};
        }
    }
    t->pts.resize(N);
    for (uint32_t p = 0; p < N; ++p) t->pts[p] = make_f4(coord(perm[p], 0), coord(perm[p], 1), coord(perm[p], 2), u2f(perm[p]));
    t->boxes.resize(2 * (size_t)nnodes);
    for (uint32_t i = 0; i < nnodes; ++i) {
        const uint32_t* b = &box[6 * (size_t)i];
        t->boxes[2 * (size_t)i] = make_f4(ord_float(b[0]), ord_float(b[1]), ord_float(b[2]), 0.f);
        t->boxes[2 * (size_t)i + 1] = make_f4(ord_float(b[3]), ord_float(b[4]), ord_float(b[5]), 0.f);
    }
    t->view.splits = t->splits.data();
    t->view.boxes = t->boxes.data();
    t->view.pts = t->pts.data();
    t->view.n = N;
    t->view.depth = D;
    return t;
}

uint32_t* g_visits_out = nullptr;  // optional per-query visit counts (analysis only)

const int32_t* g_seed = nullptr;  // optional per-query seed candidate (reference column or -1), k = 1 only

template <int KMAX>
long run_knn(const EmuTree* t, const float* T16, const float* q, int nq, int k, float max_r2, int32_t* ids, float* dists) {
    long visits = 0;
    Mat4 T;
    if (T16) std::memcpy(T.m, T16, sizeof(T.m));
    for (int i = 0; i < nq; ++i) {
        f4 p = make_f4(q[4 * (size_t)i], q[4 * (size_t)i + 1], q[4 * (size_t)i + 2], q[4 * (size_t)i + 3]);
        if (T16) p = transform_point(T, p);
        TopK<KMAX> best;
        best.init(k, max_r2);
        if (g_seed && KMAX == 1 && g_seed[i] >= 0) {
            // re-measure the previous match (knn.cu does the same with the resident ids)
            const float* r = t->orig + 4 * (size_t)g_seed[i];
            const float dd = dist2(p.x, p.y, p.z, r[0], r[1], r[2]);
            if (best.accepts(dd, g_seed[i])) best.insert(dd, g_seed[i]);
        }
        const uint32_t v = knn_search_single<KMAX>(t->view, p.x, p.y, p.z, best);
        visits += v;
        if (g_visits_out) g_visits_out[i] = v;
        for (int j = 0; j < k; ++j) {
            float bd;
            int bi;
            best.get(j, bd, bi);
            const bool valid = bi != PM_NO_ID && bd != pm_inf();
            ids[(size_t)i * k + j] = valid ? bi : -1;
            dists[(size_t)i * k + j] = valid ? bd : pm_inf();
        }
    }
    return visits;
}

}  // namespace

extern "C" {

void emu_set_visits_out(uint32_t* p) { g_visits_out = p; }
void emu_set_seed(const int32_t* p) { g_seed = p; }
void emu_stats(unsigned long long* out5, int reset) {
    out5[0] = g_emu_stats.descent_steps; out5[1] = g_emu_stats.pops; out5[2] = g_emu_stats.box_tests;
    out5[3] = g_emu_stats.redescents; out5[4] = g_emu_stats.leaves;
    if (reset) g_emu_stats = EmuStats{0, 0, 0, 0, 0};
}
void* emu_tree_build(const float* feat, int n) { return build(feat, n); }
void emu_tree_free(void* t) { delete static_cast<EmuTree*>(t); }
int emu_tree_depth(void* t) { return static_cast<EmuTree*>(t)->view.depth; }

// leaf sizes and box containment invariants; returns 0 when the structure is consistent
int emu_tree_check(void* tp) {
    const EmuTree* t = static_cast<EmuTree*>(tp);
    const TreeView& v = t->view;
    const uint32_t leaves = 1u << v.depth;
    std::vector<char> seen(v.n, 0);
    for (uint32_t leaf = 0; leaf < leaves; ++leaf) {
        const uint32_t b = seg_begin(v.depth, leaf, v.n), e = seg_begin(v.depth, leaf + 1, v.n);
        if (e <= b || e - b > PM_LEAF_MAX) return 1;
        for (uint32_t p = b; p < e; ++p) {
            if (seg_of(p, v.depth, v.n) != leaf) return 2;
            const uint32_t idx = f2u(v.pts[p].w);
            if (idx >= v.n || seen[idx]) return 3;
            seen[idx] = 1;
            // every ancestor box contains the point, and every split separates its halves
            uint32_t node = leaves + leaf;
            const float c[3] = {v.pts[p].x, v.pts[p].y, v.pts[p].z};
            while (node >= 1) {
                const f4 lo = v.boxes[2 * (size_t)node], hi = v.boxes[2 * (size_t)node + 1];
                if (c[0] < lo.x || c[0] > hi.x || c[1] < lo.y || c[1] > hi.y || c[2] < lo.z || c[2] > hi.z) return 4;
                if (node > 1) {
                    const f2 s = v.splits[node >> 1];
                    const float cd = c[f2u(s.y)];
                    if ((node & 1) ? (cd < s.x) : (cd > s.x)) return 6;
                }
                node >>= 1;
            }
        }
    }
    if (seg_begin(v.depth, leaves, v.n) != v.n) return 5;
    return 0;
}

long emu_knn(void* tp, const float* T16, const float* q, int nq, int k, float max_dist, int32_t* ids, float* dists) {
    const EmuTree* t = static_cast<EmuTree*>(tp);
    const float r2 = max_dist * max_dist;
    if (k == 1) return run_knn<1>(t, T16, q, nq, k, r2, ids, dists);
    if (k <= 4) return run_knn<4>(t, T16, q, nq, k, r2, ids, dists);
    if (k <= 8) return run_knn<8>(t, T16, q, nq, k, r2, ids, dists);
    if (k <= 10) return run_knn<10>(t, T16, q, nq, k, r2, ids, dists);
    if (k <= 16) return run_knn<16>(t, T16, q, nq, k, r2, ids, dists);
    if (k <= 20) return run_knn<20>(t, T16, q, nq, k, r2, ids, dists);
    if (k <= 32) return run_knn<32>(t, T16, q, nq, k, r2, ids, dists);
    if (k <= 64) return run_knn<64>(t, T16, q, nq, k, r2, ids, dists);
    return -1;
}

int emu_solve_psd6(const double* A, const double* b, double* x) { return solve_psd6(A, b, x); }
void emu_rotation_from_crosscov(const double* m, double* R) { rotation_from_crosscov(m, R); }
void emu_jacobi_eig3(const double* A, double* w, double* V) {
    double M[9];
    std::memcpy(M, A, sizeof(M));
    jacobi_eig3(M, w, V);
}
int emu_rank3(const float* A) {
    float M[9];
    std::memcpy(M, A, sizeof(M));
    return fullpiv_qr_rank3(M);
}
void emu_angle_axis(const float* x, float* T16) {
    Mat4 T;
    angle_axis_to_mat4(x, T);
    std::memcpy(T16, T.m, sizeof(T.m));
}
float emu_angular_distance(const float* Ta, const float* Tb) {
    Mat4 A, B;
    std::memcpy(A.m, Ta, sizeof(A.m));
    std::memcpy(B.m, Tb, sizeof(B.m));
    return quat_angular_distance(quat_from_mat4(A), quat_from_mat4(B));
}
void emu_mat4_mul(const float* A, const float* B, float* C) {
    Mat4 a, b, c;
    std::memcpy(a.m, A, 64);
    std::memcpy(b.m, B, 64);
    mat4_mul(a, b, c);
    std::memcpy(C, c.m, 64);
}
uint32_t emu_seg_begin(int level, uint32_t seg, uint32_t n) { return seg_begin(level, seg, n); }
uint32_t emu_seg_of(uint32_t p, int level, uint32_t n) { return seg_of(p, level, n); }
uint32_t emu_float_ord(float f) { return float_ord(f); }
float emu_ord_float(uint32_t o) { return ord_float(o); }

}  // extern "C"
