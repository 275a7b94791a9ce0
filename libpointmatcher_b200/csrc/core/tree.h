// core/tree.h — the search structure behind KDTreeMatcher::init / findClosests
// (MatchersImpl.cpp:77-101) and its per-query traversal.
//
// Structure (built on the device by tree_build.cu): a *complete* binary tree of depth D over the
// reference points.  Level l has 2^l nodes; node (l, s) owns the contiguous range
// [seg_begin(l, s), seg_begin(l, s + 1)) of the sorted point array, where
// seg_begin(l, s) = floor(s * N / 2^l).  Ranges of children are exact halves of the parent
// (median split along the parent's widest axis), so no child pointers, counts or leaf tables are
// stored: per inner node one (split value, axis) pair, per node one axis-aligned bounding box.
// Leaves (level D) hold floor(N / 2^D) or ceil(N / 2^D) <= PM_LEAF_MAX points.
//
// Traversal (one thread per query, no stack):
//   descent   root -> leaf by the split planes only (one 8-byte load + ~10 instructions per level).
//             The tree is complete, so every lane of a warp runs exactly D steps: no divergence.
//   backtrack a bit-trail holds one "far child pending" bit per level; the deepest pending level
//             is popped with clz, the far child's heap index is rebuilt from the current node
//             (ancestor >> shift, ^ 1), first filtered by the plane distance, then by its
//             bounding-box distance (one 32-byte load), and only then descended.
//
// Exactness.  A subtree is skipped only when a lower bound of the distance to all of its points is
// > the current k-th best distance.  Both bounds are evaluated with the same operation order and
// rounding as point distances (core/common.h dist2); float rounding is monotone, hence for every
// point p of the subtree dist2(q, p) >= bound holds *in float arithmetic*:
//   plane : the far side satisfies |q[d] - p[d]| >= |q[d] - split| (left points <= split <= right)
//   box   : per axis |q[a] - p[a]| >= gap(q[a], [lo[a], hi[a]])
// so pruning never drops a point of the answer.  Candidates are ranked lexicographically by
// (dist2, reference index): the answer is unique, independent of traversal order, and equals what
// libnabo's brute-force search returns (index-ascending scan, strict '<').
#pragma once
#include "common.h"

#ifndef PM_LEAF_MAX
#define PM_LEAF_MAX 8
#endif
#define PM_MAX_DEPTH 30

namespace pm {

PM_HD uint32_t seg_begin(int level, uint32_t seg, uint32_t n) { return (uint32_t)(((uint64_t)seg * n) >> level); }
// segment of sorted position p at `level`: the s with seg_begin(l, s) <= p < seg_begin(l, s+1)
PM_HD uint32_t seg_of(uint32_t p, int level, uint32_t n) { return (uint32_t)(((((uint64_t)p + 1) << level) - 1) / n); }

// smallest depth whose leaves hold at most PM_LEAF_MAX points
PM_HD int tree_depth_for(uint32_t n) {
    int d = 0;
    while ((((uint64_t)n + ((1ull << d) - 1)) >> d) > PM_LEAF_MAX) ++d;
    return d;
}

// Read-only view handed to the kernels (all arrays indexed by heap index, root = 1, children
// 2i and 2i + 1).
//  splits: inner node i -> {split value, bit pattern of the split axis 0/1/2}
//  boxes:  node i -> boxes[2i] = (lo.x, lo.y, lo.z, -), boxes[2i+1] = (hi.x, hi.y, hi.z, -)
//  pts:    points in leaf order, w = bit pattern of the original column index
struct TreeView {
    const f2* splits;
    const f4* boxes;
    const f4* pts;
    uint32_t n;
    int depth;
};

PM_HD f4 ldg4(const f4* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
PM_HD f2 ldg2(const f2* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
PM_HD int clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}

// distance from q to the interval [lo, hi] on one axis (0 inside)
PM_HD float axis_gap(float q, float lo, float hi) { return fmaxf(fmaxf(fsub(lo, q), fsub(q, hi)), 0.f); }

PM_HD float box_dist2(float qx, float qy, float qz, f4 lo, f4 hi) {
    const float dx = axis_gap(qx, lo.x, hi.x), dy = axis_gap(qy, lo.y, hi.y), dz = axis_gap(qz, lo.z, hi.z);
    return fadd(fadd(fmul(dx, dx), fmul(dy, dy)), fmul(dz, dz));
}

// lexicographic (dist, index) comparison
PM_HD bool cand_less(float d, int i, float bd, int bi) { return d < bd || (d == bd && i < bi); }

#define PM_NO_ID 0x7fffffff

// k best candidates, ascending, held in registers when KMAX is small (all loops unrolled).
template <int KMAX>
struct TopK {
    float d[KMAX];
    int id[KMAX];
    int k;
    float wd;  // cached k-th best (d[k-1], id[k-1])
    int wi;
    PM_HD void init(int k_, float max_r2) {
        k = k_;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) { d[j] = max_r2; id[j] = PM_NO_ID; }
        wd = max_r2;
        wi = PM_NO_ID;
    }
    PM_HD float worst_d() const { return wd; }
    PM_HD int worst_id() const { return wi; }
    // insert (nd, ni); precondition: cand_less(nd, ni, worst)
    PM_HD void insert(float nd, int ni) {
#pragma unroll
        for (int j = KMAX - 1; j >= 1; --j) {
            if (j < k) {
                if (cand_less(nd, ni, d[j - 1], id[j - 1])) { d[j] = d[j - 1]; id[j] = id[j - 1]; }
                else if (cand_less(nd, ni, d[j], id[j])) { d[j] = nd; id[j] = ni; }
            }
        }
        if (cand_less(nd, ni, d[0], id[0])) { d[0] = nd; id[0] = ni; }
#pragma unroll
        for (int j = 0; j < KMAX; ++j)
            if (j == k - 1) { wd = d[j]; wi = id[j]; }
    }
};

template <>
struct TopK<1> {
    float d[1];
    int id[1];
    int k;
    PM_HD void init(int, float max_r2) { k = 1; d[0] = max_r2; id[0] = PM_NO_ID; }
    PM_HD float worst_d() const { return d[0]; }
    PM_HD int worst_id() const { return id[0]; }
    PM_HD void insert(float nd, int ni) { d[0] = nd; id[0] = ni; }
};

// Exact k-nearest-neighbour search of one query.  Returns the number of reference points whose
// distance was evaluated (the analogue of libnabo's visit count, MatchersImpl.cpp:98).
template <int KMAX>
PM_HD uint32_t knn_search(const TreeView& t, float qx, float qy, float qz, TopK<KMAX>& best) {
    uint32_t visited = 0;
    if (t.n == 0) return 0;
    const int D = t.depth;
    const uint32_t first_leaf = 1u << D;
    uint32_t node = 1;       // current node (heap index)
    int level = 0;           // its level
    uint32_t trail = 0;      // bit l set: the far child at level l (sibling of our level-l ancestor) is pending
    {
        const float dr = box_dist2(qx, qy, qz, ldg4(t.boxes + 2), ldg4(t.boxes + 3));
        if (dr > best.worst_d()) return 0;
    }
    for (;;) {
        // ---- descent by split planes: exactly D - level steps
        while (level < D) {
            const f2 s = ldg2(t.splits + node);
            const uint32_t dim = f2u(s.y);
            const float qd = dim == 0 ? qx : (dim == 1 ? qy : qz);
            node = 2 * node + (qd >= s.x ? 1u : 0u);
            ++level;
            trail |= 1u << level;
        }
        // ---- leaf
        {
            const uint32_t leaf = node - first_leaf;
            const uint32_t b = seg_begin(D, leaf, t.n), e = seg_begin(D, leaf + 1, t.n);
#pragma unroll
            for (uint32_t j = 0; j < PM_LEAF_MAX; ++j) {
                const uint32_t p = b + j;
                if (p < e) {
                    const f4 pt = ldg4(t.pts + p);
                    const float dd = dist2(qx, qy, qz, pt.x, pt.y, pt.z);
                    const int pi = (int)f2u(pt.w);
                    if (cand_less(dd, pi, best.worst_d(), best.worst_id())) best.insert(dd, pi);
                }
            }
            visited += e - b;
        }
        // ---- backtrack: deepest pending far child that can still hold a better candidate
        bool found = false;
        while (trail != 0) {
            const int l = 31 - clz32(trail);
            trail &= ~(1u << l);
            const uint32_t far = (node >> (level - l)) ^ 1u;
            const f2 s = ldg2(t.splits + (far >> 1));
            const uint32_t dim = f2u(s.y);
            const float qd = dim == 0 ? qx : (dim == 1 ? qy : qz);
            const float diff = fsub(qd, s.x);
            const float w = best.worst_d();
            if (fmul(diff, diff) > w) continue;
            const float db = box_dist2(qx, qy, qz, ldg4(t.boxes + 2 * (size_t)far), ldg4(t.boxes + 2 * (size_t)far + 1));
            if (db > w) continue;
            node = far;
            level = l;
            found = true;
            break;
        }
        if (!found) break;
    }
    return visited;
}

}  // namespace pm
