// core/tree.h — the search structure behind KDTreeMatcher::init / findClosests
// (MatchersImpl.cpp:77-101) and its per-query traversal.
//
// Structure (built on the device by tree_build.cu): a *complete* binary tree of depth D over the
// reference points.  Level l has 2^l nodes; node (l, s) owns the contiguous range
// [seg_begin(l, s), seg_begin(l, s + 1)) of the sorted point array, where
// seg_begin(l, s) = floor(s * N / 2^l).  Ranges of children are exact halves of the parent
// (median split along the parent's widest axis), so no child pointers, counts or leaf tables are
// stored — only one axis-aligned bounding box per node.  Leaves (level D) hold
// floor(N / 2^D) or ceil(N / 2^D) <= PM_LEAF_MAX points.
//
// Exactness.  The traversal prunes a node only when its box distance is > the current k-th best
// distance.  Box distances are evaluated with the same operation order and rounding as point
// distances (core/common.h dist2); float rounding is monotone, hence for every point p inside a
// box, dist2(q, p) >= box_dist2(q, box) holds *in float arithmetic*, and pruning can never drop a
// point that belongs to the answer.  Candidates are ranked lexicographically by
// (dist2, reference index): the answer is unique, independent of traversal order, and equals
// what libnabo's brute-force search returns (index-ascending scan, strict '<').
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

// Read-only view handed to the kernels.
//  nodes: 3 f4 per inner node i (heap index, root = 1, children 2i and 2i+1), at nodes[3*(i-1)]:
//         {loL.x, hiL.x, loR.x, hiR.x}, {.. y ..}, {.. z ..}  (boxes of the two children)
//  pts:   points in leaf order, w = bit pattern of the original column index
struct TreeView {
    const f4* nodes;
    const f4* pts;
    uint32_t n;
    int depth;
    f4 root_lo, root_hi;  // box of the whole cloud (w unused)
};

PM_HD f4 ldg4(const f4* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

// distance from q to the interval [lo, hi] on one axis (0 inside)
PM_HD float axis_gap(float q, float lo, float hi) { return fmaxf(fmaxf(fsub(lo, q), fsub(q, hi)), 0.f); }

PM_HD float box_dist2(float qx, float qy, float qz, float lox, float hix, float loy, float hiy, float loz, float hiz) {
    const float dx = axis_gap(qx, lox, hix), dy = axis_gap(qy, loy, hiy), dz = axis_gap(qz, loz, hiz);
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
    uint32_t stack_node[PM_MAX_DEPTH + 1];
    float stack_d[PM_MAX_DEPTH + 1];
    int sp = 0;
    const uint32_t first_leaf = 1u << t.depth;
    uint32_t node = 1;
    {
        const float dr = box_dist2(qx, qy, qz, t.root_lo.x, t.root_hi.x, t.root_lo.y, t.root_hi.y, t.root_lo.z, t.root_hi.z);
        if (dr > best.worst_d()) return 0;
    }
    for (;;) {
        if (node >= first_leaf) {
            const uint32_t leaf = node - first_leaf;
            const uint32_t b = seg_begin(t.depth, leaf, t.n), e = seg_begin(t.depth, leaf + 1, t.n);
            for (uint32_t p = b; p < e; ++p) {
                const f4 pt = ldg4(t.pts + p);
                const float dd = dist2(qx, qy, qz, pt.x, pt.y, pt.z);
                const int pi = (int)f2u(pt.w);
                if (cand_less(dd, pi, best.worst_d(), best.worst_id())) best.insert(dd, pi);
            }
            visited += e - b;
        } else {
            const f4* nb = t.nodes + 3 * (size_t)(node - 1);
            const f4 bx = ldg4(nb), by = ldg4(nb + 1), bz = ldg4(nb + 2);
            const float dl = box_dist2(qx, qy, qz, bx.x, bx.y, by.x, by.y, bz.x, bz.y);
            const float dr = box_dist2(qx, qy, qz, bx.z, bx.w, by.z, by.w, bz.z, bz.w);
            const float w = best.worst_d();
            const bool vl = dl <= w, vr = dr <= w;
            if (vl && vr) {
                // nearer child first; ties go left (deterministic, does not affect the result)
                const bool left_first = dl <= dr;
                stack_node[sp] = left_first ? 2 * node + 1 : 2 * node;
                stack_d[sp] = left_first ? dr : dl;
                ++sp;
                node = left_first ? 2 * node : 2 * node + 1;
                continue;
            }
            if (vl) { node = 2 * node; continue; }
            if (vr) { node = 2 * node + 1; continue; }
        }
        // pop the next subtree that can still contain a better candidate
        bool found = false;
        while (sp > 0) {
            --sp;
            if (stack_d[sp] <= best.worst_d()) { node = stack_node[sp]; found = true; break; }
        }
        if (!found) break;
    }
    return visited;
}

}  // namespace pm
