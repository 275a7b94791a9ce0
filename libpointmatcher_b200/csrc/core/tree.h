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
//   descent   root -> leaf by the split planes only (one 8-byte load + ~10 instructions per level);
//             the squared plane distance of every level is cached in a per-lane scratch column.
//             The tree is complete, so every lane of a warp runs exactly D steps: no divergence.
//   backtrack a bit-trail holds one "far child pending" bit per level; the deepest pending level
//             is popped with clz and filtered by its cached plane distance; survivors rebuild the
//             far child's heap index from the current node (ancestor >> shift, ^ 1), test its
//             bounding-box distance (one 32-byte load), and only then descend into it.
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
#include <type_traits>

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

// compile-time loop: every array index below is a constant, whatever the unroller decides, so
// the candidate arrays can live in registers
template <int I, int N, typename F>
PM_HD void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// k best candidates, ascending, held in registers (all indices are compile-time constants).  A candidate is ONE 64-bit
// key, (distance bits << 32) | index: squared distances are non-negative floats, whose bit patterns order like unsigned
// integers, and indices are non-negative, so the lexicographic (dist, index) ranking is a single unsigned compare (half
// the instructions of the two-float-one-int version in the insertion chain, which is what a k > 1 search spends its
// time in).  A NaN distance is larger than +inf as a bit pattern: never accepted, as with float compares.
template <int KMAX>
struct TopK {
    unsigned long long key[KMAX];
    int k;
    unsigned long long wkey;  // cached k-th best
    PM_HD static unsigned long long pack(float d, int i) { return ((unsigned long long)f2u(d) << 32) | (unsigned long long)(uint32_t)i; }
    PM_HD void init(int k_, float max_r2) {
        k = k_;
        wkey = pack(max_r2, PM_NO_ID);
        static_for<0, KMAX>([&](auto J) { key[J] = wkey; });
    }
    PM_HD float worst_d() const { return u2f((uint32_t)(wkey >> 32)); }
    PM_HD int worst_id() const { return (int)(uint32_t)wkey; }
    PM_HD bool accepts(float d, int i) const { return pack(d, i) < wkey; }
    template <int J> PM_HD float D(std::integral_constant<int, J>) const { return u2f((uint32_t)(key[J] >> 32)); }
    template <int J> PM_HD int I(std::integral_constant<int, J>) const { return (int)(uint32_t)key[J]; }
    // insert (nd, ni); precondition: accepts(nd, ni)
    PM_HD void insert(float nd, int ni) {
        const unsigned long long c = pack(nd, ni);
        if (k == KMAX) {
            // the list is exactly as long as its registers (knn 10, 20, ...: the instantiated sizes): no per-slot `j < k`
            // test, one compare per slot carried over to its neighbour, and the k-th best is simply the last register
            bool lt_hi = true;  // c < key[KMAX - 1]: the precondition
            static_for<0, KMAX - 1>([&](auto I) {
                constexpr int j = KMAX - 1 - I;  // j = KMAX-1 .. 1; slot j - 1 is still untouched when slot j is rewritten
                const bool lt_lo = c < key[j - 1];
                key[j] = lt_lo ? key[j - 1] : (lt_hi ? c : key[j]);
                lt_hi = lt_lo;
            });
            if (lt_hi) key[0] = c;
            wkey = key[KMAX - 1];
            return;
        }
        static_for<0, KMAX - 1>([&](auto I) {
            constexpr int j = KMAX - 1 - I;
            if (j < k) key[j] = c < key[j - 1] ? key[j - 1] : (c < key[j] ? c : key[j]);
        });
        if (c < key[0]) key[0] = c;
        static_for<0, KMAX>([&](auto J) {
            if (J == k - 1) wkey = key[J];
        });
    }
    // entry j (runtime index) without dynamic array addressing
    PM_HD void get(int j, float& dj, int& ij) const {
        unsigned long long e = key[0];
        static_for<1, KMAX>([&](auto J) {
            if (J == j) e = key[J];
        });
        dj = u2f((uint32_t)(e >> 32));
        ij = (int)(uint32_t)e;
    }
    PM_HD bool contains(int q) const {
        bool f = false;
        static_for<0, KMAX>([&](auto J) { f = f || ((int)(uint32_t)key[J] == q); });
        return f;
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
    PM_HD bool accepts(float nd, int ni) const { return cand_less(nd, ni, d[0], id[0]); }
    template <int J> PM_HD float D(std::integral_constant<int, J>) const { return d[0]; }
    template <int J> PM_HD int I(std::integral_constant<int, J>) const { return id[0]; }
    PM_HD void insert(float nd, int ni) { d[0] = nd; id[0] = ni; }
    PM_HD void get(int, float& dj, int& ij) const { dj = d[0]; ij = id[0]; }
    PM_HD bool contains(int q) const { return id[0] == q; }
};

#if defined(PM_EMU_STATS) && !defined(__CUDACC__)
struct EmuStats { unsigned long long descent_steps, pops, box_tests, redescents, leaves; };
static EmuStats g_emu_stats = {0, 0, 0, 0, 0};
#define PM_STAT(field) (++g_emu_stats.field)
#else
#define PM_STAT(field) ((void)0)
#endif

// ------------------------------------------------------------------------------------------------
// Per-query search state and its three phases.  knn.cu runs the phases of the 32 lanes of a warp
// in lock step (descend* | scan leaf + filter pending levels | box test*), one step of every lane
// at a time, so the warp never diverges;
// knn_search_single() below composes the same phases sequentially (host harness, and the
// definition of the result).
//
// `plane` is a per-lane scratch column (shared memory on the device, element l at
// plane[l * stride]) that caches, for every level l, the squared distance from the query to the
// split plane crossed at that level: the cheapest valid lower bound for the pending sibling.
// It may be null: pending siblings are then judged by their boxes alone (the right trade when the
// search starts with a tight bound, so that hardly any sibling is pending at all).
// ------------------------------------------------------------------------------------------------
struct Lane {
    float qx, qy, qz;
    uint32_t node;      // current node (heap index)
    int level;          // its level
    uint32_t trail;     // bit l set: the sibling of our level-l ancestor has not been examined yet
    uint32_t visited;
};

PM_HD void lane_begin(Lane& s, float qx, float qy, float qz) {
    s.qx = qx; s.qy = qy; s.qz = qz;
    s.visited = 0;
    s.node = 1;
    s.level = 0;
    s.trail = 0;
}

PM_HD bool lane_descending(const Lane& s, const TreeView& t) { return s.level < t.depth; }

// coordinate `dim` of (x, y, z) without a branch (the ternary chain compiles to divergent branches)
PM_HD float select3(uint32_t dim, float x, float y, float z) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("{\n\t.reg .pred p0, p2;\n\tsetp.eq.u32 p0, %1, 0;\n\tsetp.eq.u32 p2, %1, 2;\n\tselp.f32 %0, %2, %3, p0;\n\tselp.f32 %0, %4, %0, p2;\n\t}"
        : "=f"(r) : "r"(dim), "f"(x), "f"(y), "f"(z));
    return r;
#else
    return dim == 0 ? x : (dim == 1 ? y : z);
#endif
}

// (Measured and dropped, round 2: the same planes stored three levels per 32-byte block, so that one load serves three
// levels and the chain of dependent loads is ~6 long instead of ~17.  Exact — same planes, same order — but the
// run-time select of one of seven plane registers per level costs more issue slots than the shorter chain saves: the
// seeded k = 1 match went from 0.173 to 0.207 ms per 1 M queries.  The kernel is issue bound before it is latency bound.)
// one step of the plane descent; `w` is the current k-th best distance: a sibling whose split
// plane is already farther than that can never be needed (w only shrinks), so it is not even
// recorded as pending
// (Measured and dropped, round 2: one level of look-ahead — the planes of both children requested as soon as the node is
// known, so that the load of level l + 1 is in flight while level l is decided.  No gain: 0.56 vs 0.50 ms for the first,
// unseeded k = 1 match of 1 M queries; the descent's loads mostly hit L1 and the extra selects and registers cost more.)
// `top` (optional): the split planes of the nodes with heap index < top_n, staged in shared memory by the kernel
PM_HD void lane_descend_step(Lane& s, const TreeView& t, float* plane, int stride, float w, const f2* top = nullptr, uint32_t top_n = 0) {
    const f2 sp = (top && s.node < top_n) ? top[s.node] : ldg2(t.splits + s.node);
    const float qd = select3(f2u(sp.y), s.qx, s.qy, s.qz);
    const float diff = fsub(qd, sp.x);
    const float pl = fmul(diff, diff);
    s.node = 2 * s.node + (qd >= sp.x ? 1u : 0u);
    ++s.level;
    if (!(pl > w)) s.trail |= 1u << s.level;
    if (plane) plane[s.level * stride] = pl;
    PM_STAT(descent_steps);
}

// distances to all points of the current leaf
template <int KMAX>
PM_HD void lane_scan_leaf(Lane& s, const TreeView& t, TopK<KMAX>& best) {
    const uint32_t leaf = s.node - (1u << t.depth);
    const uint32_t b = seg_begin(t.depth, leaf, t.n), e = seg_begin(t.depth, leaf + 1, t.n);
    if (KMAX == 1) {
#pragma unroll
        for (uint32_t j = 0; j < PM_LEAF_MAX; ++j) {
            const uint32_t p = b + j;
            if (p < e) {
                const f4 pt = ldg4(t.pts + p);
                const float dd = dist2(s.qx, s.qy, s.qz, pt.x, pt.y, pt.z);
                const int pi = (int)f2u(pt.w);
                if (best.accepts(dd, pi)) best.insert(dd, pi);
            }
        }
#ifdef PM_LEAF_ONE_PHASE
    } else if (true) {
        // (A/B builds) the one-phase scan: a real loop around the insertion
#pragma unroll 1
        for (uint32_t p = b; p < e; ++p) {
            const f4 pt = ldg4(t.pts + p);
            const float dd = dist2(s.qx, s.qy, s.qz, pt.x, pt.y, pt.z);
            const int pi = (int)f2u(pt.w);
            if (best.accepts(dd, pi)) best.insert(dd, pi);
        }
#endif
    } else {
        // k > 1, two phases.  (1) all distances of the leaf (independent loads, unrolled) -> a bit mask of the points
        // that beat the current k-th best; (2) only those are inserted, one per trip, re-derived from the (L1-resident)
        // point.  The insertion chain is ~6 k instructions and the lanes of a warp run in lock step: with the insert
        // inside the scan loop a warp paid for it at EVERY point position where ANY lane had a candidate (nearly all
        // eight), now it pays max over lanes of the number of candidates (one or two once the lists are good).
        uint32_t mask = 0;
#pragma unroll
        for (uint32_t j = 0; j < PM_LEAF_MAX; ++j) {
            const uint32_t p = b + j;
            if (p < e) {
                const f4 pt = ldg4(t.pts + p);
                const float dd = dist2(s.qx, s.qy, s.qz, pt.x, pt.y, pt.z);
                if (best.accepts(dd, (int)f2u(pt.w))) mask |= 1u << j;
            }
        }
        while (mask) {
            const uint32_t j = 31u - (uint32_t)clz32(mask & (0u - mask));  // lowest set bit: index order, like the one-phase scan
            mask &= mask - 1u;
            const f4 pt = ldg4(t.pts + b + j);
            const float dd = dist2(s.qx, s.qy, s.qz, pt.x, pt.y, pt.z);
            const int pi = (int)f2u(pt.w);
            if (best.accepts(dd, pi)) best.insert(dd, pi);  // the bound may have shrunk since phase 1
        }
    }
    s.visited += e - b;
    PM_STAT(leaves);
}

// Drop every pending level whose split plane is out of reach of the current bound.  Branch-free
// over the levels (one LDS + compare + predicated bit-clear each).  Worth its ~4 instructions per
// level only when the bound has just shrunk under many pending levels (lane_wants_filter): the
// first leaf of an unseeded search, or a seed that turned out to be far off.
template <int KMAX>
PM_HD void lane_filter_trail(Lane& s, const TreeView& t, const TopK<KMAX>& best, const float* plane, int stride) {
    const float w = best.worst_d();
    uint32_t keep = s.trail;
    for (int l = 1; l <= t.depth; ++l) {
        PM_STAT(pops);
        if (plane[l * stride] > w) keep &= ~(1u << l);  // the whole far side of that split is out of reach
    }
    s.trail = keep;
}

#define PM_FILTER_MIN_PENDING 6
PM_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
// w_before: the bound the last descent was made with
PM_HD bool lane_wants_filter(const Lane& s, float w_before, float w_now) {
    return w_now < w_before && popc32(s.trail) >= PM_FILTER_MIN_PENDING;
}

// One pending sibling, the deepest: plane test against the current bound first (cached distance,
// one LDS), then the box test.  Returns true when the sibling has to be searched (the lane then
// continues with descend steps from it); false means "rejected", the caller tries the next
// pending level while s.trail != 0.
template <int KMAX>
PM_HD bool lane_box_step(Lane& s, const TreeView& t, const TopK<KMAX>& best, const float* plane, int stride) {
    const int l = 31 - clz32(s.trail);
    s.trail &= ~(1u << l);
    PM_STAT(pops);
    if (plane && plane[l * stride] > best.worst_d()) return false;
    const uint32_t far = (s.node >> (s.level - l)) ^ 1u;
    PM_STAT(box_tests);
    const float db = box_dist2(s.qx, s.qy, s.qz, ldg4(t.boxes + 2 * (size_t)far), ldg4(t.boxes + 2 * (size_t)far + 1));
    if (db > best.worst_d()) return false;
    s.node = far;
    s.level = l;
    PM_STAT(redescents);
    return true;
}

// Exact k-nearest-neighbour search of one query.  `best` may already hold real candidates (e.g.
// the previous iteration's match re-measured, k = 1): they only tighten the bounds.  Returns the
// number of reference points whose distance was evaluated (the analogue of libnabo's visit count,
// MatchersImpl.cpp:98).
template <int KMAX>
PM_HD uint32_t knn_search_single(const TreeView& t, float qx, float qy, float qz, TopK<KMAX>& best) {
    if (t.n == 0) return 0;
    float plane[PM_MAX_DEPTH + 2];
    Lane s;
    lane_begin(s, qx, qy, qz);
    for (;;) {
        const float w0 = best.worst_d();
        while (lane_descending(s, t)) lane_descend_step(s, t, plane, 1, w0);
        lane_scan_leaf<KMAX>(s, t, best);
        if (lane_wants_filter(s, w0, best.worst_d())) lane_filter_trail<KMAX>(s, t, best, plane, 1);
        bool found = false;
        while (!found && s.trail != 0) found = lane_box_step<KMAX>(s, t, best, plane, 1);
        if (!found) break;
    }
    return s.visited;
}

}  // namespace pm
