// knn.cu — K2: RigidTransformation::compute fused with KDTreeMatcher::findClosests
// (TransformationsImpl.cpp:49-87, MatchersImpl.cpp:85-101).
//
// One thread per query; the reading is resident in Morton order, so the 32 lanes of a warp hold
// neighbouring points and walk almost the same part of the tree (the tree of a 1 M-point
// reference is 1 MB of split planes + 8 MB of boxes + 16 MB of points, resident in the 126 MB L2;
// leaf and node loads of a warp mostly hit the same L1 lines).  The search of core/tree.h is run
// as phases — plane descent | leaf scan + plane filter of the pending levels | box tests — and the
// lanes of a warp advance through them in LOCK STEP (warp-vote loops: one step of every lane at a
// time): without that, independent thread scheduling lets every lane drift into its own phase and
// the warp executes ~3 of 32 lanes per instruction (measured, profiles/).  The per-level plane
// distances of the descent are cached in shared memory (one column per lane), so rejecting a
// pending sibling costs one LDS and a compare.  T_iter is read
// from the device-resident IcpState, so no host round trip separates iterations; in iterations
// >= 2 (k = 1) the previous match, re-measured, seeds the search with a tight bound.  Queries
// still open after a budget of leaves go to a second, warp-per-query kernel (stage 2).
#include <string.h>

#include "normals.cuh"
#include "pmgpu_internal.cuh"

namespace pm {

namespace {

#ifndef PM_KNN_BLOCK
#define PM_KNN_BLOCK 128
#endif
constexpr int KNN_BLOCK = PM_KNN_BLOCK;

// Search radius of one launch and what an empty result slot is written as.  Without a cap (or with
// a cap that the caller's maxDist already undercuts) a miss is "no match within maxDist": id -1,
// dist +inf, the pair the reference's Matches carry (MatchersImpl.cpp:95-99).  Under the cap a miss
// only says "farther than the cap": it stays a finite, rejected match (id -2, dist FLT_MAX), so
// the quantile population and the rejected-point statistics are the ones the exact search gives.
struct Cap { float r2; int miss_id; float miss_d; };
#define PM_CAPPED_ID (-2)
#define PM_CAPPED_DIST 3.402823466e+38f
__device__ __forceinline__ Cap knn_cap(const IcpState* state, int use_cap, float max_r2) {
    Cap c = {max_r2, -1, pm_inf()};
    if (use_cap) {
        const float r = state->cap;
        if (r < max_r2) { c.r2 = r; c.miss_id = PM_CAPPED_ID; c.miss_d = PM_CAPPED_DIST; }
    }
    return c;
}

// NORMALS: K8 — a self-query whose epilogue turns the k neighbours into the point's surface normal (normals.cuh) instead of
// writing them out; `ids` / `dists` are then unused
template <int KMAX, bool PLANES, bool NORMALS>
__global__ void __launch_bounds__(KNN_BLOCK) knn_kernel(TreeView tree, const f4* __restrict__ queries, int nq, const IcpState* __restrict__ state,
                                                        int use_T, int gated, int self_query, int k, float max_r2,
                                                        const f4* __restrict__ ref_orig, int use_seed, int32_t* __restrict__ ids,
                                                        float* __restrict__ dists, unsigned long long* visits, int budget,
                                                        uint32_t* __restrict__ overflow, unsigned* overflow_count, int use_cap,
                                                        const float* __restrict__ var_r2, NormalsSink ns, int pos_offset, uint2* __restrict__ resume) {
    extern __shared__ float s_plane[];  // [depth + 1][KNN_BLOCK]: cached plane distances, one column per lane
    __shared__ Mat4 sT;
    pdl_wait();     // (dependent of the previous iteration's last kernel)
    pdl_release();  // stage 2 may be set up from here on; it waits for this grid's completion itself
    if (gated && state->iterate == 0) return;
#ifdef PM_TOP_SMEM
    // (A/B build) the split planes of the top PM_TOP_SMEM levels staged in shared memory by one bulk copy (TMA, 1-D): the
    // first PM_TOP_SMEM dependent loads of every descent become shared-memory loads
    __shared__ __align__(128) f2 s_top[1 << PM_TOP_SMEM];
    __shared__ __align__(8) unsigned long long s_top_bar;
    const uint32_t top_n = min(1u << PM_TOP_SMEM, 1u << tree.depth);
    {
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_top_bar), dst = (unsigned)__cvta_generic_to_shared(s_top);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned bytes = top_n * (unsigned)sizeof(f2);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(tree.splits), "r"(bytes), "r"(bar)
                         : "memory");
        }
    }
    const f2* top = s_top;
#else
    const f2* top = nullptr;
    const uint32_t top_n = 0;
#endif
    if (use_T) {
        if (threadIdx.x < 16) sT.m[threadIdx.x] = state->T_iter.m[threadIdx.x];
        __syncthreads();
    }
    float* plane = PLANES ? s_plane + threadIdx.x : nullptr;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool running = t < nq;
    uint32_t qi = (uint32_t)t;
    Lane s;
    TopK<KMAX> best;
    // fused ICP loop: the search may stop at the radius the previous iteration's outlier filters
    // make sufficient (state->cap, verified by the select kernels afterwards)
    Cap cap = knn_cap(state, use_cap, max_r2);
    if (var_r2 && t < nq) cap.r2 = var_r2[t];  // KDTreeVarDistMatcher: the query's own radius
    best.init(k, cap.r2);
    s.visited = 0;
    if (running) {
        f4 q = queries[t];
        if (self_query) {
            // K8: the query is the reference point at leaf-order position t; its original column
            // travels in w
            qi = __float_as_uint(q.w);
            q.w = 1.f;
        }
        if (use_T) q = transform_point(sT, q);
        lane_begin(s, q.x, q.y, q.z);
        if (KMAX == 1 && use_seed) {
            // ICP iterations >= 2: the previous match of this query (still resident in ids),
            // re-measured under the new T_iter, is a real candidate that makes the bound tight
            // before the first leaf is reached.  It cannot change the answer, only the work.
            const int prev = ids[t];
            if (prev >= 0) {
                const f4 r = __ldg(ref_orig + prev);
                const float dd = dist2(q.x, q.y, q.z, r.x, r.y, r.z);
                if (best.accepts(dd, prev)) best.insert(dd, prev);
            }
        }
        if (KMAX > 1 && use_seed) {
            // k > 1: the previous k matches of this query, re-measured under the new T_iter, are k distinct real points, so
            // the largest of their distances bounds the k-th nearest distance from above: the search may start with that
            // radius.  Nothing is inserted — the points themselves lie within the radius (<=) and are found again — so the
            // result is the unseeded one; only the pruning starts earlier.  Unusable when a slot was empty or capped.
            float bound = 0.f;
            bool usable = true;
            for (int j = 0; j < k; ++j) {
                const int prev = ids[(size_t)t * k + j];
                if (prev < 0) { usable = false; break; }
                const f4 r = __ldg(ref_orig + prev);
                bound = fmaxf(bound, dist2(q.x, q.y, q.z, r.x, r.y, r.z));
            }
            if (usable && bound < best.worst_d()) best.init(k, bound);
        }
    }
    // all 32 lanes stay in the loop until the slowest is done; phases re-converge the warp.
    // A lane whose query is still open after `budget` leaves hands it (with the candidates found
    // so far, written below like a final result) to the warp-cooperative stage 2: a few
    // pathological queries (e.g. equidistant to a whole scan ring) would otherwise hold their
    // warp — and the kernel — for thousands of rounds.
    int rounds = 0;
    unsigned my_slot = 0xffffffffu;  // K8: the stage-2 queue slot this query was handed over at
#ifdef PM_TOP_SMEM
    {   // the staged planes have landed (the copy ran behind the query load, the transform and the seed gather)
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_top_bar);
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
    }
#endif
    while (__any_sync(0xffffffffu, running)) {
        // descend: one level per step for every lane that is not at a leaf yet
        const float w0 = best.worst_d();
        while (__any_sync(0xffffffffu, running && lane_descending(s, tree)))
        {
            if (running && lane_descending(s, tree)) lane_descend_step(s, tree, plane, KNN_BLOCK, w0, top, top_n);
#ifdef PM_DESCENT_STEPS  // (A/B builds) more levels per warp vote
#pragma unroll
            for (int u = 1; u < PM_DESCENT_STEPS; ++u)
                if (running && lane_descending(s, tree)) lane_descend_step(s, tree, plane, KNN_BLOCK, w0, top, top_n);
#endif
        }
        // leaf; then the plane filter over all pending levels where it pays
        bool refilter = false;
        if (running) {
            lane_scan_leaf<KMAX>(s, tree, best);
            refilter = PLANES && lane_wants_filter(s, w0, best.worst_d());
        }
        if (__any_sync(0xffffffffu, refilter))
            if (refilter) lane_filter_trail<KMAX>(s, tree, best, plane, KNN_BLOCK);
        // pending siblings, deepest first, one per step, until every lane has either a subtree to
        // search or nothing left
        bool found = false;
        while (__any_sync(0xffffffffu, running && !found && s.trail != 0))
            if (running && !found && s.trail != 0) found = lane_box_step<KMAX>(s, tree, best, plane, KNN_BLOCK);
        if (running && !found) running = false;  // search complete
        ++rounds;
        if (running && rounds >= budget) {
            const unsigned slot = atomicAdd(overflow_count, 1u);
            if (!NORMALS || slot < ns.scratch_cap) {
                running = false;
                overflow[slot] = (uint32_t)t;
                // where the search stands: the subtree it was about to enter and the siblings still pending above it — the
                // whole of what is left, so stage 2 continues instead of walking down from the root again
                resume[slot] = make_uint2(s.node, s.trail);
                my_slot = slot;  // handed over: unfilled slots below carry the radius this search ran with
            } else {
                budget = 0x7fffffff;  // K8: the hand-over scratch is full — this lane finishes its query here
            }
        }
    }
    if (NORMALS) {
        bool deg = false;
        if (t < nq) {
            if (my_slot != 0xffffffffu) {
                // handed to stage 2: the candidates found so far travel in the scratch, by queue slot
                static_for<0, KMAX>([&](auto J) {
                    if (J < k) {
                        ns.scratch_ids[(size_t)my_slot * k + J] = best.I(J) != PM_NO_ID ? best.I(J) : -1;
                        ns.scratch_d[(size_t)my_slot * k + J] = best.D(J);
                    }
                });
            } else {
                deg = normals_epilogue<KMAX>(best, k, ns, ns.by_position ? (size_t)(pos_offset + t) : (size_t)qi, s.qx, s.qy, s.qz);
            }
        }
        const unsigned dm = __ballot_sync(0xffffffffu, deg);
        if ((threadIdx.x & 31) == 0 && dm) atomicAdd(ns.degenerate, __popc(dm));
    } else if (t < nq) {
        int32_t* oi = ids + (size_t)qi * k;
        float* od = dists + (size_t)qi * k;
        const bool handed = my_slot != 0xffffffffu;
        static_for<0, KMAX>([&](auto J) {
            if (J < k) {
                const bool valid = best.I(J) != PM_NO_ID && best.D(J) != pm_inf();
                oi[J] = valid ? best.I(J) : cap.miss_id;
                // a query handed to stage 2 passes on the radius it was searching with (a seeded bound, a per-point
                // distance) in its unfilled slots, so that stage 2 does not start again from the caller's maxDist
                od[J] = valid ? best.D(J) : (handed ? best.D(J) : cap.miss_d);
            }
        });
    }
    if (visits) {
        // warp-aggregated statistics (Matcher::visitCounter)
        unsigned v = (t < nq) ? s.visited : 0u;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(visits, (unsigned long long)v);
    }
}


// ---- stage 2: one warp per left-over query ------------------------------------------------------
// The 32 lanes search ONE query together: a node is expanded 5 levels at a time (32 descendants,
// one box test per lane), surviving inner nodes go on a small shared-memory stack, surviving
// leaves are scanned sixteen at a time (lane = leaf slot x point, four loads in flight per lane).
// The candidate list is replicated in every lane and updated with warp-uniform inserts, seeded with what stage 1 had found.  Same
// bounds, same ranking: the result is the one the single-lane search would have produced.
// The candidate list of stage 2 is DISTRIBUTED over the warp: slot s lives in lane s & 31 (register s >> 5), ascending in
// (dist, index) like TopK.  An insertion is then one compare per lane, a ballot that gives the position, and one shuffle
// that moves the tail up by a lane — ~15 warp instructions instead of the ~10 k of a list replicated in every lane
// (which was 60 % of this kernel's instructions at k = 10, profiles/r2_knn_overflow_k10_hot_lines.txt).
template <int KMAX>
struct WarpTopK {
    static constexpr int R = (KMAX + 31) / 32;
    static constexpr unsigned FULL = 0xffffffffu;
    unsigned long long key[R];
    unsigned long long wkey;  // the k-th best, replicated
    int k, lane;
    __device__ __forceinline__ static unsigned long long pack(float d, int i) { return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned long long)(uint32_t)i; }
    __device__ __forceinline__ float worst_d() const { return __uint_as_float((uint32_t)(wkey >> 32)); }
    __device__ __forceinline__ bool accepts(float d, int i) const { return pack(d, i) < wkey; }
    __device__ __forceinline__ void refresh_worst() {
        unsigned long long src = key[0];  // the register that holds slot k - 1, without dynamic register indexing
#pragma unroll
        for (int r = 1; r < R; ++r)
            if (((k - 1) >> 5) == r) src = key[r];
        wkey = __shfl_sync(FULL, src, (k - 1) & 31);
    }
    // all slots (radius, no point) except the real candidates the caller fills in
    __device__ __forceinline__ void init(int k_, int lane_, float r0) {
        k = k_; lane = lane_;
#pragma unroll
        for (int r = 0; r < R; ++r) key[r] = (r * 32 + lane < k) ? pack(r0, PM_NO_ID) : ~0ull;  // slots >= k: never smaller than anything
        wkey = pack(r0, PM_NO_ID);
    }
    __device__ __forceinline__ bool contains(int q) const {
        bool f = false;
#pragma unroll
        for (int r = 0; r < R; ++r) f = f || ((int)(uint32_t)key[r] == q && key[r] != ~0ull);
        return __any_sync(FULL, f);
    }
    // warp-uniform (nd, ni); precondition: accepts(nd, ni) and not contained
    __device__ __forceinline__ void insert(float nd, int ni) {
        const unsigned long long c = pack(nd, ni);
        int pos = 0;  // entries smaller than c = its slot
#pragma unroll
        for (int r = 0; r < R; ++r) pos += __popc(__ballot_sync(FULL, key[r] < c));
#pragma unroll
        for (int r = R - 1; r >= 0; --r) {
            unsigned long long up = __shfl_up_sync(FULL, key[r], 1);
            if (r > 0) {
                const unsigned long long wrap = __shfl_sync(FULL, key[r - 1], 31);
                if (lane == 0) up = wrap;
            }
            const int s = r * 32 + lane;
            key[r] = (s > pos && s < k) ? up : (s == pos ? c : key[r]);
        }
        refresh_worst();
    }
    // slot j (warp-uniform j), in every lane
    __device__ __forceinline__ unsigned long long slot(int j) const {
        unsigned long long e = __shfl_sync(FULL, key[0], j & 31);
#pragma unroll
        for (int r = 1; r < R; ++r) {
            const unsigned long long o = __shfl_sync(FULL, key[r], j & 31);
            if ((j >> 5) == r) e = o;
        }
        return e;
    }
};

constexpr int OVF_STACK = 224;  // <= 32 pushes per expansion level, <= 6 levels of expansion (depth <= 30)

template <int KMAX, bool NORMALS>
__global__ void __launch_bounds__(128) knn_overflow_kernel(TreeView tree, const f4* __restrict__ queries, const IcpState* __restrict__ state, int use_T,
                                                           int gated, int self_query, int k, float max_r2, const uint32_t* __restrict__ overflow,
                                                           unsigned* overflow_count, unsigned* next_count, int32_t* __restrict__ ids,
                                                           float* __restrict__ dists, unsigned long long* visits, int use_cap,
                                                           const float* __restrict__ var_r2, NormalsSink ns, int pos_offset, const uint2* __restrict__ resume) {
    __shared__ uint32_t s_stack[4][OVF_STACK];
    __shared__ Mat4 sT;
    pdl_wait();  // dependent of stage 1
    if (gated && state->iterate == 0) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) *next_count = 0;  // the counter the NEXT launch of stage 1 will use
    if (use_T) {
        if (threadIdx.x < 16) sT.m[threadIdx.x] = state->T_iter.m[threadIdx.x];
        __syncthreads();
    }
    unsigned count = *overflow_count;
    if (NORMALS && count > ns.scratch_cap) count = ns.scratch_cap;  // later arrivals finished in stage 1
#ifdef PM_PROFILE_NS
    if (blockIdx.x == 0 && threadIdx.x == 0) printf("stage 2: %u queries\n", count);
#endif
    const Cap cap = knn_cap(state, use_cap, max_r2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lanes_lt = (1u << lane) - 1u;
    uint32_t* stack = s_stack[warp];
    const int D = tree.depth;
    unsigned long long my_visits = 0;
    for (unsigned w = blockIdx.x * 4 + warp; w < count; w += gridDim.x * 4) {
        const uint32_t t = overflow[w];
        f4 q = queries[t];
        uint32_t qi = t;
        if (self_query) { qi = __float_as_uint(q.w); q.w = 1.f; }
        if (use_T) q = transform_point(sT, q);
        // seed: the candidates stage 1 left in the result arrays (K8: in the scratch, by queue slot) — real points,
        // ascending, so lane j takes entry j as it is; the unfilled slots (id < 0) carry the radius stage 1 searched with
        WarpTopK<KMAX> best;
        {
            int sid[WarpTopK<KMAX>::R];
            float sd[WarpTopK<KMAX>::R];
            unsigned rbits = __float_as_uint(var_r2 ? var_r2[t] : cap.r2);
#pragma unroll
            for (int r = 0; r < WarpTopK<KMAX>::R; ++r) {
                const int j = r * 32 + lane;
                sid[r] = -1;
                sd[r] = pm_inf();
                if (j < k) {
                    sid[r] = NORMALS ? ns.scratch_ids[(size_t)w * k + j] : ids[(size_t)qi * k + j];
                    sd[r] = NORMALS ? ns.scratch_d[(size_t)w * k + j] : dists[(size_t)qi * k + j];
                    if (sid[r] < 0) rbits = min(rbits, __float_as_uint(sd[r]));  // non-negative floats order like their bits
                }
            }
            rbits = __reduce_min_sync(0xffffffffu, rbits);
            best.init(k, lane, __uint_as_float(rbits));
#pragma unroll
            for (int r = 0; r < WarpTopK<KMAX>::R; ++r)
                if (sid[r] >= 0) best.key[r] = WarpTopK<KMAX>::pack(sd[r], sid[r]);
            best.refresh_worst();
        }
        // the work list: the whole tree (queries that come straight here), or what stage 1 left — the pending siblings of its
        // path, shallowest first, and on top the subtree it was about to enter
        int sp = 0;
        if (resume) {
            const uint2 rs = resume[w];
            const int lvl = 31 - __clz((int)rs.x);
            unsigned tr = rs.y;
            if (lane == 0) {
                while (tr) {
                    const int l = __ffs((int)tr) - 1;
                    tr &= tr - 1u;
                    stack[sp++] = (rs.x >> (lvl - l)) ^ 1u;
                }
                stack[sp++] = rs.x;
            }
            sp = __shfl_sync(0xffffffffu, sp, 0);
        } else {
            if (lane == 0) stack[0] = 1u;
            sp = 1;
        }
        __syncwarp();
        while (sp > 0) {
            const uint32_t n = stack[--sp];
            __syncwarp();
            const int L = 31 - __clz((int)n);
            const int step = min(5, D - L);
            const uint32_t c = (n << step) + (uint32_t)lane;
            bool pass = lane < (1 << step);
            // the node's own box (the bound may have shrunk since it was pushed) and its descendants'
            // boxes are fetched together: one L2 round trip per expansion instead of two
            const f4 nlo = ldg4(tree.boxes + 2 * (size_t)n), nhi = ldg4(tree.boxes + 2 * (size_t)n + 1);
            f4 clo = nlo, chi = nhi;
            if (pass && step > 0) { clo = ldg4(tree.boxes + 2 * (size_t)c); chi = ldg4(tree.boxes + 2 * (size_t)c + 1); }
            if (box_dist2(q.x, q.y, q.z, nlo, nhi) > best.worst_d()) continue;
            if (pass && step > 0) pass = box_dist2(q.x, q.y, q.z, clo, chi) <= best.worst_d();
            unsigned mask = __ballot_sync(0xffffffffu, pass);
            if (L + step < D) {
                if (pass) stack[sp + __popc(mask & lanes_lt)] = c;
                sp += __popc(mask);
                __syncwarp();
                continue;
            }
            // children are leaves: up to 16 leaves per pass — lane = (leaf slot, point), each lane
            // keeps U = 4 independent loads in flight (this warp is alone on its query, so
            // instruction-level parallelism is the only latency hiding it has)
            constexpr int U = 4;
            const int slot = lane >> 3, pnt = lane & 7;
            while (mask) {
                const int nset = __popc(mask);
                uint32_t b[U], e[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int sidx = slot + 4 * u;
                    b[u] = e[u] = 0;
                    if (sidx < nset) {
                        const uint32_t leaf = ((n << step) + __fns(mask, 0, sidx + 1)) - (1u << D);
                        b[u] = seg_begin(D, leaf, tree.n);
                        e[u] = seg_begin(D, leaf + 1, tree.n);
                        if (pnt == 0) my_visits += e[u] - b[u];
                    }
                }
#pragma unroll
                for (uint32_t off = 0; off < PM_LEAF_MAX; off += 8) {  // one point per lane, slot and pass
                    float dd[U];
                    int pi[U];
                    bool cand[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint32_t p = b[u] + off + pnt;
                        dd[u] = 0.f; pi[u] = 0; cand[u] = false;
                        if (p < e[u]) {
                            const f4 pt = ldg4(tree.pts + p);
                            dd[u] = dist2(q.x, q.y, q.z, pt.x, pt.y, pt.z);
                            pi[u] = (int)__float_as_uint(pt.w);
                            cand[u] = best.accepts(dd[u], pi[u]);
                        }
                    }
                    // candidates are rare once the list is good: insert them one by one, uniformly
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        unsigned cm = __ballot_sync(0xffffffffu, cand[u]);
                        while (cm) {
                            const int src = __ffs(cm) - 1;
                            cm &= cm - 1;
                            const float cd = __shfl_sync(0xffffffffu, dd[u], src);
                            const int ci = __shfl_sync(0xffffffffu, pi[u], src);
                            if (best.accepts(cd, ci) && !best.contains(ci)) best.insert(cd, ci);
                        }
                    }
                }
                for (int i = 0; i < 4 * U && mask; ++i) mask &= mask - 1;
            }
        }
        if constexpr (NORMALS) {
            // the epilogue wants the whole list in one thread: gather it (every lane takes part in the shuffles), lane 0 finishes
            TopK<KMAX> all;
            all.k = k;
            static_for<0, KMAX>([&](auto J) { all.key[J] = __shfl_sync(0xffffffffu, best.key[J / 32], J % 32); });
            if (lane == 0 && normals_epilogue<KMAX>(all, k, ns, ns.by_position ? (size_t)(pos_offset + t) : (size_t)qi, q.x, q.y, q.z))
                atomicAdd(ns.degenerate, 1);
        } else {
#pragma unroll
            for (int r = 0; r < WarpTopK<KMAX>::R; ++r) {
                const int j = r * 32 + lane;
                if (j < k) {
                    const float bd = __uint_as_float((uint32_t)(best.key[r] >> 32));
                    const int bi = (int)(uint32_t)best.key[r];
                    const bool valid = bi != PM_NO_ID && bd != pm_inf();
                    ids[(size_t)qi * k + j] = valid ? bi : cap.miss_id;
                    dists[(size_t)qi * k + j] = valid ? bd : cap.miss_d;
                }
            }
        }
        __syncwarp();
    }
    if (visits && my_visits) atomicAdd(visits, my_visits);
}

// knn > 64: every query goes straight to the warp-per-query kernel, whose distributed list holds up to 32 candidates per
// register (WarpTopK): the queue is the identity and every result slot starts unfilled at the caller's radius
__global__ void all_to_stage2_kernel(int nq, int k, float max_r2, const float* __restrict__ var_r2, uint32_t* __restrict__ overflow, unsigned* count,
                                     int32_t* __restrict__ ids, float* __restrict__ dists) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e == 0) *count = (unsigned)nq;
    if (e < (size_t)nq) overflow[e] = (uint32_t)e;
    if (e < (size_t)nq * k) {
        ids[e] = -1;
        dists[e] = var_r2 ? var_r2[e / k] : max_r2;
    }
}

template <int KMAX>
int launch_wide(pmgpu_ctx* ctx, const TreeView& tree, const f4* queries, int nq, bool use_T, bool gated, int k, float max_r2, int32_t* ids, float* dists,
                const float* var_r2) {
    if (nq == 0) return PMGPU_OK;
    PM_CUDA_TRY(ctx, ctx->overflow.reserve((size_t)nq));
    unsigned* cnt = &ctx->state->overflow_count[ctx->knn_parity];
    unsigned* cnt_next = &ctx->state->overflow_count[ctx->knn_parity ^ 1];
    ctx->knn_parity ^= 1;
    const size_t total = (size_t)nq * k;
    all_to_stage2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(nq, k, max_r2, var_r2, ctx->overflow.p, cnt, ids, dists);
    NormalsSink none;
    memset(&none, 0, sizeof(none));
    const int grid2 = min(ctx->num_sms * 8, (nq + 3) / 4);
    knn_overflow_kernel<KMAX, false><<<grid2, 128, 0, ctx->stream>>>(tree, queries, ctx->state, use_T ? 1 : 0, gated ? 1 : 0, 0, k, max_r2, ctx->overflow.p, cnt,
                                                                 cnt_next, ids, dists, &ctx->state->visits, 0, var_r2, none, 0, nullptr);
    ctx->launches += 2;
    PM_CUDA_TRY(ctx, cudaGetLastError());
    return PMGPU_OK;
}

template <int KMAX, bool NORMALS>
int launch_one(pmgpu_ctx* ctx, const TreeView& tree, const f4* queries, int nq, bool use_T, bool gated, bool self_query, int k, float max_r2,
               bool use_seed, int32_t* ids, float* dists, bool use_cap, const float* var_r2, NormalsSink ns, int pos_offset) {
    const int grid = (nq + KNN_BLOCK - 1) / KNN_BLOCK;
    if (grid == 0) return PMGPU_OK;
    const size_t smem = (size_t)(tree.depth + 2) * KNN_BLOCK * sizeof(float);
    const bool seeded = KMAX == 1 && use_seed;
    const bool planes = NORMALS || !(seeded && ctx->seeded_without_planes);
    const int budget = seeded ? ctx->knn_budget : ctx->knn_budget_unseeded;
    PM_CUDA_TRY(ctx, ctx->overflow.reserve((size_t)nq));
    PM_CUDA_TRY(ctx, ctx->overflow_resume.reserve((size_t)nq));
    if (NORMALS) {
        // hand-over scratch for the queries stage 1 gives up on: one in eight at most, the rest finish where they are
        ns.scratch_cap = (unsigned)(nq / 8 + 1024);
        PM_CUDA_TRY(ctx, ctx->ids_tmp.reserve((size_t)ns.scratch_cap * k));
        PM_CUDA_TRY(ctx, ctx->dists_tmp.reserve((size_t)ns.scratch_cap * k));
        ns.scratch_ids = ctx->ids_tmp.p;
        ns.scratch_d = ctx->dists_tmp.p;
    }
    // the two stages of one launch share counter[parity]; stage 2 clears counter[parity ^ 1] for the next launch
    unsigned* cnt = &ctx->state->overflow_count[ctx->knn_parity];
    unsigned* cnt_next = &ctx->state->overflow_count[ctx->knn_parity ^ 1];
    ctx->knn_parity ^= 1;
    if (planes)
        PM_CUDA_TRY(ctx, launch_dependent(pdl_enabled(ctx) && gated, knn_kernel<KMAX, true, NORMALS>, dim3(grid), dim3(KNN_BLOCK), smem, ctx->stream, tree, queries, nq,
                                          ctx->state, use_T ? 1 : 0, gated ? 1 : 0, self_query ? 1 : 0, k, max_r2, ctx->ref_orig.p, use_seed ? 1 : 0, ids, dists,
                                          &ctx->state->visits, budget, ctx->overflow.p, cnt, use_cap ? 1 : 0, var_r2, ns, pos_offset, ctx->overflow_resume.p));
    else if constexpr (!NORMALS)
        PM_CUDA_TRY(ctx, launch_dependent(pdl_enabled(ctx) && gated, knn_kernel<KMAX, false, false>, dim3(grid), dim3(KNN_BLOCK), 0, ctx->stream, tree, queries, nq,
                                          ctx->state, use_T ? 1 : 0, gated ? 1 : 0, self_query ? 1 : 0, k, max_r2, ctx->ref_orig.p, use_seed ? 1 : 0, ids, dists,
                                          &ctx->state->visits, budget, ctx->overflow.p, cnt, use_cap ? 1 : 0, var_r2, ns, pos_offset, ctx->overflow_resume.p));
    if (ctx->time_stage2) { ctx->stage_end(); ctx->stage_begin(3); }
    const int grid2 = min(ctx->num_sms * 10, (nq + 3) / 4);  // 48 registers: ten 128-thread blocks are resident per SM
    PM_CUDA_TRY(ctx, launch_dependent(pdl_enabled(ctx) && !ctx->time_stage2, knn_overflow_kernel<KMAX, NORMALS>, dim3(grid2), dim3(128), 0, ctx->stream, tree, queries,
                                      ctx->state, use_T ? 1 : 0, gated ? 1 : 0, self_query ? 1 : 0, k, max_r2, ctx->overflow.p, cnt, cnt_next, ids, dists,
                                      &ctx->state->visits, use_cap ? 1 : 0, var_r2, ns, pos_offset, (const uint2*)(ctx->stage2_resume ? ctx->overflow_resume.p : nullptr)));
    ctx->launches += 2;
    PM_CUDA_TRY(ctx, cudaGetLastError());
    return PMGPU_OK;
}

}  // namespace

// use_T: apply state->T_iter to every query; gated: no-op once state->iterate == 0;
// self_query: `queries` is the leaf-ordered reference itself (results indexed by original column);
// use_seed: `ids` still holds the previous matches (same k) of the same reading — a candidate for k = 1, a radius for k > 1;
// use_cap: stop at min(max_r2, state->cap)
int launch_knn(pmgpu_ctx* ctx, const TreeView& tree, const f4* queries, int nq, bool use_T, bool gated, bool self_query, int k, float max_r2,
               bool use_seed, int32_t* ids, float* dists, bool use_cap, const float* var_r2) {
    NormalsSink none;
    memset(&none, 0, sizeof(none));
#define PM_KNN_CASE(K) return launch_one<K, false>(ctx, tree, queries, nq, use_T, gated, self_query, k, max_r2, use_seed, ids, dists, use_cap, var_r2, none, 0)
    if (k == 1) PM_KNN_CASE(1);
    if (k <= 4) PM_KNN_CASE(4);
    if (k <= 8) PM_KNN_CASE(8);
    if (k <= 10) PM_KNN_CASE(10);  // BASELINE config 4
    if (k <= 16) PM_KNN_CASE(16);
    if (k <= 20) PM_KNN_CASE(20);  // BASELINE config 3 (normals)
    if (k <= 32) PM_KNN_CASE(32);
    if (k <= 64) PM_KNN_CASE(64);
#undef PM_KNN_CASE
    // larger k (MatchersImpl.h:80 allows any unsigned): the warp-per-query kernel alone, list distributed over the lanes
    if (self_query || use_cap) {
        ctx->set_error("KDTreeMatcher on GPU: knn > 64 is supported for plain matching only");
        return PMGPU_ERR_UNSUPPORTED;
    }
    if (k <= 128) return launch_wide<128>(ctx, tree, queries, nq, use_T, gated, k, max_r2, ids, dists, var_r2);
    if (k <= 256) return launch_wide<256>(ctx, tree, queries, nq, use_T, gated, k, max_r2, ids, dists, var_r2);
    if (k <= 512) return launch_wide<512>(ctx, tree, queries, nq, use_T, gated, k, max_r2, ids, dists, var_r2);
    if (k <= 1024) return launch_wide<1024>(ctx, tree, queries, nq, use_T, gated, k, max_r2, ids, dists, var_r2);
    ctx->set_error("KDTreeMatcher on GPU: knn > 1024 is not supported");
    return PMGPU_ERR_UNSUPPORTED;
}

// K8: surface normals of the leaf-order positions [pos_lo, pos_hi) of the resident cloud as the epilogue of their self-kNN
// (normals.cuh); nothing but the sink's outputs is written
int launch_knn_normals(pmgpu_ctx* ctx, const TreeView& tree, int pos_lo, int pos_hi, int k, float max_r2, const NormalsSink& sink) {
    const f4* queries = tree.pts + pos_lo;
    const int nq = pos_hi - pos_lo;
#define PM_KNN_CASE(K) return launch_one<K, true>(ctx, tree, queries, nq, false, false, true, k, max_r2, false, nullptr, nullptr, false, nullptr, sink, pos_lo)
    if (k <= 4) PM_KNN_CASE(4);
    if (k <= 8) PM_KNN_CASE(8);
    if (k <= 10) PM_KNN_CASE(10);
    if (k <= 16) PM_KNN_CASE(16);
    if (k <= 20) PM_KNN_CASE(20);  // BASELINE config 3
    if (k <= 32) PM_KNN_CASE(32);
    if (k <= 64) PM_KNN_CASE(64);
#undef PM_KNN_CASE
    ctx->set_error("SurfaceNormalDataPointsFilter on GPU: knn > 64 is not supported");
    return PMGPU_ERR_UNSUPPORTED;
}

}  // namespace pm
