// knn.cu — K2: RigidTransformation::compute fused with KDTreeMatcher::findClosests
// (TransformationsImpl.cpp:49-87, MatchersImpl.cpp:85-101).
//
// One thread per query; the reading is resident in Morton order, so the 32 lanes of a warp hold
// neighbouring points and walk almost the same part of the tree (the tree of a 1 M-point
// reference is 1 MB of split planes + 8 MB of boxes + 16 MB of points, resident in the 126 MB L2;
// leaf and node loads of a warp mostly hit the same L1 lines).  The search of core/tree.h is run
// as three phases — plane descent | leaf scan | pop pending siblings — and the lanes of a warp
// RE-CONVERGE between phases (__syncwarp): without that, independent thread scheduling lets every
// lane drift into its own phase and the warp executes ~3 of 32 lanes per instruction (measured,
// profiles/).  T_iter is read from the device-resident IcpState, so no host round trip separates
// iterations; in iterations >= 2 every query starts at the leaf of its previous match (hint),
// which skips the root descent and makes the first leaf scan produce a tight bound.
#include "pmgpu_internal.cuh"

namespace pm {

namespace {

template <int KMAX>
__global__ void __launch_bounds__(128) knn_kernel(TreeView tree, const f4* __restrict__ queries, int nq, const IcpState* __restrict__ state,
                                                  int use_T, int gated, int self_query, int k, float max_r2, uint32_t* __restrict__ hints,
                                                  int use_hints, int32_t* __restrict__ ids, float* __restrict__ dists,
                                                  unsigned long long* visits) {
    __shared__ Mat4 sT;
    if (gated && state->iterate == 0) return;
    if (use_T) {
        if (threadIdx.x < 16) sT.m[threadIdx.x] = state->T_iter.m[threadIdx.x];
        __syncthreads();
    }
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool running = t < nq;
    uint32_t qi = (uint32_t)t;
    Lane s;
    TopK<KMAX> best;
    best.init(k, max_r2);
    s.visited = 0;
    s.best_leaf = 0;
    if (running) {
        f4 q = queries[t];
        uint32_t start = 0;
        if (self_query) {
            // K8: the query is the reference point at leaf-order position t; its original column
            // travels in w and its own leaf is the perfect start
            qi = __float_as_uint(q.w);
            q.w = 1.f;
            start = (1u << tree.depth) + seg_of((uint32_t)t, tree.depth, tree.n);
        } else if (use_hints) {
            start = hints[t];
        }
        if (use_T) q = transform_point(sT, q);
        lane_begin(s, tree, q.x, q.y, q.z, start);
    }
    // all 32 lanes stay in the loop until the slowest is done; phases re-converge the warp
    while (__any_sync(0xffffffffu, running)) {
        while (running && lane_descending(s, tree)) lane_descend_step(s, tree);
        __syncwarp();
        if (running) lane_scan_leaf<KMAX>(s, tree, best);
        __syncwarp();
        if (running) running = lane_pop<KMAX>(s, tree, best);
        __syncwarp();
    }
    if (t < nq) {
        int32_t* oi = ids + (size_t)qi * k;
        float* od = dists + (size_t)qi * k;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
            if (j < k) {
                const bool valid = best.id[j] != PM_NO_ID && best.d[j] != pm_inf();
                oi[j] = valid ? best.id[j] : -1;
                od[j] = valid ? best.d[j] : pm_inf();
            }
        }
        if (hints && !self_query) hints[t] = s.best_leaf;
    }
    if (visits) {
        // warp-aggregated statistics (Matcher::visitCounter)
        unsigned v = (t < nq) ? s.visited : 0u;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(visits, (unsigned long long)v);
    }
}

template <int KMAX>
int launch_one(pmgpu_ctx* ctx, const TreeView& tree, const f4* queries, int nq, bool use_T, bool gated, bool self_query, int k, float max_r2,
               uint32_t* hints, bool use_hints, int32_t* ids, float* dists) {
    const int B = 128;
    const int grid = (nq + B - 1) / B;
    if (grid == 0) return PMGPU_OK;
    knn_kernel<KMAX><<<grid, B, 0, ctx->stream>>>(tree, queries, nq, ctx->state, use_T ? 1 : 0, gated ? 1 : 0, self_query ? 1 : 0, k, max_r2, hints,
                                                 use_hints ? 1 : 0, ids, dists, &ctx->state->visits);
    ctx->launches += 1;
    PM_CUDA_TRY(ctx, cudaGetLastError());
    return PMGPU_OK;
}

}  // namespace

// use_T: apply state->T_iter to every query; gated: no-op once state->iterate == 0;
// self_query: `queries` is the leaf-ordered reference itself (results indexed by original column);
// hints: per-query start leaves (read when use_hints, always written unless self_query / null)
int launch_knn(pmgpu_ctx* ctx, const TreeView& tree, const f4* queries, int nq, bool use_T, bool gated, bool self_query, int k, float max_r2,
               uint32_t* hints, bool use_hints, int32_t* ids, float* dists) {
#define PM_KNN_CASE(K) return launch_one<K>(ctx, tree, queries, nq, use_T, gated, self_query, k, max_r2, hints, use_hints, ids, dists)
    if (k == 1) PM_KNN_CASE(1);
    if (k <= 4) PM_KNN_CASE(4);
    if (k <= 8) PM_KNN_CASE(8);
    if (k <= 16) PM_KNN_CASE(16);
    if (k <= 32) PM_KNN_CASE(32);
    if (k <= 64) PM_KNN_CASE(64);
#undef PM_KNN_CASE
    ctx->set_error("KDTreeMatcher on GPU: knn > 64 is not supported");
    return PMGPU_ERR_UNSUPPORTED;
}

}  // namespace pm
