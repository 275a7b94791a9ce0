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
// profiles/).  The per-level plane distances of the descent are cached in shared memory (one
// column per lane), so rejecting a pending sibling costs one LDS and a compare.  T_iter is read
// from the device-resident IcpState, so no host round trip separates iterations; in iterations
// >= 2 (k = 1) the previous match, re-measured, seeds the search with a tight bound.
#include "pmgpu_internal.cuh"

namespace pm {

namespace {

constexpr int KNN_BLOCK = 128;

template <int KMAX>
__global__ void __launch_bounds__(KNN_BLOCK) knn_kernel(TreeView tree, const f4* __restrict__ queries, int nq, const IcpState* __restrict__ state,
                                                        int use_T, int gated, int self_query, int k, float max_r2,
                                                        const f4* __restrict__ ref_orig, int use_seed, int32_t* __restrict__ ids,
                                                        float* __restrict__ dists, unsigned long long* visits) {
    extern __shared__ float s_plane[];  // [depth + 1][KNN_BLOCK]: cached plane distances, one column per lane
    __shared__ Mat4 sT;
    if (gated && state->iterate == 0) return;
    if (use_T) {
        if (threadIdx.x < 16) sT.m[threadIdx.x] = state->T_iter.m[threadIdx.x];
        __syncthreads();
    }
    float* plane = s_plane + threadIdx.x;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool running = t < nq;
    uint32_t qi = (uint32_t)t;
    Lane s;
    TopK<KMAX> best;
    best.init(k, max_r2);
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
                if (cand_less(dd, prev, best.worst_d(), best.worst_id())) best.insert(dd, prev);
            }
        }
    }
    // all 32 lanes stay in the loop until the slowest is done; phases re-converge the warp
    while (__any_sync(0xffffffffu, running)) {
        while (running && lane_descending(s, tree)) lane_descend_step(s, tree, plane, KNN_BLOCK);
        __syncwarp();
        if (running) lane_scan_leaf<KMAX>(s, tree, best);
        __syncwarp();
        if (running) running = lane_pop<KMAX>(s, tree, best, plane, KNN_BLOCK);
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
               bool use_seed, int32_t* ids, float* dists) {
    const int grid = (nq + KNN_BLOCK - 1) / KNN_BLOCK;
    if (grid == 0) return PMGPU_OK;
    const size_t smem = (size_t)(tree.depth + 2) * KNN_BLOCK * sizeof(float);
    knn_kernel<KMAX><<<grid, KNN_BLOCK, smem, ctx->stream>>>(tree, queries, nq, ctx->state, use_T ? 1 : 0, gated ? 1 : 0, self_query ? 1 : 0, k, max_r2,
                                                            ctx->ref_orig.p, use_seed ? 1 : 0, ids, dists, &ctx->state->visits);
    ctx->launches += 1;
    PM_CUDA_TRY(ctx, cudaGetLastError());
    return PMGPU_OK;
}

}  // namespace

// use_T: apply state->T_iter to every query; gated: no-op once state->iterate == 0;
// self_query: `queries` is the leaf-ordered reference itself (results indexed by original column);
// use_seed (k = 1): `ids` still holds the previous matches of the same reading
int launch_knn(pmgpu_ctx* ctx, const TreeView& tree, const f4* queries, int nq, bool use_T, bool gated, bool self_query, int k, float max_r2,
               bool use_seed, int32_t* ids, float* dists) {
#define PM_KNN_CASE(K) return launch_one<K>(ctx, tree, queries, nq, use_T, gated, self_query, k, max_r2, use_seed, ids, dists)
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
