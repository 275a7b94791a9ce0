// knn.cu — K2: RigidTransformation::compute fused with KDTreeMatcher::findClosests
// (TransformationsImpl.cpp:49-87, MatchersImpl.cpp:85-101).
//
// One thread per query, queries scheduled in Morton order so the lanes of a warp walk almost the
// same root-to-leaf path (node and leaf loads become broadcasts served by L1/L2; the tree of a
// 1 M-point reference is 6 MB of boxes + 16 MB of points, resident in the 126 MB L2).  The k best
// candidates live in registers (core/tree.h TopK), the transform T_iter is read from the
// device-resident IcpState so no host round trip separates iterations.
#include "pmgpu_internal.cuh"

namespace pm {

namespace {

template <int KMAX>
__global__ void __launch_bounds__(128) knn_kernel(TreeView tree, const f4* __restrict__ queries, const uint32_t* __restrict__ order, int nq,
                                                  const IcpState* __restrict__ state, int use_T, int gated, int qi_from_w, int k,
                                                  float max_r2, int32_t* __restrict__ ids, float* __restrict__ dists, unsigned long long* visits) {
    __shared__ Mat4 sT;
    if (gated && state->iterate == 0) return;
    if (use_T) {
        if (threadIdx.x < 16) sT.m[threadIdx.x] = state->T_iter.m[threadIdx.x];
        __syncthreads();
    }
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t visited = 0;
    if (t < nq) {
        // order == queries' own w lane (self-query of the leaf-ordered reference, K8): the
        // original column index travels in w
        uint32_t qi;
        f4 q;
        if (qi_from_w) { q = queries[t]; qi = __float_as_uint(q.w); q.w = 1.f; }
        else { qi = order ? order[t] : (uint32_t)t; q = queries[qi]; }
        if (use_T) q = transform_point(sT, q);
        TopK<KMAX> best;
        best.init(k, max_r2);
        visited = knn_search<KMAX>(tree, q.x, q.y, q.z, best);
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
        unsigned v = visited;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(visits, (unsigned long long)v);
    }
}

template <int KMAX>
int launch_one(pmgpu_ctx* ctx, const TreeView& tree, const f4* queries, const uint32_t* order, int nq, bool use_T, bool gated, bool qi_from_w, int k,
               float max_r2, int32_t* ids, float* dists) {
    const int B = 128;
    const int grid = (nq + B - 1) / B;
    if (grid == 0) return PMGPU_OK;
    knn_kernel<KMAX><<<grid, B, 0, ctx->stream>>>(tree, queries, order, nq, ctx->state, use_T ? 1 : 0, gated ? 1 : 0, qi_from_w ? 1 : 0, k, max_r2, ids, dists,
                                                 &ctx->state->visits);
    ctx->launches += 1;
    PM_CUDA_TRY(ctx, cudaGetLastError());
    return PMGPU_OK;
}

}  // namespace

// use_T: apply state->T_iter to every query; gated: no-op once state->iterate == 0
int launch_knn(pmgpu_ctx* ctx, const TreeView& tree, const f4* queries, const uint32_t* order, int nq, bool use_T, bool gated, bool qi_from_w, int k,
               float max_r2, int32_t* ids, float* dists) {
    if (k == 1) return launch_one<1>(ctx, tree, queries, order, nq, use_T, gated, qi_from_w, k, max_r2, ids, dists);
    if (k <= 4) return launch_one<4>(ctx, tree, queries, order, nq, use_T, gated, qi_from_w, k, max_r2, ids, dists);
    if (k <= 8) return launch_one<8>(ctx, tree, queries, order, nq, use_T, gated, qi_from_w, k, max_r2, ids, dists);
    if (k <= 16) return launch_one<16>(ctx, tree, queries, order, nq, use_T, gated, qi_from_w, k, max_r2, ids, dists);
    if (k <= 32) return launch_one<32>(ctx, tree, queries, order, nq, use_T, gated, qi_from_w, k, max_r2, ids, dists);
    if (k <= 64) return launch_one<64>(ctx, tree, queries, order, nq, use_T, gated, qi_from_w, k, max_r2, ids, dists);
    ctx->set_error("KDTreeMatcher on GPU: knn > 64 is not supported");
    return PMGPU_ERR_UNSUPPORTED;
}

}  // namespace pm
