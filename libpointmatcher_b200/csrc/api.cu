// api.cu — the extern "C" entry points of include/pmgpu.h: argument checks, uploads/downloads,
// kernel sequencing.  All work of a context is issued on its own stream.
#include <stdlib.h>
#include <string.h>

#include <new>

#include "core/linalg.h"
#include "normals.cuh"
#include "pmgpu_internal.cuh"

using namespace pm;

thread_local cudaStream_t pm::g_alloc_stream = nullptr;

#include <atomic>
static std::atomic<int> g_live_contexts[64];  // per device
static std::atomic<int>& live_on(int device) { return g_live_contexts[device & 63]; }
bool pm::alone_on_device(const pmgpu_ctx* ctx) { return live_on(ctx->device).load(std::memory_order_relaxed) == 1; }
bool pm::pdl_enabled(const pmgpu_ctx* ctx) { return ctx->pdl && alone_on_device(ctx); }

namespace {

__global__ void pack_normals_kernel(const float* __restrict__ src, int ld, int n, f4* __restrict__ dst, int comps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* s = src + (size_t)i * ld;
    dst[i] = make_float4(s[0], s[1], comps == 3 ? s[2] : 0.f, 0.f);
}

// reading normals arrive in the caller's column order and are kept in the reading's Morton order
__global__ void pack_normals_permuted_kernel(const float* __restrict__ src, int ld, const uint32_t* __restrict__ order, int n, f4* __restrict__ dst,
                                             int comps) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float* s = src + (size_t)order[t] * ld;
    dst[t] = make_float4(s[0], s[1], comps == 3 ? s[2] : 0.f, 0.f);
}

// KDTreeVarDistMatcher: per-point search distance -> squared (maxRadius * maxRadius in float, like libnabo), Morton order
__global__ void pack_max_r2_kernel(const float* __restrict__ src, int ld, const uint32_t* __restrict__ order, int n, float* __restrict__ dst) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float r = src[(size_t)order[t] * ld];
    dst[t] = fmul(r, r);
}

// `R * inputDesc` for the "normals" descriptor (TransformationsImpl.cpp:71-84)
__global__ void rotate_normals_inplace_kernel(f4* __restrict__ nrm, int n, Mat4 T) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const f4 v = nrm[i];
    f4 r;
    r.x = fadd(fadd(fmul(T.m[0], v.x), fmul(T.m[4], v.y)), fmul(T.m[8], v.z));
    r.y = fadd(fadd(fmul(T.m[1], v.x), fmul(T.m[5], v.y)), fmul(T.m[9], v.z));
    r.z = fadd(fadd(fmul(T.m[2], v.x), fmul(T.m[6], v.y)), fmul(T.m[10], v.z));
    r.w = 0.f;
    nrm[i] = r;
}

// comps = 3 (3-D) or 2 (2-D clouds)
__global__ void unpack_f4_kernel(const f4* __restrict__ src, int n, float* __restrict__ dst, int comps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const f4 v = src[i];
    dst[comps * (size_t)i] = v.x; dst[comps * (size_t)i + 1] = v.y;
    if (comps == 3) dst[3 * (size_t)i + 2] = v.z;
}

// a 2-D cloud (3 x n: x, y, w) as resident points (x, y, 0, w): every kernel then computes what the reference computes on
// the 3-row matrices, bit for bit — the z terms are exact zeros (x + 0 and 0 * t round to themselves)
__global__ void expand_2d_kernel(const float* __restrict__ src, int n, f4* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[i] = make_float4(src[3 * (size_t)i], src[3 * (size_t)i + 1], 0.f, src[3 * (size_t)i + 2]);
}
__global__ void collapse_2d_kernel(const f4* __restrict__ src, int n, float* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const f4 v = src[i];
    dst[3 * (size_t)i] = v.x; dst[3 * (size_t)i + 1] = v.y; dst[3 * (size_t)i + 2] = v.w;
}

// Moves a structure built on the caller's coordinates into the frame centred on `mean`: every
// stored coordinate c becomes fsub(c, mean).  Rounding is monotone, so the order along every axis,
// the median splits and the tight boxes of the shifted points are exactly the shifted ones.
__global__ void center_structure_kernel(f4* __restrict__ ref_orig, f4* __restrict__ ref_sorted, int n, f2* __restrict__ splits, int nsplits,
                                        f4* __restrict__ boxes, int nboxes, float mx, float my, float mz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        f4 p = ref_orig[i];
        p.x = fsub(p.x, mx); p.y = fsub(p.y, my); p.z = fsub(p.z, mz);
        ref_orig[i] = p;
        p = ref_sorted[i];
        p.x = fsub(p.x, mx); p.y = fsub(p.y, my); p.z = fsub(p.z, mz);
        ref_sorted[i] = p;
    }
    if (i >= 1 && i < nsplits) {
        f2 sp = splits[i];
        const uint32_t dim = f2u(sp.y);
        sp.x = fsub(sp.x, dim == 0 ? mx : (dim == 1 ? my : mz));
        splits[i] = sp;
    }
    if (i < nboxes) {
        f4 b = boxes[i];
        b.x = fsub(b.x, mx); b.y = fsub(b.y, my); b.z = fsub(b.z, mz);
        boxes[i] = b;
    }
}

// sharded K8: normals arrive in leaf order (one contiguous slice per rank); position t holds the point whose original
// column travels in ref_sorted[t].w
__global__ void scatter_by_leaf_order_kernel(const f4* __restrict__ src, const f4* __restrict__ ref_sorted, int n, f4* __restrict__ dst) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dst[f2u(ref_sorted[t].w)] = src[t];
}

__global__ void transform_inplace_kernel(f4* __restrict__ pts, int n, Mat4 T) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pts[i] = transform_point(T, pts[i]);
}

// sorted position t -> caller's column order[t]; k values per column
template <typename V>
__global__ void unpermute_kernel(const V* __restrict__ src, const uint32_t* __restrict__ order, size_t n, int k, V* __restrict__ dst) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * (size_t)k) return;
    const size_t t = e / k;
    const int j = (int)(e - t * k);
    dst[(size_t)order[t] * k + j] = src[e];
}

// smoothNormals (SurfaceNormal.cpp:259-283): in place and in point order, like the reference — normal i becomes the mean of
// its valid neighbours' CURRENT normals, each flipped to the side of normal i.  `n4`: (nx, ny, nz, 0) per point.
void smooth_normals_host(f4* n4, const int32_t* ids, int knn, int n) {
    for (int i = 0; i < n; ++i) {
        const f4 cur = n4[i];
        float mx = 0.f, my = 0.f, mz = 0.f;
        int cnt = 0;
        for (int j = 0; j < knn; ++j) {
            const int r = ids[(size_t)i * knn + j];
            if (r < 0) continue;
            const f4 nb = n4[r];
            const float dot = cur.x * nb.x + cur.y * nb.y + cur.z * nb.z;
            if (dot > 0.f) { mx += nb.x; my += nb.y; mz += nb.z; }
            else { mx -= nb.x; my -= nb.y; mz -= nb.z; }
            ++cnt;
        }
        n4[i] = make_float4(mx / (float)cnt, my / (float)cnt, mz / (float)cnt, 0.f);
    }
}

// transforms cross the ABI as dimh x dimh column-major matrices (PointMatcher.h TransformationParameters): 4x4, or 3x3 for
// 2-D clouds, which live on the device embedded in a 4x4 with an identity z row / column
void T_load(int dimh, const float* T, Mat4& M) {
    if (dimh == 4) { memcpy(M.m, T, sizeof(M.m)); return; }
    mat4_identity(M);
    M.m[0] = T[0]; M.m[1] = T[1];
    M.m[4] = T[3]; M.m[5] = T[4];
    M.m[12] = T[6]; M.m[13] = T[7];
}
void T_store(int dimh, const Mat4& M, float* T) {
    if (dimh == 4) { memcpy(T, M.m, sizeof(M.m)); return; }
    T[0] = M.m[0]; T[1] = M.m[1]; T[2] = 0.f;
    T[3] = M.m[4]; T[4] = M.m[5]; T[5] = 0.f;
    T[6] = M.m[12]; T[7] = M.m[13]; T[8] = 1.f;
}

// clouds: `rows` x n column-major, rows = 4 or 3; 2-D clouds are expanded to (x, y, 0, w) on the device
int upload_cloud(pmgpu_ctx* ctx, const float* features, int rows, int n, f4* dst) {
    if (n <= 0) return PMGPU_OK;
    if (rows == 4) {
        PM_CUDA_TRY(ctx, cudaMemcpyAsync(dst, features, (size_t)n * sizeof(f4), cudaMemcpyDefault, ctx->stream));
        return PMGPU_OK;
    }
    ScopedBuf<float> staging;
    PM_CUDA_TRY(ctx, staging.reserve(3 * (size_t)n));
    PM_CUDA_TRY(ctx, cudaMemcpyAsync(staging.p, features, 3 * (size_t)n * sizeof(float), cudaMemcpyDefault, ctx->stream));
    expand_2d_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(staging.p, n, dst);
    ctx->launches += 1;
    PM_CUDA_TRY(ctx, cudaGetLastError());
    return PMGPU_OK;
}

int fail(pmgpu_ctx* ctx, int code, const char* msg) {
    ctx->set_error(msg);
    return code;
}

// keep_window: the entry point neither touches the reading nor the ordering scratch (see pmgpu_ctx::overlap_window)
int use_device(pmgpu_ctx* ctx, bool keep_window = false) {
    PM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    if (!keep_window) ctx->overlap_window = false;
    return PMGPU_OK;
}

int push_state(pmgpu_ctx* ctx) {
    // queue lengths and tickets are device-owned scratch: never push back a stale pulled copy
    IcpState* h = ctx->state_host;
    h->overflow_count[0] = h->overflow_count[1] = 0;
    h->ticket[0] = h->ticket[1] = h->ticket[2] = h->ticket[3] = 0;
    h->bar_count = 0;
    for (int f = 0; f < PM_MAX_FILTERS; ++f) h->sel_cand_count[f] = 0;
    PM_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->state, ctx->state_host, sizeof(IcpState), cudaMemcpyHostToDevice, ctx->stream));
    return PMGPU_OK;
}

int pull_state(pmgpu_ctx* ctx) {
    PM_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->state_host, ctx->state, sizeof(IcpState), cudaMemcpyDeviceToHost, ctx->stream));
    PM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return PMGPU_OK;
}

const char* status_message(int s) {
    switch (s) {
        case PMGPU_OK: return "ok";
        case PMGPU_ERR_CUDA: return "CUDA error";
        case PMGPU_ERR_BAD_ARG: return "bad argument";
        case PMGPU_ERR_UNSUPPORTED: return "GPU module: only float clouds of 2 or 3 dimensions are supported";
        case PMGPU_ERR_NO_REFERENCE: return "matcher not initialised: no reference";
        case PMGPU_ERR_NO_READING: return "no reading";
        case PMGPU_ERR_NO_MATCHES: return "no matches: findClosests must run first";
        case PMGPU_ERR_NO_OUTLIER_TO_FILTER: return "no outlier to filter";
        case PMGPU_ERR_BAD_QUANTILE: return "quantile must be between 0 and 1";
        case PMGPU_ERR_NO_POINT_TO_MINIMIZE: return "ErrorMnimizer: no point to minimize";
        case PMGPU_ERR_NO_NORMALS: return "Field normals not found";
        case PMGPU_ERR_NOT_ORTHOGONAL: return "RigidTransformation: Error, rotation matrix is not orthogonal.";
        case PMGPU_ERR_KNN_TOO_LARGE: return "knn is larger than the number of reference points";
        case PMGPU_ERR_NAN: return "abs rotation norm not a number";
        case PMGPU_ERR_COMM: return "NCCL error";
    }
    return "unknown status";
}

// status raised by a kernel -> API return value
int device_status(pmgpu_ctx* ctx) {
    const int s = ctx->state_host->status;
    if (s != PMGPU_OK) ctx->set_error(status_message(s));
    return s;
}

int upload_normals(pmgpu_ctx* ctx, const float* normals, int ld) {
    const int comps = ctx->dimh - 1;
    if (ld < comps) return fail(ctx, PMGPU_ERR_BAD_ARG, "normals_ld must be >= the cloud's dimension");
    const int n = ctx->nr;
    PM_CUDA_TRY(ctx, ctx->ref_normals.reserve(n));
    ScopedBuf<float> staging;
    PM_CUDA_TRY(ctx, staging.reserve((size_t)n * ld));
    // the last column may be shorter than ld in the caller's matrix: copy (n-1)*ld + 3 floats
    PM_CUDA_TRY(ctx, cudaMemcpyAsync(staging.p, normals, ((size_t)(n - 1) * ld + comps) * sizeof(float), cudaMemcpyDefault, ctx->stream));
    pack_normals_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(staging.p, ld, n, ctx->ref_normals.p, comps);
    ctx->launches += 1;
    PM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->has_normals = true;
    return PMGPU_OK;
}

// download a k x nq device array that lives in Morton order into the caller's column order
template <typename V>
int download_unpermuted(pmgpu_ctx* ctx, const V* src, DevBuf<V>& staging, int k, V* dst) {
    const size_t total = (size_t)k * ctx->nq;
    if (!dst || total == 0) return PMGPU_OK;
    PM_CUDA_TRY(ctx, staging.reserve(total));
    unpermute_kernel<V><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(src, ctx->q_order.p, (size_t)ctx->nq, k, staging.p);
    ctx->launches += 1;
    PM_CUDA_TRY(ctx, cudaMemcpyAsync(dst, staging.p, total * sizeof(V), cudaMemcpyDefault, ctx->stream));
    return PMGPU_OK;
}

int check_params(pmgpu_ctx* ctx, const pmgpu_icp_params* p) {
    if (!p) return fail(ctx, PMGPU_ERR_BAD_ARG, "null parameters");
    if (p->knn < 1) return fail(ctx, PMGPU_ERR_BAD_ARG, "knn must be >= 1");
    if (p->knn > ctx->nr) return fail(ctx, PMGPU_ERR_KNN_TOO_LARGE, status_message(PMGPU_ERR_KNN_TOO_LARGE));
    if (p->nfilters < 0 || p->nfilters > PM_MAX_FILTERS) return fail(ctx, PMGPU_ERR_BAD_ARG, "at most 8 outlier filters");
    if (p->minimizer < 0 || (p->minimizer & 0xff) > PMGPU_MIN_P2POINT_SIM || (p->minimizer & ~0x3ff))
        return fail(ctx, PMGPU_ERR_BAD_ARG, "unknown error minimizer");
    if ((p->minimizer & PMGPU_MIN_FORCE4DOF) && !((p->minimizer & 0xff) == PMGPU_MIN_P2PLANE || (p->minimizer & 0xff) == PMGPU_MIN_P2PLANE_COV))
        return fail(ctx, PMGPU_ERR_BAD_ARG, "force4DOF is a point-to-plane parameter");
    if ((p->minimizer & PMGPU_MIN_FORCE2D) && p->minimizer != (PMGPU_MIN_P2PLANE | PMGPU_MIN_FORCE2D))
        return fail(ctx, PMGPU_ERR_BAD_ARG, "force2D goes with PointToPlaneErrorMinimizer alone (no force4DOF, no covariance)");
    if (p->use_differential && (p->smooth_length < 0 || p->smooth_length >= PM_MAX_HISTORY))
        return fail(ctx, PMGPU_ERR_UNSUPPORTED, "DifferentialTransformationChecker: smoothLength must be < 64 on the GPU path");
    if (p->max_iterations < 0) return fail(ctx, PMGPU_ERR_BAD_ARG, "maxIterationCount must be >= 0");
    if (ctx->dimh == 3 && (p->minimizer & 0xff) != PMGPU_MIN_P2POINT && p->minimizer != PMGPU_MIN_P2PLANE)
        return fail(ctx, PMGPU_ERR_UNSUPPORTED, "2-D clouds: PointToPoint and PointToPlane error minimizers only (no covariance, similarity, force2D / force4DOF)");
    return PMGPU_OK;
}

int enqueue_iteration(pmgpu_ctx* ctx, const pmgpu_icp_params* p, bool gated) {
    const bool var_dist = p->max_dist < 0.f;  // KDTreeVarDistMatcher
    const float max_r2 = var_dist ? pm_inf() : p->max_dist * p->max_dist;
    ctx->stage_begin(0);
    SelectSpec spec;
    PM_TRY(make_select_spec(ctx, p->nfilters, p->filter_type, p->filter_param, &spec));
    // capped matching: only where nothing observable depends on the matches the filters reject —
    // the fused loop, an unbounded maxDist (a finite one decides which matches count as missing),
    // and a non-empty filter chain
    // (a RobustOutlierFilter weighs every match, however far: nothing may be cut)
    const bool use_cap = gated && ctx->cap_enabled && p->nfilters > 0 && max_r2 == pm_inf() && spec.robust_index() < 0 && spec.var_index() < 0 && !var_dist;
    PM_TRY(launch_knn(ctx, ctx->tree_view(), ctx->reading.p, ctx->nq, true, gated, false, p->knn, max_r2,
                      ctx->seed_k == p->knn && ctx->seed_enabled, ctx->ids.p, ctx->dists.p, use_cap,
                      var_dist ? ctx->reading_max_r2.p : nullptr));
    ctx->seed_k = p->knn;
    ctx->stage_end();
    if (gated && fused_select_applies(ctx, spec)) {
        // outlier-filter select + minimiser + compose / checkers as one kernel (timing slot 2)
        ctx->stage_begin(2);
        PM_TRY(launch_select_minimize(ctx, spec, p->minimizer, p, use_cap));
        ctx->stage_end();
        return PMGPU_OK;
    }
    ctx->stage_begin(1);
    PM_TRY(launch_weights(ctx, spec, gated, use_cap));
    ctx->stage_end();
    ctx->stage_begin(2);
    PM_TRY(launch_minimize(ctx, p->minimizer, true, gated, p));
    ctx->stage_end();
    return PMGPU_OK;
}

}  // namespace

extern "C" {

const char* pmgpu_status_string(int status) { return status_message(status); }

int pmgpu_ctx_create(int device, pmgpu_ctx** ctx_out) {
    if (!ctx_out) return PMGPU_ERR_BAD_ARG;
    *ctx_out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return PMGPU_ERR_CUDA;
    pmgpu_ctx* ctx = new (std::nothrow) pmgpu_ctx();
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    ctx->device = device;
    bool ok = cudaSetDevice(device) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_build, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&ctx->ev_reading, cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&ctx->state, sizeof(IcpState)) == cudaSuccess;
    ok = ok && cudaMallocHost((void**)&ctx->state_host, sizeof(IcpState)) == cudaSuccess;
    int sms = 0;
    ok = ok && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess;
    if (!ok) {
        pmgpu_ctx_destroy(ctx);
        return PMGPU_ERR_CUDA;
    }
    ctx->num_sms = sms > 0 ? sms : 148;
    {
        // keep freed buffers in the pool (see DevBuf)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t threshold = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
        }
        cudaGetLastError();
    }
    ctx->seed_enabled = getenv("PMGPU_NO_SEED") == nullptr;
    ctx->time_stage2 = getenv("PMGPU_TIME_STAGE2") != nullptr;
    ctx->cap_enabled = getenv("PMGPU_NO_CAP") == nullptr;
    ctx->fused_select = getenv("PMGPU_NO_FUSED_SELECT") == nullptr;
    ctx->overlap_enabled = getenv("PMGPU_NO_OVERLAP") == nullptr;
    if (const char* c = getenv("PMGPU_COOP")) ctx->fused_cooperative = atoi(c) != 0;
    ctx->pdl = getenv("PMGPU_NO_PDL") == nullptr;
    ctx->stage2_resume = getenv("PMGPU_NO_RESUME") == nullptr;
    ctx->build_select = getenv("PMGPU_BUILD_SORT") == nullptr;
    if (const char* c = getenv("PMGPU_DEFER_FINALIZE")) ctx->defer_finalize = atoi(c) != 0;
    ctx->seeded_without_planes = getenv("PMGPU_SEED_PLANES") == nullptr;
    if (const char* m = getenv("PMGPU_CAP_MARGIN")) ctx->cap_margin = (float)atof(m);
    // 1: (almost) everything through stage 2
    if (const char* b = getenv("PMGPU_KNN_BUDGET")) ctx->knn_budget = ctx->knn_budget_unseeded = atoi(b) > 0 ? atoi(b) : 1;
    if (const char* b = getenv("PMGPU_KNN_BUDGET_UNSEEDED")) ctx->knn_budget_unseeded = atoi(b) > 0 ? atoi(b) : 1;
    memset(ctx->state_host, 0, sizeof(IcpState));
    ctx->state_host->robust_iteration = 1;
    mat4_identity(ctx->state_host->T_iter);
    mat4_identity(ctx->state_host->T_match);
    mat4_identity(ctx->state_host->dT);
    ctx->state_host->iterate = 1;
    if (cudaMemcpy(ctx->state, ctx->state_host, sizeof(IcpState), cudaMemcpyHostToDevice) != cudaSuccess) {
        pmgpu_ctx_destroy(ctx);
        return PMGPU_ERR_CUDA;
    }
    live_on(device).fetch_add(1);
    ctx->counted = true;
    *ctx_out = ctx;
    return PMGPU_OK;
}

void pmgpu_ctx_destroy(pmgpu_ctx* ctx) {
    if (!ctx) return;
    if (ctx->counted) live_on(ctx->device).fetch_sub(1);
    cudaSetDevice(ctx->device);
    g_alloc_stream = ctx->stream;
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    pmgpu_comm_destroy(ctx);
    ctx->ref_orig.release(); ctx->ref_sorted.release(); ctx->ref_normals.release(); ctx->splits.release(); ctx->boxes.release(); ctx->gather_tmp.release();
    ctx->reading_tmp.release(); ctx->ids_tmp.release(); ctx->dists_tmp.release(); ctx->overflow.release();
    ctx->keys_a.release(); ctx->keys_b.release(); ctx->perm_a.release(); ctx->perm_b.release();
    ctx->node_box.release(); ctx->cub_tmp.release(); ctx->seg_state.release(); ctx->seg_hist.release(); ctx->seg_cnt.release();
    ctx->reading.release(); ctx->q_order.release();
    ctx->ids.release(); ctx->dists.release(); ctx->weights.release();
    ctx->hist.release(); ctx->sel_cand.release(); ctx->partials.release(); ctx->overflow_resume.release();
    ctx->reading_normals.release(); ctx->reading_max_r2.release(); ctx->var_sorted.release(); ctx->var_cum.release();
    for (auto& iv : ctx->intervals) { cudaEventDestroy(iv.a); cudaEventDestroy(iv.b); }
    for (auto e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->state) cudaFree(ctx->state);
    if (ctx->state_host) cudaFreeHost(ctx->state_host);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);  // the cudaFreeAsync calls above
    if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
    if (ctx->ev_build) cudaEventDestroy(ctx->ev_build);
    if (ctx->ev_reading) cudaEventDestroy(ctx->ev_reading);
    if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    g_alloc_stream = nullptr;
    delete ctx;
}

const char* pmgpu_last_error(const pmgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void* pmgpu_ctx_stream(pmgpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t pmgpu_launch_count(const pmgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }

int pmgpu_timing_enable(pmgpu_ctx* ctx, int on) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    ctx->timing = on != 0;
    return PMGPU_OK;
}

int pmgpu_timing_collect(pmgpu_ctx* ctx, double* ms_out, int* count_out) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    PM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int s = 0; s < 4; ++s) {
        if (ms_out) ms_out[s] = 0.0;
        if (count_out) count_out[s] = 0;
    }
    for (auto& iv : ctx->intervals) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, iv.a, iv.b) == cudaSuccess && iv.stage >= 0 && iv.stage < 4) {
            if (ms_out) ms_out[iv.stage] += ms;
            if (count_out) count_out[iv.stage] += 1;
        }
        ctx->event_pool.push_back(iv.a);
        ctx->event_pool.push_back(iv.b);
    }
    ctx->intervals.clear();
    return PMGPU_OK;
}

int pmgpu_sync(pmgpu_ctx* ctx) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    PM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return PMGPU_OK;
}

static int ref_set_impl(pmgpu_ctx* ctx, const float* features, int rows, int n, const float* normals, int normals_ld, float* mean_out);

int pmgpu_ref_set(pmgpu_ctx* ctx, const float* features, int rows, int n, const float* normals, int normals_ld) {
    return ref_set_impl(ctx, features, rows, n, normals, normals_ld, nullptr);
}

int pmgpu_ref_set_centered(pmgpu_ctx* ctx, const float* features, int rows, int n, const float* normals, int normals_ld, float* mean_out) {
    if (!mean_out) return PMGPU_ERR_BAD_ARG;
    return ref_set_impl(ctx, features, rows, n, normals, normals_ld, mean_out);
}

// Shifts the resident structure (built on the caller's coordinates) into the frame centred on the cloud's mean.
// `reference.features.rowwise().sum() / nbPtsReference` in float, column after column (ICP.cpp:292): three
// independent serial chains of float adds on the host (no reassociation without -ffast-math), which overlap
// whatever the stream is still doing (upload, build, normals).
static int center_resident_reference(pmgpu_ctx* ctx, const float* features, int n, float* mean_out) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, features) == cudaSuccess && attr.type == cudaMemoryTypeDevice) {
        ctx->nr = 0;
        return fail(ctx, PMGPU_ERR_BAD_ARG, "centring the reference needs a host pointer");
    }
    cudaGetLastError();
    float sx = 0.f, sy = 0.f, sz = 0.f;
    const int rows = ctx->dimh;
    for (int i = 0; i < n; ++i) {
        sx += features[rows * (size_t)i];
        sy += features[rows * (size_t)i + 1];
        if (rows == 4) sz += features[4 * (size_t)i + 2];
    }
    mean_out[0] = sx / (float)n; mean_out[1] = sy / (float)n;
    if (rows == 4) { mean_out[2] = sz / (float)n; mean_out[3] = 1.f; }
    else mean_out[2] = 1.f;
    const float mean_z = rows == 4 ? mean_out[2] : 0.f;
    const int nsplits = 1 << ctx->depth, nboxes = 4 << ctx->depth;
    const int m = n > nboxes ? n : nboxes;
    center_structure_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(ctx->ref_orig.p, ctx->ref_sorted.p, n, ctx->splits.p, nsplits, ctx->boxes.p, nboxes,
                                                                     mean_out[0], mean_out[1], mean_z);
    ctx->launches += 1;
    PM_CUDA_TRY(ctx, cudaGetLastError());
    return PMGPU_OK;
}

int pmgpu_ref_center(pmgpu_ctx* ctx, const float* features, int rows, int n, float* mean_out) {
    if (!ctx || !features || !mean_out) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx, true));
    if (ctx->nr == 0) return fail(ctx, PMGPU_ERR_NO_REFERENCE, status_message(PMGPU_ERR_NO_REFERENCE));
    if (rows != ctx->dimh || n != ctx->nr) return fail(ctx, PMGPU_ERR_BAD_ARG, "pmgpu_ref_center: `features` must be the cloud given to pmgpu_ref_set");
    ctx->have_matches = false;
    ctx->seed_k = 0;
    return center_resident_reference(ctx, features, n, mean_out);
}

static int ref_set_impl(pmgpu_ctx* ctx, const float* features, int rows, int n, const float* normals, int normals_ld, float* mean_out) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    if (!features) return fail(ctx, PMGPU_ERR_BAD_ARG, "null reference features");
    if (rows != 4 && rows != 3) return fail(ctx, PMGPU_ERR_UNSUPPORTED, status_message(PMGPU_ERR_UNSUPPORTED));
    if (n < 1) return fail(ctx, PMGPU_ERR_BAD_ARG, "the reference cloud is empty");
    ctx->nr = 0;
    ctx->has_normals = false;
    ctx->have_matches = false;
    ctx->seed_k = 0;
    ctx->dimh = rows;
    PM_CUDA_TRY(ctx, ctx->ref_orig.reserve(n));
    PM_TRY(upload_cloud(ctx, features, rows, n, ctx->ref_orig.p));
    ctx->nr = n;
    // the structure is built on the uploaded coordinates while, for the centred variant, the host
    // is busy with the mean; it is shifted into the centred frame afterwards
    const int s = build_tree(ctx);
    if (s != PMGPU_OK) { ctx->nr = 0; return s; }
    PM_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_build, ctx->stream));
    if (mean_out) PM_TRY(center_resident_reference(ctx, features, n, mean_out));
    if (normals) PM_TRY(upload_normals(ctx, normals, normals_ld));
    ctx->overlap_window = ctx->overlap_enabled;  // what follows on the stream (normals, centring) leaves the reading's buffers alone
    return PMGPU_OK;
}

int pmgpu_ref_set_normals(pmgpu_ctx* ctx, const float* normals, int normals_ld) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx, true));
    if (ctx->nr == 0) return fail(ctx, PMGPU_ERR_NO_REFERENCE, status_message(PMGPU_ERR_NO_REFERENCE));
    if (!normals) { ctx->has_normals = false; return PMGPU_OK; }
    return upload_normals(ctx, normals, normals_ld);
}

int pmgpu_ref_get_normals(pmgpu_ctx* ctx, float* normals_out) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    if (!normals_out) return fail(ctx, PMGPU_ERR_BAD_ARG, "null output");
    if (ctx->nr == 0) return fail(ctx, PMGPU_ERR_NO_REFERENCE, status_message(PMGPU_ERR_NO_REFERENCE));
    if (!ctx->has_normals) return fail(ctx, PMGPU_ERR_NO_NORMALS, status_message(PMGPU_ERR_NO_NORMALS));
    std::vector<f4> host((size_t)ctx->nr);
    PM_CUDA_TRY(ctx, cudaMemcpyAsync(host.data(), ctx->ref_normals.p, host.size() * sizeof(f4), cudaMemcpyDeviceToHost, ctx->stream));
    PM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const int comps = ctx->dimh - 1;
    for (int i = 0; i < ctx->nr; ++i) {
        normals_out[comps * (size_t)i + 0] = host[i].x;
        normals_out[comps * (size_t)i + 1] = host[i].y;
        if (comps == 3) normals_out[3 * (size_t)i + 2] = host[i].z;
    }
    return PMGPU_OK;
}

// shard: rank / nranks / chunk of pmgpu_reading_set_sharded (nranks <= 1: the whole cloud); `n` is then the WHOLE cloud's size
static int reading_set_impl(pmgpu_ctx* ctx, const float* features, int rows, int n_total, int rank, int nranks, int chunk) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    int n = n_total;
    int my_chunks = 0, tail = 0;
    if (nranks > 1) {
        if (rows != 4 || chunk < 1 || rank < 0 || rank >= nranks || n_total < 0) {
            ctx->set_error("pmgpu_reading_set_sharded: 3-D clouds, chunk >= 1, 0 <= rank < nranks");
            return PMGPU_ERR_BAD_ARG;
        }
        const int nchunks = n_total / chunk;
        my_chunks = nchunks > rank ? (nchunks - rank + nranks - 1) / nranks : 0;
        tail = (nchunks % nranks == rank) ? n_total - nchunks * chunk : 0;
        n = my_chunks * chunk + tail;
    }
    const bool window = ctx->overlap_window;
    PM_TRY(use_device(ctx));
    if (!features) return fail(ctx, PMGPU_ERR_BAD_ARG, "null reading features");
    if (rows != 4 && rows != 3) return fail(ctx, PMGPU_ERR_UNSUPPORTED, status_message(PMGPU_ERR_UNSUPPORTED));
    if (ctx->nr > 0 && rows != ctx->dimh) return fail(ctx, PMGPU_ERR_BAD_ARG, "reading and reference must have the same dimension");
    if (n < 0) return fail(ctx, PMGPU_ERR_BAD_ARG, "negative point count");
    if (ctx->nr == 0) ctx->dimh = rows;
    ctx->nq = 0;
    ctx->have_matches = false;
    ctx->have_weights = false;
    ctx->has_reading_normals = false;
    ctx->has_reading_max_r2 = false;
    // Upload + ordering on the second stream, behind the structure build only, while the main stream is still computing
    // the reference's normals: possible when no buffer has to grow (allocations are ordered on the main stream) and the
    // cloud needs no staging (3-D).
    const size_t need = (size_t)(n > 0 ? n : 1);
    const bool overlap = window && rows == 4 && n > 0 && ctx->reading.cap >= need && ctx->reading_tmp.cap >= need && ctx->q_order.cap >= need &&
                         ctx->perm_a.cap >= need && ctx->perm_b.cap >= need && ctx->keys_a.cap >= need && ctx->node_box.cap >= 12 && ctx->cub_tmp.cap >= morton_scratch_bytes((uint32_t)n);
    PM_CUDA_TRY(ctx, ctx->reading.reserve(need));
    PM_CUDA_TRY(ctx, ctx->reading_tmp.reserve(need));
    ctx->seed_k = 0;
    cudaStream_t main_stream = ctx->stream;
    if (overlap) {
        PM_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_build, 0));
        ctx->stream = ctx->stream2;
    }
    int s = PMGPU_OK;
    if (nranks > 1) {
        // my chunks sit at a regular stride in the caller's matrix: one 2-D copy (rows of `chunk` points), then the tail
        const size_t row_bytes = (size_t)chunk * sizeof(f4);
        if (my_chunks > 0 && cudaMemcpy2DAsync(ctx->reading_tmp.p, row_bytes, features + (size_t)rank * chunk * 4, row_bytes * nranks, row_bytes, my_chunks,
                                               cudaMemcpyDefault, ctx->stream) != cudaSuccess)
            s = fail(ctx, PMGPU_ERR_CUDA, "cudaMemcpy2DAsync");
        if (s == PMGPU_OK && tail > 0 &&
            cudaMemcpyAsync(ctx->reading_tmp.p + (size_t)my_chunks * chunk, features + (size_t)(n_total - tail) * 4, (size_t)tail * sizeof(f4), cudaMemcpyDefault,
                            ctx->stream) != cudaSuccess)
            s = fail(ctx, PMGPU_ERR_CUDA, "cudaMemcpyAsync");
    } else {
        s = upload_cloud(ctx, features, rows, n, ctx->reading_tmp.p);
    }
    if (s == PMGPU_OK && cudaEventRecord(ctx->copy_done, ctx->stream) != cudaSuccess) s = fail(ctx, PMGPU_ERR_CUDA, "cudaEventRecord");
    if (s == PMGPU_OK) {
        ctx->nq = n;
        if (n > 0) s = morton_order(ctx);
    }
    ctx->stream = main_stream;
    PM_TRY(s);
    if (overlap) {
        PM_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_reading, ctx->stream2));
        PM_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_reading, 0));  // everything queued from here on sees the reading
    }
    // the caller may release `features` on return: wait for the copy only, the ordering kernels run on
    PM_CUDA_TRY(ctx, cudaEventSynchronize(ctx->copy_done));
    return PMGPU_OK;
}

int pmgpu_reading_set(pmgpu_ctx* ctx, const float* features, int rows, int n) { return reading_set_impl(ctx, features, rows, n, 0, 1, 0); }
int pmgpu_reading_set_sharded(pmgpu_ctx* ctx, const float* features, int rows, int n, int rank, int nranks, int chunk) {
    if (nranks <= 1) return reading_set_impl(ctx, features, rows, n, 0, 1, 0);
    return reading_set_impl(ctx, features, rows, n, rank, nranks, chunk);
}

int pmgpu_reading_set_max_dists(pmgpu_ctx* ctx, const float* max_dists, int ld) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    if (ctx->nq == 0) return fail(ctx, PMGPU_ERR_NO_READING, status_message(PMGPU_ERR_NO_READING));
    if (!max_dists) { ctx->has_reading_max_r2 = false; return PMGPU_OK; }
    if (ld < 1) return fail(ctx, PMGPU_ERR_BAD_ARG, "ld must be >= 1");
    const int n = ctx->nq;
    PM_CUDA_TRY(ctx, ctx->reading_max_r2.reserve(n));
    ScopedBuf<float> staging;
    PM_CUDA_TRY(ctx, staging.reserve((size_t)n * ld));
    PM_CUDA_TRY(ctx, cudaMemcpyAsync(staging.p, max_dists, ((size_t)(n - 1) * ld + 1) * sizeof(float), cudaMemcpyDefault, ctx->stream));
    pack_max_r2_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(staging.p, ld, ctx->q_order.p, n, ctx->reading_max_r2.p);
    ctx->launches += 1;
    PM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->has_reading_max_r2 = true;
    ctx->seed_k = 0;
    return PMGPU_OK;
}

int pmgpu_reading_set_normals(pmgpu_ctx* ctx, const float* normals, int ld) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    if (ctx->nq == 0) return fail(ctx, PMGPU_ERR_NO_READING, status_message(PMGPU_ERR_NO_READING));
    if (!normals) { ctx->has_reading_normals = false; return PMGPU_OK; }
    const int comps = ctx->dimh - 1;
    if (ld < comps) return fail(ctx, PMGPU_ERR_BAD_ARG, "normals_ld must be >= the cloud's dimension");
    const int n = ctx->nq;
    PM_CUDA_TRY(ctx, ctx->reading_normals.reserve(n));
    ScopedBuf<float> staging;
    PM_CUDA_TRY(ctx, staging.reserve((size_t)n * ld));
    PM_CUDA_TRY(ctx, cudaMemcpyAsync(staging.p, normals, ((size_t)(n - 1) * ld + comps) * sizeof(float), cudaMemcpyDefault, ctx->stream));
    pack_normals_permuted_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(staging.p, ld, ctx->q_order.p, n, ctx->reading_normals.p, comps);
    ctx->launches += 1;
    PM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->has_reading_normals = true;
    return PMGPU_OK;
}

int pmgpu_reading_apply_transform(pmgpu_ctx* ctx, const float* T) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    if (!T) return fail(ctx, PMGPU_ERR_BAD_ARG, "null transform");
    Mat4 M;
    T_load(ctx->dimh, T, M);
    if (!mat4_is_rigid(M)) return fail(ctx, PMGPU_ERR_NOT_ORTHOGONAL, status_message(PMGPU_ERR_NOT_ORTHOGONAL));
    if (ctx->nq > 0) {
        transform_inplace_kernel<<<(ctx->nq + 255) / 256, 256, 0, ctx->stream>>>(ctx->reading.p, ctx->nq, M);
        ctx->launches += 1;
        if (ctx->has_reading_normals) {
            rotate_normals_inplace_kernel<<<(ctx->nq + 255) / 256, 256, 0, ctx->stream>>>(ctx->reading_normals.p, ctx->nq, M);
            ctx->launches += 1;
        }
        PM_CUDA_TRY(ctx, cudaGetLastError());
    }
    ctx->have_matches = false;
    return PMGPU_OK;
}

int pmgpu_reading_get(pmgpu_ctx* ctx, float* features_out) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    if (!features_out) return fail(ctx, PMGPU_ERR_BAD_ARG, "null output");
    if (ctx->dimh == 4) {
        PM_TRY(download_unpermuted<f4>(ctx, ctx->reading.p, ctx->reading_tmp, 1, reinterpret_cast<f4*>(features_out)));
    } else if (ctx->nq > 0) {
        PM_CUDA_TRY(ctx, ctx->reading_tmp.reserve(ctx->nq));
        unpermute_kernel<f4><<<(ctx->nq + 255) / 256, 256, 0, ctx->stream>>>(ctx->reading.p, ctx->q_order.p, (size_t)ctx->nq, 1, ctx->reading_tmp.p);
        ScopedBuf<float> flat;
        PM_CUDA_TRY(ctx, flat.reserve(3 * (size_t)ctx->nq));
        collapse_2d_kernel<<<(ctx->nq + 255) / 256, 256, 0, ctx->stream>>>(ctx->reading_tmp.p, ctx->nq, flat.p);
        ctx->launches += 2;
        PM_CUDA_TRY(ctx, cudaMemcpyAsync(features_out, flat.p, 3 * (size_t)ctx->nq * sizeof(float), cudaMemcpyDefault, ctx->stream));
        PM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return PMGPU_OK;
    }
    PM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return PMGPU_OK;
}

int pmgpu_knn(pmgpu_ctx* ctx, const float* T, int k, float epsilon, float max_dist, int32_t* ids_out, float* dists_out, uint64_t* visit_out) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    if (ctx->nr == 0) return fail(ctx, PMGPU_ERR_NO_REFERENCE, status_message(PMGPU_ERR_NO_REFERENCE));
    if (k < 1) return fail(ctx, PMGPU_ERR_BAD_ARG, "knn must be >= 1");
    if (k > ctx->nr) return fail(ctx, PMGPU_ERR_KNN_TOO_LARGE, status_message(PMGPU_ERR_KNN_TOO_LARGE));
    if (!(epsilon >= 0.f)) return fail(ctx, PMGPU_ERR_BAD_ARG, "epsilon must be >= 0");
    const bool var_dist = max_dist < 0.f;  // KDTreeVarDistMatcher: per-point distances (pmgpu_reading_set_max_dists)
    if (var_dist && !ctx->has_reading_max_r2) return fail(ctx, PMGPU_ERR_BAD_ARG, "maxDist < 0 needs pmgpu_reading_set_max_dists first");
    if (!var_dist && !(max_dist >= 0.f)) return fail(ctx, PMGPU_ERR_BAD_ARG, "maxDist must be >= 0");
    IcpState* h = ctx->state_host;
    if (T) {
        T_load(ctx->dimh, T, h->T_iter);
        if (!mat4_is_rigid(h->T_iter)) return fail(ctx, PMGPU_ERR_NOT_ORTHOGONAL, status_message(PMGPU_ERR_NOT_ORTHOGONAL));
    } else {
        mat4_identity(h->T_iter);
    }
    h->T_match = h->T_iter;
    h->status = 0;
    h->iterate = 1;
    h->visits = 0;
    PM_TRY(push_state(ctx));
    const size_t total = (size_t)k * (ctx->nq > 0 ? ctx->nq : 1);
    PM_CUDA_TRY(ctx, ctx->ids.reserve(total));
    PM_CUDA_TRY(ctx, ctx->dists.reserve(total));
    ctx->k = k;
    ctx->have_weights = false;
    const float max_r2 = var_dist ? pm_inf() : max_dist * max_dist;
    ctx->stage_begin(0);
    PM_TRY(launch_knn(ctx, ctx->tree_view(), ctx->reading.p, ctx->nq, T != nullptr, false, false, k, max_r2,
                      ctx->seed_k == k && ctx->seed_enabled, ctx->ids.p, ctx->dists.p, false, var_dist ? ctx->reading_max_r2.p : nullptr));
    ctx->seed_k = k;
    ctx->stage_end();
    ctx->have_matches = true;
    PM_TRY(download_unpermuted<int32_t>(ctx, ctx->ids.p, ctx->ids_tmp, k, ids_out));
    PM_TRY(download_unpermuted<float>(ctx, ctx->dists.p, ctx->dists_tmp, k, dists_out));
    PM_TRY(pull_state(ctx));
    if (visit_out) *visit_out = ctx->state_host->visits;
    return PMGPU_OK;
}

int pmgpu_weights(pmgpu_ctx* ctx, int nfilters, const int* types, const float* params, float* weights_out, float* limits_out) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    if (!ctx->have_matches) return fail(ctx, PMGPU_ERR_NO_MATCHES, status_message(PMGPU_ERR_NO_MATCHES));
    if (nfilters > 0 && (!types || !params)) return fail(ctx, PMGPU_ERR_BAD_ARG, "null filter arrays");
    SelectSpec spec;
    PM_TRY(make_select_spec(ctx, nfilters, types, params, &spec));
    ctx->stage_begin(1);
    PM_TRY(launch_weights(ctx, spec, false, false));
    ctx->stage_end();
    const size_t total = (size_t)ctx->k * ctx->nq;
    if (weights_out && total) {
        PM_TRY(launch_materialize_weights(ctx));
        PM_TRY(download_unpermuted<float>(ctx, ctx->weights.p, ctx->dists_tmp, ctx->k, weights_out));
    }
    PM_TRY(pull_state(ctx));
    const int s = device_status(ctx);
    if (s != PMGPU_OK) {
        ctx->have_weights = false;
        ctx->state_host->status = 0;
        ctx->state_host->iterate = 1;
        PM_TRY(push_state(ctx));
        PM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return s;
    }
    if (limits_out)
        for (int f = 0; f < nfilters; ++f) limits_out[f] = spec.is_robust(f) ? ctx->state_host->robust_scale : ctx->state_host->limit[f];
    return PMGPU_OK;
}

int pmgpu_matches_get(pmgpu_ctx* ctx, int32_t* ids_out, float* dists_out, float* weights_out, float* T_match_out) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    if (!ctx->have_matches) return fail(ctx, PMGPU_ERR_NO_MATCHES, status_message(PMGPU_ERR_NO_MATCHES));
    if (weights_out && !ctx->have_weights) return fail(ctx, PMGPU_ERR_NO_MATCHES, "no outlier weights have been evaluated");
    const size_t total = (size_t)ctx->k * ctx->nq;
    if (ids_out && total) PM_TRY(download_unpermuted<int32_t>(ctx, ctx->ids.p, ctx->ids_tmp, ctx->k, ids_out));
    if (dists_out && total) PM_TRY(download_unpermuted<float>(ctx, ctx->dists.p, ctx->dists_tmp, ctx->k, dists_out));
    if (weights_out && total) {
        PM_TRY(launch_materialize_weights(ctx));
        PM_TRY(download_unpermuted<float>(ctx, ctx->weights.p, ctx->dists_tmp, ctx->k, weights_out));
    }
    PM_TRY(pull_state(ctx));
    if (T_match_out) T_store(ctx->dimh, ctx->state_host->T_match, T_match_out);
    return PMGPU_OK;
}

int pmgpu_set_var_trimmed_ratios(pmgpu_ctx* ctx, float min_ratio, float max_ratio) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    if (!(min_ratio > 0.f) || !(max_ratio <= 1.f) || !(min_ratio < max_ratio))
        return fail(ctx, PMGPU_ERR_BAD_ARG, "VarTrimmedDistOutlierFilter: 0 < minRatio < maxRatio <= 1");  // OutlierFiltersImpl.cpp:161-164, .h:156-157
    ctx->var_min_ratio = min_ratio;
    ctx->var_max_ratio = max_ratio;
    return PMGPU_OK;
}

int pmgpu_set_robust_approximation(pmgpu_ctx* ctx, float approximation) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    if (!(approximation >= 0.f)) return fail(ctx, PMGPU_ERR_BAD_ARG, "RobustOutlierFilter: approximation must be >= 0");
    // squaredApproximation(pow(get<T>("approximation"), 2)): the square is taken in double (OutlierFiltersImpl.cpp:400)
    ctx->robust_approx2 = (float)((double)approximation * (double)approximation);
    return PMGPU_OK;
}

int pmgpu_var_trimmed_ratio(pmgpu_ctx* ctx, float* ratio_out) {
    if (!ctx || !ratio_out) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    if (!ctx->have_weights) return fail(ctx, PMGPU_ERR_NO_MATCHES, "no outlier weights have been evaluated");
    PM_TRY(pull_state(ctx));
    *ratio_out = ctx->state_host->var_ratio;
    return PMGPU_OK;
}

int pmgpu_minimize(pmgpu_ctx* ctx, int minimizer, float sensor_std_dev, float* T_out, float* cov_out, float* stats_out) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    if (!ctx->have_matches) return fail(ctx, PMGPU_ERR_NO_MATCHES, status_message(PMGPU_ERR_NO_MATCHES));
    if (minimizer < 0 || (minimizer & 0xff) > PMGPU_MIN_P2POINT_SIM || (minimizer & ~0x3ff)) return fail(ctx, PMGPU_ERR_BAD_ARG, "unknown error minimizer");
    if ((minimizer & PMGPU_MIN_FORCE4DOF) && (minimizer & 0xff) != PMGPU_MIN_P2PLANE && (minimizer & 0xff) != PMGPU_MIN_P2PLANE_COV)
        return fail(ctx, PMGPU_ERR_BAD_ARG, "force4DOF is a point-to-plane parameter");
    if ((minimizer & PMGPU_MIN_FORCE2D) && minimizer != (PMGPU_MIN_P2PLANE | PMGPU_MIN_FORCE2D))
        return fail(ctx, PMGPU_ERR_BAD_ARG, "force2D goes with PointToPlaneErrorMinimizer alone (no force4DOF, no covariance)");
    if (ctx->dimh == 3 && minimizer != PMGPU_MIN_P2POINT && minimizer != PMGPU_MIN_P2PLANE)
        return fail(ctx, PMGPU_ERR_UNSUPPORTED, "2-D clouds: PointToPoint and PointToPlane error minimizers only (no covariance, similarity, force2D / force4DOF)");
    if (!ctx->have_weights) PM_TRY(launch_weights(ctx, SelectSpec(), false, false));  // empty chain
    ctx->stage_begin(2);
    PM_TRY(launch_minimize(ctx, minimizer, false, false, nullptr));
    ctx->stage_end();
    const bool with_cov = (minimizer & 0xff) == PMGPU_MIN_P2POINT_COV || (minimizer & 0xff) == PMGPU_MIN_P2PLANE_COV;
    if (with_cov) {
        ctx->stage_begin(3);
        PM_TRY(launch_covariance(ctx, minimizer, sensor_std_dev));
        ctx->stage_end();
    }
    PM_TRY(pull_state(ctx));
    const int s = device_status(ctx);
    if (s != PMGPU_OK) {
        ctx->state_host->status = 0;
        ctx->state_host->iterate = 1;
        PM_TRY(push_state(ctx));
        PM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        return s;
    }
    if (T_out) T_store(ctx->dimh, ctx->state_host->dT, T_out);
    if (cov_out && with_cov) memcpy(cov_out, ctx->state_host->cov, sizeof(float) * 36);
    if (stats_out) memcpy(stats_out, ctx->state_host->stats, sizeof(float) * 5);
    return PMGPU_OK;
}

int pmgpu_icp_reset(pmgpu_ctx* ctx, const float* T_iter_init) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    IcpState* h = ctx->state_host;
    PM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (T_iter_init) T_load(ctx->dimh, T_iter_init, h->T_iter);
    else mat4_identity(h->T_iter);
    if (!mat4_is_rigid(h->T_iter)) return fail(ctx, PMGPU_ERR_NOT_ORTHOGONAL, status_message(PMGPU_ERR_NOT_ORTHOGONAL));
    h->T_match = h->T_iter;
    mat4_identity(h->dT);
    h->status = 0;
    h->iterate = 1;
    h->iterations = 0;
    h->counter = 0;
    h->visits = 0;
    h->robust_iteration = 1;  // a fresh RobustOutlierFilter (OutlierFiltersImpl.cpp:420-438)
    h->robust_scale = 0.f;
    h->cap = pm_inf();
    h->cap_need = 0.f;
    h->redo = 0;
    h->redo_count = 0;
    for (int f = 0; f < PM_MAX_FILTERS; ++f) { h->sel_guess[f] = 0; h->sel_prev[f] = 0; h->sel_inner[f] = 64; }  // no previous order statistic: the first select is the generic one
    h->sel_passes = 0;
    // a fresh registration never starts from an earlier run's matches (they would only be a seed, but a seed a first
    // iteration does not have)
    ctx->seed_k = 0;
    // TransformationCheckers::init (TransformationCheckersImpl.cpp:107-124): history starts with T_iter
    const Quat q = quat_from_mat4(h->T_iter);
    h->hist_q[0][0] = q.w; h->hist_q[0][1] = q.x; h->hist_q[0][2] = q.y; h->hist_q[0][3] = q.z;
    h->hist_t[0][0] = h->T_iter.m[12]; h->hist_t[0][1] = h->T_iter.m[13]; h->hist_t[0][2] = h->T_iter.m[14];
    h->hist_len = 1;
    PM_TRY(push_state(ctx));
    PM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return PMGPU_OK;
}

int pmgpu_icp_enqueue(pmgpu_ctx* ctx, const pmgpu_icp_params* params, int n_iterations) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    if (ctx->nr == 0) return fail(ctx, PMGPU_ERR_NO_REFERENCE, status_message(PMGPU_ERR_NO_REFERENCE));
    // a rank of a sharded registration may hold an empty slice: its kernels still run (one block each) and take part in
    // the exchanges, or its peers would wait for it
    if (ctx->nq == 0 && ctx->nranks <= 1) return fail(ctx, PMGPU_ERR_NO_READING, status_message(PMGPU_ERR_NO_READING));
    PM_TRY(check_params(ctx, params));
    if (params->max_dist < 0.f && !ctx->has_reading_max_r2) return fail(ctx, PMGPU_ERR_BAD_ARG, "maxDist < 0 needs pmgpu_reading_set_max_dists first");
    const bool plane = (params->minimizer & 0xff) == PMGPU_MIN_P2PLANE || (params->minimizer & 0xff) == PMGPU_MIN_P2PLANE_COV;
    if (plane && !ctx->has_normals) return fail(ctx, PMGPU_ERR_NO_NORMALS, status_message(PMGPU_ERR_NO_NORMALS));
    const size_t total = (size_t)params->knn * ctx->nq;
    PM_CUDA_TRY(ctx, ctx->ids.reserve(total));
    PM_CUDA_TRY(ctx, ctx->dists.reserve(total));
    ctx->k = params->knn;
    for (int it = 0; it < n_iterations; ++it) PM_TRY(enqueue_iteration(ctx, params, true));
    ctx->have_matches = true;
    ctx->have_weights = true;
    return PMGPU_OK;
}

int pmgpu_icp_result(pmgpu_ctx* ctx, float* T_iter_out, int* iterations_out, float* cov_out, float* stats_out) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    PM_TRY(pull_state(ctx));
    if (T_iter_out) T_store(ctx->dimh, ctx->state_host->T_iter, T_iter_out);
    if (iterations_out) *iterations_out = ctx->state_host->iterations;
    if (cov_out) memcpy(cov_out, ctx->state_host->cov, sizeof(float) * 36);
    if (stats_out) memcpy(stats_out, ctx->state_host->stats, sizeof(float) * 5);
    return device_status(ctx);
}

int pmgpu_host_pin(void* ptr, size_t bytes) {
    if (!ptr || bytes == 0) return PMGPU_ERR_BAD_ARG;
    return cudaHostRegister(ptr, bytes, cudaHostRegisterDefault) == cudaSuccess ? PMGPU_OK : (cudaGetLastError(), PMGPU_ERR_CUDA);
}
int pmgpu_host_unpin(void* ptr) {
    if (!ptr) return PMGPU_ERR_BAD_ARG;
    return cudaHostUnregister(ptr) == cudaSuccess ? PMGPU_OK : (cudaGetLastError(), PMGPU_ERR_CUDA);
}

int pmgpu_icp_cap_redos(const pmgpu_ctx* ctx) { return ctx ? ctx->state_host->redo_count : 0; }

int pmgpu_icp_step(pmgpu_ctx* ctx, const pmgpu_icp_params* params, float* T_iter_out, int* iterations_out, float* cov_out, float* stats_out) {
    if (!ctx || !params) return PMGPU_ERR_BAD_ARG;
    // exact matching for this slot: whoever asks for single iterations is going to look at the matches
    const bool cap = ctx->cap_enabled;
    ctx->cap_enabled = false;
    const int rc = pmgpu_icp_enqueue(ctx, params, 1);
    ctx->cap_enabled = cap;
    PM_TRY(rc);
    const bool with_cov = (params->minimizer & 0xff) == PMGPU_MIN_P2POINT_COV || (params->minimizer & 0xff) == PMGPU_MIN_P2PLANE_COV;
    if (with_cov) {
        PM_TRY(pull_state(ctx));
        if (ctx->state_host->status == PMGPU_OK) PM_TRY(launch_covariance(ctx, params->minimizer, params->sensor_std_dev));
    }
    return pmgpu_icp_result(ctx, T_iter_out, iterations_out, cov_out, stats_out);
}

int pmgpu_icp_run(pmgpu_ctx* ctx, const pmgpu_icp_params* params, const float* T_iter_init, float* T_iter_out, int* iterations_out, float* cov_out,
                  float* stats_out) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    if (ctx->nr == 0) return fail(ctx, PMGPU_ERR_NO_REFERENCE, status_message(PMGPU_ERR_NO_REFERENCE));
    PM_TRY(check_params(ctx, params));
    PM_TRY(pmgpu_icp_reset(ctx, T_iter_init));
    // ICP.cpp:371: `while (iterate)` runs the body at least once, Counter stops it after
    // maxIterationCount checks (max(1, maxIterationCount) iterations)
    const int n = params->max_iterations > 1 ? params->max_iterations : 1;
    // Iteration slots are enqueued in bounded chunks with a look at the state in between: a chain without a Counter
    // (max_iterations = INT_MAX, Differential only) or with a large one must not queue millions of gated no-op launches
    // before the host learns that the checkers have stopped the loop.  A slot whose capped match was void
    // (IcpState::redo) does not count as an iteration, so slots are topped up until the checkers stop the loop.
    const int chunk = 64;
    for (;;) {
        const IcpState* h = ctx->state_host;
        const int left = n - h->iterations;
        PM_TRY(pmgpu_icp_enqueue(ctx, params, left < chunk ? left : chunk));
        PM_TRY(pull_state(ctx));
        if (h->status != PMGPU_OK || !h->iterate || h->iterations >= n) break;
    }
    const bool with_cov = (params->minimizer & 0xff) == PMGPU_MIN_P2POINT_COV || (params->minimizer & 0xff) == PMGPU_MIN_P2PLANE_COV;
    if (with_cov) {
        PM_TRY(pull_state(ctx));
        if (ctx->state_host->status == PMGPU_OK) PM_TRY(launch_covariance(ctx, params->minimizer, params->sensor_std_dev));
    }
    return pmgpu_icp_result(ctx, T_iter_out, iterations_out, cov_out, stats_out);
}

int pmgpu_ref_compute_normals(pmgpu_ctx* ctx, int knn, float epsilon, float max_dist, int flags) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx, true));
    if (ctx->nr == 0) return fail(ctx, PMGPU_ERR_NO_REFERENCE, status_message(PMGPU_ERR_NO_REFERENCE));
    if (knn < 1) return fail(ctx, PMGPU_ERR_BAD_ARG, "knn must be >= 1");
    if (knn > ctx->nr) return fail(ctx, PMGPU_ERR_KNN_TOO_LARGE, status_message(PMGPU_ERR_KNN_TOO_LARGE));
    if (!(epsilon >= 0.f) || !(max_dist >= 0.f)) return fail(ctx, PMGPU_ERR_BAD_ARG, "epsilon and maxDist must be >= 0");
    const bool smooth = (flags & PMGPU_NORMALS_SMOOTH) != 0;
    const int n = ctx->nr;
    PM_CUDA_TRY(ctx, ctx->ref_normals.reserve(n));
    NormalsSink sink;
    memset(&sink, 0, sizeof(sink));
    sink.pts = ctx->ref_orig.p;
    sink.degenerate = &ctx->state->degenerate;
    sink.dim2 = ctx->dimh == 3 ? 1 : 0;
    const float max_r2 = max_dist * max_dist;
    ScopedBuf<int32_t> nids;
    if (smooth) {
        PM_CUDA_TRY(ctx, nids.reserve((size_t)knn * n));
        sink.ids_i32 = nids.p;
    }
    if (ctx->nranks > 1 && ctx->nccl_comm && !smooth) {
        // every rank holds the whole structure and computes the normals of one slice of it (leaf-order
        // positions, so the slice is spatially compact and its output contiguous); one all-gather of float4[N / G] and a
        // local scatter into the caller's column order complete every rank's copy
        const int chunk = (n + ctx->nranks - 1) / ctx->nranks;
        const int lo = ctx->rank * chunk < n ? ctx->rank * chunk : n;
        const int hi = lo + chunk < n ? lo + chunk : n;
        PM_CUDA_TRY(ctx, ctx->gather_tmp.reserve((size_t)chunk * ctx->nranks));
        sink.normals4 = ctx->gather_tmp.p;
        sink.by_position = 1;
        ctx->stage_begin(0);
        PM_TRY(launch_knn_normals(ctx, ctx->tree_view(), lo, hi, knn, max_r2, sink));
        ctx->stage_end();
        PM_TRY(comm_allgather_bytes(ctx, ctx->gather_tmp.p, (size_t)chunk * sizeof(f4)));
        scatter_by_leaf_order_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->gather_tmp.p, ctx->ref_sorted.p, n, ctx->ref_normals.p);
        ctx->launches += 1;
    } else {
        sink.normals4 = ctx->ref_normals.p;
        ctx->stage_begin(0);
        PM_TRY(launch_knn_normals(ctx, ctx->tree_view(), 0, n, knn, max_r2, sink));
        ctx->stage_end();
    }
    PM_CUDA_TRY(ctx, cudaGetLastError());
    if (smooth) {
        // the serial pass of the reference, on the host (see normals.cuh)
        std::vector<f4> h_n((size_t)n);
        std::vector<int32_t> h_ids((size_t)knn * n);
        PM_CUDA_TRY(ctx, cudaMemcpyAsync(h_n.data(), ctx->ref_normals.p, h_n.size() * sizeof(f4), cudaMemcpyDeviceToHost, ctx->stream));
        PM_CUDA_TRY(ctx, cudaMemcpyAsync(h_ids.data(), nids.p, h_ids.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        PM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        smooth_normals_host(h_n.data(), h_ids.data(), knn, n);
        PM_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->ref_normals.p, h_n.data(), h_n.size() * sizeof(f4), cudaMemcpyHostToDevice, ctx->stream));
        PM_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->has_normals = true;
    return PMGPU_OK;
}

int pmgpu_normals(pmgpu_ctx* ctx, const float* features, int rows, int n, int knn, float epsilon, float max_dist, int flags,
                  const pmgpu_normals_out* out, int* degenerate_out) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    PM_TRY(use_device(ctx));
    if (!features || !out) return fail(ctx, PMGPU_ERR_BAD_ARG, "null argument");
    if (rows != 4 && rows != 3) return fail(ctx, PMGPU_ERR_UNSUPPORTED, status_message(PMGPU_ERR_UNSUPPORTED));
    if (n < 1) return fail(ctx, PMGPU_ERR_BAD_ARG, "empty cloud");
    const int dn = rows - 1;  // spans: normals dn, eig_values dn, eig_vectors dn * dn
    if (knn < 1) return fail(ctx, PMGPU_ERR_BAD_ARG, "knn must be >= 1");
    if (knn > n) return fail(ctx, PMGPU_ERR_KNN_TOO_LARGE, status_message(PMGPU_ERR_KNN_TOO_LARGE));
    if (!(epsilon >= 0.f) || !(max_dist >= 0.f)) return fail(ctx, PMGPU_ERR_BAD_ARG, "epsilon and maxDist must be >= 0");
    const bool smooth = (flags & PMGPU_NORMALS_SMOOTH) != 0 && out->normals;
    // a private context holds the cloud's own search structure (the filter builds its own
    // KDTreeMatcher, SurfaceNormal.cpp:153-162); it shares nothing with the ICP reference
    pmgpu_ctx* sub = nullptr;
    int s = pmgpu_ctx_create(ctx->device, &sub);
    if (s != PMGPU_OK) return fail(ctx, s, "cannot create the filter's private context");
    struct Guard {
        pmgpu_ctx* c;
        pmgpu_ctx* parent;
        ~Guard() { parent->launches += c->launches; pmgpu_ctx_destroy(c); g_alloc_stream = parent->stream; }
    } guard{sub, ctx};
    s = pmgpu_ref_set(sub, features, rows, n, nullptr, 0);
    if (s != PMGPU_OK) return fail(ctx, s, sub->err.c_str());
    ScopedBuf<float> scratch;
    ScopedBuf<f4> n4;
    ScopedBuf<int32_t> nids;
    if (out->normals) PM_CUDA_TRY(ctx, n4.reserve(n));
    if (smooth) PM_CUDA_TRY(ctx, nids.reserve((size_t)knn * n));
    // scratch (laid out for the 3-D spans): densities n | eig_values 3n | eig_vectors 9n | mean_dists n | normals 3n | ids-as-float knn*n
    const size_t off_den = 0, off_val = (size_t)n, off_vec = 4 * (size_t)n, off_md = 13 * (size_t)n, off_n3 = 14 * (size_t)n, off_ids = 17 * (size_t)n;
    PM_CUDA_TRY(ctx, scratch.reserve(17 * (size_t)n + (out->matched_ids ? (size_t)knn * n : 0)));
    cudaStream_t st = sub->stream;
    IcpState* h = sub->state_host;
    h->degenerate = 0;
    s = push_state(sub);
    // K8 as the epilogue of the self-kNN: only what the caller asked for is computed and written
    NormalsSink sink;
    memset(&sink, 0, sizeof(sink));
    sink.pts = sub->ref_orig.p;
    sink.normals4 = out->normals ? n4.p : nullptr;
    sink.densities = out->densities ? scratch.p + off_den : nullptr;
    sink.eig_values = out->eig_values ? scratch.p + off_val : nullptr;
    sink.eig_vectors = out->eig_vectors ? scratch.p + off_vec : nullptr;
    sink.mean_dists = out->mean_dists ? scratch.p + off_md : nullptr;
    sink.matched_ids = out->matched_ids ? scratch.p + off_ids : nullptr;
    sink.degenerate = &sub->state->degenerate;
    sink.dim2 = rows == 3 ? 1 : 0;
    sink.ids_i32 = smooth ? nids.p : nullptr;
    if (s == PMGPU_OK) s = launch_knn_normals(sub, sub->tree_view(), 0, n, knn, max_dist * max_dist, sink);
    if (s != PMGPU_OK) return fail(ctx, s, sub->err.c_str());
    if (smooth) {
        std::vector<f4> h_n((size_t)n);
        std::vector<int32_t> h_ids((size_t)knn * n);
        PM_CUDA_TRY(ctx, cudaMemcpyAsync(h_n.data(), n4.p, h_n.size() * sizeof(f4), cudaMemcpyDeviceToHost, sub->stream));
        PM_CUDA_TRY(ctx, cudaMemcpyAsync(h_ids.data(), nids.p, h_ids.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, sub->stream));
        PM_CUDA_TRY(ctx, cudaStreamSynchronize(sub->stream));
        smooth_normals_host(h_n.data(), h_ids.data(), knn, n);
        PM_CUDA_TRY(ctx, cudaMemcpyAsync(n4.p, h_n.data(), h_n.size() * sizeof(f4), cudaMemcpyHostToDevice, sub->stream));
        PM_CUDA_TRY(ctx, cudaStreamSynchronize(sub->stream));
    }
    auto copy_out = [&](float* dst, int ld, const float* src, int span) -> cudaError_t {
        if (!dst) return cudaSuccess;
        if (ld == span) return cudaMemcpyAsync(dst, src, (size_t)span * n * sizeof(float), cudaMemcpyDefault, st);
        return cudaMemcpy2DAsync(dst, (size_t)ld * sizeof(float), src, (size_t)span * sizeof(float), (size_t)span * sizeof(float), n, cudaMemcpyDefault, st);
    };
    if (out->normals) {
        unpack_f4_kernel<<<(n + 255) / 256, 256, 0, st>>>(n4.p, n, scratch.p + off_n3, dn);
        sub->launches += 1;
        PM_CUDA_TRY(ctx, copy_out(out->normals, out->normals_ld, scratch.p + off_n3, dn));
    }
    PM_CUDA_TRY(ctx, copy_out(out->densities, out->densities_ld, scratch.p + off_den, 1));
    PM_CUDA_TRY(ctx, copy_out(out->eig_values, out->eig_values_ld, scratch.p + off_val, dn));
    PM_CUDA_TRY(ctx, copy_out(out->eig_vectors, out->eig_vectors_ld, scratch.p + off_vec, dn * dn));
    PM_CUDA_TRY(ctx, copy_out(out->mean_dists, out->mean_dists_ld, scratch.p + off_md, 1));
    PM_CUDA_TRY(ctx, copy_out(out->matched_ids, out->matched_ids_ld, scratch.p + off_ids, knn));
    s = pull_state(sub);
    if (s != PMGPU_OK) return fail(ctx, s, sub->err.c_str());
    PM_CUDA_TRY(ctx, cudaGetLastError());
    if (degenerate_out) *degenerate_out = h->degenerate;
    return PMGPU_OK;
}

}  // extern "C"
