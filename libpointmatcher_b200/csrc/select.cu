// select.cu — K3: Matches::getDistsQuantile (Matches.cpp:60-87) as an exact 3-pass radix select,
// and the outlier-filter chain OutlierFilters::compute (OutlierFilter.cpp:63-103) over
// MaxDist / MedianDist / TrimmedDist (OutlierFiltersImpl.cpp:66-81, 109-147).
//
// Squared distances are non-negative floats, so their bit patterns order like unsigned integers:
// pass 0 histograms bits 31..21, pass 1 bits 20..10 inside the selected bucket, pass 2 bits 9..0.
// Every pass is one streaming read of the k x N distance matrix (4 B per match, HBM bound) into a
// shared-memory histogram; a one-block kernel scans the 2048 bins and narrows the bucket.  The
// three in-scope filters all produce weights of the form (dist <= limit_f), and the chain
// multiplies them, so the whole chain collapses to ONE threshold limit_all = min_f limit_f that
// the minimiser kernels apply on the fly: the weight matrix is never written unless a caller asks
// for it.  With queries sharded over GPUs the histograms are all-reduced between the two kernels
// of a pass, so every rank selects the same global order statistic.
#include "pmgpu_internal.cuh"

namespace pm {

namespace {

__global__ void init_limits_kernel(IcpState* state, int gated, int nfilters, const int* __restrict__ types_unused, int t0, int t1, int t2, int t3,
                                   int t4, int t5, int t6, int t7, float l0, float l1, float l2, float l3, float l4, float l5, float l6, float l7) {
    (void)types_unused;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (gated && state->iterate == 0) return;
    const int types[8] = {t0, t1, t2, t3, t4, t5, t6, t7};
    const float lims[8] = {l0, l1, l2, l3, l4, l5, l6, l7};
    float all = pm_inf();
    for (int f = 0; f < nfilters; ++f) {
        if (types[f] == PMGPU_FILTER_MAXDIST) {
            state->limit[f] = lims[f];
            all = fminf(all, lims[f]);
        } else {
            state->limit[f] = pm_inf();
        }
    }
    state->limit_all = all;
    state->has_filters = nfilters > 0 ? 1 : 0;
}

// pass: 0, 1, 2.  Finite distances only (Matches.cpp:70).
__global__ void __launch_bounds__(256) hist_kernel(const float* __restrict__ dists, size_t total, int pass, const IcpState* __restrict__ state,
                                                   int gated, unsigned* __restrict__ hist) {
    __shared__ unsigned sh[PM_HIST_BINS];
    if (gated && state->iterate == 0) return;
    for (int i = threadIdx.x; i < PM_HIST_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const unsigned prefix = state->prefix;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const unsigned b = __float_as_uint(__ldg(dists + i));
        if (b == PM_INF_BITS) continue;
        if (pass == 0) {
            atomicAdd(&sh[b >> 21], 1u);
        } else if (pass == 1) {
            if ((b >> 21) == prefix) atomicAdd(&sh[(b >> 10) & 0x7ffu], 1u);
        } else {
            if ((b >> 10) == prefix) atomicAdd(&sh[b & 0x3ffu], 1u);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < PM_HIST_BINS; i += blockDim.x) {
        const unsigned v = sh[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// One block of 1024 threads: inclusive scan of the 2048 bins, pick the bin holding `rank`.
// pass 0 also derives the rank from the number of finite distances:
//   rank = size_t(float(n_valid) * quantile)   (Matches.cpp:85-86; quantile == 1 -> max element)
// pass 2 finishes the filter: limit[f] = value (* factor for MedianDist), limit_all = min.
__global__ void __launch_bounds__(1024) pick_kernel(unsigned* __restrict__ hist, int pass, float quantile, int filter_index, float factor,
                                                    IcpState* state, int gated) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long s_rank;
    __shared__ int s_abort;
    if (gated && state->iterate == 0) return;
    const int t = threadIdx.x;
    const unsigned h0 = hist[2 * t], h1 = hist[2 * t + 1];
    hist[2 * t] = 0;  // leave the histogram clean for the next pass
    hist[2 * t + 1] = 0;
    unsigned long long mine = (unsigned long long)h0 + h1;
    // inclusive scan of `mine` over the block
    unsigned long long incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((t & 31) >= o) incl += v;
    }
    if ((t & 31) == 31) warp_tot[t >> 5] = incl;
    __syncthreads();
    if (t < 32) {
        unsigned long long w = warp_tot[t];
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, w, o);
            if (t >= o) w += v;
        }
        warp_tot[t] = w;  // inclusive over warps
    }
    __syncthreads();
    const unsigned long long before_warp = (t >> 5) ? warp_tot[(t >> 5) - 1] : 0ull;
    const unsigned long long excl = before_warp + incl - mine;  // elements in bins < 2t
    const unsigned long long total = warp_tot[31];
    if (t == 0) {
        s_abort = 0;
        if (pass == 0) {
            state->n_valid = total;
            if (total == 0) {
                if (state->status == 0) state->status = PMGPU_ERR_NO_OUTLIER_TO_FILTER;
                state->iterate = 0;
                s_abort = 1;
                s_rank = 0;
            } else {
                unsigned long long r;
                if (quantile == 1.0f) r = total - 1;
                else {
                    r = (unsigned long long)(__ull2float_rn(total) * quantile);
                    if (r > total - 1) r = total - 1;
                }
                s_rank = r;
            }
        } else {
            s_rank = state->rank;
        }
    }
    __syncthreads();
    if (s_abort) return;
    const unsigned long long rank = s_rank;
    // the bin b with excl(b) <= rank < excl(b) + count(b)
    int found = -1;
    unsigned long long rem = 0;
    if (rank >= excl && rank < excl + h0) { found = 2 * t; rem = rank - excl; }
    else if (rank >= excl + h0 && rank < excl + mine) { found = 2 * t + 1; rem = rank - excl - h0; }
    if (found >= 0) {
        if (pass == 0) { state->prefix = (unsigned)found; state->rank = rem; }
        else if (pass == 1) { state->prefix = (state->prefix << 11) | (unsigned)found; state->rank = rem; }
        else {
            const unsigned bits = (state->prefix << 10) | (unsigned)found;
            const float value = __uint_as_float(bits);
            const float lim = factor != 0.f ? __fmul_rn(factor, value) : value;
            state->limit[filter_index] = lim;
            state->limit_all = fminf(state->limit_all, lim);
        }
    }
}

__global__ void weights_kernel(const float* __restrict__ dists, size_t total, const IcpState* __restrict__ state, float* __restrict__ w) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float d = dists[i];
    // empty chain: OutlierFilter.cpp:70-85; otherwise product of (dist <= limit_f)
    w[i] = state->has_filters ? ((d <= state->limit_all) ? 1.f : 0.f) : ((d == pm_inf()) ? 0.f : 1.f);
}

}  // namespace

int launch_weights(pmgpu_ctx* ctx, int nfilters, const int* types, const float* params, bool gated) {
    if (nfilters < 0 || nfilters > PM_MAX_FILTERS) {
        ctx->set_error("at most 8 outlier filters are supported in one chain");
        return PMGPU_ERR_BAD_ARG;
    }
    cudaStream_t st = ctx->stream;
    PM_CUDA_TRY(ctx, ctx->hist.reserve(PM_HIST_BINS));
    bool any_quantile = false;
    for (int f = 0; f < nfilters; ++f) any_quantile |= (types[f] != PMGPU_FILTER_MAXDIST);
    if (any_quantile) PM_CUDA_TRY(ctx, cudaMemsetAsync(ctx->hist.p, 0, PM_HIST_BINS * sizeof(unsigned), st));
    int t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float l[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int f = 0; f < nfilters; ++f) {
        t[f] = types[f];
        if (types[f] == PMGPU_FILTER_MAXDIST) {
            // maxDist(pow(get<T>("maxDist"), 2)): the square is taken in double (OutlierFiltersImpl.cpp:69)
            l[f] = (float)((double)params[f] * (double)params[f]);
        } else if (types[f] == PMGPU_FILTER_TRIMMEDDIST) {
            if (params[f] < 0.f || params[f] > 1.f) {
                ctx->set_error("quantile must be between 0 and 1");
                return PMGPU_ERR_BAD_QUANTILE;
            }
        } else if (types[f] != PMGPU_FILTER_MEDIANDIST) {
            ctx->set_error("unknown outlier filter type");
            return PMGPU_ERR_BAD_ARG;
        }
    }
    init_limits_kernel<<<1, 32, 0, st>>>(ctx->state, gated ? 1 : 0, nfilters, nullptr, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], l[0], l[1], l[2],
                                         l[3], l[4], l[5], l[6], l[7]);
    ctx->launches += 1;
    const size_t total = (size_t)ctx->k * ctx->nq;
    const int B = 256;
    const int grid = grid_for((int)((total + 3) / 4 > 0x7fffffff ? 0x7fffffff : (total + 3) / 4), B, ctx->num_sms, 8);
    for (int f = 0; f < nfilters; ++f) {
        if (types[f] == PMGPU_FILTER_MAXDIST) continue;
        const float q = types[f] == PMGPU_FILTER_MEDIANDIST ? 0.5f : params[f];
        const float factor = types[f] == PMGPU_FILTER_MEDIANDIST ? params[f] : 0.f;
        for (int pass = 0; pass < 3; ++pass) {
            hist_kernel<<<grid, B, 0, st>>>(ctx->dists.p, total, pass, ctx->state, gated ? 1 : 0, ctx->hist.p);
            ctx->launches += 1;
            if (ctx->nranks > 1) PM_TRY(comm_allreduce_u32(ctx, ctx->hist.p, PM_HIST_BINS));
            pick_kernel<<<1, 1024, 0, st>>>(ctx->hist.p, pass, q, f, factor, ctx->state, gated ? 1 : 0);
            ctx->launches += 1;
        }
    }
    PM_CUDA_TRY(ctx, cudaGetLastError());
    ctx->have_weights = true;
    return PMGPU_OK;
}

int launch_materialize_weights(pmgpu_ctx* ctx) {
    const size_t total = (size_t)ctx->k * ctx->nq;
    PM_CUDA_TRY(ctx, ctx->weights.reserve(total));
    const int B = 256;
    weights_kernel<<<(unsigned)((total + B - 1) / B), B, 0, ctx->stream>>>(ctx->dists.p, total, ctx->state, ctx->weights.p);
    ctx->launches += 1;
    PM_CUDA_TRY(ctx, cudaGetLastError());
    return PMGPU_OK;
}

}  // namespace pm
