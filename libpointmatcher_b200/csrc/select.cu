// select.cu — K3: Matches::getDistsQuantile (Matches.cpp:60-87) as an exact 3-pass radix select,
// and the outlier-filter chain OutlierFilters::compute (OutlierFilter.cpp:63-103) over
// MaxDist / MedianDist / TrimmedDist (OutlierFiltersImpl.cpp:66-81, 109-147).
//
// Every pass is one streaming read of the k x N distance matrix (4 B per match, HBM/L2 bound)
// into a shared-memory histogram; the last block to finish scans the 2048 bins and narrows the
// bucket (select.cuh), so a pass is one kernel.  The three in-scope filters all produce weights of the
// form (dist <= limit_f), and the chain multiplies them, so the whole chain collapses to ONE
// threshold limit_all = min_f limit_f that the minimiser kernels apply on the fly: the weight
// matrix is never written unless a caller asks for it.  With queries sharded over GPUs the
// histograms are all-reduced between the histogram kernel and a one-block pick kernel, so every
// rank selects the same global order statistic.
#include <cub/device/device_radix_sort.cuh>

#include "comm.cuh"
#include "select.cuh"

namespace pm {

namespace {

constexpr int HIST_BLOCK = 1024;

__device__ __forceinline__ void hist_add(unsigned* sh, float d, int pass, unsigned prefix) {
    const int bin = select_bin(__float_as_uint(d), pass, prefix);
    if (bin >= 0) atomicAdd(&sh[bin], 1u);
}

// pass 0: one histogram serves every quantile filter (slot 0); pass 1, 2: one per filter.
// One 1024-thread block per SM: every block flushes its 2048 bins with global atomics, so the fewer
// blocks the shorter the serialised tail on the hot bins; the distances are read four at a time.
__global__ void __launch_bounds__(HIST_BLOCK) hist_kernel(const float* __restrict__ dists, size_t total, int pass, SelectSpec spec, IcpState* state,
                                                          int gated, int do_init, int do_pick, unsigned* __restrict__ hist, int cap_active,
                                                          float cap_margin, PeerComm pc) {
    __shared__ unsigned sh[PM_HIST_BINS];
    if (gated && state->iterate == 0) return;
    if (do_init && blockIdx.x == 0 && threadIdx.x == 0) select_init_limits(state, spec);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t quads = total / 4;
    const float4* __restrict__ d4 = reinterpret_cast<const float4*>(dists);
    int slot = 0;
    for (int f = 0; f < spec.nfilters; ++f) {
        if (!spec.is_quantile(f)) continue;
        for (int i = threadIdx.x; i < PM_HIST_BINS; i += blockDim.x) sh[i] = 0;
        __syncthreads();
        const unsigned prefix = pass == 0 ? 0u : state->sel_prefix[f];
        for (size_t i = gtid; i < quads; i += stride) {
            const float4 v = __ldg(d4 + i);
            hist_add(sh, v.x, pass, prefix);
            hist_add(sh, v.y, pass, prefix);
            hist_add(sh, v.z, pass, prefix);
            hist_add(sh, v.w, pass, prefix);
        }
        if (gtid < total - 4 * quads) hist_add(sh, __ldg(dists + 4 * quads + gtid), pass, prefix);
        __syncthreads();
        select_flush(sh, hist + (size_t)slot * PM_HIST_BINS);
        __syncthreads();
        ++slot;
        if (pass == 0) break;
    }
    if (!do_pick || slot == 0) return;
    if (!select_last_block(&state->ticket[0])) return;
    const int nq = spec.n_quantile();
    // sharded reading: this rank's histograms become the sum over all ranks, exchanged over the peer mailboxes right here
    if (pc.nranks > 1 && !peer_allreduce<false>(pc, hist, (pass == 0 ? 1 : nq) * PM_HIST_BINS, state)) return;
    slot = 0;
    for (int f = 0; f < spec.nfilters; ++f) {
        if (!spec.is_quantile(f)) continue;
        const bool last = slot == nq - 1;
        select_pick(hist + (size_t)(pass == 0 ? 0 : slot) * PM_HIST_BINS, pass, spec.quantile(f), f, spec.factor(f), state, pass != 0 || last);
        ++slot;
    }
    if (pass == 2 && threadIdx.x == 0) select_finish(state, cap_active, cap_margin);
}

// RobustOutlierFilter scale = sqrt(MAD): the same three passes twice, phase 0 on the distances
// (median), phase 1 on |dist - median| (Matches.cpp:88-122); histogram slot PM_MAX_FILTERS
// `want`: the estimator these passes belong to (robust_recompute: 1 mad, 2 berg's first median)
__global__ void __launch_bounds__(HIST_BLOCK) robust_hist_kernel(const float* __restrict__ dists, size_t total, int pass, int phase, IcpState* state,
                                                                 int gated, unsigned* __restrict__ hist, int want) {
    __shared__ unsigned sh[PM_HIST_BINS];
    if (gated && state->iterate == 0) return;
    if (state->robust_recompute != want) return;  // the scale is frozen (nbIterationForScale), or not this estimator's turn
    for (int i = threadIdx.x; i < PM_HIST_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const unsigned prefix = pass == 0 ? 0u : state->robust_prefix;
    const float median = state->robust_median;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const float d = __ldg(dists + i);
        if (d == pm_inf()) continue;
        hist_add(sh, phase == 0 ? d : fabsf(__fsub_rn(d, median)), pass, prefix);
    }
    __syncthreads();
    select_flush(sh, hist);
    __syncthreads();
    if (!select_last_block(&state->ticket[0])) return;
    select_pick(hist, pass, 0.5f, 0, 0.f, state, true, want == 2 ? 3 : phase + 1);
}

// scaleEstimator "std": scale = sqrt(Matches::getStandardDeviation()) = sqrt(sqrt(sum (d - mean)^2 / (size - 1))) over ALL
// entries of the distance matrix (Matches.cpp:124-129).  Two reductions: phase 0 the mean, phase 1 the squared deviations
// (terms in float like Eigen's array expression, sums in fp64, block partials added in block order by the last block).
__global__ void __launch_bounds__(HIST_BLOCK) robust_std_kernel(const float* __restrict__ dists, size_t total, int phase, IcpState* state, int gated,
                                                                double* __restrict__ partials) {
    __shared__ double s_warp[HIST_BLOCK / 32];
    if (gated && state->iterate == 0) return;
    if (state->robust_recompute != 3) return;
    const float mean = state->robust_median;  // phase 1: the mean of phase 0
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const float d = __ldg(dists + i);
        if (phase == 0) acc += (double)d;
        else { const float t = __fsub_rn(d, mean); acc += (double)__fmul_rn(t, t); }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int w = 0; w < HIST_BLOCK / 32; ++w) v += s_warp[w];
        partials[blockIdx.x] = v;
    }
    if (!select_last_block(&state->ticket[0])) return;
    if (threadIdx.x != 0) return;
    double sum = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) sum += __ldcg(partials + b);
    if (phase == 0) state->robust_median = (float)(sum / (double)total);
    else state->robust_scale = sqrtf(sqrtf((float)(sum / (double)(total - 1))));
}

// sharded reading: the scan runs after the histograms have been all-reduced
__global__ void __launch_bounds__(1024) pick_kernel(unsigned* __restrict__ hist, int pass, SelectSpec spec, IcpState* state, int gated,
                                                    int cap_active, float cap_margin) {
    if (gated && state->iterate == 0) return;
    const int nq = spec.n_quantile();
    int slot = 0;
    for (int f = 0; f < spec.nfilters; ++f) {
        if (!spec.is_quantile(f)) continue;
        const bool last = slot == nq - 1;
        select_pick(hist + (size_t)(pass == 0 ? 0 : slot) * PM_HIST_BINS, pass, spec.quantile(f), f, spec.factor(f), state, pass != 0 || last);
        ++slot;
    }
    if (pass == 2 && threadIdx.x == 0) select_finish(state, cap_active, cap_margin);
}

// ---- VarTrimmedDistOutlierFilter::optimizeInlierRatio (OutlierFiltersImpl.cpp:177-218) --------------------------
// `sorted`: the bit patterns of all k x N distances in ascending order (zeros first, +inf last).
__device__ __forceinline__ size_t lower_bound_bits(const unsigned* __restrict__ a, size_t n, unsigned key) {
    size_t lo = 0, hi = n;
    while (lo < hi) {
        const size_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}
struct VarRange {
    size_t n_zero, n_finite;  // the qualifying distances (finite, > 0) are sorted[n_zero .. n_finite)
    long long min_el, end;    // FRMS candidates: cum[min_el .. end)
    int points_nbr;
};
__device__ __forceinline__ VarRange var_range(const unsigned* sorted, size_t total, float min_ratio, float max_ratio) {
    VarRange r;
    r.n_zero = lower_bound_bits(sorted, total, 1u);
    r.n_finite = lower_bound_bits(sorted, total, PM_INF_BITS);
    r.points_nbr = (int)total;
    r.min_el = (long long)floorf(__fmul_rn(min_ratio, (float)r.points_nbr));
    const long long max_el = (long long)floorf(__fmul_rn(max_ratio, (float)r.points_nbr));
    const long long count = (long long)(r.n_finite - r.n_zero);
    r.end = max_el < count ? max_el : count;  // past `count` the reference reads uninitialised memory
    return r;
}
// std::partial_sum in float: cum[i] = fl(cum[i - 1] + d[i]).  The rounding of every step depends on the one before, so
// the sum is a serial chain of dependent adds by definition (4 cycles each).  One thread walks it; everything else is
// taken off that thread: the other warps of the block stage the next tile of sorted distances into shared memory and
// write the previous tile of running sums back (coalesced), so the chain thread issues one 16-byte shared load, four
// adds and one 16-byte shared store per four elements.  (A single warp doing loads, shuffles and selects itself was
// issue-bound at ~10 cycles per element: 6.5 ms per 1 M.)
constexpr int VT_TILE = 2048;
constexpr int VT_COPY = 128;               // threads that stage tiles (warps 1..4); warp 0 holds the chain thread
constexpr int VT_THREADS = 32 + VT_COPY;
__global__ void __launch_bounds__(VT_THREADS) vartrim_scan_kernel(const unsigned* __restrict__ sorted, size_t total, float min_ratio, float max_ratio,
                                                                  IcpState* state, int gated, float* __restrict__ cum) {
    __shared__ __align__(16) float s_in[2][VT_TILE];
    __shared__ __align__(16) float s_out[2][VT_TILE];
    if (gated && state->iterate == 0) return;
    if (threadIdx.x == 0) state->var_best = ~0ull;  // the pick kernel behind this one takes its minimum into it
    const VarRange r = var_range(sorted, total, min_ratio, max_ratio);
    if (r.n_finite == r.n_zero) {
        if (threadIdx.x == 0) {
            if (state->status == 0) state->status = PMGPU_ERR_NO_OUTLIER_TO_FILTER;
            state->iterate = 0;
        }
        return;
    }
    const unsigned* __restrict__ src = sorted + r.n_zero;
    const long long ntiles = (r.end + VT_TILE - 1) / VT_TILE;
    const int tid = threadIdx.x;
    // fixed trip counts, fully unrolled: the 16 loads of a thread are in flight together (a rolled loop issues load, dependent
    // shared store, next load ... and pays the memory latency 16 times per tile — more than the chain takes)
    auto stage_in = [&](long long tile, int c) {
        float v[VT_TILE / VT_COPY];
#pragma unroll
        for (int k = 0; k < VT_TILE / VT_COPY; ++k) {
            const long long g = tile * VT_TILE + c + VT_COPY * k;
            v[k] = g < r.end ? __uint_as_float(__ldg(src + g)) : 0.f;  // + 0 leaves the sum as it is
        }
#pragma unroll
        for (int k = 0; k < VT_TILE / VT_COPY; ++k) s_in[tile & 1][c + VT_COPY * k] = v[k];
    };
    auto stage_out = [&](long long tile, int c) {
#pragma unroll
        for (int k = 0; k < VT_TILE / VT_COPY; ++k) {
            const long long g = tile * VT_TILE + c + VT_COPY * k;
            if (g < r.end) cum[g] = s_out[tile & 1][c + VT_COPY * k];
        }
    };
    if (ntiles > 0 && tid >= 32) stage_in(0, tid - 32);
    __syncthreads();
    float acc = 0.f;
    for (long long t = 0; t < ntiles; ++t) {
        if (tid == 0) {
            const float4* in4 = reinterpret_cast<const float4*>(s_in[t & 1]);
            float4* out4 = reinterpret_cast<float4*>(s_out[t & 1]);
            // batches of 32 elements, the next batch's shared loads issued before this batch's adds: left to the compiler, every
            // 16-byte load sat between the store before it and the adds after it, and its latency was on the chain
            constexpr int B = 8;
            float4 cur[B], nxt[B];
#pragma unroll
            for (int k = 0; k < B; ++k) cur[k] = in4[k];
            for (int i = 0; i < VT_TILE / 4; i += B) {
                const int ni = (i + B < VT_TILE / 4) ? i + B : i;  // the last batch re-reads itself, unused
#pragma unroll
                for (int k = 0; k < B; ++k) nxt[k] = in4[ni + k];
#pragma unroll
                for (int k = 0; k < B; ++k) {
                    acc = __fadd_rn(acc, cur[k].x); cur[k].x = acc;
                    acc = __fadd_rn(acc, cur[k].y); cur[k].y = acc;
                    acc = __fadd_rn(acc, cur[k].z); cur[k].z = acc;
                    acc = __fadd_rn(acc, cur[k].w); cur[k].w = acc;
                }
#pragma unroll
                for (int k = 0; k < B; ++k) out4[i + k] = cur[k];
#pragma unroll
                for (int k = 0; k < B; ++k) cur[k] = nxt[k];
            }
        } else if (tid >= 32) {  // warps 1..4: the next tile in, the previous tile out
            if (t + 1 < ntiles) stage_in(t + 1, tid - 32);
            if (t > 0) stage_out(t - 1, tid - 32);
        }
        __syncthreads();
    }
    if (ntiles > 0 && tid >= 32) stage_out(ntiles - 1, tid - 32);
}
// FRMS = trunkSortedDist * ids.inverse() * deno.inverse().square(), per coefficient in float; minCoeff = first minimum;
// then limit = getDistsQuantile(optRatio) read from the sorted array (Matches.cpp:60-87).  pow: Eigen calls powf; the
// double-precision pow rounded to float is the correctly rounded value glibc's powf returns in all but rare cases.
__global__ void __launch_bounds__(1024) vartrim_pick_kernel(const unsigned* __restrict__ sorted, size_t total, float min_ratio, float max_ratio,
                                                            float lambda, int f, IcpState* state, int gated, const float* __restrict__ cum) {
    if (gated && state->iterate == 0) return;
    __shared__ unsigned long long s_best[32];
    const VarRange r = var_range(sorted, total, min_ratio, max_ratio);
    const float n_f = (float)r.points_nbr;
    // FRMS >= 0, so its bit pattern orders like an unsigned integer: the minimum of (bits << 32 | offset) over all candidates
    // is the smallest FRMS and, among equals, the smallest index — Eigen's minCoeff (first minimum)
    unsigned long long best = ~0ull;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = r.min_el + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < r.end; e += stride) {
        const float id = __fadd_rn((float)(r.min_el + 1), __fmul_rn((float)(e - r.min_el), 1.f));  // LinSpaced, step 1
        const float ratio = __fdiv_rn(id, n_f);
        const float deno = (float)pow((double)ratio, (double)lambda);
        const float inv = __fdiv_rn(1.f, deno);
        const float frms = __fmul_rn(__fmul_rn(cum[e], __fdiv_rn(1.f, id)), __fmul_rn(inv, inv));
        const unsigned long long key = ((unsigned long long)__float_as_uint(frms) << 32) | (unsigned long long)(unsigned)(e - r.min_el);
        if (frms == frms && key < best) best = key;  // a NaN is never smaller (minCoeff skips it)
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_down_sync(0xffffffffu, best, o);
        if (other < best) best = other;
    }
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (s_best[w] < best) best = s_best[w];
        if (best != ~0ull) atomicMin(&state->var_best, best);
    }
    if (!select_last_block(&state->ticket[0])) return;
    if (threadIdx.x != 0) return;
    const unsigned long long winner = __ldcg(&state->var_best);
    const long long min_index = winner == ~0ull ? 0 : (long long)(winner & 0xffffffffull);
    const float opt = __fdiv_rn((float)(min_index + r.min_el), n_f);
    state->var_ratio = opt;
    const unsigned long long n_valid = r.n_finite;  // zeros count for the quantile
    if (n_valid == 0 || !(opt >= 0.f && opt <= 1.f)) {
        if (state->status == 0) state->status = n_valid == 0 ? PMGPU_ERR_NO_OUTLIER_TO_FILTER : PMGPU_ERR_BAD_QUANTILE;
        state->iterate = 0;
        return;
    }
    unsigned long long rank = opt == 1.0f ? n_valid - 1 : (unsigned long long)(__ull2float_rn(n_valid) * opt);
    if (rank > n_valid - 1) rank = n_valid - 1;
    const float lim = __uint_as_float(sorted[rank]);
    state->n_valid = n_valid;
    state->limit[f] = lim;
    state->limit_all = fminf(state->limit_all, lim);
}

__global__ void init_limits_kernel(IcpState* state, SelectSpec spec, int gated, int cap_active, float cap_margin) {
    if (gated && state->iterate == 0) return;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        select_init_limits(state, spec);
        select_finish(state, cap_active, cap_margin);
    }
}

__global__ void weights_kernel(const float* __restrict__ dists, const int32_t* __restrict__ ids, int k, size_t total, const IcpState* __restrict__ state,
                               const f4* __restrict__ reading_normals, const f4* __restrict__ ref_normals, float* __restrict__ w,
                               const f4* __restrict__ reading, const f4* __restrict__ ref) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float d = dists[i];
    // empty chain: OutlierFilter.cpp:70-85; otherwise the product of the filters' weights
    float wt = pm_pair_weight(state, d);
    // the reading normal turns with the transform the MATCHES were made with: after a fused iteration T_iter is already
    // the composed one, T_match the one the minimiser used
    if (wt != 0.f && state->sn_on) wt = __fmul_rn(wt, pm_sn_weight(state->T_match, reading_normals[i / k], __ldg(ref_normals + ids[i]), state->sn_eps));
    // Robust distanceType point2plane: the reading point as the matches saw it, the matched point and its normal
    if (wt != 0.f && state->robust_on && state->robust_p2plane)
        wt = __fmul_rn(wt, pm_robust_p2plane_weight(state, transform_point(state->T_match, reading[i / k]), __ldg(ref + ids[i]), __ldg(ref_normals + ids[i])));
    w[i] = wt;
}

}  // namespace

int make_select_spec(pmgpu_ctx* ctx, int nfilters, const int* types, const float* params, SelectSpec* spec) {
    if (nfilters < 0 || nfilters > PM_MAX_FILTERS) {
        ctx->set_error("at most 8 outlier filters are supported in one chain");
        return PMGPU_ERR_BAD_ARG;
    }
    *spec = SelectSpec();
    spec->nfilters = nfilters;
    for (int f = 0; f < nfilters; ++f) {
        spec->type[f] = types[f];
        spec->param[f] = params[f];
        if (types[f] == PMGPU_FILTER_MAXDIST) {
            // maxDist(pow(get<T>("maxDist"), 2)): the square is taken in double (OutlierFiltersImpl.cpp:69)
            spec->param[f] = (float)((double)params[f] * (double)params[f]);
        } else if (types[f] == PMGPU_FILTER_MINDIST) {
            // minDist(pow(get<T>("minDist"), 2)), likewise (OutlierFiltersImpl.cpp:90)
            spec->param[f] = (float)((double)params[f] * (double)params[f]);
        } else if (types[f] == PMGPU_FILTER_TRIMMEDDIST) {
            if (params[f] < 0.f || params[f] > 1.f) {
                ctx->set_error("quantile must be between 0 and 1");
                return PMGPU_ERR_BAD_QUANTILE;
            }
        } else if (types[f] == PMGPU_FILTER_SURFACENORMAL) {
            spec->param[f] = cosf(params[f]);  // eps(cos(maxAngle)) on the host, OutlierFiltersImpl.cpp:227
        } else if ((types[f] & 0xff) == PMGPU_FILTER_ROBUST) {
            const int fct = (types[f] >> 8) & 0xff, est = (types[f] >> 16) & 0xf;
            if (fct > PMGPU_ROBUST_STUDENT) {
                ctx->set_error("Invalid robust function name.");
                return PMGPU_ERR_BAD_ARG;
            }
            if (est > PMGPU_SCALE_STD) {
                ctx->set_error("Invalid scale estimator name.");
                return PMGPU_ERR_BAD_ARG;
            }
            if (types[f] & PMGPU_ROBUST_P2PLANE) {
                if (!ctx->has_normals) {
                    ctx->set_error("Field normals not found");  // getDescriptorViewByName("normals"), OutlierFiltersImpl.cpp:477
                    return PMGPU_ERR_NO_NORMALS;
                }
                if (ctx->dimh != 4) {
                    ctx->set_error("RobustOutlierFilter on GPU: distanceType point2plane needs 3-D clouds");
                    return PMGPU_ERR_UNSUPPORTED;
                }
            }
        } else if (types[f] == PMGPU_FILTER_VARTRIMMEDDIST) {
            // lambda: any value, like the reference's parameter table (OutlierFiltersImpl.h:158)
        } else if (types[f] != PMGPU_FILTER_MEDIANDIST) {
            ctx->set_error("unknown outlier filter type");
            return PMGPU_ERR_BAD_ARG;
        }
    }
    spec->sn_active = (ctx->has_normals && ctx->has_reading_normals) ? 1 : 0;
    spec->robust_approx2 = ctx->robust_approx2;
    int nsn = 0;
    for (int f = 0; f < nfilters; ++f) nsn += spec->kind(f) == PMGPU_FILTER_SURFACENORMAL ? 1 : 0;
    if (nsn > 1) {
        ctx->set_error("at most one SurfaceNormalOutlierFilter per chain on the GPU");
        return PMGPU_ERR_UNSUPPORTED;
    }
    int nrobust = 0;
    for (int f = 0; f < nfilters; ++f) nrobust += spec->is_robust(f) ? 1 : 0;
    if (nrobust > 1) {
        ctx->set_error("at most one RobustOutlierFilter per chain on the GPU");
        return PMGPU_ERR_UNSUPPORTED;
    }
    int nvar = 0;
    for (int f = 0; f < nfilters; ++f) nvar += spec->kind(f) == PMGPU_FILTER_VARTRIMMEDDIST ? 1 : 0;
    if (nvar > 1) {
        ctx->set_error("at most one VarTrimmedDistOutlierFilter per chain on the GPU");
        return PMGPU_ERR_UNSUPPORTED;
    }
    if (nvar && ctx->nranks > 1) {
        ctx->set_error("VarTrimmedDistOutlierFilter is not supported with a sharded reading");
        return PMGPU_ERR_UNSUPPORTED;
    }
    if (nrobust && ctx->nranks > 1) {
        ctx->set_error("RobustOutlierFilter is not supported with a sharded reading");
        return PMGPU_ERR_UNSUPPORTED;
    }
    return PMGPU_OK;
}

int select_reserve(pmgpu_ctx* ctx) {
    if (ctx->hist.cap < (size_t)(PM_MAX_FILTERS + 1) * PM_HIST_BINS) {
        PM_CUDA_TRY(ctx, ctx->hist.reserve((size_t)(PM_MAX_FILTERS + 1) * PM_HIST_BINS));
        PM_CUDA_TRY(ctx, cudaMemsetAsync(ctx->hist.p, 0, (size_t)(PM_MAX_FILTERS + 1) * PM_HIST_BINS * sizeof(unsigned), ctx->stream));
    }
    return PMGPU_OK;
}

// cap_active: the distances come from a capped match (fused ICP loop) — verify the cap and set the next
int launch_weights(pmgpu_ctx* ctx, const SelectSpec& spec, bool gated, bool cap_active) {
    cudaStream_t st = ctx->stream;
    PM_TRY(select_reserve(ctx));
    const int g = gated ? 1 : 0;
    const int nquant = spec.n_quantile();
    const int ca = cap_active ? 1 : 0;
    if (nquant == 0) {
        init_limits_kernel<<<1, 32, 0, st>>>(ctx->state, spec, g, ca, ctx->cap_margin);
        ctx->launches += 1;
    } else {
        const size_t total = (size_t)ctx->k * ctx->nq;
        const size_t want = (total + HIST_BLOCK * 4 - 1) / (HIST_BLOCK * 4);
        const int grid = grid_for((int)(want > 0x1fffff ? 0x1fffff : want) * HIST_BLOCK, HIST_BLOCK, ctx->num_sms, 1);
        // sharded reading: fused peer exchange in the last block (comm.cuh) when the mailboxes are up, else NCCL between
        // the histogram kernel and a one-block pick kernel
        const bool split = ctx->nranks > 1 && !ctx->peer_on;
        for (int pass = 0; pass < 3; ++pass) {
            hist_kernel<<<grid, HIST_BLOCK, 0, st>>>(ctx->dists.p, total, pass, spec, ctx->state, g, pass == 0 ? 1 : 0, split ? 0 : 1, ctx->hist.p, ca,
                                                     ctx->cap_margin, comm_peers(ctx));
            ctx->launches += 1;
            if (split) {
                PM_TRY(comm_allreduce_u32(ctx, ctx->hist.p, (size_t)(pass == 0 ? 1 : nquant) * PM_HIST_BINS));
                pick_kernel<<<1, 1024, 0, st>>>(ctx->hist.p, pass, spec, ctx->state, g, ca, ctx->cap_margin);
                ctx->launches += 1;
            }
        }
    }
    const int r = spec.robust_index();
    const int est = r >= 0 ? ((spec.type[r] >> 16) & 0xf) : PMGPU_SCALE_NONE;
    if (est == PMGPU_SCALE_MAD || est == PMGPU_SCALE_BERG) {
        const size_t total = (size_t)ctx->k * ctx->nq;
        const size_t want = (total + HIST_BLOCK * 4 - 1) / (HIST_BLOCK * 4);
        const int grid = grid_for((int)(want > 0x1fffff ? 0x1fffff : want) * HIST_BLOCK, HIST_BLOCK, ctx->num_sms, 1);
        // mad: median, then median of the absolute deviations; berg: the median alone (used at the filter's first iteration)
        for (int phase = 0; phase < (est == PMGPU_SCALE_MAD ? 2 : 1); ++phase)
            for (int pass = 0; pass < 3; ++pass) {
                robust_hist_kernel<<<grid, HIST_BLOCK, 0, st>>>(ctx->dists.p, total, pass, phase, ctx->state, g, ctx->hist.p + (size_t)PM_MAX_FILTERS * PM_HIST_BINS,
                                                                est == PMGPU_SCALE_MAD ? 1 : 2);
                ctx->launches += 1;
            }
    } else if (est == PMGPU_SCALE_STD) {
        const size_t total = (size_t)ctx->k * ctx->nq;
        PM_CUDA_TRY(ctx, ctx->partials.reserve((size_t)(ctx->num_sms * 4 + 2) * 42));
        for (int phase = 0; phase < 2; ++phase) {
            robust_std_kernel<<<ctx->num_sms, HIST_BLOCK, 0, st>>>(ctx->dists.p, total, phase, ctx->state, g, ctx->partials.p);
            ctx->launches += 1;
        }
    }
    const int v = spec.var_index();
    if (v >= 0) {
        // sort all k x N distance bit patterns (non-negative floats order like unsigned integers), then the serial
        // running sum and the parallel FRMS minimum; runs after the other filters have set their limits
        const size_t total = (size_t)ctx->k * ctx->nq;
        PM_CUDA_TRY(ctx, ctx->var_sorted.reserve(total));
        PM_CUDA_TRY(ctx, ctx->var_cum.reserve(total));
        const unsigned* keys = reinterpret_cast<const unsigned*>(ctx->dists.p);
        size_t tmp_bytes = 0;
        PM_CUDA_TRY(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys, ctx->var_sorted.p, (int)total, 0, 32, st));
        PM_CUDA_TRY(ctx, ctx->cub_tmp.reserve(tmp_bytes));
        size_t tb = ctx->cub_tmp.cap;
        PM_CUDA_TRY(ctx, cub::DeviceRadixSort::SortKeys(ctx->cub_tmp.p, tb, keys, ctx->var_sorted.p, (int)total, 0, 32, st));
        vartrim_scan_kernel<<<1, VT_THREADS, 0, st>>>(ctx->var_sorted.p, total, ctx->var_min_ratio, ctx->var_max_ratio, ctx->state, g, ctx->var_cum.p);
        vartrim_pick_kernel<<<ctx->num_sms, 1024, 0, st>>>(ctx->var_sorted.p, total, ctx->var_min_ratio, ctx->var_max_ratio, spec.param[v], v, ctx->state, g,
                                                ctx->var_cum.p);
        ctx->launches += 3;
    }
    PM_CUDA_TRY(ctx, cudaGetLastError());
    ctx->have_weights = true;
    return PMGPU_OK;
}

int launch_materialize_weights(pmgpu_ctx* ctx) {
    const size_t total = (size_t)ctx->k * ctx->nq;
    PM_CUDA_TRY(ctx, ctx->weights.reserve(total));
    const int B = 256;
    weights_kernel<<<(unsigned)((total + B - 1) / B), B, 0, ctx->stream>>>(ctx->dists.p, ctx->ids.p, ctx->k, total, ctx->state, ctx->reading_normals.p,
                                                                            ctx->ref_normals.p, ctx->weights.p, ctx->reading.p, ctx->ref_orig.p);
    ctx->launches += 1;
    PM_CUDA_TRY(ctx, cudaGetLastError());
    return PMGPU_OK;
}

}  // namespace pm
