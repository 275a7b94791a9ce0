// select.cuh — device-side pieces of the exact radix select (K3) and the "last block" epilogue
// pattern shared by select.cu and minimize.cu.  (Histogramming pass 0 inside the kNN kernel was
// tried and measured slower: zeroing + flushing 2048 shared bins per 128-query block costs more
// than the 7 us pass it saves.)
//
// Squared distances are non-negative floats, so their bit patterns order like unsigned integers:
// pass 0 histograms bits 31..21, pass 1 bits 20..10 inside the selected bucket, pass 2 bits 9..0.
// On one GPU the block that finishes last ("last block", ticket counter + __threadfence) scans
// the 2048 bins and narrows the bucket for every quantile filter, so a pass is ONE kernel.  With
// the reading sharded over GPUs the histograms are all-reduced first and a separate one-block
// kernel does the scan (select.cu).
#pragma once
#include "pmgpu_internal.cuh"

namespace pm {

// MaxDist limits are known up front; quantile filters start at +inf and are lowered by pass 2
__device__ __forceinline__ void select_init_limits(IcpState* st, const SelectSpec& sp) {
    float all = pm_inf();
    for (int f = 0; f < sp.nfilters; ++f) {
        if (sp.type[f] == PMGPU_FILTER_MAXDIST) {
            st->limit[f] = sp.param[f];
            all = fminf(all, sp.param[f]);
        } else {
            st->limit[f] = pm_inf();
        }
    }
    st->limit_all = all;
    st->has_filters = sp.nfilters > 0 ? 1 : 0;
    st->cap_need = 0.f;
    const int sn = sp.sn_index();
    st->sn_on = (sn >= 0 && sp.sn_active) ? 1 : 0;
    if (sn >= 0) { st->sn_eps = sp.param[sn]; st->limit[sn] = sp.param[sn]; }
    // RobustOutlierFilter::robustFiltering preamble (OutlierFiltersImpl.cpp:508-540)
    const int r = sp.robust_index();
    st->robust_on = r >= 0 ? 1 : 0;
    if (r >= 0) {
        const int word = sp.type[r], nb = (word >> 20) & 0xff, est = (word >> 16) & 0xf;
        st->robust_fct = (word >> 8) & 0xff;
        st->robust_k = sp.param[r];
        st->robust_recompute = (est == PMGPU_SCALE_MAD && (st->robust_iteration <= nb || nb == 0)) ? 1 : 0;
        if (est == PMGPU_SCALE_NONE) st->robust_scale = 1.f;
        st->robust_iteration += 1;
    }
}

// Weight of one match: product of the chain's filters (OutlierFilter.cpp:63-103).  The distance
// filters are one threshold; a RobustOutlierFilter multiplies its M-estimator weight
// (OutlierFiltersImpl.cpp:545-583), in float, e^2 = dist / (scale * scale).
__device__ __forceinline__ float pm_robust_weight(const IcpState* st, float d) {
    const float k = st->robust_k, k2 = __fmul_rn(k, k), sc = st->robust_scale;
    const float e2 = __fdiv_rn(d, __fmul_rn(sc, sc));
    float w;
    switch (st->robust_fct) {
        case PMGPU_ROBUST_CAUCHY: w = __fdiv_rn(1.f, __fadd_rn(1.f, __fdiv_rn(e2, k2))); break;
        case PMGPU_ROBUST_WELSCH: w = expf(-__fdiv_rn(e2, k2)); break;
        case PMGPU_ROBUST_SC: { const float a = __fadd_rn(k, e2); w = (e2 >= k) ? __fmul_rn(__fmul_rn(4.0f, k2), __fdiv_rn(1.f, __fmul_rn(a, a))) : 1.f; break; }
        case PMGPU_ROBUST_GM: { const float a = __fadd_rn(k, e2); w = __fmul_rn(k2, __fdiv_rn(1.f, __fmul_rn(a, a))); break; }
        case PMGPU_ROBUST_TUKEY: { const float a = __fsub_rn(1.f, __fdiv_rn(e2, k2)); w = (e2 >= k2) ? 0.f : __fmul_rn(a, a); break; }
        case PMGPU_ROBUST_HUBER: w = (e2 >= k2) ? __fmul_rn(k, __fdiv_rn(1.f, sqrtf(e2))) : 1.f; break;
        case PMGPU_ROBUST_L1: w = __fdiv_rn(1.f, sqrtf(e2)); break;
        default: { const float dd = 3.f; w = powf(1.f + e2 / k, -(k + dd) / 2.f) * (k + dd) * (1.f / (k + e2)); break; }  // Student
    }
    return w <= 1e-50f ? 0.f : w;  // `w <= 1e-50 -> 1e-50` stored in a float array
}
// SurfaceNormalOutlierFilter (OutlierFiltersImpl.cpp:248-265): rn = the reading normal as stored, turned by the
// rotation block of T like `R * inputDesc`; both normals `.normalized()`; weight 0 where |dot| < eps
__device__ __forceinline__ void pm_normalized(float& x, float& y, float& z) {
    const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    if (n2 > 0.f) {
        const float n = __fsqrt_rn(n2);
        x = __fdiv_rn(x, n); y = __fdiv_rn(y, n); z = __fdiv_rn(z, n);
    }
}
__device__ __forceinline__ float pm_sn_weight(const Mat4& T, f4 rn, f4 qn, float eps) {
    float ax = __fadd_rn(__fadd_rn(__fmul_rn(T.m[0], rn.x), __fmul_rn(T.m[4], rn.y)), __fmul_rn(T.m[8], rn.z));
    float ay = __fadd_rn(__fadd_rn(__fmul_rn(T.m[1], rn.x), __fmul_rn(T.m[5], rn.y)), __fmul_rn(T.m[9], rn.z));
    float az = __fadd_rn(__fadd_rn(__fmul_rn(T.m[2], rn.x), __fmul_rn(T.m[6], rn.y)), __fmul_rn(T.m[10], rn.z));
    pm_normalized(ax, ay, az);
    float bx = qn.x, by = qn.y, bz = qn.z;
    pm_normalized(bx, by, bz);
    const float value = fabsf(__fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz)));
    return value < eps ? 0.f : 1.f;
}

__device__ __forceinline__ float pm_pair_weight(const IcpState* st, float d) {
    if (d == pm_inf()) return 0.f;  // ErrorMinimizer.cpp:103-106: invalid matches are skipped
    float w = st->has_filters ? ((d <= st->limit_all) ? 1.f : 0.f) : 1.f;
    if (st->robust_on && w != 0.f) w = __fmul_rn(w, pm_robust_weight(st, d));
    return w;
}

// Capped matching (fused ICP loop).  The matcher of this iteration stopped at squared radius
// state->cap, so every distance above it is only known to be "larger than the cap".  The filter
// chain is exact as long as everything it had to know exactly — each order statistic and each
// limit — lies below that radius; `need` is the largest of them.  If it does not, the iteration is
// void: the minimiser kernels skip it (redo) and the next match runs without a cap.  Otherwise
// the next match may stop at margin x need.  Called by one thread once all limits are final.
__device__ __forceinline__ void select_finish(IcpState* st, int cap_active, float margin) {
    if (!cap_active) {
        st->redo = 0;
        return;
    }
    const float need = st->has_filters ? fmaxf(st->cap_need, st->limit_all) : pm_inf();
    if (st->cap != pm_inf() && !(need < st->cap)) {
        st->cap = pm_inf();
        st->redo = 1;
        st->redo_count += 1;
    } else {
        st->cap = __fmul_rn(need, margin);  // +inf stays +inf
        st->redo = 0;
    }
}

// bin of a distance in `pass`, or -1 when it lies outside the bucket selected so far
__device__ __forceinline__ int select_bin(unsigned bits, int pass, unsigned prefix) {
    if (bits == PM_INF_BITS) return -1;  // finite distances only (Matches.cpp:70)
    if (pass == 0) return (int)(bits >> 21);
    if (pass == 1) return ((bits >> 21) == prefix) ? (int)((bits >> 10) & 0x7ffu) : -1;
    return ((bits >> 10) == prefix) ? (int)(bits & 0x3ffu) : -1;
}

// add the block's shared histogram to the global one
__device__ __forceinline__ void select_flush(const unsigned* sh, unsigned* hist) {
    for (int i = threadIdx.x; i < PM_HIST_BINS; i += blockDim.x) {
        const unsigned v = sh[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// true (for every thread of the block) in the block that arrives last; resets the ticket
__device__ __forceinline__ bool select_last_block(unsigned* ticket) {
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1) ? 1 : 0;
        if (s_last) *ticket = 0;
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

// Scan the 2048 bins with the whole block (blockDim.x a multiple of 32, <= 1024, dividing 2048)
// and pick the bin that holds `rank`.  pass 0 derives the rank from the number of finite
// distances:  rank = size_t(float(n_valid) * quantile)  (Matches.cpp:85-86; quantile == 1 -> max
// element); pass 2 finishes filter f: limit[f] = value (* factor for MedianDist), limit_all = min.
// `clear`: zero the histogram afterwards.
// target: 0 = quantile filter f; 1 = the median of a robust filter's MAD estimate, 2 = its median of
// absolute deviations (both at position size / 2, Matches.cpp:107-120).
__device__ __forceinline__ void select_pick(unsigned* hist, int pass, float quantile, int f, float factor, IcpState* state, bool clear,
                                            int target = 0) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long s_rank;
    __shared__ int s_abort;
    const int t = threadIdx.x;
    const int per = PM_HIST_BINS / blockDim.x;  // consecutive bins per thread
    unsigned long long mine = 0;
    for (int j = 0; j < per; ++j) mine += __ldcg(hist + t * per + j);
    unsigned long long incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((t & 31) >= o) incl += v;
    }
    __syncthreads();  // warp_tot may still be read by a previous call
    if ((t & 31) == 31) warp_tot[t >> 5] = incl;
    __syncthreads();
    const int nwarps = blockDim.x >> 5;
    if (t < 32) {
        unsigned long long w = t < nwarps ? warp_tot[t] : 0ull;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, w, o);
            if (t >= o) w += v;
        }
        warp_tot[t] = w;  // inclusive over warps
    }
    __syncthreads();
    const unsigned long long before_warp = (t >> 5) ? warp_tot[(t >> 5) - 1] : 0ull;
    unsigned long long excl = before_warp + incl - mine;  // elements in bins < t * per
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
                if (target != 0) r = total / 2;
                else if (quantile == 1.0f) r = total - 1;
                else {
                    r = (unsigned long long)(__ull2float_rn(total) * quantile);
                    if (r > total - 1) r = total - 1;
                }
                s_rank = r;
            }
        } else {
            s_rank = target != 0 ? state->robust_rank : state->sel_rank[f];
        }
    }
    __syncthreads();
    if (!s_abort) {
        const unsigned long long rank = s_rank;
        for (int j = 0; j < per; ++j) {
            const unsigned h = __ldcg(hist + t * per + j);
            if (rank >= excl && rank < excl + h) {
                const unsigned found = (unsigned)(t * per + j);
                const unsigned long long rem = rank - excl;
                if (target != 0) {
                    if (pass == 0) { state->robust_prefix = found; state->robust_rank = rem; }
                    else if (pass == 1) { state->robust_prefix = (state->robust_prefix << 11) | found; state->robust_rank = rem; }
                    else {
                        const float value = __uint_as_float((state->robust_prefix << 10) | found);
                        if (target == 1) state->robust_median = value;
                        else state->robust_scale = sqrtf(value);  // scale = sqrt(MAD), OutlierFiltersImpl.cpp:512
                    }
                } else if (pass == 0) {
                    state->sel_prefix[f] = found;
                    state->sel_rank[f] = rem;
                } else if (pass == 1) { state->sel_prefix[f] = (state->sel_prefix[f] << 11) | found; state->sel_rank[f] = rem; }
                else {
                    const unsigned bits = (state->sel_prefix[f] << 10) | found;
                    const float value = __uint_as_float(bits);
                    const float lim = factor != 0.f ? __fmul_rn(factor, value) : value;
                    state->limit[f] = lim;
                    atomicMin(reinterpret_cast<int*>(&state->limit_all), __float_as_int(lim));  // non-negative floats order like ints
                    atomicMax(reinterpret_cast<int*>(&state->cap_need), __float_as_int(fmaxf(value, lim)));
                }
            }
            excl += h;
        }
    }
    __syncthreads();
    if (clear)
        for (int j = 0; j < per; ++j) hist[t * per + j] = 0;
}

}  // namespace pm
