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
    float all = pm_inf(), lo = 0.f;
    for (int f = 0; f < sp.nfilters; ++f) {
        if (sp.type[f] == PMGPU_FILTER_MAXDIST) {
            st->limit[f] = sp.param[f];
            all = fminf(all, sp.param[f]);
        } else if (sp.type[f] == PMGPU_FILTER_MINDIST) {
            st->limit[f] = sp.param[f];
            lo = fmaxf(lo, sp.param[f]);
        } else {
            st->limit[f] = pm_inf();
        }
    }
    st->limit_all = all;
    st->limit_lo = lo;
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
        const int fct = (word >> 8) & 0xff;
        st->robust_fct = fct;
        st->robust_p2plane = (word & PMGPU_ROBUST_P2PLANE) ? 1 : 0;
        st->robust_k = sp.param[r];
        st->robust_approx2 = sp.robust_approx2;
        const bool re = st->robust_iteration <= nb || nb == 0;  // nbIterationForScale
        st->robust_recompute = 0;
        if (est == PMGPU_SCALE_MAD) {
            st->robust_recompute = re ? 1 : 0;
        } else if (est == PMGPU_SCALE_BERG) {
            // OutlierFiltersImpl.cpp:420-432: the tuning given is the target scale, the tuning constant is Bergstrom's
            st->robust_target = sp.param[r];
            if (fct == PMGPU_ROBUST_CAUCHY) st->robust_k = 4.3040f;
            else if (fct == PMGPU_ROBUST_TUKEY) st->robust_k = 7.0589f;
            else if (fct == PMGPU_ROBUST_HUBER) st->robust_k = 2.0138f;
            if (re) {
                if (st->robust_iteration == 1) st->robust_recompute = 2;  // 1.9 sqrt(median): one select, then select_pick
                else st->robust_scale = __fadd_rn(__fmul_rn(0.85f, __fsub_rn(st->robust_scale, st->robust_target)), st->robust_target);
            }
        } else if (est == PMGPU_SCALE_STD) {
            st->robust_recompute = re ? 3 : 0;
        } else {
            st->robust_scale = 1.f;
        }
        st->robust_iteration += 1;
    }
}

// Weight of one match: product of the chain's filters (OutlierFilter.cpp:63-103).  The distance
// filters are one threshold; a RobustOutlierFilter multiplies its M-estimator weight
// (OutlierFiltersImpl.cpp:545-583), in float, e^2 = dist / (scale * scale).
// (S: IcpState, or the PairW copy a kernel makes of the fields below when the state changes under it)
template <typename S>
__device__ __forceinline__ float pm_robust_weight(const S* st, float d) {
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
    if (e2 >= st->robust_approx2 && st->robust_approx2 != pm_inf()) return 0.f;  // `approximation`, OutlierFiltersImpl.cpp:591-595
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

// what the weight of a match depends on, copied out of the IcpState with coherent loads: select_accumulate_kernel
// (minimize.cu) computes the limits in the same launch that applies them, so its blocks must not read them through a
// const __restrict__ pointer (non-coherent loads, hoistable above the grid barrier)
struct PairW {
    float limit_all, limit_lo;
    int has_filters, robust_on, robust_fct;
    float robust_k, robust_scale, robust_approx2;
    int robust_p2plane;
    int sn_on;
    float sn_eps;
};
__device__ __forceinline__ void load_pairw(const IcpState* st, PairW* w) {
    w->limit_all = __ldcg(&st->limit_all);
    w->limit_lo = __ldcg(&st->limit_lo);
    w->has_filters = __ldcg(&st->has_filters);
    w->robust_on = __ldcg(&st->robust_on);
    w->robust_fct = __ldcg(&st->robust_fct);
    w->robust_k = __ldcg(&st->robust_k);
    w->robust_scale = __ldcg(&st->robust_scale);
    w->robust_approx2 = __ldcg(&st->robust_approx2);
    w->robust_p2plane = __ldcg(&st->robust_p2plane);
    w->sn_on = __ldcg(&st->sn_on);
    w->sn_eps = __ldcg(&st->sn_eps);
}

template <typename S>
__device__ __forceinline__ float pm_pair_weight(const S* st, float d) {
    if (d == pm_inf()) return 0.f;  // ErrorMinimizer.cpp:103-106: invalid matches are skipped
    float w = st->has_filters ? ((d <= st->limit_all && d >= st->limit_lo) ? 1.f : 0.f) : 1.f;
    // (distanceType point2plane: the robust factor needs the matched point and its normal — pm_robust_p2plane_weight, applied
    // by the caller once it has gathered them)
    if (st->robust_on && !st->robust_p2plane && w != 0.f) w = __fmul_rn(w, pm_robust_weight(st, d));
    return w;
}
// RobustOutlierFilter::computePointToPlaneDistance (OutlierFiltersImpl.cpp:468-500): dot(n.normalized(), p - q)^2 in float
template <typename S>
__device__ __forceinline__ float pm_robust_p2plane_weight(const S* st, const f4& p, const f4& q, const f4& n) {
    float nx = n.x, ny = n.y, nz = n.z;
    pm_normalized(nx, ny, nz);
    const float dot = __fadd_rn(__fadd_rn(__fmul_rn(nx, __fsub_rn(p.x, q.x)), __fmul_rn(ny, __fsub_rn(p.y, q.y))), __fmul_rn(nz, __fsub_rn(p.z, q.z)));
    return pm_robust_weight(st, __fmul_rn(dot, dot));
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
    // all five values requested before the first is used: one round trip, not one per branch
    const int has_filters = st->has_filters, redo_count = st->redo_count;
    const float cap_need = st->cap_need, limit_all = st->limit_all, cap = st->cap;
    const float need = has_filters ? fmaxf(cap_need, limit_all) : pm_inf();
    if (cap != pm_inf() && !(need < cap)) {
        st->cap = pm_inf();
        st->redo = 1;
        st->redo_count = redo_count + 1;
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
// absolute deviations (both at position size / 2, Matches.cpp:107-120); 3 = the median of the berg estimator's first
// iteration (getDistsQuantile(0.5)).
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
                if (target == 1 || target == 2) r = total / 2;  // Matches.cpp:107-120; target 3 (berg) is getDistsQuantile(0.5)
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
                        else if (target == 3) state->robust_scale = (float)(1.9 * (double)sqrtf(value));  // berg, OutlierFiltersImpl.cpp:529
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

// ---- the select as a phase of another kernel (minimize.cu select_accumulate_kernel) -----------------------------------------
// A histogram pass is described by a plan: bin = (bits - lo) >> shift, valid for lo <= bits < lo + (nb << shift).  The generic
// select is the plan sequence (0, 21, 2048) -> (bucket, 10, 2048) -> (bucket, 0, 1024) of the three-pass kernels above.  Between
// ICP iterations the order statistic moves little, so the first plan of an iteration is a WINDOW of 2046 bins of 2^11 bit
// patterns centred on the previous value (about +-25 % of it), with the finite distances below / above the window counted in
// bins 2046 / 2047.  The same pass also COLLECTS the distances that fall into the few bins around the centre (the inner window,
// as wide as the value moved last time, at most PM_SEL_CAND_CAP distances): if the rank lands in one of those bins — verified
// against the histogram, never assumed — the block that picks finds the exact value among the collected distances and the
// select is ONE pass; if it lands elsewhere in the window, one more pass over the 2^11 patterns of its bin finishes it; if it
// left the window, the generic sequence runs.  Exact in every case.
struct SelPlan { unsigned lo; int shift, nb, outside, c0, c1; };  // [c0, c1): bins whose distances are collected (window plan)
#define PM_SEL_WINDOW_BINS 2046
#define PM_SEL_CENTRE (PM_SEL_WINDOW_BINS / 2)
#define PM_SEL_CAND_CAP 16384   // collected distances per filter (global scratch)
#define PM_SEL_BLOCK_CAND 512   // ... per block and pass (shared staging)
__device__ __forceinline__ SelPlan select_first_plan(unsigned guess, int inner) {
    SelPlan p;
    p.c0 = p.c1 = 0;
    if (guess == 0u) { p.lo = 0u; p.shift = 21; p.nb = PM_HIST_BINS; p.outside = 0; return p; }
    const unsigned half = (unsigned)PM_SEL_CENTRE << 11;
    p.lo = guess > half ? guess - half : 0u;
    p.shift = 11; p.nb = PM_SEL_WINDOW_BINS; p.outside = 1;
    if (inner > 0) {
        const int centre = (int)((guess - p.lo) >> 11);
        p.c0 = centre - inner < 0 ? 0 : centre - inner;
        p.c1 = centre + inner + 1 > PM_SEL_WINDOW_BINS ? PM_SEL_WINDOW_BINS : centre + inner + 1;
    }
    return p;
}

struct SelScratch {
    unsigned hist[PM_HIST_BINS];
    unsigned cand[PM_SEL_BLOCK_CAND];
    unsigned ncand;
};

// one block's share of one pass: shared histogram, then added to the global slot; collected distances appended to `cand`
__device__ __forceinline__ void select_plan_add(SelScratch* sh, const SelPlan& p, float d, unsigned& below, unsigned& above) {
    const unsigned bits = __float_as_uint(d);
    if (bits == PM_INF_BITS) return;  // finite distances only (Matches.cpp:70)
    if (bits < p.lo) { ++below; return; }
    const unsigned off = (bits - p.lo) >> p.shift;
    if (off < (unsigned)p.nb) {
        atomicAdd(&sh->hist[off], 1u);
        if ((int)off >= p.c0 && (int)off < p.c1) {
            const unsigned pos = atomicAdd(&sh->ncand, 1u);
            if (pos < PM_SEL_BLOCK_CAND) sh->cand[pos] = bits;
        }
    } else {
        ++above;
    }
}
__device__ __forceinline__ void select_pass_block(const float* __restrict__ dists, size_t total, const SelPlan p, SelScratch* sh, unsigned* hist_slot,
                                                  unsigned* cand, unsigned* cand_count) {
    for (int i = threadIdx.x; i < PM_HIST_BINS; i += blockDim.x) sh->hist[i] = 0;
    if (threadIdx.x == 0) sh->ncand = 0;
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t quads = total / 4;
    const float4* __restrict__ d4 = reinterpret_cast<const float4*>(dists);
    unsigned below = 0, above = 0;
    // four independent 16-byte loads in flight per thread: at 1 M distances that is the thread's whole share, one round trip
    for (size_t i0 = gtid; i0 < quads; i0 += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + u * stride < quads) v[u] = __ldg(d4 + i0 + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + u * stride < quads) {
                select_plan_add(sh, p, v[u].x, below, above);
                select_plan_add(sh, p, v[u].y, below, above);
                select_plan_add(sh, p, v[u].z, below, above);
                select_plan_add(sh, p, v[u].w, below, above);
            }
    }
    if (gtid < total - 4 * quads) select_plan_add(sh, p, __ldg(dists + 4 * quads + gtid), below, above);
    if (p.outside) {
        below = __reduce_add_sync(0xffffffffu, below);
        above = __reduce_add_sync(0xffffffffu, above);
        if ((threadIdx.x & 31) == 0) {
            if (below) atomicAdd(&sh->hist[PM_SEL_WINDOW_BINS], below);
            if (above) atomicAdd(&sh->hist[PM_SEL_WINDOW_BINS + 1], above);
        }
    }
    __syncthreads();
    select_flush(sh->hist, hist_slot);
    if (p.c1 > p.c0) {
        // the block's collected distances: one reservation in the global list.  A block that saw more than it can stage
        // poisons the count, so the picking block falls back to the second pass
        __shared__ unsigned s_base;
        const unsigned n = sh->ncand;
        if (threadIdx.x == 0 && n) s_base = atomicAdd(cand_count, n > PM_SEL_BLOCK_CAND ? (unsigned)PM_SEL_CAND_CAP + 1u : n);
        __syncthreads();
        if (n && n <= PM_SEL_BLOCK_CAND)
            for (unsigned i = threadIdx.x; i < n; i += blockDim.x)
                if (s_base + i < PM_SEL_CAND_CAP) cand[s_base + i] = sh->cand[i];
    }
    __syncthreads();
}

// the plans of a fresh select (one thread)
__device__ __forceinline__ void select_plans_begin(IcpState* st, const SelectSpec& sp, int collect) {
    int pending = 0;
    for (int f = 0; f < sp.nfilters; ++f) {
        if (!sp.is_quantile(f)) continue;
        const SelPlan p = select_first_plan(st->sel_guess[f], collect ? st->sel_inner[f] : 0);
        st->sel_prefix[f] = p.lo; st->sel_shift[f] = p.shift; st->sel_nb[f] = p.nb; st->sel_outside[f] = p.outside;
        st->sel_c0[f] = p.c0; st->sel_c1[f] = p.c1;
        st->sel_have_rank[f] = 0; st->sel_done[f] = 0;
        ++pending;
    }
    st->sel_pending = pending;
    st->sel_passes = 0;
}

// Whole block, two steps: (1) every thread reads its PM_HIST_BINS / blockDim.x consecutive bins of `bins[0 .. nb)` (counts,
// global or shared) and the block scans them — exclusive prefix per thread, total in *total; (2) once the rank is known, the
// thread whose bins hold it reports (bin, count of that bin, rank inside it).
struct SelScan { unsigned h[8]; unsigned long long excl; };  // blockDim.x >= 256
struct SelLocate { int bin; unsigned count; unsigned long long rem, total; };
template <bool SHARED>
__device__ __forceinline__ void select_scan(const unsigned* bins, int nb, SelScan& sc, SelLocate* out, unsigned long long* warp_tot) {
    const int t = threadIdx.x;
    const int per = PM_HIST_BINS / blockDim.x;
    unsigned long long mine = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int bin = t * per + j;
        sc.h[j] = (j < per && bin < nb) ? (SHARED ? bins[bin] : __ldcg(bins + bin)) : 0u;
        mine += sc.h[j];
    }
    unsigned long long incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((t & 31) >= o) incl += v;
    }
    __syncthreads();  // warp_tot / out may still be read by a previous call
    if ((t & 31) == 31) warp_tot[t >> 5] = incl;
    if (t == 0) out->bin = -1;
    __syncthreads();
    const int nwarps = blockDim.x >> 5;
    if (t < 32) {
        unsigned long long w = t < nwarps ? warp_tot[t] : 0ull;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, w, o);
            if (t >= o) w += v;
        }
        warp_tot[t] = w;
    }
    __syncthreads();
    sc.excl = ((t >> 5) ? warp_tot[(t >> 5) - 1] : 0ull) + incl - mine;
    if (t == 0) out->total = warp_tot[31];
    __syncthreads();
}
__device__ __forceinline__ void select_find(const SelScan& sc, unsigned long long rank, SelLocate* out) {
    const int per = PM_HIST_BINS / blockDim.x;
    unsigned long long excl = sc.excl;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (rank >= excl && rank < excl + sc.h[j]) { out->bin = threadIdx.x * per + j; out->count = sc.h[j]; out->rem = rank - excl; }
        excl += sc.h[j];
    }
    __syncthreads();
}

// one thread: filter f's order statistic is `bits` — limit[f], limit_all, cap_need exactly as select_pick's last pass — and
// what the next iteration's first pass is centred on
// (`pre`: limit_all, cap_need, sel_prev[f], sel_pending as the picking thread read them on entry — one round trip to the L2
// for all four instead of one each, on the critical path of every iteration)
struct SelFinishPre { float limit_all, cap_need; unsigned prev; int pending; };
__device__ __forceinline__ void select_finish_filter(IcpState* st, int f, float factor, unsigned bits, int inner_next, const SelFinishPre& pre) {
    const float value = __uint_as_float(bits);
    const float lim = factor != 0.f ? __fmul_rn(factor, value) : value;
    st->limit[f] = lim;
    st->limit_all = fminf(pre.limit_all, lim);  // one thread at a time: the filters are picked in turn
    st->cap_need = fmaxf(pre.cap_need, fmaxf(value, lim));
    // the next iteration's window is centred where the value is heading: this value plus its last move (in bit patterns, i.e.
    // roughly geometric — the limit of a converging registration shrinks by a few per cent per iteration, hundreds of window
    // bins, but steadily: extrapolated, the first pass lands within a few dozen bins)
    const unsigned prev = pre.prev;
    long long next = (long long)bits;
    if (prev != 0u) next += (long long)bits - (long long)prev;
    if (next < 1) next = 1;
    if (next >= (long long)PM_INF_BITS) next = (long long)PM_INF_BITS - 1;
    st->sel_prev[f] = bits;
    st->sel_guess[f] = (unsigned)next;
    if (inner_next >= 0) st->sel_inner[f] = inner_next;  // < 0: set by the window pass before this one, or unknown (kept)
    st->sel_done[f] = 1;
    st->sel_pending = pre.pending - 1;
}
// half-width (in window bins) of the next inner window: twice the move just seen, as far as the collected list can hold
__device__ __forceinline__ int select_next_inner(int moved_bins, unsigned bin_count, int half_now) {
    int half = 3 * moved_bins + 64;
    if (half < (3 * half_now) / 4) half = (3 * half_now) / 4;  // narrows slowly: a miss costs a whole pass, a wide list little
    const int fit = (int)(PM_SEL_CAND_CAP / (3u * (bin_count + 1u)));  // bins of this density that fit (with a margin), half each side
    if (half > fit) half = fit;
    if (half > PM_SEL_CENTRE) half = PM_SEL_CENTRE;
    return half < 1 ? 1 : half;
}

// Whole block: advance filter f by the pass whose (all-rank) histogram sits in `hist` — next plan, or the final value.
// Clears the slot and the collected list.  `sh`: the block's scratch (free between passes).
__device__ __forceinline__ void select_pick_plan(unsigned* hist, int f, float quantile, float factor, IcpState* st, SelScratch* sh, const unsigned* cand,
                                                 unsigned* cand_count) {
    __shared__ unsigned long long warp_tot[32];
    __shared__ SelLocate s_loc;
    __shared__ unsigned long long s_rank;
    __shared__ int s_mode;  // 0 find the bin, 1 window missed (generic plan next), 2 nothing to select from
    const int t = threadIdx.x;
    const unsigned lo = st->sel_prefix[f];
    const int shift = st->sel_shift[f], nb = st->sel_nb[f], outside = st->sel_outside[f], c0 = st->sel_c0[f], c1 = st->sel_c1[f];
    const int have_rank = st->sel_have_rank[f];
    const unsigned long long rank_in = st->sel_rank[f];
    const unsigned long long below = outside ? __ldcg(hist + PM_SEL_WINDOW_BINS) : 0ull;
    const unsigned long long above = outside ? __ldcg(hist + PM_SEL_WINDOW_BINS + 1) : 0ull;
    const unsigned ncand = c1 > c0 ? __ldcg(cand_count) : 0u;
    // requested now, used at the end: what the finishing thread needs, and the first batch of collected distances (the list's
    // length is not known yet; the buffer is)
    SelFinishPre pre = {0.f, 0.f, 0u, 0};
    if (t == 0) { pre.limit_all = st->limit_all; pre.cap_need = st->cap_need; pre.prev = st->sel_prev[f]; pre.pending = st->sel_pending; }
    unsigned c_first[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) c_first[u] = c1 > c0 ? __ldcg(cand + (unsigned)u * blockDim.x + t) : 0u;
    SelScan sc;
    select_scan<false>(hist, nb, sc, &s_loc, warp_tot);
    if (t == 0) {
        s_mode = 0;
        s_rank = rank_in;
        if (!have_rank) {
            // first pass of this filter: the number of finite distances and, from it, the rank
            const unsigned long long inside = s_loc.total, total = below + inside + above;
            st->n_valid = total;
            if (total == 0) {
                if (st->status == 0) st->status = PMGPU_ERR_NO_OUTLIER_TO_FILTER;
                st->iterate = 0;
                st->sel_pending = 0;
                s_mode = 2;
            } else {
                unsigned long long r;  // Matches.cpp:85-86
                if (quantile == 1.0f) r = total - 1;
                else {
                    r = (unsigned long long)(__ull2float_rn(total) * quantile);
                    if (r > total - 1) r = total - 1;
                }
                if (r < below || r >= below + inside) {
                    // the order statistic left the window: the generic sequence from the top
                    st->sel_prefix[f] = 0u; st->sel_shift[f] = 21; st->sel_nb[f] = PM_HIST_BINS; st->sel_outside[f] = 0;
                    st->sel_c0[f] = 0; st->sel_c1[f] = 0;
                    s_mode = 1;
                } else {
                    st->sel_have_rank[f] = 1;
                    s_rank = r - below;
                }
            }
        }
    }
    __syncthreads();
    if (s_mode == 0) {
        select_find(sc, s_rank, &s_loc);
        const int bin = s_loc.bin;
        const unsigned nlo = lo + ((unsigned)bin << shift);
        const int moved = bin > PM_SEL_CENTRE ? bin - PM_SEL_CENTRE : PM_SEL_CENTRE - bin;
        const int inner_next = outside ? select_next_inner(moved, s_loc.count, (c1 - c0) / 2) : -1;
        if (shift == 0) {
            if (t == 0) select_finish_filter(st, f, factor, nlo, -1, pre);
        } else if (bin >= c0 && bin < c1 && ncand <= PM_SEL_CAND_CAP) {
            // the bin was collected: its distances, by their low 11 bits, into the shared histogram -> the exact value now.
            // (ncand <= cap also says no block overflowed its staging, so the list holds every distance of bins [c0, c1).)
            const unsigned long long rem = s_loc.rem;
            for (int i = t; i < PM_HIST_BINS; i += blockDim.x) sh->hist[i] = 0;
            __syncthreads();
            for (unsigned base = 0; base < ncand; base += 8u * blockDim.x) {  // eight loads in flight per thread
                unsigned c[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const unsigned i = base + (unsigned)u * blockDim.x + t;
                    c[u] = i < ncand ? (base == 0 ? c_first[u] : __ldcg(cand + i)) : 0xffffffffu;  // not the bit pattern of a distance
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const unsigned d = c[u] - lo;
                    if (c[u] != 0xffffffffu && (int)(d >> 11) == bin) atomicAdd(&sh->hist[d & 2047u], 1u);
                }
            }
            __syncthreads();
            select_scan<true>(sh->hist, PM_HIST_BINS, sc, &s_loc, warp_tot);
            select_find(sc, rem, &s_loc);
            if (t == 0) select_finish_filter(st, f, factor, nlo + (unsigned)s_loc.bin, inner_next, pre);
        } else if (t == 0) {
            const int nshift = shift > 11 ? shift - 11 : 0;
            st->sel_prefix[f] = nlo;
            st->sel_shift[f] = nshift;
            st->sel_nb[f] = 1 << (shift - nshift);
            st->sel_outside[f] = 0;
            st->sel_c0[f] = 0; st->sel_c1[f] = 0;
            st->sel_rank[f] = s_loc.rem;
            if (inner_next >= 0) st->sel_inner[f] = inner_next;
        }
    }
    __syncthreads();
    const int per = PM_HIST_BINS / blockDim.x;
    for (int j = 0; j < per; ++j) hist[t * per + j] = 0;
    if (t == 0 && c1 > c0) *cand_count = 0;
}

// ---- grid barrier of a kernel whose blocks are all resident (cooperative launch) -----------------------------------------
// arrive: true in the block that arrives last (every thread of it) — it does the serial work and then releases the others.
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ bool grid_bar_arrive(IcpState* st, unsigned& gen) {
    __shared__ int s_last;
    __shared__ unsigned s_gen;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_gen = ld_acquire_gpu(&st->bar_gen);  // read before arriving: the release of THIS barrier cannot have happened yet
        const unsigned t = atomicAdd(&st->bar_count, 1u);
        s_last = (t == gridDim.x - 1) ? 1 : 0;
        if (s_last) atomicExch(&st->bar_count, 0u);  // everybody has arrived; nobody arrives again before the release
    }
    __syncthreads();
    gen = s_gen;
    if (s_last) __threadfence();
    return s_last != 0;
}
__device__ __forceinline__ void grid_bar_release(IcpState* st, unsigned gen) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(&st->bar_gen, gen + 1u);
}
__device__ __forceinline__ void grid_bar_wait(IcpState* st, unsigned gen) {
    if (threadIdx.x == 0)
        while (ld_acquire_gpu(&st->bar_gen) == gen) {}
    __syncthreads();
}

}  // namespace pm
