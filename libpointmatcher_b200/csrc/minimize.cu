// minimize.cu — K4-K7: ErrorMinimizer::compute without ever building ErrorElements.
//
// Reference path per iteration (ErrorMinimizer.cpp:58-193, then PointToPlane.cpp:171-312 or
// PointToPoint.cpp:61-101): compact the kept pairs into new matrices, gather the matched reference
// columns, build F / wF / deltas temporaries, run a 6xM.Mx6 GEMM, solve.  Here one streaming
// kernel re-derives the transformed reading point, reads (id, dist), applies the collapsed outlier
// threshold (select.cu), gathers the matched reference point (+ normal) and accumulates the
// normal-equation sums in fp64 registers; warp shuffles -> shared memory -> one partial row per
// block; a one-block kernel reduces the rows in a fixed order (deterministic), solves, converts to
// a 4x4, composes T_iter <- dT * T_iter and evaluates the Counter / Differential checkers, all on
// the device.  Per-pair terms (cross product, residual dot product, w*F) are computed in float
// with the reference's operation order; only the long sums are carried in fp64.
//
// Algorithmic bytes per match: reading 16 + id 4 + dist 4 + reference gather 16 (+ normal gather
// 16 for point-to-plane) = 40 / 56 B.
#include "comm.cuh"
#include "core/linalg.h"
#include "select.cuh"

namespace pm {

namespace {

// columns of a partial row
//  point-to-plane : [0..20] upper triangle of A (row-major over i <= j), [21..26] sum wF*dot
//  point-to-point : [0] sum w, [1..3] sum w p, [4..6] sum w q, [7..15] sum (w q_r) p_c  (r + 3 c)
//  both           : [NS-4] kept pairs, [NS-3] rejected matches, [NS-2] rejected points, [NS-1] points seen
constexpr int NS_PLANE = 28 + 4;  // 21 A + 6 b + sum of weights, then the 4 counters
constexpr int NS_POINT = 17 + 4;  // W, 3 + 3 weighted sums, 9 cross terms, sum w |p|^2, then the 4 counters
constexpr int NS_COV = 42;
constexpr int NS_MAX = 42;
constexpr int ACC_BLOCK = 256;

template <int NS>
__device__ __forceinline__ void block_reduce_store(double* acc, double* __restrict__ out_row) {
    __shared__ double sh[ACC_BLOCK / 32][NS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < NS; ++c) {
        double v = acc[c];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sh[warp][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        double v = 0.0;
        for (int w = 0; w < ACC_BLOCK / 32; ++w) v += sh[w][threadIdx.x];
        out_row[threadIdx.x] = v;
    }
}

#ifdef PM_PROFILE_NS
__device__ __forceinline__ unsigned long long pm_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#endif

template <int MODE>
__device__ void finalize_body(const double* partials, int nblocks, double* sums, int phase, IcpState* state, int compose,
                              const pmgpu_icp_params& ck, const PeerComm& pc);

// The streaming part of K4 / K5: this thread's share of the (reading point, match) pairs into its fp64 sums.
// W: where the weight of a match is read from — the IcpState itself, or the PairW copy of a kernel that computed the limits
// in the same launch (select_accumulate_kernel).
template <int MODE, typename W>
__device__ __forceinline__ void accumulate_pairs(double* acc, const Mat4& sT, const W* wst, const f4* __restrict__ reading, int nq, int k,
                                                 const int32_t* __restrict__ ids, const float* __restrict__ dists, const f4* __restrict__ ref,
                                                 const f4* __restrict__ normals, const pmgpu_icp_params& ck, const f4* __restrict__ reading_normals) {
    constexpr int NS = MODE == 1 ? NS_PLANE : NS_POINT;
    const int stride = gridDim.x * blockDim.x;
    // A (reading point, match) pair in two steps, so that the gathers of SEVERAL pairs are in flight before the first one is
    // consumed — the kernel waits on 16-byte random reads of the matched reference point and normal, nothing else:
    //   pair_begin : weight from the distance alone (limits, M-estimator); where it is not zero, the gathers are issued
    //   pair_end   : SurfaceNormal factor, then the sums
    struct Pend { float w; f4 q, n; bool live; };
    int kept = 0, rej_matches = 0, rej_points = 0, seen = 0;  // the four counters, integers until the end
    auto pair_begin = [&](Pend& e, const float d, const int id) {
        e.live = d != pm_inf();  // an invalid match is not even counted as rejected (ErrorMinimizer.cpp:103-106)
        e.w = 0.f;
        if (!e.live) return;
        e.w = pm_pair_weight(wst, d);
        if (e.w != 0.f) {
            e.q = __ldg(ref + id);
            // SurfaceNormalOutlierFilter / Robust point2plane: the reference's descriptor whenever it exists
            if (MODE == 1 || wst->sn_on || (wst->robust_on && wst->robust_p2plane)) e.n = __ldg(normals + id);
        }
    };
    auto pair_end = [&](const Pend& e, const f4& p, const int i, bool& match_exist) {
        if (!e.live) return;
        float w = e.w;
        if (w != 0.f && wst->sn_on) w = __fmul_rn(w, pm_sn_weight(sT, reading_normals[i], e.n, wst->sn_eps));
        if (w != 0.f && wst->robust_on && wst->robust_p2plane) w = __fmul_rn(w, pm_robust_p2plane_weight(wst, p, e.q, e.n));
        if (w == 0.f) { ++rej_matches; return; }
        match_exist = true;
        ++kept;
        const f4 q = e.q;
        if (MODE == 1) {
            const f4 n = e.n;
            float F[6], wF[6];
            F[0] = fsub(fmul(p.y, n.z), fmul(p.z, n.y));  // crossProduct, ErrorMinimizer.cpp:304-306
            F[1] = fsub(fmul(p.z, n.x), fmul(p.x, n.z));
            F[2] = fsub(fmul(p.x, n.y), fmul(p.y, n.x));
            F[3] = n.x; F[4] = n.y; F[5] = n.z;
#pragma unroll
            for (int a = 0; a < 6; ++a) wF[a] = fmul(w, F[a]);
            // dot(deltas, normals) accumulated from zero, PointToPlane.cpp:233-240
            float dot = fmul(fsub(p.x, q.x), n.x);
            dot = fadd(dot, fmul(fsub(p.y, q.y), n.y));
            if (!(ck.minimizer & PMGPU_MIN_FORCE2D)) dot = fadd(dot, fmul(fsub(p.z, q.z), n.z));  // force2D: clouds are [x, y, 1]
            int c = 0;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = a; b < 6; ++b) acc[c++] += (double)wF[a] * (double)F[b];
#pragma unroll
            for (int a = 0; a < 6; ++a) acc[21 + a] += (double)wF[a] * (double)dot;
            acc[27] += (double)w;
        } else {
            acc[0] += (double)w;
            const float wp[3] = {fmul(p.x, w), fmul(p.y, w), fmul(p.z, w)};
            const float wq[3] = {fmul(q.x, w), fmul(q.y, w), fmul(q.z, w)};
            const float pc[3] = {p.x, p.y, p.z};
#pragma unroll
            for (int a = 0; a < 3; ++a) { acc[1 + a] += (double)wp[a]; acc[4 + a] += (double)wq[a]; }
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int r = 0; r < 3; ++r) acc[7 + r + 3 * c] += (double)wq[r] * (double)pc[c];
            acc[16] += (double)w * ((double)p.x * (double)p.x + (double)p.y * (double)p.y + (double)p.z * (double)p.z);
        }
    };
    if (k == 1) {
        // the common case: two reading points per trip — both gathers issued before either is consumed — and the
        // (point, distance, id) of the next two already requested
        int i = blockIdx.x * blockDim.x + threadIdx.x;
        f4 rp[2];
        float d[2];
        int id[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            rp[u] = make_float4(0.f, 0.f, 0.f, 0.f); d[u] = pm_inf(); id[u] = 0;
            if (i + u * stride < nq) { rp[u] = reading[i + u * stride]; d[u] = dists[i + u * stride]; id[u] = ids[i + u * stride]; }
        }
        while (i < nq) {
            Pend e[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) pair_begin(e[u], d[u], id[u]);
            const int inext = i + 2 * stride;
            f4 rp_n[2];
            float d_n[2];
            int id_n[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                rp_n[u] = rp[u]; d_n[u] = pm_inf(); id_n[u] = 0;
                if (inext + u * stride < nq) { rp_n[u] = reading[inext + u * stride]; d_n[u] = dists[inext + u * stride]; id_n[u] = ids[inext + u * stride]; }
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (i + u * stride < nq) {
                    const f4 p = transform_point(sT, rp[u]);
                    bool match_exist = false;
                    pair_end(e[u], p, i + u * stride, match_exist);
                    if (!match_exist) ++rej_points;
                    ++seen;
                }
            }
            i = inext;
#pragma unroll
            for (int u = 0; u < 2; ++u) { rp[u] = rp_n[u]; d[u] = d_n[u]; id[u] = id_n[u]; }
        }
    } else {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride) {
            const f4 p = transform_point(sT, reading[i]);
            bool match_exist = false;
            for (int kk = 0; kk < k; kk += 2) {
                Pend e[2];
                const bool two = kk + 1 < k;
                pair_begin(e[0], dists[(size_t)i * k + kk], ids[(size_t)i * k + kk]);
                if (two) pair_begin(e[1], dists[(size_t)i * k + kk + 1], ids[(size_t)i * k + kk + 1]);
                pair_end(e[0], p, i, match_exist);
                if (two) pair_end(e[1], p, i, match_exist);
            }
            if (!match_exist) ++rej_points;
            ++seen;
        }
    }
    acc[NS - 4] = (double)kept; acc[NS - 3] = (double)rej_matches; acc[NS - 2] = (double)rej_points; acc[NS - 1] = (double)seen;
}

// `fuse`: the last block to finish reduces the partial rows, solves, composes T_iter and runs the
// checkers (finalize_body), so the whole minimisation is one kernel
template <int MODE>  // 0 point-to-point, 1 point-to-plane
__global__ void __launch_bounds__(ACC_BLOCK, MODE == 1 ? 2 : 3) accumulate_kernel(const f4* __restrict__ reading, int nq, int k, const int32_t* __restrict__ ids,
                                                               const float* __restrict__ dists, const f4* __restrict__ ref,
                                                               const f4* __restrict__ normals, const IcpState* __restrict__ state, int gated,
                                                               double* __restrict__ partials, int fuse, IcpState* state_rw, double* sums,
                                                               int compose, pmgpu_icp_params ck, const f4* __restrict__ reading_normals, PeerComm pc) {
    constexpr int NS = MODE == 1 ? NS_PLANE : NS_POINT;
    __shared__ Mat4 sT;
    // redo: this iteration's capped match was void (select_finish) — leave T_iter as it is
    if (gated && (state->iterate == 0 || state->redo)) return;
    if (threadIdx.x < 16) sT.m[threadIdx.x] = state->T_iter.m[threadIdx.x];
    __syncthreads();
    double acc[NS];
#pragma unroll
    for (int c = 0; c < NS; ++c) acc[c] = 0.0;
    accumulate_pairs<MODE>(acc, sT, state, reading, nq, k, ids, dists, ref, normals, ck, reading_normals);
#ifdef PM_PROFILE_NS
    __shared__ unsigned long long s_t[3];
    if (threadIdx.x == 0) s_t[0] = pm_globaltimer();
#endif
    block_reduce_store<NS>(acc, partials + (size_t)blockIdx.x * NS_MAX);
#ifdef PM_PROFILE_NS
    if (threadIdx.x == 0) s_t[1] = pm_globaltimer();
#endif
    if (fuse && select_last_block(&state_rw->ticket[1])) {
#ifdef PM_PROFILE_NS
        if (threadIdx.x == 0) s_t[2] = pm_globaltimer();
#endif
        finalize_body<MODE>(partials, gridDim.x, sums, 3, state_rw, compose, ck, pc);
#ifdef PM_PROFILE_NS
        if (threadIdx.x == 0)
            printf("accumulate<%d> last block: block reduce %llu ns, ticket %llu ns, finalize %llu ns\n", MODE, s_t[1] - s_t[0], s_t[2] - s_t[1],
                   pm_globaltimer() - s_t[2]);
#endif
    }
}

// K3 + K4/K5 + K7 as ONE kernel of the fused loop: the exact quantile select runs as the first phase of the minimiser kernel,
// whose grid is one resident wave (cooperative launch), so a pass is "histogram my share, grid barrier" instead of a kernel:
// the block that arrives last at the barrier scans the bins (sharded reading: after the peer exchange of the histograms),
// writes the next plan or the limits, and releases the others.  With the window plan of select.cuh an iteration needs two
// passes (three or four when the order statistic left the window or there is none yet, e.g. the first iteration).  Limits,
// weights, sums and T are bit for bit those of hist_kernel x 3 + accumulate_kernel; what is gone is three launches, their
// drains and the DRAM round trips between them (0.037 + 0.043 ms -> see DESIGN.md K3).
#ifndef PM_FUSED_MIN_BLOCKS
#define PM_FUSED_MIN_BLOCKS(MODE) ((MODE) == 1 ? 2 : 3)
#endif
template <int MODE>
__global__ void __launch_bounds__(ACC_BLOCK, PM_FUSED_MIN_BLOCKS(MODE)) select_accumulate_kernel(const f4* __restrict__ reading, int nq, int k, const int32_t* __restrict__ ids,
                                                               const float* __restrict__ dists, const f4* __restrict__ ref,
                                                               const f4* __restrict__ normals, IcpState* state, double* __restrict__ partials,
                                                               double* sums, pmgpu_icp_params ck, const f4* __restrict__ reading_normals, PeerComm pc,
                                                               SelectSpec spec, unsigned* __restrict__ hist, int cap_active, float cap_margin, unsigned* __restrict__ cand, int defer_finalize) {
    constexpr int NS = MODE == 1 ? NS_PLANE : NS_POINT;
    __shared__ SelScratch s_sel;
    __shared__ Mat4 sT;
    __shared__ PairW s_w;
    __shared__ int s_flags[3];
    pdl_release();  // the one-block finalize kernel behind this one may be set up; it waits for this grid's completion itself
    if (__ldcg(&state->iterate) == 0) return;  // the fused loop is always gated; written by an earlier kernel
#ifdef PM_PROFILE_NS
    unsigned long long tp[16];
    int ntp = 0;
    tp[ntp++] = pm_globaltimer();
#endif
    const size_t total = (size_t)nq * k;
    const int nquant = spec.n_quantile();
    // sharded reading: the collected distances travel with the histogram when the chain has one quantile filter (one mailbox
    // slot holds both); with several, the histograms alone are exchanged and the select takes its second pass
    const int collect = (pc.nranks <= 1 || nquant == 1) ? 1 : 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        select_init_limits(state, spec);
        select_plans_begin(state, spec, collect);
    }
    for (int pass = 0;; ++pass) {
        int slot = 0;
        for (int f = 0; f < spec.nfilters; ++f) {
            if (!spec.is_quantile(f)) continue;
            SelPlan p;
            bool done = false;
            if (pass == 0) {
                // what block 0 writes into the state: no need to wait for it
                p = select_first_plan(__ldcg(&state->sel_guess[f]), collect ? __ldcg(&state->sel_inner[f]) : 0);
            } else {
                done = __ldcg(&state->sel_done[f]) != 0;
                p.lo = __ldcg(&state->sel_prefix[f]); p.shift = __ldcg(&state->sel_shift[f]);
                p.nb = __ldcg(&state->sel_nb[f]); p.outside = __ldcg(&state->sel_outside[f]);
                p.c0 = p.c1 = 0;
            }
            if (!done) select_pass_block(dists, total, p, &s_sel, hist + (size_t)slot * PM_HIST_BINS, cand + (size_t)f * PM_SEL_CAND_CAP, &state->sel_cand_count[f]);
            ++slot;
        }
#ifdef PM_PROFILE_NS
        if (ntp < 14) tp[ntp++] = pm_globaltimer();
#endif
        unsigned gen;
        if (grid_bar_arrive(state, gen)) {
            bool ok = true;
            // sharded reading: this rank's histograms become the sums over all ranks (comm.cuh), right here
            if (nquant > 0 && pc.nranks > 1) {
                if (nquant == 1) {
                    int fq = 0;
                    while (!spec.is_quantile(fq)) ++fq;
                    ok = peer_select_exchange(pc, hist, cand + (size_t)fq * PM_SEL_CAND_CAP, &state->sel_cand_count[fq], PM_SEL_CAND_CAP, state);
                } else {
                    ok = peer_allreduce<false>(pc, hist, nquant * PM_HIST_BINS, state);
                }
            }
            if (ok) {
                slot = 0;
                for (int f = 0; f < spec.nfilters; ++f) {
                    if (!spec.is_quantile(f)) continue;
                    if (!state->sel_done[f] && state->sel_pending > 0)
                        select_pick_plan(hist + (size_t)slot * PM_HIST_BINS, f, spec.quantile(f), spec.factor(f), state, &s_sel,
                                         cand + (size_t)f * PM_SEL_CAND_CAP, &state->sel_cand_count[f]);
                    ++slot;
                    __syncthreads();
                }
            }
            if (threadIdx.x == 0) {
                // (loads first, decisions after: the picking thread's round trips to the L2 are the critical path here)
                const int pending = state->sel_pending, iterate = state->iterate, passes = state->sel_passes;
                if (!ok) state->sel_pending = 0;  // the peers never arrived: status is raised, the loop has stopped
                if (nquant > 0) state->sel_passes = passes + 1;
                if (pending == 0 && ok && iterate) select_finish(state, cap_active, cap_margin);
            }
            grid_bar_release(state, gen);
        } else {
            grid_bar_wait(state, gen);
        }
#ifdef PM_PROFILE_NS
        if (ntp < 14) tp[ntp++] = pm_globaltimer();
#endif
        // everything the rest of the kernel reads from the state, requested at once (used only when the select is complete)
        if (threadIdx.x == 0) { s_flags[0] = __ldcg(&state->iterate); s_flags[1] = __ldcg(&state->redo); s_flags[2] = __ldcg(&state->sel_pending); }
        if (threadIdx.x >= 1 && threadIdx.x <= 16) sT.m[threadIdx.x - 1] = __ldcg(&state->T_iter.m[threadIdx.x - 1]);
        if (threadIdx.x == 32) load_pairw(state, &s_w);
        __syncthreads();
        if (s_flags[2] == 0) break;
        __syncthreads();  // s_flags is rewritten after the next pass
    }
    // redo: this iteration's capped match was void (select_finish) — leave T_iter as it is
    if (s_flags[0] == 0 || s_flags[1]) return;
    double acc[NS];
#pragma unroll
    for (int c = 0; c < NS; ++c) acc[c] = 0.0;
    accumulate_pairs<MODE>(acc, sT, &s_w, reading, nq, k, ids, dists, ref, normals, ck, reading_normals);
#ifdef PM_PROFILE_NS
    const unsigned long long t_acc = pm_globaltimer();
#endif
    block_reduce_store<NS>(acc, partials + (size_t)blockIdx.x * NS_MAX);
#ifdef PM_PROFILE_NS
    const unsigned long long t_red = pm_globaltimer();
#endif
    if (defer_finalize) return;  // the reduction of the rows, the solve and the composition follow as finalize_kernel
    if (select_last_block(&state->ticket[1])) {
#ifdef PM_PROFILE_NS
        const unsigned long long t_last = pm_globaltimer();
#endif
        finalize_body<MODE>(partials, gridDim.x, sums, 3, state, 1, ck, pc);
#ifdef PM_PROFILE_NS
        if (threadIdx.x == 0)
            printf("select_accumulate<%d> LAST block %d (ns since ITS start): select done %llu, accumulated %llu, block-reduced %llu, ticket %llu, finalized %llu, passes %d\n",
                   MODE, blockIdx.x, tp[ntp - 1] - tp[0], t_acc - tp[0], t_red - tp[0], t_last - tp[0], pm_globaltimer() - tp[0], state->sel_passes);
#endif
    }
#ifdef PM_PROFILE_NS
    if (blockIdx.x == 0 && threadIdx.x == 0)
        printf("select_accumulate<%d> block 0: select done %llu, accumulated %llu, block-reduced %llu\n", MODE, tp[ntp - 1] - tp[0], t_acc - tp[0], t_red - tp[0]);
#endif
}

// Censi covariance sums (PointToPlaneWithCov.cpp:100-150, PointToPointWithCov.cpp:84-135):
// [0..20] J_hessian upper triangle, [21..41] (d2J_dZdX d2J_dZdX^T) upper triangle.
template <int MODE>
__global__ void __launch_bounds__(ACC_BLOCK) cov_accumulate_kernel(const f4* __restrict__ reading, int nq, int k, const int32_t* __restrict__ ids,
                                                                   const float* __restrict__ dists, const f4* __restrict__ ref,
                                                                   const f4* __restrict__ normals, const IcpState* __restrict__ state,
                                                                   double* __restrict__ partials, const f4* __restrict__ reading_normals,
                                                                   const f4* __restrict__ sn_normals) {
    __shared__ Mat4 sT;
    __shared__ float sPar[12];  // alpha beta gamma tx ty tz | mean_reading | mean_reference
    if (threadIdx.x < 16) sT.m[threadIdx.x] = state->T_match.m[threadIdx.x];
    if (threadIdx.x == 0) {
        const float* t = state->dT.m;
        const float beta = -asinf(t[2]);
        const float alpha = atan2f(t[2 + 4], t[2 + 8]);
        const float gamma = atan2f(t[1] / cosf(beta), t[0] / cosf(beta));
        sPar[0] = alpha; sPar[1] = beta; sPar[2] = gamma;
        sPar[3] = t[12]; sPar[4] = t[13]; sPar[5] = t[14];
        for (int a = 0; a < 3; ++a) {
            sPar[6 + a] = MODE == 0 ? (float)state->mean_reading[a] : 0.f;
            sPar[9 + a] = MODE == 0 ? (float)state->mean_reference[a] : 0.f;
        }
    }
    __syncthreads();
    const float alpha = sPar[0], beta = sPar[1], gamma = sPar[2], t_x = sPar[3], t_y = sPar[4], t_z = sPar[5];
    double acc[NS_COV];
#pragma unroll
    for (int c = 0; c < NS_COV; ++c) acc[c] = 0.0;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride) {
        const f4 p = transform_point(sT, reading[i]);
        for (int kk = 0; kk < k; ++kk) {
            const float d = dists[(size_t)i * k + kk];
            if (pm_pair_weight(state, d) == 0.f) continue;
            const int id = ids[(size_t)i * k + kk];
            if (state->sn_on && pm_sn_weight(sT, reading_normals[i], __ldg(sn_normals + id), state->sn_eps) == 0.f) continue;
            const f4 q = __ldg(ref + id);
            if (state->robust_on && state->robust_p2plane && pm_robust_p2plane_weight(state, p, q, __ldg(sn_normals + id)) == 0.f) continue;
            float rp[3] = {p.x, p.y, p.z}, fp[3] = {q.x, q.y, q.z}, nrm[3] = {1.f, 1.f, 1.f};
            if (MODE == 1) {
                const f4 n = __ldg(normals + id);
                nrm[0] = n.x; nrm[1] = n.y; nrm[2] = n.z;
            } else {
                // the reference evaluates the point-to-point covariance on the de-meaned clouds
                for (int a = 0; a < 3; ++a) { rp[a] = fsub(rp[a], sPar[6 + a]); fp[a] = fsub(fp[a], sPar[9 + a]); }
            }
            const float reading_range = sqrtf(rp[0] * rp[0] + rp[1] * rp[1] + rp[2] * rp[2]);
            const float rd[3] = {rp[0] / reading_range, rp[1] / reading_range, rp[2] / reading_range};
            const float reference_range = sqrtf(fp[0] * fp[0] + fp[1] * fp[1] + fp[2] * fp[2]);
            const float fd[3] = {fp[0] / reference_range, fp[1] / reference_range, fp[2] / reference_range};
            const float n_alpha = nrm[2] * rd[1] - nrm[1] * rd[2];
            const float n_beta = nrm[0] * rd[2] - nrm[2] * rd[0];
            const float n_gamma = nrm[1] * rd[0] - nrm[0] * rd[1];
            float E = nrm[0] * (rp[0] - gamma * rp[1] + beta * rp[2] + t_x - fp[0]);
            E += nrm[1] * (gamma * rp[0] + rp[1] - alpha * rp[2] + t_y - fp[1]);
            E += nrm[2] * (-beta * rp[0] + alpha * rp[1] + rp[2] + t_z - fp[2]);
            float N_reading = nrm[0] * (rd[0] - gamma * rd[1] + beta * rd[2]);
            N_reading += nrm[1] * (gamma * rd[0] + rd[1] - alpha * rd[2]);
            N_reading += nrm[2] * (-beta * rd[0] + alpha * rd[1] + rd[2]);
            const float N_reference = -(nrm[0] * fd[0] + nrm[1] * fd[1] + nrm[2] * fd[2]);
            const float v[6] = {nrm[0], nrm[1], nrm[2], reading_range * n_alpha, reading_range * n_beta, reading_range * n_gamma};
            const float er = E + reading_range * N_reading;
            const float d1[6] = {nrm[0] * N_reading, nrm[1] * N_reading, nrm[2] * N_reading, n_alpha * er, n_beta * er, n_gamma * er};
            const float d2[6] = {nrm[0] * N_reference, nrm[1] * N_reference, nrm[2] * N_reference, reference_range * n_alpha * N_reference,
                                 reference_range * n_beta * N_reference, reference_range * n_gamma * N_reference};
            int c = 0;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = a; b < 6; ++b) {
                    acc[c] += (double)v[a] * (double)v[b];
                    acc[21 + c] += (double)d1[a] * (double)d1[b] + (double)d2[a] * (double)d2[b];
                    ++c;
                }
        }
    }
    block_reduce_store<NS_COV>(acc, partials + (size_t)blockIdx.x * NS_MAX);
}

// fixed-order reduction of the per-block rows: 256 threads = 32 columns x 8 slices (two rounds
// when there are more than 32 columns).  The loads of a slice are issued sixteen at a time before
// they are added (in row order), so the reduction costs a few L2 round trips, not one per row.
__device__ void reduce_rows(const double* __restrict__ partials, int nblocks, int ns, double* sums) {
    __shared__ double sh[8][NS_MAX];
    const int slice = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = lane; c < ns; c += 32) {
        double v = 0.0;
        for (int b = slice; b < nblocks; b += 128) {  // sixteen rows requested at once (rows past the end read as + 0)
            double r[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) r[u] = (b + 8 * u < nblocks) ? __ldcg(partials + (size_t)(b + 8 * u) * NS_MAX + c) : 0.0;
#pragma unroll
            for (int u = 0; u < 16; ++u) v += r[u];
        }
        sh[slice][c] = v;
    }
    __syncthreads();
    if (threadIdx.x < ns) {
        double v = 0.0;
        for (int s = 0; s < 8; ++s) v += sh[s][threadIdx.x];
        sums[threadIdx.x] = v;
    }
    __syncthreads();
}

__device__ void expand_sym6(const double* tri, double* A) {
    int c = 0;
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b) { A[a + 6 * b] = tri[c]; A[b + 6 * a] = tri[c]; ++c; }
}

// CounterTransformationChecker + DifferentialTransformationChecker
// (TransformationCheckersImpl.cpp:45-158), in the chain order the reference configs use.
__device__ void run_checkers(IcpState* st, const pmgpu_icp_params& ck) {
    st->counter += 1;
    if (st->counter >= ck.max_iterations) { st->iterate = 0; return; }  // MaxNumIterationsReached, ICP.cpp:423-427
    if (!ck.use_differential) return;
    const int len = st->hist_len;
    Mat4 Tq = st->T_iter;
    if (ck.minimizer & PM_MIN_DIM2) {
        // 2-D: check() builds the quaternion from topLeftCorner(3, 3) of the 3x3 HOMOGENEOUS matrix, translation column
        // included (TransformationCheckersImpl.cpp:131) — reproduced as it is (init() embeds the 2x2 properly, :116-121)
        Tq.m[8] = st->T_iter.m[12]; Tq.m[9] = st->T_iter.m[13]; Tq.m[10] = 1.f;
        Tq.m[2] = 0.f; Tq.m[6] = 0.f;
    }
    const Quat q = quat_from_mat4(Tq);
    float* hq = st->hist_q[len % PM_MAX_HISTORY];
    float* ht = st->hist_t[len % PM_MAX_HISTORY];
    hq[0] = q.w; hq[1] = q.x; hq[2] = q.y; hq[3] = q.z;
    ht[0] = st->T_iter.m[12]; ht[1] = st->T_iter.m[13]; ht[2] = st->T_iter.m[14];
    st->hist_len = len + 1;
    const int size = len + 1;
    float c0 = 0.f, c1 = 0.f;
    if (size > ck.smooth_length) {
        for (int i = size - 1; i >= size - ck.smooth_length; --i) {
            const float* a = st->hist_q[i % PM_MAX_HISTORY];
            const float* b = st->hist_q[(i - 1) % PM_MAX_HISTORY];
            const Quat qa = {a[0], a[1], a[2], a[3]}, qb = {b[0], b[1], b[2], b[3]};
            c0 += fabsf(quat_angular_distance(qa, qb));
            const float* ta = st->hist_t[i % PM_MAX_HISTORY];
            const float* tb = st->hist_t[(i - 1) % PM_MAX_HISTORY];
            const float dx = ta[0] - tb[0], dy = ta[1] - tb[1], dz = ta[2] - tb[2];
            c1 += fabsf(sqrtf(dx * dx + dy * dy + dz * dz));
        }
        c0 /= (float)ck.smooth_length;
        c1 /= (float)ck.smooth_length;
        if (c0 < ck.min_diff_rot_err && c1 < ck.min_diff_trans_err) st->iterate = 0;
    }
    if (c0 != c0 || c1 != c1) {
        if (st->status == 0) st->status = PMGPU_ERR_NAN;
        st->iterate = 0;
    }
}

// solve_psd (core/linalg.h) by one warp: the same diagonally pivoted Cholesky, element for element the same operations (so
// the same bits), but the pivot search, the column of L and the rank-1 update of every step run across the lanes on
// shared-memory matrices.  The serial routine is ~1 900 dependent instructions of ONE thread on an otherwise idle SM — 13 us,
// a fifth of the iteration's minimiser kernel — of which the arithmetic (6 square roots, 27 divisions, ~150 fused
// multiply-adds; tools/fp64_latency.cu: 99 / 132 / 7 cycles each on this GPU) is a small part: the rest is local-memory
// indexing.  N is a template parameter so that all index arithmetic is constant folding.  A, b, x in shared memory; every
// lane of warp 0 calls it.  A rank-deficient system takes the one-thread routine.
__device__ __noinline__ int solve_psd_serial(const double* A, const double* b, double* x, int n) { return solve_psd(A, b, x, n); }

template <int N>
__device__ __forceinline__ int solve_psd_warp(const double* A, const double* b, double* x) {
    __shared__ double sM[36], sL[36], sc[6];
    __shared__ int sperm[6];
    const int lane = threadIdx.x & 31;
    // the two matrix elements of this lane: e0 = lane, e1 = lane + 32 (column-major i + N j)
    const int e0 = lane, e1 = lane + 32;
    const int i0 = e0 % N, j0 = e0 / N, i1 = e1 % N, j1 = e1 / N;
    const bool in0 = e0 < N * N, in1 = e1 < N * N;
    if (in0) { sM[e0] = A[e0]; sL[e0] = 0.0; }
    if (in1) { sM[e1] = A[e1]; sL[e1] = 0.0; }
    if (lane < N) sperm[lane] = lane;
    __syncwarp();
    const double thr = (double)N * PM_FLT_EPS;
    double d0 = 0.0;
    int rank = 0;
#pragma unroll 1
    for (int k = 0; k < N; ++k) {
        // the first largest remaining diagonal entry
        double dv = (lane >= k && lane < N) ? sM[lane + N * lane] : -1.0;
        int pi = lane;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, dv, o);
            const int oi = __shfl_down_sync(0xffffffffu, pi, o);
            const bool mine_valid = pi >= k && pi < N, other_valid = oi >= k && oi < N;
            if (other_valid && (!mine_valid || ov > dv || (ov == dv && oi < pi))) { dv = ov; pi = oi; }
        }
        const int p = __shfl_sync(0xffffffffu, pi, 0);
        const double dmax = __shfl_sync(0xffffffffu, dv, 0);
        if (k == 0) d0 = dmax;
        if (!(dmax > thr * d0) || !(dmax > 0.0)) break;
        if (p != k) {
            if (lane < N) { const double t = sM[k + N * lane]; sM[k + N * lane] = sM[p + N * lane]; sM[p + N * lane] = t; }
            __syncwarp();
            if (lane < N) { const double t = sM[lane + N * k]; sM[lane + N * k] = sM[lane + N * p]; sM[lane + N * p] = t; }
            if (lane < k) { const double t = sL[k + N * lane]; sL[k + N * lane] = sL[p + N * lane]; sL[p + N * lane] = t; }
            if (lane == 0) { const int t = sperm[k]; sperm[k] = sperm[p]; sperm[p] = t; }
            __syncwarp();
        }
        const double lkk = sqrt(dmax);
        if (lane == k) sL[k + N * k] = lkk;
        if (lane > k && lane < N) sL[lane + N * k] = sM[lane + N * k] / lkk;
        __syncwarp();
        if (in0 && i0 > k && j0 > k) sM[e0] -= sL[i0 + N * k] * sL[j0 + N * k];
        if (in1 && i1 > k && j1 > k) sM[e1] -= sL[i1 + N * k] * sL[j1 + N * k];
        __syncwarp();
        ++rank;
    }
    if (rank < N) {
        if (lane == 0) solve_psd_serial(A, b, x, N);  // minimum-norm solution: rare
        __syncwarp();
        return rank;
    }
    if (lane < N) sc[lane] = b[sperm[lane]];
    __syncwarp();
    if (lane == 0) {
        double y[N], z[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double s = sc[i];
#pragma unroll
            for (int j = 0; j < i; ++j) s -= sL[i + N * j] * y[j];
            y[i] = s / sL[i + N * i];
        }
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            double s = y[i];
#pragma unroll
            for (int j = i + 1; j < N; ++j) s -= sL[j + N * i] * z[j];
            z[i] = s / sL[i + N * i];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) x[sperm[i]] = z[i];
    }
    __syncwarp();
    return rank;
}

// phase bits: 1 = reduce partial rows into sums, 2 = solve from sums.  Runs in one block of 256
// threads: the last block of the accumulate kernel (one GPU) or finalize_kernel (sharded reading,
// where the sums are all-reduced between the two phases).
template <int MODE>
__device__ void finalize_body(const double* partials, int nblocks, double* sums, int phase, IcpState* state, int compose,
                              const pmgpu_icp_params& ck, const PeerComm& pc) {
    constexpr int NS = MODE == 1 ? NS_PLANE : NS_POINT;
#ifdef PM_PROFILE_NS
    const unsigned long long t_a = pm_globaltimer();
#endif
    if (phase & 1) reduce_rows(partials, nblocks, NS, sums);
    // sharded reading: this rank's sums become the sums over all ranks, in rank order on every rank (comm.cuh), so every
    // rank solves the same system to the same bits
    if (phase == 3 && pc.nranks > 1 && !peer_allreduce<true>(pc, sums, (2 * NS + 3) & ~3, state)) return;
    if (!(phase & 2) || threadIdx.x >= 32) return;  // warp 0: the solve is a warp's work, everything around it lane 0's
    const bool lead = threadIdx.x == 0;
#ifdef PM_PROFILE_NS
    const unsigned long long t_b = pm_globaltimer();
#endif
    const double kept = sums[NS - 4], rej_matches = sums[NS - 3], rej_points = sums[NS - 2], seen = sums[NS - 1];
    // ErrorMinimizer.cpp:139-140: ratios over knn * number of reading points (all ranks)
    const float denom = (float)(seen * (double)ck.knn);
    if (lead) {
        state->stats[0] = (float)kept / denom;
        state->stats[2] = (float)rej_matches;
        state->stats[3] = (float)rej_points;
        state->stats[4] = (float)kept;
        state->T_match = state->T_iter;
    }
    if (!(kept > 0.0)) {
        if (lead) {
            if (state->status == 0) state->status = PMGPU_ERR_NO_POINT_TO_MINIMIZE;
            state->iterate = 0;
        }
        return;
    }
    Mat4 dT;
    if (MODE == 1) {
        __shared__ double sA[36], sb[6], sx[6];
        if (lead) state->stats[1] = (float)sums[27] / denom;  // sum of weights (== kept for the 0/1 filters)
        // rows / columns of the system actually solved: all six (rotation vector, translation); force4DOF: (cross_z, nx, ny,
        // nz) (PointToPlane.cpp:203-214); force2D / 2-D clouds: (cross_z, nx, ny) — the pseudo cross product of
        // ErrorMinimizer.cpp:308-313 is cross_z, and the accumulate kernel left z out of the residual (PointToPlane.cpp:177-186, 294-310)
        const int n = (ck.minimizer & PMGPU_MIN_FORCE2D) ? 3 : ((ck.minimizer & PMGPU_MIN_FORCE4DOF) ? 4 : 6), off = n == 6 ? 0 : 2;
        for (int e = threadIdx.x; e < n * n; e += 32) {
            const int r = off + e % n, c = off + e / n;
            const int a = r < c ? r : c, bb = r < c ? c : r;
            sA[e] = sums[a * 6 - a * (a - 1) / 2 + (bb - a)];  // upper triangle, row-major over a <= bb
        }
        if ((int)threadIdx.x < n) sb[threadIdx.x] = -sums[21 + off + threadIdx.x];
        __syncwarp();
#ifdef PM_PROFILE_NS
        const unsigned long long t_s0 = pm_globaltimer();
#endif
        const int rank_ = n == 6 ? solve_psd_warp<6>(sA, sb, sx) : (n == 4 ? solve_psd_warp<4>(sA, sb, sx) : solve_psd_warp<3>(sA, sb, sx));
#ifdef PM_PROFILE_NS
        if (lead) printf("   finalize: prologue %llu ns, solve_psd_warp %llu ns (n %d rank %d)\n", t_s0 - t_b, pm_globaltimer() - t_s0, n, rank_);
#else
        (void)rank_;
#endif
        if (!lead) return;
        float xf[6];
        if (ck.minimizer & PMGPU_MIN_FORCE2D) {
            const float ang = (float)sx[0];
            const float sn = sinf(ang), cs = cosf(ang);  // Eigen::Rotation2D<float>
            mat4_identity(dT);
            dT.m[0] = cs; dT.m[4] = -sn;
            dT.m[1] = sn; dT.m[5] = cs;
            dT.m[12] = (float)sx[1]; dT.m[13] = (float)sx[2];
        } else if (ck.minimizer & PMGPU_MIN_FORCE4DOF) {
            xf[0] = 0.f; xf[1] = 0.f;
            for (int a = 0; a < 4; ++a) xf[2 + a] = (float)sx[a];  // AngleAxis(x(0), unitZ), translation x(1..3)
        } else {
            for (int a = 0; a < 6; ++a) xf[a] = (float)sx[a];
        }
        if (!(ck.minimizer & PMGPU_MIN_FORCE2D)) angle_axis_to_mat4(xf, dT);
    } else {
        if (!lead) return;
        const double W = sums[0];
        state->stats[1] = (float)W / denom;
        const double inv = 1.0 / W;
        double mp[3], mq[3], m[9], R[9];
        for (int a = 0; a < 3; ++a) { mp[a] = sums[1 + a] * inv; mq[a] = sums[4 + a] * inv; }
        // the reference keeps the centroids in float (PointToPoint.cpp:67-72)
        float mpf[3], mqf[3];
        for (int a = 0; a < 3; ++a) { mpf[a] = (float)mp[a]; mqf[a] = (float)mq[a]; state->mean_reading[a] = mpf[a]; state->mean_reference[a] = mqf[a]; }
        // m = sum (w (q - mq)) (p - mp)^T = sum (w q) p^T - mq (sum w p)^T - (sum w q) mp^T + W mq mp^T
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r)
                m[r + 3 * c] = sums[7 + r + 3 * c] - (double)mqf[r] * sums[1 + c] - sums[4 + r] * (double)mpf[c] + W * (double)mqf[r] * (double)mpf[c];
        double sv[3], scale = 1.0;
        if (ck.minimizer & PM_MIN_DIM2) {
            // 2-D clouds: the rotation that maximises trace(R^T m) of the 2x2 cross-covariance in closed form — what
            // U V^T with the reflection fix of PointToPoint.cpp:88-93 (row dimCount - 2 of V^T) evaluates to
            const double cs = m[0] + m[4], sn = m[1] - m[3];
            const double h = sqrt(cs * cs + sn * sn);
            const double c = h > 0.0 ? cs / h : 1.0, s_ = h > 0.0 ? sn / h : 0.0;
            for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
            R[0] = c; R[3] = -s_;
            R[1] = s_; R[4] = c;
            sv[0] = sv[1] = sv[2] = 0.0;
        } else {
            rotation_from_crosscov(m, R, sv);
        }
        if ((ck.minimizer & 0xff) == PMGPU_MIN_P2POINT_SIM) {
            // sigma = sum w |p - mp|^2 from the raw sums (PointToPointSimilarity.cpp:69)
            double sigma = sums[16] + W * ((double)mpf[0] * mpf[0] + (double)mpf[1] * mpf[1] + (double)mpf[2] * mpf[2]);
            for (int a = 0; a < 3; ++a) sigma -= 2.0 * (double)mpf[a] * sums[1 + a];
            scale = (sv[0] + sv[1] + sv[2]) / sigma;
            if (sigma < 0.0001) scale = 1.0;
        }
        mat4_identity(dT);
        for (int c = 0; c < 3; ++c)
            for (int r = 0; r < 3; ++r) dT.m[r + 4 * c] = (float)(scale * R[r + 3 * c]);
        for (int r = 0; r < 3; ++r) {
            double acc = 0.0;
            for (int c = 0; c < 3; ++c) acc += (double)dT.m[r + 4 * c] * (double)mpf[c];
            dT.m[12 + r] = (float)((double)mqf[r] - acc);  // PointToPoint.cpp:94
        }
    }
#ifdef PM_PROFILE_NS
    const unsigned long long t_c = pm_globaltimer();
#endif
    state->dT = dT;
    if (compose) {
        Mat4 Tn;
        mat4_mul(dT, state->T_iter, Tn);  // ICP.cpp:411-412
        state->T_iter = Tn;
        state->iterations += 1;
        run_checkers(state, ck);
        // RigidTransformation::compute would throw on the next iteration (TransformationsImpl.cpp:62)
        if (state->iterate && (ck.minimizer & 0xff) != PMGPU_MIN_P2POINT_SIM && !mat4_is_rigid(Tn)) {
            if (state->status == 0) state->status = PMGPU_ERR_NOT_ORTHOGONAL;
            state->iterate = 0;
        }
    }
#ifdef PM_PROFILE_NS
    printf("finalize<%d>: reduce rows %llu ns, solve %llu ns, compose + checkers %llu ns\n", MODE, t_b - t_a, t_c - t_b, pm_globaltimer() - t_c);
#endif
}

template <int MODE>
__global__ void __launch_bounds__(256) finalize_kernel(const double* __restrict__ partials, int nblocks, double* __restrict__ sums, int phase,
                                                       IcpState* state, int gated, int compose, pmgpu_icp_params ck, PeerComm pc) {
    pdl_wait();     // dependent of the accumulate kernel when launched as such
    pdl_release();  // the next iteration's match may be set up
    if (gated && (state->iterate == 0 || state->redo)) return;
    finalize_body<MODE>(partials, nblocks, sums, phase, state, compose, ck, pc);
}

// inverse of a 6x6 by Gauss-Jordan with partial pivoting (J_hessian.inverse())
__device__ void inverse6(const double* A, double* Inv) {
    double M[36];
    for (int i = 0; i < 36; ++i) { M[i] = A[i]; Inv[i] = (i % 7 == 0) ? 1.0 : 0.0; }
    for (int k = 0; k < 6; ++k) {
        int piv = k;
        for (int i = k + 1; i < 6; ++i)
            if (fabs(M[i + 6 * k]) > fabs(M[piv + 6 * k])) piv = i;
        if (piv != k)
            for (int j = 0; j < 6; ++j) {
                double t = M[k + 6 * j]; M[k + 6 * j] = M[piv + 6 * j]; M[piv + 6 * j] = t;
                t = Inv[k + 6 * j]; Inv[k + 6 * j] = Inv[piv + 6 * j]; Inv[piv + 6 * j] = t;
            }
        const double d = M[k + 6 * k];
        for (int j = 0; j < 6; ++j) { M[k + 6 * j] /= d; Inv[k + 6 * j] /= d; }
        for (int i = 0; i < 6; ++i) {
            if (i == k) continue;
            const double f = M[i + 6 * k];
            if (f == 0.0) continue;
            for (int j = 0; j < 6; ++j) { M[i + 6 * j] -= f * M[k + 6 * j]; Inv[i + 6 * j] -= f * Inv[k + 6 * j]; }
        }
    }
}

__global__ void __launch_bounds__(256) cov_finalize_kernel(const double* __restrict__ partials, int nblocks, double* __restrict__ sums, int phase,
                                                           IcpState* state, float sensor_std_dev, PeerComm pc) {
    if (phase & 1) reduce_rows(partials, nblocks, NS_COV, sums);
    if (phase == 3 && pc.nranks > 1 && !peer_allreduce<true>(pc, sums, (2 * NS_COV + 3) & ~3, state)) return;
    if (!(phase & 2) || threadIdx.x != 0) return;
    double J[36], D[36], Ji[36], T1[36];
    expand_sym6(sums, J);
    expand_sym6(sums + 21, D);
    inverse6(J, Ji);
    for (int c = 0; c < 6; ++c)
        for (int r = 0; r < 6; ++r) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += Ji[r + 6 * k] * D[k + 6 * c];
            T1[r + 6 * c] = s;
        }
    const double s2 = (double)(sensor_std_dev * sensor_std_dev);
    for (int c = 0; c < 6; ++c)
        for (int r = 0; r < 6; ++r) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += T1[r + 6 * k] * Ji[k + 6 * c];
            state->cov[r + 6 * c] = (float)(s2 * s);
        }
}

}  // namespace

int launch_minimize(pmgpu_ctx* ctx, int minimizer_word, bool compose_and_check, bool gated, const pmgpu_icp_params* checks) {
    const int minimizer = minimizer_word & 0xff;
    const bool plane = (minimizer == PMGPU_MIN_P2PLANE || minimizer == PMGPU_MIN_P2PLANE_COV);
    if (plane && !ctx->has_normals) {
        ctx->set_error("Field normals not found");
        return PMGPU_ERR_NO_NORMALS;
    }
    cudaStream_t st = ctx->stream;
    // the kernel holds 20 (point) / 27 (plane) fp64 sums per thread: three / two 256-thread blocks are
    // resident per SM (launch bounds), so that is the whole grid — one wave, and only that many
    // partial rows for the last block to reduce
    const int grid = grid_for(ctx->nq, ACC_BLOCK, ctx->num_sms, plane ? 2 : 3);
    PM_CUDA_TRY(ctx, ctx->partials.reserve((size_t)(ctx->num_sms * 4 + 2) * NS_MAX));
    double* sums = ctx->partials.p + (size_t)ctx->num_sms * 4 * NS_MAX;
    pmgpu_icp_params ck;
    if (checks) ck = *checks;
    else {
        ck = pmgpu_icp_params();
        ck.max_iterations = 0x7fffffff;
    }
    ck.knn = ctx->k;
    ck.minimizer = minimizer_word;
    // 2-D clouds: point-to-plane is the force2D system (cross_z, nx, ny) with the xy residual (PointToPlane.cpp:171-312 with
    // dim == 3), point-to-point the 2x2 rotation, the Differential checker its 3x3 quirk
    if (ctx->dimh == 3) ck.minimizer |= PM_MIN_DIM2 | (plane ? PMGPU_MIN_FORCE2D : 0);
    const int g = gated ? 1 : 0, comp = compose_and_check ? 1 : 0;
    // sharded reading: the exchange of the sums is the last block's epilogue over the peer mailboxes (comm.cuh); without
    // mailboxes, NCCL between two finalize kernels
    const int fuse = (ctx->nranks > 1 && !ctx->peer_on) ? 0 : 1;
    const PeerComm pc = fuse ? comm_peers(ctx) : PeerComm{0, 1, {}};
    if (plane) accumulate_kernel<1><<<grid, ACC_BLOCK, 0, st>>>(ctx->reading.p, ctx->nq, ctx->k, ctx->ids.p, ctx->dists.p, ctx->ref_orig.p,
            ctx->ref_normals.p, ctx->state, g, ctx->partials.p, fuse, ctx->state, sums, comp, ck, ctx->reading_normals.p, pc);
    else accumulate_kernel<0><<<grid, ACC_BLOCK, 0, st>>>(ctx->reading.p, ctx->nq, ctx->k, ctx->ids.p, ctx->dists.p, ctx->ref_orig.p,
            ctx->has_normals ? ctx->ref_normals.p : nullptr, ctx->state, g, ctx->partials.p, fuse, ctx->state, sums, comp, ck,
            ctx->reading_normals.p, pc);
    ctx->launches += 1;
    if (!fuse) {
        const int ns = plane ? NS_PLANE : NS_POINT;
        if (plane) finalize_kernel<1><<<1, 256, 0, st>>>(ctx->partials.p, grid, sums, 1, ctx->state, g, comp, ck, pc);
        else finalize_kernel<0><<<1, 256, 0, st>>>(ctx->partials.p, grid, sums, 1, ctx->state, g, comp, ck, pc);
        PM_TRY(comm_allreduce_f64(ctx, sums, ns));
        if (plane) finalize_kernel<1><<<1, 256, 0, st>>>(ctx->partials.p, grid, sums, 2, ctx->state, g, comp, ck, pc);
        else finalize_kernel<0><<<1, 256, 0, st>>>(ctx->partials.p, grid, sums, 2, ctx->state, g, comp, ck, pc);
        ctx->launches += 2;
    }
    PM_CUDA_TRY(ctx, cudaGetLastError());
    return PMGPU_OK;
}

// The fused loop's K3 + K4/K5 + K7 in one cooperative launch (select_accumulate_kernel).  Applies to chains of distance
// filters (MaxDist / MedianDist / TrimmedDist / SurfaceNormal); chains with a Robust or VarTrimmedDist filter, and a sharded
// reading without peer mailboxes, keep the separate kernels.
bool fused_select_applies(const pmgpu_ctx* ctx, const SelectSpec& spec) {
    if (!ctx->fused_select) return false;
    // a sharded registration never consults the neighbourhood: every rank must take the same path (the two paths exchange
    // different messages), and a rank cannot know how many contexts its peers' GPUs hold
    if (ctx->nranks <= 1 && !alone_on_device(ctx)) return false;
    if (spec.robust_index() >= 0 || spec.var_index() >= 0) return false;
    if (ctx->nranks > 1 && !ctx->peer_on) return false;
    return true;
}

template <int MODE>
static int launch_select_minimize_mode(pmgpu_ctx* ctx, const SelectSpec& spec, const pmgpu_icp_params& ck, bool cap_active, const f4* normals) {
    auto kernel = select_accumulate_kernel<MODE>;
    if (ctx->fused_grid[MODE] == 0) {
        int per_sm = 0;
        PM_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, ACC_BLOCK, 0));
        if (per_sm < 1) {
            ctx->set_error("select_accumulate_kernel does not fit on an SM");
            return PMGPU_ERR_CUDA;
        }
        const int want = PM_FUSED_MIN_BLOCKS(MODE);  // the launch bounds
        ctx->fused_grid[MODE] = ctx->num_sms * (per_sm < want ? per_sm : want);
    }
    int grid = (ctx->nq + ACC_BLOCK - 1) / ACC_BLOCK;
    if (grid > ctx->fused_grid[MODE]) grid = ctx->fused_grid[MODE];
    if (grid < 1) grid = 1;
    PM_CUDA_TRY(ctx, ctx->partials.reserve((size_t)(ctx->num_sms * 4 + 2) * NS_MAX));
    double* sums = ctx->partials.p + (size_t)ctx->num_sms * 4 * NS_MAX;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(ACC_BLOCK);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // all blocks resident at once: the in-kernel grid barrier cannot deadlock
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx->fused_cooperative ? 1 : 0;
    PM_CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, kernel, (const f4*)ctx->reading.p, ctx->nq, ctx->k, (const int32_t*)ctx->ids.p, (const float*)ctx->dists.p,
                                        (const f4*)ctx->ref_orig.p, normals, ctx->state, ctx->partials.p, sums, ck, (const f4*)ctx->reading_normals.p,
                                        comm_peers(ctx), spec, ctx->hist.p, cap_active ? 1 : 0, ctx->cap_margin, ctx->sel_cand.p, ctx->defer_finalize ? 1 : 0));
    ctx->launches += 1;
    if (ctx->defer_finalize) {
        PM_CUDA_TRY(ctx, launch_dependent(pdl_enabled(ctx), finalize_kernel<MODE>, dim3(1), dim3(256), 0, ctx->stream, (const double*)ctx->partials.p, grid, sums, 3,
                                          ctx->state, 1, 1, ck, comm_peers(ctx)));
        ctx->launches += 1;
    }
    return PMGPU_OK;
}

int launch_select_minimize(pmgpu_ctx* ctx, const SelectSpec& spec, int minimizer_word, const pmgpu_icp_params* checks, bool cap_active) {
    const int minimizer = minimizer_word & 0xff;
    const bool plane = (minimizer == PMGPU_MIN_P2PLANE || minimizer == PMGPU_MIN_P2PLANE_COV);
    if (plane && !ctx->has_normals) {
        ctx->set_error("Field normals not found");
        return PMGPU_ERR_NO_NORMALS;
    }
    PM_TRY(select_reserve(ctx));
    PM_CUDA_TRY(ctx, ctx->sel_cand.reserve((size_t)PM_MAX_FILTERS * PM_SEL_CAND_CAP));
    pmgpu_icp_params ck = *checks;
    ck.knn = ctx->k;
    ck.minimizer = minimizer_word;
    if (ctx->dimh == 3) ck.minimizer |= PM_MIN_DIM2 | (plane ? PMGPU_MIN_FORCE2D : 0);  // as launch_minimize
    ctx->have_weights = true;
    if (plane) return launch_select_minimize_mode<1>(ctx, spec, ck, cap_active, ctx->ref_normals.p);
    return launch_select_minimize_mode<0>(ctx, spec, ck, cap_active, ctx->has_normals ? ctx->ref_normals.p : nullptr);
}

int launch_covariance(pmgpu_ctx* ctx, int minimizer, float sensor_std_dev) {
    const bool plane = ((minimizer & 0xff) == PMGPU_MIN_P2PLANE_COV);
    cudaStream_t st = ctx->stream;
    const int grid = grid_for(ctx->nq, ACC_BLOCK, ctx->num_sms, 4);
    PM_CUDA_TRY(ctx, ctx->partials.reserve((size_t)(ctx->num_sms * 4 + 2) * NS_MAX));
    double* sums = ctx->partials.p + (size_t)ctx->num_sms * 4 * NS_MAX;
    if (plane) cov_accumulate_kernel<1><<<grid, ACC_BLOCK, 0, st>>>(ctx->reading.p, ctx->nq, ctx->k, ctx->ids.p, ctx->dists.p, ctx->ref_orig.p,
            ctx->ref_normals.p, ctx->state, ctx->partials.p, ctx->reading_normals.p, ctx->ref_normals.p);
    else cov_accumulate_kernel<0><<<grid, ACC_BLOCK, 0, st>>>(ctx->reading.p, ctx->nq, ctx->k, ctx->ids.p, ctx->dists.p, ctx->ref_orig.p, nullptr,
            ctx->state, ctx->partials.p, ctx->reading_normals.p, ctx->ref_normals.p);
    ctx->launches += 1;
    const PeerComm none{0, 1, {}};
    if (ctx->nranks > 1 && !ctx->peer_on) {
        cov_finalize_kernel<<<1, 256, 0, st>>>(ctx->partials.p, grid, sums, 1, ctx->state, sensor_std_dev, none);
        PM_TRY(comm_allreduce_f64(ctx, sums, NS_COV));
        cov_finalize_kernel<<<1, 256, 0, st>>>(ctx->partials.p, grid, sums, 2, ctx->state, sensor_std_dev, none);
        ctx->launches += 2;
    } else {
        cov_finalize_kernel<<<1, 256, 0, st>>>(ctx->partials.p, grid, sums, 3, ctx->state, sensor_std_dev, comm_peers(ctx));
        ctx->launches += 1;
    }
    PM_CUDA_TRY(ctx, cudaGetLastError());
    return PMGPU_OK;
}

}  // namespace pm
