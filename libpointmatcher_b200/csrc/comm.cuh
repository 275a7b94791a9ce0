// comm.cuh — the cross-GPU exchange of a sharded registration, done INSIDE the producing kernel.
//
// With the reading's queries sharded over G GPUs against a replicated reference,
// every rank needs the same global order statistic and the same normal equations each iteration:
// the select histograms (<= 8 x 2048 u32) and the reduced sums (<= 42 f64) have to be summed over
// the ranks.  Those messages are <= 64 KB and the exchange sits on the critical path of a ~0.1 ms
// iteration, so it is not a library collective between kernels but the epilogue of the kernel that
// produced the data: the block that finishes last ("last block", select.cuh) stores its rank's
// contribution straight into a mailbox slot in EVERY peer's memory over NVLink (peer-mapped
// pointers: cudaIpcOpenMemHandle across processes, plain UVA pointers inside one), publishes an
// epoch flag per peer, spins on its own mailbox until all G flags of this epoch have arrived, and
// sums the G slots in rank order — the same order on every rank, so all ranks hold bit-identical
// sums, solve redundantly and stay in lock step without a broadcast.  One kernel per stage, as on
// one GPU; no NCCL launch between kernels.
//
// Mailbox discipline.  Every rank counts the exchanges it has EXECUTED (Mailbox::seq, device side: a
// gated kernel that returns early — loop finished, voided slot — executes none, on every rank
// alike, because all ranks hold identical state); that count is the epoch and its parity selects one
// of two slot banks.  A rank can only be one exchange ahead of a peer (it needs the peer's flag of
// exchange e to leave e), so bank (e & 1) is never rewritten (exchange e + 2) while a peer still
// reads it (exchange e).
// The wait is bounded: a rank whose peers never arrive (a failed launch elsewhere) raises
// PMGPU_ERR_COMM in its IcpState and stops iterating instead of hanging the device.
#pragma once
#include "pmgpu_internal.cuh"

namespace pm {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_volatile_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long pm_globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

#ifndef PM_COMM_TIMEOUT_NS
#define PM_COMM_TIMEOUT_NS 2000000000ull  // 2 s: far beyond any skew between ranks of one enqueue sequence
#endif

// Sum `words` 32-bit words (a multiple of 4, <= PM_MAILBOX_SLOT_WORDS) held in local global memory `data` over all
// ranks, in place.  Called by EVERY thread of ONE block per rank (the last block of the producing kernel), after the
// data is complete and visible (select_last_block).  F64: the words are doubles (added as doubles), else unsigned.
// Returns false on every thread when the peers did not arrive in time (status raised, iteration stopped).
template <bool F64>
__device__ __forceinline__ bool peer_allreduce(const PeerComm& pc, void* data, int words, IcpState* state) {
    __shared__ int s_ok;
    __shared__ unsigned s_epoch;
    if (threadIdx.x == 0) {
        s_epoch = pc.box[pc.rank]->seq + 1u;
        pc.box[pc.rank]->seq = s_epoch;
        s_ok = 1;
    }
    __syncthreads();
    const unsigned epoch = s_epoch;
    const int bank = (int)(epoch & 1u);
    const int quads = words >> 2;
    const uint4* src = reinterpret_cast<const uint4*>(data);
    // 1. my contribution into slot[bank][rank] of every rank (NVLink stores; own mailbox included)
    for (int i = threadIdx.x; i < quads; i += blockDim.x) {
        const uint4 v = __ldcg(src + i);
        for (int r = 0; r < pc.nranks; ++r) reinterpret_cast<uint4*>(pc.box[r]->slot[bank][pc.rank])[i] = v;
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish: one flag per peer, released after the data
    if ((int)threadIdx.x < pc.nranks) st_release_sys(&pc.box[threadIdx.x]->flag[bank][pc.rank], epoch);
    // 3. wait for every rank's flag of this epoch in MY mailbox
    if ((int)threadIdx.x < pc.nranks) {
        const unsigned* f = &pc.box[pc.rank]->flag[bank][threadIdx.x];
        const unsigned long long t0 = pm_globaltimer_ns();
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(f) - epoch) < 0) {
            if ((++spins & 0x3ffu) == 0 && pm_globaltimer_ns() - t0 > PM_COMM_TIMEOUT_NS) { s_ok = 0; break; }
        }
    }
    __syncthreads();
    if (!s_ok) {
        if (threadIdx.x == 0) {
            if (state->status == 0) state->status = PMGPU_ERR_COMM;
            state->iterate = 0;
        }
        return false;
    }
    // 4. sum the slots in rank order: the same order everywhere, so every rank holds identical bits
    const Mailbox* mine = pc.box[pc.rank];
    for (int i = threadIdx.x; i < quads; i += blockDim.x) {
        uint4 acc = ld_volatile_u4(reinterpret_cast<const uint4*>(mine->slot[bank][0]) + i);
        for (int r = 1; r < pc.nranks; ++r) {
            const uint4 v = ld_volatile_u4(reinterpret_cast<const uint4*>(mine->slot[bank][r]) + i);
            if (F64) {
                const double a0 = __hiloint2double((int)acc.y, (int)acc.x) + __hiloint2double((int)v.y, (int)v.x);
                const double a1 = __hiloint2double((int)acc.w, (int)acc.z) + __hiloint2double((int)v.w, (int)v.z);
                acc.x = (unsigned)__double2loint(a0); acc.y = (unsigned)__double2hiint(a0);
                acc.z = (unsigned)__double2loint(a1); acc.w = (unsigned)__double2hiint(a1);
            } else {
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
        reinterpret_cast<uint4*>(data)[i] = acc;
    }
    __threadfence();
    __syncthreads();
    return true;
}

// The in-kernel select of a sharded registration (select.cuh, minimize.cu): one exchange carries this rank's histogram slot
// (PM_HIST_BINS words, summed over the ranks in rank order) AND the distances its window pass collected (concatenated in rank
// order), so that a sharded iteration finds its order statistic in one pass like a single GPU does.  Message layout in a
// mailbox slot: [0, 2048) histogram | [2048] number of collected distances (0xffffffff: more than fit) | [2052, ...) distances.
// Same protocol and guarantees as peer_allreduce; called by every thread of the one picking block.  On return `hist` holds
// the summed histogram, cand[0 .. *cand_count) all ranks' collected distances (*cand_count > cap: the list is incomplete).
#define PM_SEL_MSG_CAND_OFFSET (PM_HIST_BINS + 4)
#define PM_SEL_MSG_CAND_MAX (PM_MAILBOX_SLOT_WORDS - PM_SEL_MSG_CAND_OFFSET)
__device__ __forceinline__ bool peer_select_exchange(const PeerComm& pc, unsigned* hist, unsigned* cand, unsigned* cand_count, unsigned cand_cap,
                                                     IcpState* state) {
    __shared__ int s_ok;
    __shared__ unsigned s_epoch, s_total;
    if (threadIdx.x == 0) {
        s_epoch = pc.box[pc.rank]->seq + 1u;
        pc.box[pc.rank]->seq = s_epoch;
        s_ok = 1;
    }
    __syncthreads();
    const unsigned epoch = s_epoch;
    const int bank = (int)(epoch & 1u);
    const unsigned n_mine = __ldcg(cand_count);
    const bool fits = n_mine <= (unsigned)PM_SEL_MSG_CAND_MAX - 4u && n_mine <= cand_cap - 4u;
    // lists travel in 16-byte units: the tail of the last unit is padded with 0xffffffff, which is not the bit pattern of a
    // distance and is skipped by the pick
    const unsigned n_pad = (n_mine + 3u) & ~3u;
    if (fits && threadIdx.x < n_pad - n_mine) cand[n_mine + threadIdx.x] = 0xffffffffu;
    __syncthreads();
    // 1. my histogram, count and collected distances into slot[bank][rank] of every rank (16-byte stores, loads issued ahead)
    const uint4* hist4 = reinterpret_cast<const uint4*>(hist);
    const uint4* cand4 = reinterpret_cast<const uint4*>(cand);
    const unsigned nq4 = fits ? n_pad >> 2 : 0u;
    for (unsigned i0 = threadIdx.x; i0 < (unsigned)(PM_HIST_BINS / 4) + nq4; i0 += 4u * blockDim.x) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned i = i0 + (unsigned)u * blockDim.x;
            if (i < (unsigned)(PM_HIST_BINS / 4)) v[u] = __ldcg(hist4 + i);
            else if (i < (unsigned)(PM_HIST_BINS / 4) + nq4) v[u] = __ldcg(cand4 + (i - PM_HIST_BINS / 4));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned i = i0 + (unsigned)u * blockDim.x;
            if (i >= (unsigned)(PM_HIST_BINS / 4) + nq4) continue;
            const unsigned w = i < (unsigned)(PM_HIST_BINS / 4) ? 4u * i : (unsigned)PM_SEL_MSG_CAND_OFFSET + 4u * (i - PM_HIST_BINS / 4);
            for (int r = 0; r < pc.nranks; ++r) *reinterpret_cast<uint4*>(&pc.box[r]->slot[bank][pc.rank][w]) = v[u];
        }
    }
    if (threadIdx.x == 0)
        for (int r = 0; r < pc.nranks; ++r) pc.box[r]->slot[bank][pc.rank][PM_HIST_BINS] = fits ? n_pad : 0xffffffffu;
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < pc.nranks) st_release_sys(&pc.box[threadIdx.x]->flag[bank][pc.rank], epoch);
    if ((int)threadIdx.x < pc.nranks) {
        const unsigned* f = &pc.box[pc.rank]->flag[bank][threadIdx.x];
        const unsigned long long t0 = pm_globaltimer_ns();
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(f) - epoch) < 0) {
            if ((++spins & 0x3ffu) == 0 && pm_globaltimer_ns() - t0 > PM_COMM_TIMEOUT_NS) { s_ok = 0; break; }
        }
    }
    __syncthreads();
    if (!s_ok) {
        if (threadIdx.x == 0) {
            if (state->status == 0) state->status = PMGPU_ERR_COMM;
            state->iterate = 0;
        }
        return false;
    }
    // 2. histogram summed in rank order; collected distances concatenated in rank order
    const Mailbox* mine = pc.box[pc.rank];
    for (int i = threadIdx.x; i < PM_HIST_BINS / 4; i += blockDim.x) {
        uint4 acc = ld_volatile_u4(reinterpret_cast<const uint4*>(mine->slot[bank][0]) + i);
        for (int r = 1; r < pc.nranks; ++r) {
            const uint4 v = ld_volatile_u4(reinterpret_cast<const uint4*>(mine->slot[bank][r]) + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<uint4*>(hist)[i] = acc;
    }
    unsigned base = 0;
    bool complete = true;
    for (int r = 0; r < pc.nranks; ++r) {
        const unsigned n_r = *(volatile const unsigned*)&mine->slot[bank][r][PM_HIST_BINS];  // a multiple of 4
        if (n_r == 0xffffffffu) { complete = false; break; }
        if (base + n_r <= cand_cap) {
            const uint4* src = reinterpret_cast<const uint4*>(&mine->slot[bank][r][PM_SEL_MSG_CAND_OFFSET]);
            uint4* dst = reinterpret_cast<uint4*>(cand + base);
            for (unsigned i0 = threadIdx.x; i0 < (n_r >> 2); i0 += 4u * blockDim.x) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i0 + (unsigned)u * blockDim.x < (n_r >> 2)) v[u] = ld_volatile_u4(src + i0 + (unsigned)u * blockDim.x);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i0 + (unsigned)u * blockDim.x < (n_r >> 2)) dst[i0 + (unsigned)u * blockDim.x] = v[u];
            }
        }
        base += n_r;
    }
    if (threadIdx.x == 0) s_total = complete ? base : cand_cap + 1u;
    __syncthreads();
    if (threadIdx.x == 0) *cand_count = s_total;
    __threadfence();
    __syncthreads();
    return true;
}

}  // namespace pm
