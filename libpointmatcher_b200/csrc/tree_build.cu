// tree_build.cu — K1: builds the search structure of KDTreeMatcher::init (MatchersImpl.cpp:77-83)
// on the device.  See core/tree.h for the layout.
//
// Upper levels, level by level, top down: (1) every node's bounding box by warp-aggregated integer
// atomics on order-preserving float keys, (2) the split axis = widest box extent, (3) one stable
// radix sort of (segment id << 32 | ordered coordinate) which median-splits every segment of the
// level at once (segments are position ranges, so the sort cannot move a point across segments).
// As soon as a segment fits one thread block (<= 4096 points) the rest of its subtree — boxes,
// axes, median splits, leaf order — is finished by that block in shared memory (subtree_kernel,
// a bitonic network on (node, coordinate, position) keys: the same total order as the stable
// radix sort, so the tree is the one the level-by-level build produces).
// The only library call is cub::DeviceRadixSort (part of the CUDA toolkit) — the build runs once
// per init(); the per-iteration kernels are all hand-written.
#include <cub/device/device_radix_sort.cuh>

#include "pmgpu_internal.cuh"
#include "select.cuh"

namespace pm {

namespace {

__global__ void iota_kernel(uint32_t* perm, uint32_t n) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) perm[p] = p;
}

__global__ void init_boxes_kernel(uint32_t* box, uint32_t nnodes) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnodes) {
        box[6 * i + 0] = 0xffffffffu; box[6 * i + 1] = 0xffffffffu; box[6 * i + 2] = 0xffffffffu;
        box[6 * i + 3] = 0u; box[6 * i + 4] = 0u; box[6 * i + 5] = 0u;
    }
}

// Boxes of all nodes of `level` (node = 2^level + segment).  Precondition: every segment of the
// level is longer than BOX_CHUNK (the upper levels; shorter segments belong to subtree_kernel), so
// the BOX_CHUNK consecutive positions of a block touch at most two segments: min/max in
// registers, one warp reduction, shared atomics, and 12 global atomics per block.
constexpr int BOX_CHUNK = 4096;
constexpr int BOX_THREADS = 256;
__global__ void __launch_bounds__(BOX_THREADS) level_boxes_kernel(const f4* __restrict__ pts, const uint32_t* __restrict__ perm, uint32_t n, int level,
                                                                  uint32_t* __restrict__ box) {
    __shared__ uint32_t sbox[12];
    const uint32_t c0 = blockIdx.x * BOX_CHUNK;
    const uint32_t c1 = min(n, c0 + BOX_CHUNK);
    const uint32_t seg0 = seg_of(c0, level, n);
    const uint32_t nb = seg_begin(level, seg0 + 1, n);  // first position of the next segment
    if (threadIdx.x < 12) sbox[threadIdx.x] = (threadIdx.x % 6 < 3) ? 0xffffffffu : 0u;
    __syncthreads();
    uint32_t lo[2][3], hi[2][3];
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[k][a] = 0xffffffffu; hi[k][a] = 0u; }
    for (uint32_t p = c0 + threadIdx.x; p < c1; p += BOX_THREADS) {
        const f4 pt = pts[perm ? perm[p] : p];
        const uint32_t o[3] = {float_ord(pt.x), float_ord(pt.y), float_ord(pt.z)};
        const bool second = p >= nb;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            lo[0][a] = min(lo[0][a], second ? 0xffffffffu : o[a]);
            hi[0][a] = max(hi[0][a], second ? 0u : o[a]);
            lo[1][a] = min(lo[1][a], second ? o[a] : 0xffffffffu);
            hi[1][a] = max(hi[1][a], second ? o[a] : 0u);
        }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const uint32_t l = __reduce_min_sync(0xffffffffu, lo[k][a]), h = __reduce_max_sync(0xffffffffu, hi[k][a]);
            if ((threadIdx.x & 31) == 0) { atomicMin(&sbox[6 * k + a], l); atomicMax(&sbox[6 * k + 3 + a], h); }
        }
    __syncthreads();
    if (threadIdx.x < 12) {
        const uint32_t k = threadIdx.x / 6, a = threadIdx.x % 6;
        if (k == 0 || nb < c1) {
            uint32_t* dst = box + 6 * (size_t)((1u << level) + seg0 + k) + a;
            if (a < 3) atomicMin(dst, sbox[threadIdx.x]); else atomicMax(dst, sbox[threadIdx.x]);
        }
    }
}

__device__ __forceinline__ int widest_axis(const uint32_t* b);

// sort key of every point for the split of `level`: (segment, coordinate along the widest axis)
__global__ void level_keys_kernel(const f4* __restrict__ pts, const uint32_t* __restrict__ perm, uint32_t n, int level,
                                  const uint32_t* __restrict__ box, uint64_t* __restrict__ keys) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t seg = seg_of(p, level, n);
    const int dim = widest_axis(box + 6 * (size_t)((1u << level) + seg));
    const f4 pt = pts[perm[p]];
    const float c = dim == 0 ? pt.x : (dim == 1 ? pt.y : pt.z);
    keys[p] = ((uint64_t)seg << 32) | (uint64_t)float_ord(c);
}

// ---- upper levels: exact median split of every segment of a level by radix SELECT + partition ------------------------------
// A level only has to put the (mid - lo) smallest coordinates of every segment into its left half — a full sort (round 1: one
// cub::DeviceRadixSort::SortPairs of 64-bit keys per level, five passes over 12 bytes per point) orders far more than that.  Here:
// 32-bit ordered coordinate per point (seg_keys_kernel), three histogram passes that find every segment's median key exactly
// (bits 31..21, 20..10, 9..0; the chunk's block that completes a segment's histogram also scans it — per-segment ticket),
// a count + a scatter pass.  Points equal to the median key fill the left half up to its exact size, so left <= split <= right
// holds and both halves have their exact sizes; the order is fixed by position (no atomics decide anything), so the same cloud
// gives the same tree on every GPU and in every run.
struct SegState {
    unsigned prefix, rank;        // radix-select state: selected bucket so far, remaining rank inside it
    unsigned ticket;              // chunks that have added their histogram
    unsigned key, ties_left;      // the median's ordered key; how many points equal to it go left
    unsigned eq_total;            // points equal to the median key
    unsigned pad[2];
};
constexpr int SEG_CHUNK = 4096;   // positions per block: touches at most two segments of an upper level (their length is >= 4096)
constexpr int SEG_THREADS = 256;

__global__ void seg_keys_kernel(const f4* __restrict__ pts, const uint32_t* __restrict__ perm, uint32_t n, int level, const uint32_t* __restrict__ box,
                                uint32_t* __restrict__ k32) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t seg = seg_of(p, level, n);
    const int dim = widest_axis(box + 6 * (size_t)((1u << level) + seg));
    const f4 pt = pts[perm[p]];
    k32[p] = float_ord(dim == 0 ? pt.x : (dim == 1 ? pt.y : pt.z));
}

__device__ __forceinline__ uint32_t seg_chunks(uint32_t lo, uint32_t hi) { return (hi - 1) / SEG_CHUNK - lo / SEG_CHUNK + 1; }

template <int PASS>
__global__ void __launch_bounds__(SEG_THREADS) seg_hist_kernel(const uint32_t* __restrict__ k32, uint32_t n, int level, SegState* __restrict__ seg,
                                                               unsigned* __restrict__ hist, const uint32_t* __restrict__ box, f2* __restrict__ splits) {
    __shared__ unsigned sh[2][PM_HIST_BINS];
    __shared__ unsigned long long warp_tot[32];
    __shared__ SelLocate s_loc;
    __shared__ int s_last[2];
    const uint32_t c0 = blockIdx.x * SEG_CHUNK, c1 = min(n, c0 + SEG_CHUNK);
    const uint32_t seg0 = seg_of(c0, level, n);
    const uint32_t nb = seg_begin(level, seg0 + 1, n);  // first position of the next segment
    const bool two = nb < c1;
    for (int i = threadIdx.x; i < 2 * PM_HIST_BINS; i += SEG_THREADS) (&sh[0][0])[i] = 0;
    __syncthreads();
    const unsigned pf0 = PASS == 0 ? 0u : seg[seg0].prefix, pf1 = (PASS == 0 || !two) ? 0u : seg[seg0 + 1].prefix;
    for (uint32_t p = c0 + threadIdx.x; p < c1; p += SEG_THREADS) {
        const uint32_t key = k32[p];
        const int s = p >= nb ? 1 : 0;
        const unsigned pf = s ? pf1 : pf0;
        if (PASS == 0) atomicAdd(&sh[s][key >> 21], 1u);
        else if (PASS == 1) { if ((key >> 21) == pf) atomicAdd(&sh[s][(key >> 10) & 0x7ffu], 1u); }
        else { if ((key >> 10) == pf) atomicAdd(&sh[s][key & 0x3ffu], 1u); }
    }
    __syncthreads();
    for (int s = 0; s < (two ? 2 : 1); ++s) select_flush(sh[s], hist + (size_t)(seg0 + s) * PM_HIST_BINS);
    __threadfence();
    __syncthreads();
    if (threadIdx.x < 2) {
        s_last[threadIdx.x] = 0;
        if (threadIdx.x == 0 || two) {
            const uint32_t sg = seg0 + threadIdx.x;
            const uint32_t lo = seg_begin(level, sg, n), hi = seg_begin(level, sg + 1, n);
            const unsigned t = atomicAdd(&seg[sg].ticket, 1u);
            if (t == seg_chunks(lo, hi) - 1) { s_last[threadIdx.x] = 1; seg[sg].ticket = 0; }
        }
    }
    __syncthreads();
    for (int s = 0; s < 2; ++s) {
        if (!s_last[s]) continue;  // block-uniform
        __threadfence();
        const uint32_t sg = seg0 + s;
        unsigned* h = hist + (size_t)sg * PM_HIST_BINS;
        const uint32_t lo = seg_begin(level, sg, n), mid = seg_begin(level + 1, 2 * sg + 1, n);
        const unsigned long long rank = PASS == 0 ? (unsigned long long)(mid - lo) : (unsigned long long)seg[sg].rank;
        const unsigned prefix = PASS == 0 ? 0u : seg[sg].prefix;
        SelScan sc;
        select_scan<false>(h, PASS == 2 ? 1024 : PM_HIST_BINS, sc, &s_loc, warp_tot);
        select_find(sc, rank, &s_loc);
        if (threadIdx.x == 0) {
            const unsigned bin = (unsigned)s_loc.bin;
            if (PASS == 0) { seg[sg].prefix = bin; seg[sg].rank = (unsigned)s_loc.rem; }
            else if (PASS == 1) { seg[sg].prefix = (prefix << 11) | bin; seg[sg].rank = (unsigned)s_loc.rem; }
            else {
                const unsigned key = (prefix << 10) | bin;
                seg[sg].key = key;
                seg[sg].ties_left = (unsigned)s_loc.rem;  // points equal to the median key that belong to the left half
                seg[sg].eq_total = s_loc.count;
                // split value of node (level, sg) = coordinate of the first point of its right half (left <= split <= right)
                const uint32_t node = (1u << level) + sg;
                splits[node] = make_float2(ord_float(key), __uint_as_float((uint32_t)widest_axis(box + 6 * (size_t)node)));
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < PM_HIST_BINS; i += SEG_THREADS) h[i] = 0;
        __syncthreads();
    }
}

// Every point to its half: < median key left, > right, == left while the left half has room.  DETERMINISTIC: the same cloud gives
// the same tree on every GPU (a sharded registration computes map normals per slice of leaf-order positions, so the ranks' trees
// must agree position by position) and in every run.  Two kernels: seg_count_kernel counts, per chunk and segment, the points below /
// equal to / above the median key, and the block that completes a segment's counts turns them into exclusive offsets in chunk
// order; seg_scatter_kernel classifies again and writes every point to  lo + [less | first ties_left equal] and
// mid + [other equal | greater],  ranked by chunk, then by thread, then by position within the thread.
constexpr int SEG_PER = SEG_CHUNK / SEG_THREADS;  // 16 positions per thread, strided

__global__ void __launch_bounds__(SEG_THREADS) seg_count_kernel(const uint32_t* __restrict__ k32, uint32_t n, int level, SegState* __restrict__ seg,
                                                                unsigned* __restrict__ chunk_cnt) {
    __shared__ unsigned s_tot[SEG_THREADS / 32][6];
    __shared__ int s_last[2];
    __shared__ unsigned s_scan[SEG_THREADS];
    const uint32_t c0 = blockIdx.x * SEG_CHUNK, c1 = min(n, c0 + SEG_CHUNK);
    const uint32_t seg0 = seg_of(c0, level, n);
    const uint32_t nb = seg_begin(level, seg0 + 1, n);
    const bool two = nb < c1;
    const unsigned m0 = seg[seg0].key, m1 = two ? seg[seg0 + 1].key : 0u;
    unsigned cnt[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < SEG_PER; ++j) {
        const uint32_t p = c0 + threadIdx.x + SEG_THREADS * j;
        if (p >= c1) continue;
        const uint32_t key = k32[p];
        const unsigned s = p >= nb ? 1u : 0u;
        const unsigned m = s ? m1 : m0;
        const unsigned c = key < m ? 0u : (key == m ? 1u : 2u);
#pragma unroll
        for (int q = 0; q < 6; ++q) cnt[q] += (s * 3 + c == (unsigned)q) ? 1u : 0u;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const unsigned v = __reduce_add_sync(0xffffffffu, cnt[q]);
        if (lane == 0) s_tot[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        unsigned v = 0;
        for (int w = 0; w < SEG_THREADS / 32; ++w) v += s_tot[w][threadIdx.x];
        chunk_cnt[(size_t)blockIdx.x * 6 + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x < 2) {
        s_last[threadIdx.x] = 0;
        if (threadIdx.x == 0 || two) {
            const uint32_t sg = seg0 + threadIdx.x;
            const uint32_t lo = seg_begin(level, sg, n), hi = seg_begin(level, sg + 1, n);
            const unsigned t = atomicAdd(&seg[sg].ticket, 1u);
            if (t == seg_chunks(lo, hi) - 1) { s_last[threadIdx.x] = 1; seg[sg].ticket = 0; }
        }
    }
    __syncthreads();
    for (int s = 0; s < 2; ++s) {
        if (!s_last[s]) continue;  // block-uniform
        __threadfence();
        // exclusive offsets of this segment's three classes over its chunks, in chunk order
        const uint32_t sg = seg0 + s;
        const uint32_t lo = seg_begin(level, sg, n), hi = seg_begin(level, sg + 1, n);
        const uint32_t cf = lo / SEG_CHUNK, nch = seg_chunks(lo, hi);
        const uint32_t per = (nch + SEG_THREADS - 1) / SEG_THREADS;
        const uint32_t i0 = min(nch, threadIdx.x * per), i1 = min(nch, i0 + per);
        for (int c = 0; c < 3; ++c) {
            // the segment is slot 0 of every chunk that starts inside it and slot 1 of the chunk that straddles its lower end
            auto slot_of = [&](uint32_t i) { return (i == 0 && (uint32_t)(cf * SEG_CHUNK) < lo) ? 1u : 0u; };
            unsigned sum = 0;
            for (uint32_t i = i0; i < i1; ++i) sum += __ldcg(chunk_cnt + (size_t)(cf + i) * 6 + slot_of(i) * 3 + c);
            // block exclusive scan of the threads' sums (warp shuffles, then the eight warp totals)
            unsigned incl = sum;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            if (lane == 31) s_scan[warp] = incl;
            __syncthreads();
            unsigned wbase = 0;
            for (int w = 0; w < warp; ++w) wbase += s_scan[w];
            __syncthreads();
            s_scan[threadIdx.x] = wbase + incl - sum;
            __syncthreads();
            unsigned off = s_scan[threadIdx.x];
            for (uint32_t i = i0; i < i1; ++i) {
                unsigned* e = chunk_cnt + (size_t)(cf + i) * 6 + slot_of(i) * 3 + c;
                const unsigned v = __ldcg(e);
                *e = off;
                off += v;
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(SEG_THREADS) seg_scatter_kernel(const uint32_t* __restrict__ k32, const uint32_t* __restrict__ perm_in, uint32_t n, int level,
                                                                  const SegState* __restrict__ seg, const unsigned* __restrict__ chunk_off,
                                                                  uint32_t* __restrict__ perm_out) {
    __shared__ unsigned s_warp[SEG_THREADS / 32][6];
    const uint32_t c0 = blockIdx.x * SEG_CHUNK, c1 = min(n, c0 + SEG_CHUNK);
    const uint32_t seg0 = seg_of(c0, level, n);
    const uint32_t nb = seg_begin(level, seg0 + 1, n);
    const bool two = nb < c1;
    const unsigned m0 = seg[seg0].key, m1 = two ? seg[seg0 + 1].key : 0u;
    unsigned long long cls = 0;  // 3 bits per position
    unsigned cnt[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < SEG_PER; ++j) {
        const uint32_t p = c0 + threadIdx.x + SEG_THREADS * j;
        if (p >= c1) continue;
        const uint32_t key = k32[p];
        const unsigned s = p >= nb ? 1u : 0u;
        const unsigned m = s ? m1 : m0;
        const unsigned c = key < m ? 0u : (key == m ? 1u : 2u);
        cls |= (unsigned long long)(s * 3 + c) << (3 * j);
#pragma unroll
        for (int q = 0; q < 6; ++q) cnt[q] += (s * 3 + c == (unsigned)q) ? 1u : 0u;
    }
    // rank of this thread's first point of every class inside the block: warps in order, lanes in order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned excl[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        unsigned v = cnt[q];
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        excl[q] = v - cnt[q];
        if (lane == 31) s_warp[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        unsigned run = 0;
        for (int w = 0; w < SEG_THREADS / 32; ++w) { const unsigned v = s_warp[w][threadIdx.x]; s_warp[w][threadIdx.x] = run; run += v; }
    }
    __syncthreads();
    // where class q of segment slot s starts for this thread
    unsigned pos[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) pos[q] = chunk_off[(size_t)blockIdx.x * 6 + q] + s_warp[warp][q] + excl[q];
    uint32_t lo[2], mid[2];
    unsigned ties[2], eq_total[2];
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const uint32_t sg = seg0 + s;
        lo[s] = seg_begin(level, sg, n);
        mid[s] = seg_begin(level + 1, 2 * sg + 1, n);
        ties[s] = (s == 0 || two) ? seg[sg].ties_left : 0u;
        eq_total[s] = (s == 0 || two) ? seg[sg].eq_total : 0u;
    }
#pragma unroll
    for (int j = 0; j < SEG_PER; ++j) {
        const uint32_t p = c0 + threadIdx.x + SEG_THREADS * j;
        if (p >= c1) continue;
        const unsigned q = (unsigned)((cls >> (3 * j)) & 7ull);
        unsigned r = 0;
#pragma unroll
        for (int qq = 0; qq < 6; ++qq)
            if ((unsigned)qq == q) { r = pos[qq]; pos[qq] += 1; }
        const unsigned s = q >= 3 ? 1u : 0u, c = q - 3 * s;
        const unsigned n_less = (mid[s] - lo[s]) - ties[s];
        uint32_t dst;
        if (c == 0) dst = lo[s] + r;
        else if (c == 1) dst = r < ties[s] ? lo[s] + n_less + r : mid[s] + (r - ties[s]);
        else dst = mid[s] + (eq_total[s] - ties[s]) + r;
        perm_out[dst] = perm_in[p];
    }
}

__global__ void gather_sorted_kernel(const f4* __restrict__ pts, const uint32_t* __restrict__ perm, uint32_t n, f4* __restrict__ out) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t src = perm[p];
    const f4 pt = pts[src];
    out[p] = make_float4(pt.x, pt.y, pt.z, __uint_as_float(src));
}

__device__ __forceinline__ int widest_axis(const uint32_t* b) {
    const float ex = fsub(ord_float(b[3]), ord_float(b[0]));
    const float ey = fsub(ord_float(b[4]), ord_float(b[1]));
    const float ez = fsub(ord_float(b[5]), ord_float(b[2]));
    int dim = 0;
    float best = ex;
    if (ey > best) { dim = 1; best = ey; }
    if (ez > best) { dim = 2; }
    return dim;
}

// after the sort of `level`: split value of node (level, s) = coordinate of the first point of its
// right half (left points <= split <= right points along the split axis)
__global__ void level_splits_kernel(const f4* __restrict__ pts, const uint32_t* __restrict__ perm, uint32_t n, int level,
                                    const uint32_t* __restrict__ box, f2* __restrict__ splits) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= (1u << level)) return;
    const uint32_t node = (1u << level) + s;
    const int dim = widest_axis(box + 6 * (size_t)node);
    const uint32_t mid = seg_begin(level + 1, 2 * s + 1, n);
    const f4 pt = pts[perm[mid]];
    const float c = dim == 0 ? pt.x : (dim == 1 ? pt.y : pt.z);
    splits[node] = make_float2(c, __uint_as_float((uint32_t)dim));
}

// node i (heap index) -> boxes[2i] = lo, boxes[2i+1] = hi
__global__ void pack_boxes_kernel(const uint32_t* __restrict__ box, uint32_t nnodes, f4* __restrict__ boxes) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnodes) return;
    const uint32_t* b = box + 6 * (size_t)i;
    boxes[2 * (size_t)i] = make_float4(ord_float(b[0]), ord_float(b[1]), ord_float(b[2]), 0.f);
    boxes[2 * (size_t)i + 1] = make_float4(ord_float(b[3]), ord_float(b[4]), ord_float(b[5]), 0.f);
}

// 30-bit Morton code of a point inside the cloud's bounding box
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void morton_keys_kernel(const f4* __restrict__ pts, uint32_t n, const uint32_t* __restrict__ box, uint32_t* __restrict__ keys,
                                   uint32_t* __restrict__ vals) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float lx = ord_float(box[6 + 0]), ly = ord_float(box[6 + 1]), lz = ord_float(box[6 + 2]);
    const float hx = ord_float(box[6 + 3]), hy = ord_float(box[6 + 4]), hz = ord_float(box[6 + 5]);
    const f4 pt = pts[p];
    const float sx = hx > lx ? 1023.f / (hx - lx) : 0.f, sy = hy > ly ? 1023.f / (hy - ly) : 0.f, sz = hz > lz ? 1023.f / (hz - lz) : 0.f;
    const uint32_t ix = (uint32_t)fminf(fmaxf((pt.x - lx) * sx, 0.f), 1023.f);
    const uint32_t iy = (uint32_t)fminf(fmaxf((pt.y - ly) * sy, 0.f), 1023.f);
    const uint32_t iz = (uint32_t)fminf(fmaxf((pt.z - lz) * sz, 0.f), 1023.f);
    keys[p] = spread10(ix) | (spread10(iy) << 1) | (spread10(iz) << 2);
    vals[p] = p;
}

__global__ void gather_f4_kernel(const f4* __restrict__ src, const uint32_t* __restrict__ order, uint32_t n, f4* __restrict__ dst) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dst[t] = src[order[t]];
}

// ---- lower levels: one block per subtree ---------------------------------------------------------
constexpr int SUB_MAX = 4096;      // points of one block's subtree
constexpr int SUB_THREADS = 1024;
constexpr int SUB_NODES = 1024;    // nodes of one level inside a block: 4096 points / leaves of >= 4

struct SubSmem {
    union {
        unsigned long long key[SUB_MAX];  // sort keys ...
        uint32_t box[6 * SUB_NODES];      // ... and, before they are generated, the ordered-uint boxes of the level's nodes
    };
    float x[SUB_MAX], y[SUB_MAX], z[SUB_MAX];
    uint16_t slot_a[SUB_MAX], slot_b[SUB_MAX];  // local position -> slot of the point (ping-pong)
    uint32_t nbeg[SUB_NODES + 1];               // local begin of every node of the current level
    uint8_t dim[SUB_NODES];                     // split axis of every node of the current level
};
static_assert(sizeof(SubSmem) <= 110 * 1024, "two subtree blocks per SM");

// node (0-based inside the block) that owns local position i at the current level
__device__ __forceinline__ uint32_t sub_node_of(const uint32_t* nbeg, uint32_t i, uint32_t nodes, uint32_t cnt) {
    // estimate (i * nodes < 2^22 is exact in float), then walk to the exact range
    uint32_t k = (uint32_t)(__fdividef((float)(i * nodes), (float)cnt));
    if (k >= nodes) k = nodes - 1;
    while (i < nbeg[k]) --k;
    while (i >= nbeg[k + 1]) ++k;
    return k;
}

__device__ __forceinline__ int widest_axis(const uint32_t* b);

// Block b finishes the subtree of node (L0, b): levels L0 .. D.  perm_in: order after the upper
// levels (null: identity, L0 == 0).
__global__ void __launch_bounds__(SUB_THREADS) subtree_kernel(const f4* __restrict__ pts, const uint32_t* __restrict__ perm_in, uint32_t n, int L0, int D,
                                                              uint32_t* __restrict__ node_box, f2* __restrict__ splits, f4* __restrict__ ref_sorted) {
    extern __shared__ __align__(16) unsigned char sub_raw[];
    SubSmem& sm = *reinterpret_cast<SubSmem*>(sub_raw);
    const uint32_t tid = threadIdx.x, b = blockIdx.x;
    const uint32_t gbeg = seg_begin(L0, b, n), cnt = seg_begin(L0, b + 1, n) - gbeg;
    for (uint32_t i = tid; i < cnt; i += SUB_THREADS) {
        const f4 p = pts[perm_in ? perm_in[gbeg + i] : gbeg + i];
        sm.x[i] = p.x; sm.y[i] = p.y; sm.z[i] = p.z;
        sm.slot_a[i] = (uint16_t)i;
    }
    uint16_t* cur = sm.slot_a;
    uint16_t* nxt = sm.slot_b;
    const uint32_t cnt32 = (cnt + 31u) & ~31u;
    for (int l = L0; l <= D; ++l) {
        const uint32_t nodes = 1u << (l - L0), first = b << (l - L0);
        for (uint32_t k = tid; k <= nodes; k += SUB_THREADS) sm.nbeg[k] = seg_begin(l, first + k, n) - gbeg;
        for (uint32_t k = tid; k < 6 * nodes; k += SUB_THREADS) sm.box[k] = (k % 6 < 3) ? 0xffffffffu : 0u;
        __syncthreads();
        // (1) boxes of the level's nodes
        for (uint32_t i = tid; i < cnt32; i += SUB_THREADS) {
            const bool active = i < cnt;
            uint32_t k = 0xffffffffu, ox = 0, oy = 0, oz = 0;
            if (active) {
                const uint32_t s = cur[i];
                k = sub_node_of(sm.nbeg, i, nodes, cnt);
                ox = float_ord(sm.x[s]); oy = float_ord(sm.y[s]); oz = float_ord(sm.z[s]);
            }
            const unsigned amask = __ballot_sync(0xffffffffu, active);
            if (active) {
                const unsigned group = __match_any_sync(amask, k);
                const uint32_t lx = __reduce_min_sync(group, ox), ly = __reduce_min_sync(group, oy), lz = __reduce_min_sync(group, oz);
                const uint32_t hx = __reduce_max_sync(group, ox), hy = __reduce_max_sync(group, oy), hz = __reduce_max_sync(group, oz);
                if ((tid & 31u) == (unsigned)(__ffs(group) - 1)) {
                    uint32_t* bx = sm.box + 6 * k;
                    atomicMin(bx + 0, lx); atomicMin(bx + 1, ly); atomicMin(bx + 2, lz);
                    atomicMax(bx + 3, hx); atomicMax(bx + 4, hy); atomicMax(bx + 5, hz);
                }
            }
        }
        __syncthreads();
        for (uint32_t k = tid; k < 6 * nodes; k += SUB_THREADS) node_box[6 * (size_t)((1u << l) + first) + k] = sm.box[k];
        if (l == D) break;
        for (uint32_t k = tid; k < nodes; k += SUB_THREADS) sm.dim[k] = (uint8_t)widest_axis(sm.box + 6 * k);
        __syncthreads();  // the boxes are dead from here: the keys take their place
        // (2) keys: coordinate along the node's widest axis | position (= stable)
        for (uint32_t i = tid; i < cnt; i += SUB_THREADS) {
            const uint32_t s = cur[i];
            const int dim = sm.dim[sub_node_of(sm.nbeg, i, nodes, cnt)];
            const float c = dim == 0 ? sm.x[s] : (dim == 1 ? sm.y[s] : sm.z[s]);
            sm.key[i] = ((unsigned long long)float_ord(c) << 12) | (unsigned long long)i;
        }
        __syncthreads();
        // (3) every node's range sorted ascending by a bitonic network of P = 2^m >= node size
        // virtual elements in which every compare-exchange points the same way (the first step of
        // a merge mirrors its partner): a partner beyond the node's size is a virtual +inf and the
        // exchange is a no-op, so ranges of any length sort in place without padding.
        uint32_t P = 2;
        while (P < sm.nbeg[1] + 1u) P <<= 1;  // node sizes of a level differ by at most one
        const uint32_t halfP = P >> 1, pairs = nodes * halfP;
        const int hs = 31 - __clz((int)halfP);
        for (uint32_t kk = 2; kk <= P; kk <<= 1) {
            for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
                const int js = 31 - __clz((int)j);
                for (uint32_t v = tid; v < pairs; v += SUB_THREADS) {
                    const uint32_t node = v >> hs, w = v & (halfP - 1);
                    uint32_t r_lo, r_hi;
                    if (j == (kk >> 1)) {  // flip step
                        const uint32_t blk = w >> js, off = w & (j - 1);
                        r_lo = blk * kk + off;
                        r_hi = blk * kk + kk - 1 - off;
                    } else {
                        r_lo = 2 * w - (w & (j - 1));
                        r_hi = r_lo + j;
                    }
                    const uint32_t base = sm.nbeg[node];
                    if (base + r_hi < sm.nbeg[node + 1]) {
                        const unsigned long long a = sm.key[base + r_lo], c = sm.key[base + r_hi];
                        if (a > c) { sm.key[base + r_lo] = c; sm.key[base + r_hi] = a; }
                    }
                }
                __syncthreads();
            }
        }
        // (4) the new order, (5) split values of the level's nodes
        for (uint32_t i = tid; i < cnt; i += SUB_THREADS) nxt[i] = cur[(uint32_t)(sm.key[i] & 0xfffull)];
        __syncthreads();
        { uint16_t* t = cur; cur = nxt; nxt = t; }
        for (uint32_t k = tid; k < nodes; k += SUB_THREADS) {
            const int dim = sm.dim[k];
            const uint32_t mid = seg_begin(l + 1, 2 * (first + k) + 1, n) - gbeg;
            const uint32_t s = cur[mid];
            const float c = dim == 0 ? sm.x[s] : (dim == 1 ? sm.y[s] : sm.z[s]);
            splits[(1u << l) + first + k] = make_float2(c, __uint_as_float((uint32_t)dim));
        }
        __syncthreads();
    }
    // leaf order: the point and, in w, its original column
    for (uint32_t i = tid; i < cnt; i += SUB_THREADS) {
        const uint32_t s = cur[i];
        const uint32_t src = perm_in ? perm_in[gbeg + s] : gbeg + s;
        ref_sorted[gbeg + i] = make_float4(sm.x[s], sm.y[s], sm.z[s], __uint_as_float(src));
    }
}

// first level whose segments fit one block
inline int subtree_level(uint32_t n, int D) {
    int l = 0;
    while (l < D && (((uint64_t)n + ((1ull << l) - 1)) >> l) > (uint64_t)SUB_MAX) ++l;
    return l;
}

inline unsigned blocks_for(uint32_t n, int block) { return (unsigned)((n + (uint32_t)block - 1) / (uint32_t)block); }

}  // namespace

int build_tree(pmgpu_ctx* ctx) {
    const uint32_t n = (uint32_t)ctx->nr;
    cudaStream_t st = ctx->stream;
    const int D = tree_depth_for(n);
    ctx->depth = D;
    const uint32_t nnodes = 2u << D;  // heap indices 1 .. 2^(D+1)-1
    PM_CUDA_TRY(ctx, ctx->node_box.reserve(6 * (size_t)nnodes));
    PM_CUDA_TRY(ctx, ctx->perm_a.reserve(n));
    PM_CUDA_TRY(ctx, ctx->perm_b.reserve(n));
    PM_CUDA_TRY(ctx, ctx->keys_a.reserve(n));
    PM_CUDA_TRY(ctx, ctx->keys_b.reserve(n));
    PM_CUDA_TRY(ctx, ctx->ref_sorted.reserve(n));
    PM_CUDA_TRY(ctx, ctx->splits.reserve((size_t)1 << D));
    PM_CUDA_TRY(ctx, ctx->boxes.reserve(2 * (size_t)nnodes));
    size_t tmp_bytes = 0;
    PM_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, ctx->keys_a.p, ctx->keys_b.p, ctx->perm_a.p, ctx->perm_b.p, (int)n, 0, 64, st));
    PM_CUDA_TRY(ctx, ctx->cub_tmp.reserve(tmp_bytes));

    const int B = 256;
    init_boxes_kernel<<<blocks_for(nnodes, B), B, 0, st>>>(ctx->node_box.p, nnodes);
    iota_kernel<<<blocks_for(n, B), B, 0, st>>>(ctx->perm_a.p, n);
    ctx->launches += 2;
    uint32_t* perm = ctx->perm_a.p;
    uint32_t* perm_alt = ctx->perm_b.p;
    const int L0 = subtree_level(n, D);
    if (L0 > 0 && ctx->build_select) {
        // upper levels by radix select + partition (see SegState)
        const size_t nseg = (size_t)1 << (L0 - 1);
        PM_CUDA_TRY(ctx, ctx->seg_state.reserve(nseg * sizeof(SegState)));
        PM_CUDA_TRY(ctx, ctx->seg_hist.reserve(nseg * PM_HIST_BINS));
        PM_CUDA_TRY(ctx, ctx->seg_cnt.reserve((size_t)blocks_for(n, SEG_CHUNK) * 6));
        PM_CUDA_TRY(ctx, cudaMemsetAsync(ctx->seg_state.p, 0, nseg * sizeof(SegState), st));
        PM_CUDA_TRY(ctx, cudaMemsetAsync(ctx->seg_hist.p, 0, nseg * PM_HIST_BINS * sizeof(unsigned), st));
        SegState* seg = reinterpret_cast<SegState*>(ctx->seg_state.p);
        uint32_t* k32 = reinterpret_cast<uint32_t*>(ctx->keys_a.p);
        const unsigned chunks = blocks_for(n, SEG_CHUNK);
        for (int l = 0; l < L0; ++l) {
            level_boxes_kernel<<<blocks_for(n, BOX_CHUNK), BOX_THREADS, 0, st>>>(ctx->ref_orig.p, perm, n, l, ctx->node_box.p);
            seg_keys_kernel<<<blocks_for(n, B), B, 0, st>>>(ctx->ref_orig.p, perm, n, l, ctx->node_box.p, k32);
            seg_hist_kernel<0><<<chunks, SEG_THREADS, 0, st>>>(k32, n, l, seg, ctx->seg_hist.p, ctx->node_box.p, ctx->splits.p);
            seg_hist_kernel<1><<<chunks, SEG_THREADS, 0, st>>>(k32, n, l, seg, ctx->seg_hist.p, ctx->node_box.p, ctx->splits.p);
            seg_hist_kernel<2><<<chunks, SEG_THREADS, 0, st>>>(k32, n, l, seg, ctx->seg_hist.p, ctx->node_box.p, ctx->splits.p);
            seg_count_kernel<<<chunks, SEG_THREADS, 0, st>>>(k32, n, l, seg, ctx->seg_cnt.p);
            seg_scatter_kernel<<<chunks, SEG_THREADS, 0, st>>>(k32, perm, n, l, seg, ctx->seg_cnt.p, perm_alt);
            uint32_t* t = perm; perm = perm_alt; perm_alt = t;
            ctx->launches += 7;
        }
    } else {
    for (int l = 0; l < L0; ++l) {
        level_boxes_kernel<<<blocks_for(n, BOX_CHUNK), BOX_THREADS, 0, st>>>(ctx->ref_orig.p, perm, n, l, ctx->node_box.p);
        level_keys_kernel<<<blocks_for(n, B), B, 0, st>>>(ctx->ref_orig.p, perm, n, l, ctx->node_box.p, ctx->keys_a.p);
        size_t tb = ctx->cub_tmp.cap;
        PM_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tb, ctx->keys_a.p, ctx->keys_b.p, perm, perm_alt, (int)n, 0, 32 + l, st));
        uint32_t* t = perm; perm = perm_alt; perm_alt = t;
        level_splits_kernel<<<blocks_for(1u << l, B), B, 0, st>>>(ctx->ref_orig.p, perm, n, l, ctx->node_box.p, ctx->splits.p);
        ctx->launches += 4;
    }
    }
    PM_CUDA_TRY(ctx, cudaFuncSetAttribute(subtree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SubSmem)));
    subtree_kernel<<<1u << L0, SUB_THREADS, sizeof(SubSmem), st>>>(ctx->ref_orig.p, L0 > 0 ? perm : nullptr, n, L0, D, ctx->node_box.p, ctx->splits.p,
                                                                   ctx->ref_sorted.p);
    ctx->launches += 1;
    pack_boxes_kernel<<<blocks_for(nnodes, B), B, 0, st>>>(ctx->node_box.p, nnodes, ctx->boxes.p);
    ctx->launches += 1;
    PM_CUDA_TRY(ctx, cudaGetLastError());
    return PMGPU_OK;
}

// Query schedule: the reading is static during an ICP run (only T_iter changes), so it is stored
// once in Morton order of its *untransformed* coordinates (ctx->reading; q_order maps a sorted
// position back to the caller's column).  A rigid motion preserves locality, hence the 32 queries
// of a warp walk nearly the same tree path every iteration and all loads/stores are coalesced.
// Expects the uploaded cloud in ctx->reading_tmp.
// bytes of sort scratch (ctx->cub_tmp) morton_order needs for n points
size_t morton_scratch_bytes(uint32_t n) {
    size_t tmp_bytes = 0;
    uint32_t* none = nullptr;
    if (cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, none, none, none, none, (int)n, 0, 30, nullptr) != cudaSuccess) return ~(size_t)0;
    return tmp_bytes;
}

int morton_order(pmgpu_ctx* ctx) {
    const uint32_t n = (uint32_t)ctx->nq;
    cudaStream_t st = ctx->stream;
    PM_CUDA_TRY(ctx, ctx->q_order.reserve(n));
    PM_CUDA_TRY(ctx, ctx->perm_a.reserve(n));
    PM_CUDA_TRY(ctx, ctx->perm_b.reserve(n));
    PM_CUDA_TRY(ctx, ctx->keys_a.reserve(n));  // reused as 2 x uint32 key arrays
    PM_CUDA_TRY(ctx, ctx->node_box.reserve(12));
    uint32_t* keys_in = reinterpret_cast<uint32_t*>(ctx->keys_a.p);
    uint32_t* keys_out = keys_in + n;
    size_t tmp_bytes = 0;
    PM_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, ctx->perm_a.p, ctx->q_order.p, (int)n, 0, 30, st));
    PM_CUDA_TRY(ctx, ctx->cub_tmp.reserve(tmp_bytes));
    const int B = 256;
    init_boxes_kernel<<<1, 32, 0, st>>>(ctx->node_box.p, 2);
    level_boxes_kernel<<<blocks_for(n, BOX_CHUNK), BOX_THREADS, 0, st>>>(ctx->reading_tmp.p, nullptr, n, 0, ctx->node_box.p);
    morton_keys_kernel<<<blocks_for(n, B), B, 0, st>>>(ctx->reading_tmp.p, n, ctx->node_box.p, keys_in, ctx->perm_a.p);
    size_t tb = ctx->cub_tmp.cap;
    PM_CUDA_TRY(ctx, cub::DeviceRadixSort::SortPairs(ctx->cub_tmp.p, tb, keys_in, keys_out, ctx->perm_a.p, ctx->q_order.p, (int)n, 0, 30, st));
    // the reading itself is kept in that order so every per-iteration access is coalesced
    gather_f4_kernel<<<blocks_for(n, B), B, 0, st>>>(ctx->reading_tmp.p, ctx->q_order.p, n, ctx->reading.p);
    ctx->launches += 5;
    PM_CUDA_TRY(ctx, cudaGetLastError());
    return PMGPU_OK;
}

}  // namespace pm
