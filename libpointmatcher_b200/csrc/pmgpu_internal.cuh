// pmgpu_internal.cuh — context, device-buffer bookkeeping and the kernel-launcher prototypes
// shared by the translation units of libpmgpu.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/pmgpu.h"
#include "core/common.h"
#include "core/tree.h"

namespace pm {

// ---- error plumbing -------------------------------------------------------------------------
#define PM_CUDA_TRY(ctx, expr)                                                                        \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                             std::to_string(__LINE__) + ")");                                         \
            return PMGPU_ERR_CUDA;                                                                    \
        }                                                                                             \
    } while (0)

#define PM_TRY(expr)                     \
    do {                                 \
        int _s = (expr);                 \
        if (_s != PMGPU_OK) return _s;   \
    } while (0)

// Stream the current API call allocates on (set by every entry point).  Buffers come from the
// device's default memory pool (cudaMallocAsync) whose release threshold is raised at context
// creation, so the buffers of a finished registration are recycled by the next one instead of
// going back to the driver (cudaMalloc/cudaFree cost ~0.1-0.2 ms each and synchronise).
extern thread_local cudaStream_t g_alloc_stream;

// growable device buffer
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeAsync(p, g_alloc_stream);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocAsync((void**)&p, n * sizeof(T), g_alloc_stream);
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (p) cudaFreeAsync(p, g_alloc_stream);
        p = nullptr;
        cap = 0;
    }
};

// a temporary device buffer of one API call: released on every return path
template <typename T>
struct ScopedBuf : DevBuf<T> {
    ScopedBuf() = default;
    ScopedBuf(const ScopedBuf&) = delete;
    ScopedBuf& operator=(const ScopedBuf&) = delete;
    ~ScopedBuf() { this->release(); }
};

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------------
// The kernels of an iteration form a strict chain on one stream.  A kernel launched with launch_dependent() may be set up and
// have its blocks made resident while its predecessor is still running (the predecessor says when with pdl_release(), at its
// top: by then all of its own blocks have been scheduled); it must call pdl_wait() before it touches anything the
// predecessor writes — the wait returns when the predecessor has completed and its writes are visible.  Launched normally, both
// calls are no-ops.  What it buys is the launch latency of a ~2 us boundary, twice or three times per iteration.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- device-side state of one ICP registration -------------------------------------------------
// Lives in device memory so that a whole iteration (match -> select -> weights -> minimise ->
// compose -> check) runs without a host round trip.
#define PM_MAX_FILTERS 8
#define PM_HIST_BINS 2048
#define PM_MAX_HISTORY 64   // ring of T_iter history for the Differential checker
// internal bit of the minimiser word: the clouds are 2-D (features.rows() == 3), embedded with z = 0
#define PM_MIN_DIM2 0x400
#define PM_MAX_RANKS 8      // GPUs one registration can be sharded over (one NVSwitch node)
#define PM_MAILBOX_SLOT_WORDS (PM_MAX_FILTERS * PM_HIST_BINS)  // 64 KB: the largest message (one histogram per quantile filter)

struct IcpState {
    Mat4 T_iter;                 // cumulative transform applied to the reading (ICP.cpp:381)
    Mat4 T_match;                // the transform the resident matches were computed with
    Mat4 dT;                     // last incremental transform
    int status;                  // first PMGPU_* error raised on the device, 0 if none
    int iterate;                 // 1 while the checkers want another iteration
    int iterations;              // iterations executed
    int counter;                 // CounterTransformationChecker state
    // select / weights
    float limit[PM_MAX_FILTERS]; // per-filter squared-distance limit
    float limit_all;             // min over filters: weight = dist <= limit_all
    float limit_lo;              // max over the MinDist filters (0 without one): ... and dist >= limit_lo
    int has_filters;             // 0: empty chain (weight = dist != inf)
    unsigned long long n_valid;  // number of finite distances (all ranks)
    unsigned sel_prefix[PM_MAX_FILTERS];           // radix-select state per quantile filter
    unsigned long long sel_rank[PM_MAX_FILTERS];   // remaining rank inside the selected bucket
    unsigned ticket[4];          // "last block" counters: 0 select, 1 minimise, 2 covariance
    // select inside the minimiser kernel (minimize.cu select_accumulate_kernel): grid barrier words, and per quantile filter
    // the plan of the next histogram pass — bins (bits - sel_prefix) >> sel_shift, sel_nb of them — plus the bit pattern
    // of the previous iteration's order statistic, around which the first pass of the next iteration opens its window
    unsigned bar_count, bar_gen;
    unsigned sel_guess[PM_MAX_FILTERS], sel_prev[PM_MAX_FILTERS];  // centre of the next window; the last order statistic (0: none)
    int sel_shift[PM_MAX_FILTERS], sel_nb[PM_MAX_FILTERS], sel_outside[PM_MAX_FILTERS];
    int sel_inner[PM_MAX_FILTERS];            // half-width (window bins) of the bins whose distances the first pass collects
    int sel_c0[PM_MAX_FILTERS], sel_c1[PM_MAX_FILTERS];
    unsigned sel_cand_count[PM_MAX_FILTERS];  // collected distances (device-owned scratch counter)
    int sel_have_rank[PM_MAX_FILTERS], sel_done[PM_MAX_FILTERS];
    int sel_pending;             // quantile filters whose order statistic is not final yet
    int sel_passes;              // histogram passes of the last in-kernel select (statistics)
    unsigned overflow_count[2];  // kNN stage-2 queue lengths (ping-pong between consecutive launches)
    int sn_on;                   // SurfaceNormalOutlierFilter active (both clouds have normals)
    float sn_eps;                // cos(maxAngle)
    // RobustOutlierFilter (one per chain): set by select_init_limits / the robust select passes
    int robust_on;               // chain has a robust filter
    int robust_fct;              // PMGPU_ROBUST_*
    float robust_k;              // tuning
    float robust_scale;          // sqrt(MAD) or 1
    int robust_iteration;        // RobustOutlierFilter::iteration (starts at 1)
    int robust_recompute;        // this call re-estimates the scale (nbIterationForScale): 1 mad, 2 berg's first median, 3 std
    int robust_p2plane;          // distanceType point2plane: the weight function sees dot(n / |n|, p - q)^2
    float robust_target;         // berg: the scale the estimator converges to (the tuning the caller gave)
    float robust_approx2;        // weight 0 where e^2 >= this (`approximation` squared; +inf: none)
    float robust_median;
    unsigned robust_prefix;
    unsigned long long robust_rank;
    float var_ratio;             // VarTrimmedDistOutlierFilter: the optimised inlier ratio of the last evaluation
    unsigned long long var_best; // (FRMS bits << 32 | candidate index): atomicMin over the blocks = the first minimum
    // adaptive search radius of the fused loop (DESIGN.md "capped matching"): squared radius the
    // NEXT match may stop at, the largest distance the filters of THIS iteration needed to know
    // exactly, and the flag that voids an iteration whose cap turned out too small
    float cap;
    float cap_need;
    int redo;
    int redo_count;
    // minimiser outputs
    float cov[36];
    float stats[5];              // pointUsedRatio, weightedPointUsedRatio, nbRejectedMatches, nbRejectedPoints, nbKept
    double mean_reading[3], mean_reference[3]; // weighted centroids of the last point-to-point solve
    // Differential checker history
    float hist_q[PM_MAX_HISTORY][4];
    float hist_t[PM_MAX_HISTORY][3];
    int hist_len;
    unsigned long long visits;   // reference points examined (Matcher::visitCounter)
    int degenerate;              // degenerate normals counter (K8)
};

// the outlier-filter chain, by value in kernel arguments
struct SelectSpec {
    int nfilters;
    int type[PM_MAX_FILTERS];
    float param[PM_MAX_FILTERS];  // MaxDist / MinDist: squared limit; MedianDist: factor; TrimmedDist: ratio
    __host__ __device__ int kind(int f) const { return type[f] & 0xff; }
    __host__ __device__ bool is_robust(int f) const { return kind(f) == PMGPU_FILTER_ROBUST; }
    __host__ __device__ bool is_quantile(int f) const { return kind(f) == PMGPU_FILTER_MEDIANDIST || kind(f) == PMGPU_FILTER_TRIMMEDDIST; }
    __host__ __device__ float quantile(int f) const { return type[f] == PMGPU_FILTER_MEDIANDIST ? 0.5f : param[f]; }
    __host__ __device__ float factor(int f) const { return type[f] == PMGPU_FILTER_MEDIANDIST ? param[f] : 0.f; }
    int sn_active;  // the chain's SurfaceNormalOutlierFilter has normals on both sides to work with
    float robust_approx2;  // RobustOutlierFilter: squared `approximation`
    __host__ __device__ int sn_index() const {
        for (int f = 0; f < nfilters; ++f)
            if (kind(f) == PMGPU_FILTER_SURFACENORMAL) return f;
        return -1;
    }
    __host__ __device__ int robust_index() const {
        for (int f = 0; f < nfilters; ++f)
            if (is_robust(f)) return f;
        return -1;
    }
    __host__ __device__ int var_index() const {
        for (int f = 0; f < nfilters; ++f)
            if (kind(f) == PMGPU_FILTER_VARTRIMMEDDIST) return f;
        return -1;
    }
    __host__ __device__ int n_quantile() const {
        int n = 0;
        for (int f = 0; f < nfilters; ++f) n += is_quantile(f) ? 1 : 0;
        return n;
    }
};

// peer mailboxes of a sharded registration (protocol and device side: comm.cuh)
struct Mailbox {
    unsigned flag[2][PM_MAX_RANKS];                         // epoch of the last exchange rank r completed into bank b
    unsigned seq;                                           // exchanges THIS rank has executed (local; the epoch counter)
    unsigned pad[31 - 2 * PM_MAX_RANKS];
    unsigned slot[2][PM_MAX_RANKS][PM_MAILBOX_SLOT_WORDS];  // rank r's contribution, bank b
};
static_assert(2 * PM_MAX_RANKS <= 31, "flag block");
// by value in the kernel arguments
struct PeerComm {
    int rank, nranks;            // nranks <= 1: no exchange
    Mailbox* box[PM_MAX_RANKS];  // box[r]: rank r's mailbox as mapped here; box[rank] is local memory
};

}  // namespace pm

struct pmgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    int num_sms = 148;

    // reference (K1)
    int dimh = 4;                // features.rows() of the resident clouds: 4 (3-D) or 3 (2-D, held as (x, y, 0, w))
    int nr = 0;
    int depth = 0;
    pm::DevBuf<f4> ref_orig;     // original order, (x, y, z, w)
    pm::DevBuf<f4> ref_sorted;   // leaf order, w = original index bits
    pm::DevBuf<f4> ref_normals;  // original order, (nx, ny, nz, 0)
    pm::DevBuf<f4> reading_normals;  // Morton order like `reading`
    bool has_reading_normals = false;
    pm::DevBuf<float> reading_max_r2;  // KDTreeVarDistMatcher: squared per-point search radius, Morton order
    bool has_reading_max_r2 = false;
    pm::DevBuf<f2> splits;       // inner node (heap index) -> {split value, axis bits}
    pm::DevBuf<f4> boxes;        // node (heap index) -> 2 f4 (lo, hi)
    bool has_normals = false;
    // build scratch
    pm::DevBuf<uint64_t> keys_a, keys_b;
    pm::DevBuf<uint32_t> perm_a, perm_b;
    pm::DevBuf<uint32_t> node_box;  // 6 ordered-uint per node (lo xyz, hi xyz), heap index
    pm::DevBuf<uint8_t> cub_tmp;
    pm::DevBuf<uint8_t> seg_state;   // per-segment radix-select / partition state of the level being split (tree_build.cu SegState)
    pm::DevBuf<unsigned> seg_hist;   // ... and its histograms: PM_HIST_BINS per segment
    pm::DevBuf<unsigned> seg_cnt;    // ... and, per chunk of 4096 positions, the counts / offsets of its two segments' three classes
    bool build_select = true;        // PMGPU_BUILD_SORT=1: the upper levels by cub::DeviceRadixSort as in round 1 (A/B)

    // reading
    int nq = 0;
    pm::DevBuf<f4> reading;          // Morton order (position t holds original column q_order[t])
    pm::DevBuf<f4> reading_tmp;      // upload staging (original order)
    pm::DevBuf<uint32_t> q_order;    // sorted position -> original column
    int seed_k = 0;                  // k of the matches resident in `ids` for this reading / reference (0: none)
    pm::DevBuf<uint32_t> overflow;   // kNN stage-2 queue: sorted positions of the queries stage 1 did not finish
    pm::DevBuf<uint2> overflow_resume;  // ... and where each of them stood: (node to search next, pending-sibling trail)
    int knn_parity = 0;
    int knn_budget = 16;             // leaves a lane may scan before its query goes to stage 2
    bool seed_enabled = true;        // PMGPU_NO_SEED=1 switches the seeding off (A/B profiling)
    bool cap_enabled = true;         // PMGPU_NO_CAP=1: the fused loop matches without the adaptive radius
    float cap_margin = 1.5f;         // cap = margin x (largest squared distance the last filters needed); PMGPU_CAP_MARGIN
    bool seeded_without_planes = true;   // seeded searches skip the plane cache (measured 18 % faster); PMGPU_SEED_PLANES=1 reverts
    int knn_budget_unseeded = 32;        // same budget for searches that start without a bound (first iteration, k > 1)
    bool time_stage2 = false;        // PMGPU_TIME_STAGE2=1: report kNN stage 2 in the "covariance" timing slot (profiling)
    pm::DevBuf<int32_t> ids_tmp;     // un-permute staging for downloads
    pm::DevBuf<float> dists_tmp;

    // matches (K2)
    int k = 0;
    bool have_matches = false;
    bool have_weights = false;
    pm::DevBuf<int32_t> ids;     // k x nq
    pm::DevBuf<float> dists;     // k x nq
    pm::DevBuf<float> weights;   // k x nq, only materialised on request

    // select (K3)
    pm::DevBuf<unsigned> hist;   // PM_HIST_BINS
    pm::DevBuf<unsigned> sel_cand;   // distances collected by the in-kernel select's first pass, PM_SEL_CAND_CAP per filter
    // VarTrimmedDistOutlierFilter: minRatio / maxRatio (pmgpu_set_var_trimmed_ratios), sorted distance bits, running sums
    float var_min_ratio = 0.05f, var_max_ratio = 0.99f;
    float robust_approx2 = __builtin_inff();   // RobustOutlierFilter `approximation`, squared (pmgpu_set_robust_approximation)
    pm::DevBuf<unsigned> var_sorted;
    pm::DevBuf<float> var_cum;
    // minimiser partial sums (K4-K6)
    pm::DevBuf<double> partials;
    bool fused_select = true;        // fused loop: quantile select inside the minimiser kernel; PMGPU_NO_FUSED_SELECT=1 reverts
    bool defer_finalize = true;      // PMGPU_DEFER_FINALIZE=1: rows + solve + compose as a second, one-block kernel (A/B)
    bool counted = false;            // in the process-wide count of live contexts (api.cu)
    bool stage2_resume = true;       // PMGPU_NO_RESUME=1: stage 2 restarts every handed-over query from the root (A/B)
    bool pdl = true;                 // PMGPU_NO_PDL=1: every kernel waits for its predecessor's completion before it is set up
    bool fused_cooperative = true;   // PMGPU_COOP=0: plain launch of the same one-wave grid (A/B)
    int fused_grid[2] = {0, 0};      // co-resident blocks of select_accumulate_kernel<MODE> (occupancy query, once)

    pm::IcpState* state = nullptr;   // device
    pm::IcpState* state_host = nullptr;  // pinned host mirror

    // multi-GPU
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;
    // peer mailboxes (comm.cuh): mine, and every rank's as mapped into this process
    void* mailbox = nullptr;
    void* peer_box[PM_MAX_RANKS] = {};
    bool peer_opened[PM_MAX_RANKS] = {};   // mapped with cudaIpcOpenMemHandle (to be closed)
    bool peer_on = false;
    pm::DevBuf<f4> gather_tmp;             // sharded K8: leaf-order normals of all ranks before the un-permute

    // optional per-stage event timing (pmgpu_timing_enable)
    struct Interval { cudaEvent_t a, b; int stage; };
    bool timing = false;
    std::vector<Interval> intervals;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t copy_done = nullptr;  // host buffers handed to *_set may be released once this has fired
    // The reading's upload and Morton ordering run on a second stream when the main stream is still busy with work that
    // does not touch the reading or the ordering scratch (the reference's normals, the centring): `overlap_window` is
    // opened by pmgpu_ref_set after the structure build (ev_build) and closed by every other entry point.
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_build = nullptr, ev_reading = nullptr;
    bool overlap_window = false;
    bool overlap_enabled = true;      // PMGPU_NO_OVERLAP=1: everything on the one stream
    cudaEvent_t take_event() {
        cudaEvent_t e = nullptr;
        if (!event_pool.empty()) { e = event_pool.back(); event_pool.pop_back(); }
        else cudaEventCreate(&e);
        return e;
    }
    void stage_begin(int stage) {
        if (!timing) return;
        Interval iv{take_event(), take_event(), stage};
        cudaEventRecord(iv.a, stream);
        intervals.push_back(iv);
    }
    void stage_end() {
        if (!timing || intervals.empty()) return;
        cudaEventRecord(intervals.back().b, stream);
    }

    void set_error(const std::string& e) { err = e; }
    pm::TreeView tree_view() const {
        pm::TreeView t;
        t.splits = splits.p;
        t.boxes = boxes.p;
        t.pts = ref_sorted.p;
        t.n = (uint32_t)nr;
        t.depth = depth;
        return t;
    }
};

namespace pm {

// api.cu: is this the process's only live context on its GPU?  Two things are reserved for that case, because they hold SM
// slots while they wait and that only costs nothing when nothing else could have used them (config 5 runs three contexts per
// GPU): programmatic dependent launch (a pre-launched dependent's blocks: 420 vs 528 pairs/s) and the select inside the
// minimiser kernel (a whole-GPU cooperative grid spinning at its barrier: 525 vs 627 pairs/s).  Results do not depend on it.
bool alone_on_device(const pmgpu_ctx* ctx);
bool pdl_enabled(const pmgpu_ctx* ctx);
// tree_build.cu
int build_tree(pmgpu_ctx* ctx);
int morton_order(pmgpu_ctx* ctx);
size_t morton_scratch_bytes(uint32_t n);
// knn.cu
int launch_knn(pmgpu_ctx* ctx, const TreeView& tree, const f4* queries, int nq, bool use_T, bool gated, bool self_query, int k, float max_r2,
               bool use_seed, int32_t* ids, float* dists, bool use_cap = false, const float* var_r2 = nullptr);
// select.cu
int make_select_spec(pmgpu_ctx* ctx, int nfilters, const int* types, const float* params, SelectSpec* spec);
int select_reserve(pmgpu_ctx* ctx);
int launch_weights(pmgpu_ctx* ctx, const SelectSpec& spec, bool gated, bool cap_active);
int launch_materialize_weights(pmgpu_ctx* ctx);
// minimize.cu
int launch_minimize(pmgpu_ctx* ctx, int minimizer, bool compose_and_check, bool gated, const pmgpu_icp_params* checks);
int launch_covariance(pmgpu_ctx* ctx, int minimizer, float sensor_std_dev);
bool fused_select_applies(const pmgpu_ctx* ctx, const SelectSpec& spec);
int launch_select_minimize(pmgpu_ctx* ctx, const SelectSpec& spec, int minimizer, const pmgpu_icp_params* checks, bool cap_active);
// comm.cu
int comm_allreduce_u32(pmgpu_ctx* ctx, unsigned* buf, size_t count);
int comm_allreduce_f64(pmgpu_ctx* ctx, double* buf, size_t count);
int comm_allgather_bytes(pmgpu_ctx* ctx, void* buf, size_t bytes_per_rank);  // in place: rank r's block at r * bytes_per_rank
PeerComm comm_peers(pmgpu_ctx* ctx);  // the kernel argument of a fused exchange

inline int grid_for(int n, int block, int num_sms, int per_sm) {
    long g = ((long)n + block - 1) / block;
    const long cap = (long)num_sms * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace pm
