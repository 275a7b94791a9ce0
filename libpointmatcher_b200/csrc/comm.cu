// comm.cu — host side of the cross-GPU exchanges of the path.
//
// Sharded registration (queries split over ranks, reference replicated): every rank needs the same
// global quantile and the same normal equations each iteration.  The per-iteration exchanges (select
// histograms, reduced sums) run INSIDE the producing kernels over peer-mapped mailboxes (comm.cuh);
// this file allocates a rank's mailbox, exports it (cudaIpcGetMemHandle) and maps the peers'
// (cudaIpcOpenMemHandle across processes, direct UVA pointers + cudaDeviceEnablePeerAccess inside one
// process).  NCCL stays for what is a plain bandwidth collective — the one-off all-gather of the map
// normals each rank computed for its slice — and as the fallback exchange when no
// mailboxes were set up (in-stream all-reduces between two kernels).  NCCL is resolved at run time
// (dlopen) so that a process that already carries an NCCL (e.g. torch's bundled one) shares it and
// single-GPU users need none.
#include <dlfcn.h>
#include <string.h>
#include <unistd.h>
#include <nccl.h>

#include "pmgpu_internal.cuh"

namespace pm {

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
    bool load() {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            err = std::string("cannot load NCCL: ") + dlerror();
            return false;
        }
        GetUniqueId = (decltype(GetUniqueId))dlsym(handle, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(handle, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
        AllGather = (decltype(AllGather))dlsym(handle, "ncclAllGather");
        GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !AllGather || !GetErrorString) {
            err = "NCCL library lacks a required symbol";
            return false;
        }
        return true;
    }
};

NcclApi& api() {
    static NcclApi a;
    return a;
}

int allreduce(pmgpu_ctx* ctx, void* buf, size_t count, ncclDataType_t type) {
    if (ctx->nranks <= 1) return PMGPU_OK;
    if (!ctx->nccl_comm) {
        ctx->set_error("sharded reading without a communicator: call pmgpu_comm_peer_init or pmgpu_comm_init first");
        return PMGPU_ERR_COMM;
    }
    ncclResult_t r = api().AllReduce(buf, buf, count, type, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != ncclSuccess) {
        ctx->set_error(std::string("ncclAllReduce: ") + api().GetErrorString(r));
        return PMGPU_ERR_COMM;
    }
    return PMGPU_OK;
}

// what pmgpu_comm_peer_handle hands out: enough for a peer in this or another process to map the mailbox
struct PeerHandle {
    uint64_t magic;
    int32_t pid, device;
    uint64_t ptr;
    cudaIpcMemHandle_t ipc;
};
static_assert(sizeof(PeerHandle) <= 128, "peer handles travel as 128-byte blobs");
constexpr uint64_t PEER_MAGIC = 0x706d6770755f6d62ull;  // "pmgpu_mb"

void peer_close(pmgpu_ctx* ctx) {
    for (int r = 0; r < PM_MAX_RANKS; ++r) {
        if (ctx->peer_opened[r] && ctx->peer_box[r]) cudaIpcCloseMemHandle(ctx->peer_box[r]);
        ctx->peer_box[r] = nullptr;
        ctx->peer_opened[r] = false;
    }
    ctx->peer_on = false;
}

}  // namespace

int comm_allreduce_u32(pmgpu_ctx* ctx, unsigned* buf, size_t count) { return allreduce(ctx, buf, count, ncclUint32); }
int comm_allreduce_f64(pmgpu_ctx* ctx, double* buf, size_t count) { return allreduce(ctx, buf, count, ncclFloat64); }

int comm_allgather_bytes(pmgpu_ctx* ctx, void* buf, size_t bytes_per_rank) {
    if (ctx->nranks <= 1) return PMGPU_OK;
    if (!ctx->nccl_comm) {
        ctx->set_error("sharded surface normals need the NCCL communicator (pmgpu_comm_init) for their all-gather");
        return PMGPU_ERR_COMM;
    }
    const char* mine = (const char*)buf + (size_t)ctx->rank * bytes_per_rank;
    ncclResult_t r = api().AllGather(mine, buf, bytes_per_rank, ncclInt8, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != ncclSuccess) {
        ctx->set_error(std::string("ncclAllGather: ") + api().GetErrorString(r));
        return PMGPU_ERR_COMM;
    }
    return PMGPU_OK;
}

PeerComm comm_peers(pmgpu_ctx* ctx) {
    PeerComm pc;
    memset(&pc, 0, sizeof(pc));
    pc.rank = ctx->rank;
    pc.nranks = ctx->peer_on ? ctx->nranks : 1;
    if (ctx->peer_on)
        for (int r = 0; r < ctx->nranks; ++r) pc.box[r] = (Mailbox*)ctx->peer_box[r];
    return pc;
}

}  // namespace pm

extern "C" {

int pmgpu_comm_unique_id(void* unique_id_128) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (!unique_id_128 || !pm::api().load()) return PMGPU_ERR_COMM;
    ncclUniqueId id;
    if (pm::api().GetUniqueId(&id) != ncclSuccess) return PMGPU_ERR_COMM;
    memcpy(unique_id_128, &id, sizeof(id));
    return PMGPU_OK;
}

int pmgpu_comm_init(pmgpu_ctx* ctx, const void* unique_id_128, int rank, int nranks) {
    if (!ctx || !unique_id_128 || nranks < 1 || rank < 0 || rank >= nranks) return PMGPU_ERR_BAD_ARG;
    if (!pm::api().load()) {
        ctx->set_error(pm::api().err);
        return PMGPU_ERR_COMM;
    }
    PM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, unique_id_128, sizeof(id));
    ncclComm_t comm;
    ncclResult_t r = pm::api().CommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) {
        ctx->set_error(std::string("ncclCommInitRank: ") + pm::api().GetErrorString(r));
        return PMGPU_ERR_COMM;
    }
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    return PMGPU_OK;
}

int pmgpu_comm_peer_handle(pmgpu_ctx* ctx, void* handle_128) {
    if (!ctx || !handle_128) return PMGPU_ERR_BAD_ARG;
    PM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->mailbox) {
        // cudaMalloc, not the stream-ordered pool: legacy IPC handles only export plain allocations
        PM_CUDA_TRY(ctx, cudaMalloc(&ctx->mailbox, sizeof(pm::Mailbox)));
        PM_CUDA_TRY(ctx, cudaMemset(ctx->mailbox, 0, sizeof(pm::Mailbox)));
        PM_CUDA_TRY(ctx, cudaDeviceSynchronize());
    }
    pm::PeerHandle h;
    memset(&h, 0, sizeof(h));
    h.magic = pm::PEER_MAGIC;
    h.pid = (int32_t)getpid();
    h.device = ctx->device;
    h.ptr = (uint64_t)(uintptr_t)ctx->mailbox;
    PM_CUDA_TRY(ctx, cudaIpcGetMemHandle(&h.ipc, ctx->mailbox));
    memset(handle_128, 0, 128);
    memcpy(handle_128, &h, sizeof(h));
    return PMGPU_OK;
}

int pmgpu_comm_peer_init(pmgpu_ctx* ctx, const void* handles, int rank, int nranks) {
    if (!ctx || !handles || nranks < 1 || rank < 0 || rank >= nranks) return PMGPU_ERR_BAD_ARG;
    if (nranks > PM_MAX_RANKS) {
        ctx->set_error("a registration can be sharded over at most 8 GPUs");
        return PMGPU_ERR_UNSUPPORTED;
    }
    if (!ctx->mailbox) {
        ctx->set_error("pmgpu_comm_peer_init: call pmgpu_comm_peer_handle on this context first");
        return PMGPU_ERR_BAD_ARG;
    }
    PM_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    pm::peer_close(ctx);
    for (int r = 0; r < nranks; ++r) {
        pm::PeerHandle h;
        memcpy(&h, (const char*)handles + 128 * (size_t)r, sizeof(h));
        if (h.magic != pm::PEER_MAGIC) {
            ctx->set_error("pmgpu_comm_peer_init: not a peer handle");
            return PMGPU_ERR_BAD_ARG;
        }
        if (r == rank) {
            if ((void*)(uintptr_t)h.ptr != ctx->mailbox) {
                ctx->set_error("pmgpu_comm_peer_init: handles[rank] is not this context's handle");
                return PMGPU_ERR_BAD_ARG;
            }
            ctx->peer_box[r] = ctx->mailbox;
        } else if (h.pid == (int32_t)getpid()) {
            // same process: the peer's pointer is valid here (UVA); make its device reachable from ours
            if (h.device != ctx->device) {
                int can = 0;
                PM_CUDA_TRY(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, h.device));
                if (!can) {
                    ctx->set_error("pmgpu_comm_peer_init: no peer access between the GPUs of this registration");
                    return PMGPU_ERR_COMM;
                }
                cudaError_t e = cudaDeviceEnablePeerAccess(h.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) PM_CUDA_TRY(ctx, e);
                cudaGetLastError();
            }
            ctx->peer_box[r] = (void*)(uintptr_t)h.ptr;
        } else {
            void* p = nullptr;
            PM_CUDA_TRY(ctx, cudaIpcOpenMemHandle(&p, h.ipc, cudaIpcMemLazyEnablePeerAccess));
            ctx->peer_box[r] = p;
            ctx->peer_opened[r] = true;
        }
    }
    ctx->rank = rank;
    ctx->nranks = nranks;
    ctx->peer_on = nranks > 1;
    return PMGPU_OK;
}

int pmgpu_comm_destroy(pmgpu_ctx* ctx) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    pm::peer_close(ctx);
    if (ctx->mailbox) {
        cudaFree(ctx->mailbox);
        ctx->mailbox = nullptr;
    }
    if (ctx->nccl_comm) {
        pm::api().CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->rank = 0;
    ctx->nranks = 1;
    return PMGPU_OK;
}

}  // extern "C"
