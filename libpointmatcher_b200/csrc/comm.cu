// comm.cu — the only cross-GPU exchange of the path: when the reading is sharded over ranks
// against a replicated reference, every rank needs the same global quantile and the same normal
// equations, so the select histograms (2048 x u32) and the reduced sums (<= 42 x f64) are
// all-reduced in-stream between the two kernels that produce and consume them.  Messages are
// <= 8 KB: latency-bound, NVSwitch makes the algorithm choice irrelevant, so this is plain NCCL
// in the context's stream.  NCCL is resolved at run time (dlopen) so that a process that already
// carries an NCCL (e.g. torch's bundled one) shares it and single-GPU users need none.
#include <dlfcn.h>
#include <string.h>
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
        GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !GetErrorString) {
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
    ncclResult_t r = api().AllReduce(buf, buf, count, type, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != ncclSuccess) {
        ctx->set_error(std::string("ncclAllReduce: ") + api().GetErrorString(r));
        return PMGPU_ERR_COMM;
    }
    return PMGPU_OK;
}

}  // namespace

int comm_allreduce_u32(pmgpu_ctx* ctx, unsigned* buf, size_t count) { return allreduce(ctx, buf, count, ncclUint32); }
int comm_allreduce_f64(pmgpu_ctx* ctx, double* buf, size_t count) { return allreduce(ctx, buf, count, ncclFloat64); }

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

int pmgpu_comm_destroy(pmgpu_ctx* ctx) {
    if (!ctx) return PMGPU_ERR_BAD_ARG;
    if (ctx->nccl_comm) {
        pm::api().CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->rank = 0;
    ctx->nranks = 1;
    return PMGPU_OK;
}

}  // extern "C"
