// Data-parallel gradient exchange: one NCCL communicator per process/GPU.  The reference has no
// counterpart (single process, single GPU: currennt/src/main.cpp:526-541).
//
// Default schedule: bl_allreduce_sum_f32 only queues the buffer; bl_comm_join issues ONE grouped all-reduce of all queued
// buffers on the compute stream.  Measured on B200: the persistent recurrent kernels occupy 144 of 148 SMs with one CTA
// each, so an all-reduce launched on a side stream during the backward pass does not overlap -- it competes with the
// cooperative launches for SMs (2 GPUs: 15.95 ms/step overlapped vs 15.58 ms grouped at the end, 15.33 ms single GPU).
// BLSTM_COMM_MODE=overlap restores the per-layer side-stream schedule.
//
// NCCL is resolved at run time with dlopen("libnccl.so.2"): inside a Python process that already imported torch
// this binds to the NCCL torch loaded (one NCCL per process); in a plain C++ host it binds to the system library.
// Only the stable core API (ncclGetUniqueId / ncclCommInitRank / ncclAllReduce / ncclCommDestroy) is used.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl(bl_ctx *ctx)
{
    if (g_nccl.lib) return 0;
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return bl::fail(ctx, "cannot load libnccl.so.2: %s", dlerror());
#define BL_SYM(field, name) \
    *(void **)(&g_nccl.field) = dlsym(lib, name); \
    if (!g_nccl.field) return bl::fail(ctx, "libnccl.so.2 lacks %s", name);
    BL_SYM(GetUniqueId, "ncclGetUniqueId")
    BL_SYM(CommInitRank, "ncclCommInitRank")
    BL_SYM(AllReduce, "ncclAllReduce")
    BL_SYM(CommDestroy, "ncclCommDestroy")
    BL_SYM(GroupStart, "ncclGroupStart")
    BL_SYM(GroupEnd, "ncclGroupEnd")
    BL_SYM(GetErrorString, "ncclGetErrorString")
#undef BL_SYM
    g_nccl.lib = lib;
    return 0;
}
} // namespace

struct bl_comm {
    bl_ctx      *ctx;
    ncclComm_t   comm;
    cudaStream_t stream;
    cudaEvent_t  ready, done;
    int          rank, world;
    bool         pending;
    bool         deferred;                                      // default: one grouped all-reduce at join time; BLSTM_COMM_MODE=overlap: per call, side stream
    std::vector<std::pair<float *, size_t>> queue;
};

#define BL_NCCL(ctx, expr)                                                                      \
    do {                                                                                        \
        ncclResult_t r__ = (expr);                                                              \
        if (r__ != ncclSuccess)                                                                 \
            return bl::fail((ctx), "%s failed: %s", #expr, g_nccl.GetErrorString(r__));         \
    } while (0)

extern "C" {

int bl_comm_unique_id(void *id128)
{
    BL_CHECK(load_nccl(nullptr));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    BL_NCCL(nullptr, g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int bl_comm_create(bl_ctx *ctx, int rank, int world, const void *id128, bl_comm **out)
{
    if (!ctx || !out) return bl::fail(ctx, "bl_comm_create: NULL argument");
    *out = nullptr;
    BL_CHECK(load_nccl(ctx));
    BL_CUDA(ctx, cudaSetDevice(ctx->device));
    bl_comm *c = new bl_comm();
    c->ctx = ctx; c->rank = rank; c->world = world; c->pending = false;
    { const char *m = getenv("BLSTM_COMM_MODE"); c->deferred = !(m && !strcmp(m, "overlap")); }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) { delete c; return bl::fail(ctx, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); }
    cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&c->ready, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming);
    *out = c;
    return 0;
}

void bl_comm_info(const bl_comm *comm, int *rank, int *world)
{
    if (rank) *rank = comm ? comm->rank : 0;
    if (world) *world = comm ? comm->world : 1;
}

void bl_comm_destroy(bl_comm *c)
{
    if (!c) return;
    cudaStreamSynchronize(c->stream);
    g_nccl.CommDestroy(c->comm);
    cudaEventDestroy(c->ready); cudaEventDestroy(c->done);
    cudaStreamDestroy(c->stream);
    delete c;
}

int bl_allreduce_sum_f32(bl_comm *c, float *buf, size_t count)
{
    bl_ctx *ctx = c->ctx;
    if (!count) return 0;
    if (c->deferred) { c->queue.emplace_back(buf, count); c->pending = true; return 0; }
    // the reduction may start once everything enqueued so far on the compute stream (the layer's backward) is done
    BL_CUDA(ctx, cudaEventRecord(c->ready, ctx->stream));
    BL_CUDA(ctx, cudaStreamWaitEvent(c->stream, c->ready, 0));
    BL_NCCL(ctx, g_nccl.AllReduce(buf, buf, count, ncclFloat, ncclSum, c->comm, c->stream));
    ctx->launches++;
    c->pending = true;
    return 0;
}

int bl_comm_join(bl_comm *c)
{
    bl_ctx *ctx = c->ctx;
    if (!c->pending) return 0;
    if (c->deferred) {
        BL_NCCL(ctx, g_nccl.GroupStart());
        for (auto &q : c->queue) BL_NCCL(ctx, g_nccl.AllReduce(q.first, q.first, q.second, ncclFloat, ncclSum, c->comm, ctx->stream));
        BL_NCCL(ctx, g_nccl.GroupEnd());
        ctx->launches++;
        c->queue.clear(); c->pending = false;
        return 0;
    }
    BL_CUDA(ctx, cudaEventRecord(c->done, c->stream));
    BL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, c->done, 0));
    c->pending = false;
    return 0;
}

} // extern "C"
