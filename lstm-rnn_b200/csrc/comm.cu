// Data-parallel gradient exchange: one NCCL communicator per process/GPU.  The reference has no
// counterpart (single process, single GPU: currennt/src/main.cpp:526-541).
//
// Two schedules (BLSTM_COMM_MODE):
//   overlap  (default) bl_allreduce_sum_f32 issues the reduction at once on a side stream, ordered after the layer's backward pass, through
//            a communicator split off with ncclConfig_t::maxCTAs = BLSTM_COMM_MAX_CTAS (default 4): the persistent recurrent kernels
//            hold one CTA on each of 144 of the 148 SMs (cooperative launch, one tensor-memory allocation per SM), so an all-reduce
//            only overlaps with the backward pass of the layers below if it fits the SMs they leave free -- uncapped (round 1) it took
//            SMs the next cooperative launch then had to wait for.  The LAST gradient of a step (the first hidden layer's: nothing is
//            left to hide it behind) goes through bl_allreduce_sum_f32_last: full-width communicator, compute stream.  bl_comm_join
//            makes the compute stream wait for the side stream before the weight update.
//   grouped  bl_allreduce_sum_f32 only queues the buffer; bl_comm_join issues ONE grouped all-reduce of all queued buffers on the
//            compute stream.
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
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, ncclConfig_t *) = nullptr;                // NCCL >= 2.18; optional
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
    *(void **)(&g_nccl.CommSplit) = dlsym(lib, "ncclCommSplit");
    g_nccl.lib = lib;
    return 0;
}
} // namespace

struct bl_comm {
    bl_ctx      *ctx;
    ncclComm_t   comm;                                          // full width: grouped mode, the last gradient of a step, statistics
    ncclComm_t   side;                                          // overlap mode: capped to max_ctas CTAs (== comm when NCCL cannot split)
    cudaStream_t stream;
    cudaEvent_t  ready, done;
    int          rank, world;
    bool         pending;
    bool         deferred;                                      // BLSTM_COMM_MODE=grouped: one grouped all-reduce at join time; default: per call, side stream
    int          max_ctas;
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
    { const char *m = getenv("BLSTM_COMM_MODE"); c->deferred = (m && !strcmp(m, "grouped")); }
    { const char *m = getenv("BLSTM_COMM_MAX_CTAS"); c->max_ctas = m ? atoi(m) : 4; }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) { delete c; return bl::fail(ctx, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); }
    c->side = c->comm;
    if (!c->deferred && c->max_ctas > 0 && g_nccl.CommSplit) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.minCTAs = 1;
        cfg.maxCTAs = c->max_ctas;          // stay inside the SMs the persistent recurrent kernels leave free
        r = g_nccl.CommSplit(c->comm, 0, rank, &c->side, &cfg);
        if (r != ncclSuccess) { g_nccl.CommDestroy(c->comm); delete c; return bl::fail(ctx, "ncclCommSplit failed: %s", g_nccl.GetErrorString(r)); }
    }
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
    cudaStreamSynchronize(c->ctx->stream);
    if (c->side != c->comm) g_nccl.CommDestroy(c->side);
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
    BL_NCCL(ctx, g_nccl.AllReduce(buf, buf, count, ncclFloat, ncclSum, c->side, c->stream));
    ctx->launches++;
    c->pending = true;
    return 0;
}

int bl_allreduce_sum_f32_last(bl_comm *c, float *buf, size_t count)
{
    bl_ctx *ctx = c->ctx;
    if (!count) return 0;
    if (c->deferred) { c->queue.emplace_back(buf, count); c->pending = true; return 0; }
    BL_NCCL(ctx, g_nccl.AllReduce(buf, buf, count, ncclFloat, ncclSum, c->comm, ctx->stream));
    ctx->launches++;
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
