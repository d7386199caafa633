// Context, memory and error plumbing of the C ABI (include/blstm_b200.h).
#include "common.cuh"
#include <cstring>
#include <cstdlib>

namespace bl {
thread_local std::string g_err;

int fail(bl_ctx *ctx, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_err = buf;
    return 1;
}

int ensure_scratch(bl_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->scratch_bytes) return 0;
    if (ctx->scratch) { BL_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); BL_CUDA(ctx, cudaFree(ctx->scratch)); ctx->scratch = nullptr; ctx->scratch_bytes = 0; }
    size_t want = bytes + bytes / 4;
    BL_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
    ctx->scratch_bytes = want;
    return 0;
}
static cudaEvent_t get_event(bl_ctx *ctx)
{
    if (!ctx->tpool.empty()) { cudaEvent_t e = ctx->tpool.back(); ctx->tpool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}

TimedRegion::TimedRegion(bl_ctx *c, int k) : ctx(c), cls(k), a(nullptr), b(nullptr), on(c->timing)
{
    if (on) { a = get_event(ctx); b = get_event(ctx); cudaEventRecord(a, ctx->stream); }
}
TimedRegion::~TimedRegion()
{
    if (on) { cudaEventRecord(b, ctx->stream); ctx->tev[cls].emplace_back(a, b); }
}
int ensure_scratch2(bl_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->scratch2_bytes) return 0;
    if (ctx->scratch2) { BL_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); BL_CUDA(ctx, cudaFree(ctx->scratch2)); ctx->scratch2 = nullptr; ctx->scratch2_bytes = 0; }
    size_t want = bytes + bytes / 4;
    BL_CUDA(ctx, cudaMalloc(&ctx->scratch2, want));
    ctx->scratch2_bytes = want;
    return 0;
}
} // namespace bl

extern "C" {

int bl_ctx_timing_enable(bl_ctx *ctx, int on) { ctx->timing = (on != 0); return 0; }

int bl_ctx_timing_read(bl_ctx *ctx, double *ms4, long *count4)
{
    BL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < BL_TIMING_CLASSES; ++k) {
        double total = 0;
        for (auto &pr : ctx->tev[k]) {
            float ms = 0; cudaEventElapsedTime(&ms, pr.first, pr.second); total += ms;
            ctx->tpool.push_back(pr.first); ctx->tpool.push_back(pr.second);
        }
        if (ms4) ms4[k] = total;
        if (count4) count4[k] = (long)ctx->tev[k].size();
        ctx->tev[k].clear();
    }
    return 0;
}

int bl_ctx_create(int device, void *stream, bl_ctx **out)
{
    if (!out) return bl::fail(nullptr, "bl_ctx_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return bl::fail(nullptr, "bl_ctx_create: no CUDA device (%s); this library has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count) return bl::fail(nullptr, "bl_ctx_create: bad device %d of %d", device, count);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bl::fail(nullptr, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bl::fail(nullptr, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return bl::fail(nullptr, "bl_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    bl_ctx *ctx = new bl_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    ctx->gemm_mode = BL_GEMM_STRICT;
    { const char *e = getenv("BLSTM_GEMM_BACKEND"); ctx->gemm_backend = e ? atoi(e) : 0; }
    ctx->launches = 0;
    ctx->timing = false;
    ctx->scratch = nullptr;
    ctx->scratch_bytes = 0;
    ctx->scratch2 = nullptr;
    ctx->scratch2_bytes = 0;
    for (int i = 0; i < 8; ++i) { ctx->up_ev[i] = nullptr; ctx->up_ticket[i] = 0; }
    ctx->up_next = 0;
    if (stream) { ctx->stream = (cudaStream_t)stream; ctx->own_stream = false; }
    else {
        if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
            delete ctx; return bl::fail(nullptr, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        ctx->own_stream = true;
    }
    *out = ctx;
    return 0;
}

void bl_ctx_destroy(bl_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->scratch2) cudaFree(ctx->scratch2);
    for (int i = 0; i < 8; ++i) if (ctx->up_ev[i]) cudaEventDestroy(ctx->up_ev[i]);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *bl_last_error(const bl_ctx *ctx) { return ctx ? ctx->err.c_str() : bl::g_err.c_str(); }

int bl_sync(bl_ctx *ctx) { BL_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); return 0; }

int bl_ctx_set_gemm_mode(bl_ctx *ctx, int mode)
{
    if (mode != BL_GEMM_STRICT && mode != BL_GEMM_FAST) return bl::fail(ctx, "bad gemm mode %d", mode);
    ctx->gemm_mode = mode;
    return 0;
}

int bl_ctx_set_gemm_backend(bl_ctx *ctx, int backend)
{
    if (backend < 0 || backend > 2) return bl::fail(ctx, "bad gemm backend %d", backend);
    ctx->gemm_backend = backend;
    return 0;
}

int bl_ctx_num_sms(const bl_ctx *ctx) { return ctx->num_sms; }
int bl_device_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }
long bl_ctx_launch_count(const bl_ctx *ctx) { return ctx->launches; }

int bl_malloc(bl_ctx *ctx, void **ptr, size_t bytes)
{
    BL_CUDA(ctx, cudaSetDevice(ctx->device));
    BL_CUDA(ctx, cudaMalloc(ptr, bytes ? bytes : 1));
    return 0;
}
int bl_free(bl_ctx *ctx, void *ptr) { if (ptr) BL_CUDA(ctx, cudaFree(ptr)); return 0; }
int bl_memset(bl_ctx *ctx, void *ptr, int value, size_t bytes) { if (bytes) BL_CUDA(ctx, cudaMemsetAsync(ptr, value, bytes, ctx->stream)); return 0; }
int bl_memcpy_h2d(bl_ctx *ctx, void *dst, const void *src, size_t bytes) { if (bytes) BL_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream)); return 0; }
int bl_memcpy_d2h(bl_ctx *ctx, void *dst, const void *src, size_t bytes) { if (bytes) BL_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream)); return 0; }
int bl_memcpy_d2d(bl_ctx *ctx, void *dst, const void *src, size_t bytes) { if (bytes) BL_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream)); return 0; }
int bl_memcpy2d_h2d(bl_ctx *ctx, void *dst, size_t dp, const void *src, size_t sp, size_t rb, size_t rows)
{
    if (rb && rows) BL_CUDA(ctx, cudaMemcpy2DAsync(dst, dp, src, sp, rb, rows, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
int bl_memcpy2d_d2h(bl_ctx *ctx, void *dst, size_t dp, const void *src, size_t sp, size_t rb, size_t rows)
{
    if (rb && rows) BL_CUDA(ctx, cudaMemcpy2DAsync(dst, dp, src, sp, rb, rows, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}
int bl_malloc_host(bl_ctx *ctx, void **ptr, size_t bytes)
{
    if (ctx) BL_CUDA(ctx, cudaSetDevice(ctx->device));      // may be called from a prefetch thread that has no current device yet
    BL_CUDA(ctx, cudaMallocHost(ptr, bytes ? bytes : 1));
    return 0;
}
int bl_free_host(bl_ctx *ctx, void *ptr) { if (ptr) BL_CUDA(ctx, cudaFreeHost(ptr)); return 0; }

int bl_upload_mark(bl_ctx *ctx, unsigned long long *ticket)
{
    const unsigned long long t = ++ctx->up_next;
    const int slot = (int)(t % 8);
    if (!ctx->up_ev[slot]) BL_CUDA(ctx, cudaEventCreateWithFlags(&ctx->up_ev[slot], cudaEventDisableTiming));
    else if (ctx->up_ticket[slot]) BL_CUDA(ctx, cudaEventSynchronize(ctx->up_ev[slot]));     // the ring wrapped: the old ticket is done after this
    BL_CUDA(ctx, cudaEventRecord(ctx->up_ev[slot], ctx->stream));
    ctx->up_ticket[slot] = t;
    *ticket = t;
    return 0;
}

int bl_upload_wait(bl_ctx *ctx, unsigned long long ticket)
{
    if (!ticket) return 0;
    const int slot = (int)(ticket % 8);
    if (ctx->up_ticket[slot] != ticket) return 0;      // slot reused: the ticket was waited for when it was overwritten
    BL_CUDA(ctx, cudaEventSynchronize(ctx->up_ev[slot]));
    return 0;
}

} // extern "C"
