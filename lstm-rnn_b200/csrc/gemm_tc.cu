// tcgen05 / TMEM / TMA GEMM for the time-parallel contractions of the LSTM layer (north_star item 1):
// input-to-gate projection, input-error and weight-gradient GEMMs (reference: the cublasSgemm calls behind
// helpers::Matrix, layers/LstmLayer.cu:774-784, 996-1006, 1038-1043; layers/FeedForwardLayer.cu:152, 196, 206).
//
// Kernel: C[M x N] (row-major, ldc) = A[M x K] * B[N x K]^T, both operands K-major fp32 in global memory.
//   warp 0      TMA producer: cp.async.bulk.tensor 2D boxes (32 floats = 128 B along K) into 128B-swizzled shared tiles
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8 per instruction),
//               accumulating in TMEM (fp32); tcgen05.commit releases shared stages and signals the epilogue
//   warp 2      TMEM allocation / deallocation
//   warps 4-7   epilogue: tcgen05.ld (32 lanes x 32 columns per call) -> registers -> global (optionally += C)
// Precision modes (include/blstm_b200.h):
//   BL_GEMM_FAST    one TF32 MMA per k-step (10-bit mantissa products, fp32 accumulate)          -> <= 2e-3 class
//   BL_GEMM_STRICT  error-compensated 3xTF32: every operand is split in global memory into hi = tf32(x) and
//                   lo = x - hi (exact), and the kernel accumulates hi*hi + hi*lo + lo*hi in TMEM       -> fp32 class
// Operands that are not K-major / 16-byte aligned are first transposed into scratch (bandwidth-bound pass).
// Split-K (grid.z) with ordered partial-sum reduction keeps the result deterministic.
#include "common.cuh"
#include "gemm_tc.cuh"
#include <cuda.h>
#include <cstdint>
#include <cstdlib>

namespace bl {

constexpr int TC_BM = 128, TC_BK = 32, TC_UMMA_K = 8, TC_THREADS = 256;

struct GemmTcParams {
    CUtensorMap tmA, tmAlo, tmB, tmBlo;     // lo maps unused in fast mode
    float *C; int ldc;
    int M, N, K;
    int a_k0, b_k0;                          // element offsets added to the K coordinate of the A / B boxes (time-shifted operands)
    int batches, mt_per_batch;               // grid.y = batches * mt_per_batch: batch b multiplies A rows [b*a_batch_rows, +M) by the same B
    int a_batch_rows; long long c_batch_stride;   // and writes its [M x N] block at C + b*c_batch_stride
    int kblocks_per_split;
    int accumulate;
    float *partial; int ldp;                 // split-K: slice z at partial + z*M*ldp
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// K-major, 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start address >> 4 in [0,14), LBO (unused for one swizzle atom along K) in [16,30), SBO = 1024 B (8 rows x 128 B) >> 4 in [32,46),
// version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(const void *smem_ptr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem_ptr) & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 [4,6)=1, A/B tf32 [7,10)=[10,13)=2, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar)
{ asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------------ the GEMM kernel
template <int BN, bool STRICT, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tf32_tcgen05_kernel(const __grid_constant__ GemmTcParams p)
{
    constexpr int A_BYTES = TC_BM * TC_BK * 4, B_BYTES = BN * TC_BK * 4;
    constexpr int STAGE_BYTES = (STRICT ? 2 : 1) * (A_BYTES + B_BYTES);
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024 B alignment
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *tmem_full = empty + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int batch = blockIdx.y / p.mt_per_batch;
    const int n0 = blockIdx.x * BN, m0 = (blockIdx.y - batch * p.mt_per_batch) * TC_BM;      // m0: row inside the batch's block
    const int a_row0 = batch * p.a_batch_rows + m0;                                            // row in the A tensor map
    const int kb_total = (p.K + TC_BK - 1) / TC_BK;
    const int kb_begin = blockIdx.z * p.kblocks_per_split;
    const int kb_end = min(kb_total, kb_begin + p.kblocks_per_split);
    const int nkb = kb_end - kb_begin;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&p.tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&p.tmB) : "memory");
        if (STRICT) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(&p.tmAlo) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(&p.tmBlo) : "memory");
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // TMEM: BN fp32 accumulator columns (power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES, round = i / STAGES;
                mbar_wait(&empty[s], (round & 1) ^ 1);                 // first pass over the ring succeeds immediately
                uint8_t *st = smem + s * STAGE_BYTES;
                mbar_expect_tx(&full[s], STAGE_BYTES);
                const int kc = (kb_begin + i) * TC_BK;
                tma_load_2d(st, &p.tmA, &full[s], kc + p.a_k0, a_row0);
                tma_load_2d(st + A_BYTES, &p.tmB, &full[s], kc + p.b_k0, n0);
                if (STRICT) {
                    tma_load_2d(st + A_BYTES + B_BYTES, &p.tmAlo, &full[s], kc + p.a_k0, a_row0);
                    tma_load_2d(st + 2 * A_BYTES + B_BYTES, &p.tmBlo, &full[s], kc + p.b_k0, n0);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = make_idesc(TC_BM, BN);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % STAGES, round = i / STAGES;
            mbar_wait(&full[s], round & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint8_t *st = smem + s * STAGE_BYTES;
                const uint64_t a_hi = make_smem_desc(st), b_hi = make_smem_desc(st + A_BYTES);
                const uint64_t a_lo = make_smem_desc(st + A_BYTES + B_BYTES), b_lo = make_smem_desc(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
                for (int kk = 0; kk < TC_BK / TC_UMMA_K; ++kk) {
                    const uint64_t adv = (uint64_t)((kk * TC_UMMA_K * 4) >> 4);      // advance the start address inside the swizzle atom
                    const uint32_t acc0 = (i > 0 || kk > 0) ? 1u : 0u;
                    if (STRICT) {
                        // small terms first, then the leading term
                        umma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, acc0);
                        umma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc, 1u);
                        umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, 1u);
                    } else {
                        umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, acc0);
                    }
                }
                umma_commit(&empty[s]);                                   // frees the stage once these MMAs have read it
                if (i == nkb - 1) umma_commit(tmem_full);                 // accumulator complete
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> global =====
        const int ew = warp & 3;                                          // TMEM lane quarter this warp may access
        const int row = m0 + ew * 32 + lane;
        if (nkb > 0) {
            mbar_wait(tmem_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        float *out = p.partial ? p.partial + ((size_t)blockIdx.z * p.batches + batch) * p.M * p.ldp : p.C + (size_t)batch * p.c_batch_stride;
        const int ldo = p.partial ? p.ldp : p.ldc;
        const bool acc = (!p.partial) && p.accumulate;
        const bool vec_ok = ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            float v[32];
            if (nkb > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)c, v);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.0f;
            }
            if (row < p.M) {
                float *dst = out + (size_t)row * ldo + n0 + c;
                if (vec_ok && n0 + c + 32 <= p.N) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        if (acc) { const float4 old = *reinterpret_cast<const float4 *>(dst + i); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                        *reinterpret_cast<float4 *>(dst + i) = o;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (n0 + c + i < p.N) dst[i] = acc ? dst[i] + v[i] : v[i];
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(BN) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ operand preparation
// One fused, bandwidth-bound pass per operand: bring it into K-major [rows][ld] form (transposing if needed) and, in strict
// mode, split it into hi = tf32_rna(x) (low 13 mantissa bits zero) and lo = x - hi (exact in fp32).
__device__ __forceinline__ void split_tf32(float v, float &hi, float &lo)
{
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi = __uint_as_float(h);
    lo = __fsub_rn(v, hi);
}

// src [rows][K] (lds) -> hi/lo [rows][ldd]; grid (rows, ceil(ldd/512)), 128 threads x 4 consecutive k
template <bool STRICT>
__global__ void prep_rows_kernel(int K, const float *__restrict__ src, size_t lds, float *__restrict__ hi, float *__restrict__ lo, size_t ldd)
{
    const size_t r = blockIdx.x;
    const int k = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
    if ((size_t)k >= ldd) return;
    const float *s = src + r * lds + k;
    float v[4], h[4], l[4];
    const bool vec = ((lds & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (k + 4 <= K);
    if (vec) { const float4 q = *reinterpret_cast<const float4 *>(s); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
    else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (k + i < K) ? s[i] : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { if (STRICT) split_tf32(v[i], h[i], l[i]); else { h[i] = v[i]; l[i] = 0.0f; } }
    *reinterpret_cast<float4 *>(hi + r * ldd + k) = make_float4(h[0], h[1], h[2], h[3]);
    if (STRICT) *reinterpret_cast<float4 *>(lo + r * ldd + k) = make_float4(l[0], l[1], l[2], l[3]);
}

// src [K][rows] (lds) -> hi/lo [rows][ldd]: 32x32 tiles through shared memory, both sides coalesced
template <bool STRICT>
__global__ void prep_transpose_kernel(int K, int rows, const float *__restrict__ src, size_t lds, float *__restrict__ hi,
                                      float *__restrict__ lo, size_t ldd)
{
    __shared__ float t[32][33];
    const int r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int k = k0 + i, r = r0 + threadIdx.x;
        t[i][threadIdx.x] = (k < K && r < rows) ? src[(size_t)k * lds + r] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, k = k0 + threadIdx.x;
        if (r < rows && k < K) {
            const float v = t[threadIdx.x][i];
            if (STRICT) { float h, l; split_tf32(v, h, l); hi[(size_t)r * ldd + k] = h; lo[(size_t)r * ldd + k] = l; }
            else hi[(size_t)r * ldd + k] = v;
        }
    }
}

__global__ void sum_slices_kernel(int M, int N, int nsplit, int batches, const float *__restrict__ partial, int ldp,
                                  float *__restrict__ C, int ldc, long long c_batch_stride, int accumulate)
{
    const size_t total = (size_t)M * N;
    const int b = blockIdx.y;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t r = e / N, c = e % N;
        float s = 0.0f;
        for (int z = 0; z < nsplit; ++z) s += partial[(((size_t)z * batches + b) * M + r) * ldp + c];      // fixed order: deterministic
        float *o = C + (size_t)b * c_batch_stride + r * ldc + c;
        *o = accumulate ? *o + s : s;
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn(bl_ctx *ctx)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
            fail(ctx, "cuTensorMapEncodeTiled is not available from the driver");
            return nullptr;
        }
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// K-major operand [rows][K] with leading dimension ld (floats): box = 32 floats (128 B, one swizzle span) x box_rows
static int make_map(bl_ctx *ctx, CUtensorMap *map, const float *base, int rows, int K, size_t ld, int box_rows)
{
    EncodeTiledFn fn = encode_fn(ctx);
    if (!fn) return 1;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, "cuTensorMapEncodeTiled failed (%d) rows=%d K=%d ld=%zu", (int)r, rows, K, ld);
    return 0;
}

template <int BN, bool STRICT, int STAGES>
static int launch_tc(bl_ctx *ctx, const GemmTcParams &p, dim3 grid)
{
    constexpr int STAGE_BYTES = (STRICT ? 2 : 1) * (TC_BM * TC_BK * 4 + BN * TC_BK * 4);
    const int smem = STAGES * STAGE_BYTES + 1024 + 256;
    auto kernel = gemm_tf32_tcgen05_kernel<BN, STRICT, STAGES>;
    BL_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kernel<<<grid, TC_THREADS, smem, ctx->stream>>>(p);
    BL_LAUNCHED(ctx);
    return 0;
}

size_t tc_operand_ld(int K) { return ((size_t)K + 3) & ~(size_t)3; }

// Fills `out` with a K-major view of src ([rows][K] if kmajor else [K][rows], leading dimension ld_src).  hi/lo are caller-owned
// buffers of rows*tc_operand_ld(K) floats (lo unused in fast mode).  In fast mode an already aligned K-major source is used in place.
int tc_prepare(bl_ctx *ctx, const float *src, int rows, int K, size_t ld_src, bool kmajor, bool strict, float *hi, float *lo, TcOperand *out)
{
    TimedRegion timed(ctx, 0);
    out->rows = rows; out->K = K; out->strict = strict;
    const bool aligned = kmajor && ((ld_src & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (!strict && aligned) { out->hi = src; out->lo = nullptr; out->ld = ld_src; return 0; }
    const size_t ldd = tc_operand_ld(K);
    if (kmajor) {
        dim3 grid(rows, cdiv((int)ldd, 512));
        if (strict) prep_rows_kernel<true><<<grid, 128, 0, ctx->stream>>>(K, src, ld_src, hi, lo, ldd);
        else        prep_rows_kernel<false><<<grid, 128, 0, ctx->stream>>>(K, src, ld_src, hi, lo, ldd);
    } else {
        dim3 grid(cdiv(rows, 32), cdiv(K, 32));
        if (grid.y > 65535) return fail(ctx, "tc_prepare: K too large for the transpose grid");
        if (strict) prep_transpose_kernel<true><<<grid, dim3(32, 8), 0, ctx->stream>>>(K, rows, src, ld_src, hi, lo, ldd);
        else        prep_transpose_kernel<false><<<grid, dim3(32, 8), 0, ctx->stream>>>(K, rows, src, ld_src, hi, lo, ldd);
    }
    BL_LAUNCHED(ctx);
    out->hi = hi; out->lo = strict ? lo : nullptr; out->ld = ldd;
    return 0;
}

// C[M x N] row-major (ldc) (+)= A[a_row0 .. +M][a_k0 .. +K] * B[b_row0 .. +N][b_k0 .. +K]^T on prepared operands
int tc_gemm(bl_ctx *ctx, int M, int N, int K, const TcOperand &A, int a_row0, int a_k0, const TcOperand &B, int b_row0, int b_k0,
            float *C, int ldc, int accumulate)
{
    return tc_gemm_batched(ctx, M, N, K, A, a_row0, a_k0, B, b_row0, b_k0, C, ldc, accumulate, 1, 0, 0);
}

// `batches` products sharing B: batch b uses A rows [a_row0 + b*a_batch_rows, +M) and writes C + b*c_batch_stride
int tc_gemm_batched(bl_ctx *ctx, int M, int N, int K, const TcOperand &A, int a_row0, int a_k0, const TcOperand &B, int b_row0, int b_k0,
                    float *C, int ldc, int accumulate, int batches, int a_batch_rows, long long c_batch_stride)
{
    TimedRegion timed(ctx, 0);
    const bool strict = A.strict;
    if (A.strict != B.strict) return fail(ctx, "tc_gemm: operands prepared for different precision modes");
    const int a_rows_total = (batches - 1) * a_batch_rows + M;
    if (a_row0 + a_rows_total > A.rows || b_row0 + N > B.rows || a_k0 + K > A.K || b_k0 + K > B.K) return fail(ctx, "tc_gemm: sub-view out of range");
    if ((a_k0 | b_k0) & 3) return fail(ctx, "tc_gemm: K offsets must be multiples of 4 floats (TMA box starts are 16-byte aligned)");
    // strict tiles: BLSTM_TC_BN=256 selects 128x256 tiles with a 2-stage ring (less L2 traffic per MMA, half the per-tile overhead)
    static const int bn_strict_env = getenv("BLSTM_TC_BN") ? atoi(getenv("BLSTM_TC_BN")) : 128;
    const int BN_STRICT = (bn_strict_env == 256) ? 256 : 128; constexpr int BN_FAST = 256;
    const int BN = strict ? BN_STRICT : BN_FAST;
    const int tiles = batches * cdiv(M, TC_BM) * cdiv(N, BN);
    const int kb_total = cdiv(K, TC_BK);
    int nsplit = 1;
    if (tiles < ctx->num_sms && kb_total >= 16) {
        nsplit = cdiv(2 * ctx->num_sms, tiles);
        if (nsplit > kb_total / 8) nsplit = kb_total / 8;
        if (nsplit > 64) nsplit = 64;
        if (nsplit < 1) nsplit = 1;
    }
    const int kbs = cdiv(kb_total, nsplit);
    nsplit = cdiv(kb_total, kbs);
    const size_t ldp = tc_operand_ld(N);
    GemmTcParams p;
    p.partial = nullptr;
    if (nsplit > 1) {
        BL_CHECK(ensure_scratch2(ctx, (size_t)nsplit * batches * M * ldp * sizeof(float) + 64));
        p.partial = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(ctx->scratch2) + 15) & ~(uintptr_t)15);
    }
    // the maps cover rows [row0, row0+M) and K extent [0, k0+K): boxes running past either edge are zero-filled
    BL_CHECK(make_map(ctx, &p.tmA, A.hi + (size_t)a_row0 * A.ld, a_rows_total, a_k0 + K, A.ld, TC_BM));
    BL_CHECK(make_map(ctx, &p.tmB, B.hi + (size_t)b_row0 * B.ld, N, b_k0 + K, B.ld, BN));
    if (strict) {
        BL_CHECK(make_map(ctx, &p.tmAlo, A.lo + (size_t)a_row0 * A.ld, a_rows_total, a_k0 + K, A.ld, TC_BM));
        BL_CHECK(make_map(ctx, &p.tmBlo, B.lo + (size_t)b_row0 * B.ld, N, b_k0 + K, B.ld, BN));
    } else { p.tmAlo = p.tmA; p.tmBlo = p.tmB; }
    p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.a_k0 = a_k0; p.b_k0 = b_k0;
    p.kblocks_per_split = kbs; p.accumulate = accumulate; p.ldp = (int)ldp;
    p.batches = batches; p.mt_per_batch = cdiv(M, TC_BM); p.a_batch_rows = a_batch_rows; p.c_batch_stride = c_batch_stride;
    dim3 grid(cdiv(N, BN), batches * cdiv(M, TC_BM), nsplit);
    if (grid.y > 65535) return fail(ctx, "tc_gemm: M too large");
    if (strict && BN_STRICT == 256) BL_CHECK((launch_tc<256, true, 2>(ctx, p, grid)));
    else if (strict) BL_CHECK((launch_tc<128, true, 3>(ctx, p, grid)));
    else        BL_CHECK((launch_tc<BN_FAST, false, 4>(ctx, p, grid)));
    if (nsplit > 1) {
        int blocks = (int)cdivz((size_t)M * N, 256); if (blocks > 2048) blocks = 2048;
        sum_slices_kernel<<<dim3(blocks, batches), 256, 0, ctx->stream>>>(M, N, nsplit, batches, p.partial, (int)ldp, C, ldc, c_batch_stride, accumulate);
        BL_LAUNCHED(ctx);
    }
    return 0;
}

// Generic entry used by bl_gemm_f32: prepares both operands in the context's scratch, then multiplies.
// C[M x N] row-major (ldc) (+)= A[M x K] * B[N x K]^T.  a_kmajor: A is given as [M][K] (lda) else as [K][M] (lda); same for B.
int gemm_tf32_tc(bl_ctx *ctx, int M, int N, int K, const float *A, size_t lda, bool a_kmajor, const float *B, size_t ldb, bool b_kmajor,
                 float *C, int ldc, int accumulate, int mode)
{
    const bool strict = (mode == BL_GEMM_STRICT);
    const size_t ldk = tc_operand_ld(K);
    const size_t a_elems = (size_t)M * ldk, b_elems = (size_t)N * ldk;
    BL_CHECK(ensure_scratch(ctx, 2 * (a_elems + b_elems) * sizeof(float) + 64));
    float *S = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(ctx->scratch) + 15) & ~(uintptr_t)15);
    TcOperand a, b;
    BL_CHECK(tc_prepare(ctx, A, M, K, lda, a_kmajor, strict, S, S + a_elems, &a));
    BL_CHECK(tc_prepare(ctx, B, N, K, ldb, b_kmajor, strict, S + 2 * a_elems, S + 2 * a_elems + b_elems, &b));
    return tc_gemm(ctx, M, N, K, a, 0, 0, b, 0, 0, C, ldc, accumulate);
}

} // namespace bl
