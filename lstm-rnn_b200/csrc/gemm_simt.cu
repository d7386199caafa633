// Strict-fp32 SIMT GEMM: the drop-in for cublas::multiplyMatrices (reference helpers/cublas.cu:60-88).
//
// C[m x n] (column-major, ldc) = op(A) * op(B) (+ C).  128x128x16 CTA tiles, 256 threads, 8x8 register
// tiles read as float4 from k-major shared tiles, register-staged double buffering, deterministic
// split-K (partials + ordered reduce) for the weight-gradient shapes (tiny m x n, k = T*S).
// Used for every contraction in strict mode until the tcgen05 path takes the aligned TN shapes.
#include "common.cuh"
#include "gemm_tc.cuh"

namespace bl {

constexpr int BM = 128, BN = 128, BK = 16, GT = 256;
constexpr int LDS_A = BM + 4, LDS_B = BN + 4;      // +4 floats: keeps 16 B alignment, breaks the store conflicts

template <bool KCONTIG>
__device__ __forceinline__ void load_tile(float (&reg)[8], const float *__restrict__ M, int ld, int row0, int nrows,
                                          int k0, int kend, int tid)
{
    // tile element (r, kk): r in [0,128) over the m/n index, kk in [0,16)
    if (KCONTIG) {              // M[(k0+kk) + (row0+r)*ld]
        const int kk = tid & 15, rb = tid >> 4;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = rb + 16 * q;
            const bool ok = (row0 + r < nrows) && (k0 + kk < kend);
            reg[q] = ok ? __ldg(M + (size_t)(row0 + r) * ld + (k0 + kk)) : 0.0f;
        }
    } else {                    // M[(row0+r) + (k0+kk)*ld]
        const int r = tid & 127, kb = tid >> 7;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int kk = kb + 2 * q;
            const bool ok = (row0 + r < nrows) && (k0 + kk < kend);
            reg[q] = ok ? __ldg(M + (size_t)(k0 + kk) * ld + (row0 + r)) : 0.0f;
        }
    }
}

template <bool KCONTIG, int LDS>
__device__ __forceinline__ void store_tile(const float (&reg)[8], float *sm, int tid)
{
    if (KCONTIG) {
        const int kk = tid & 15, rb = tid >> 4;
#pragma unroll
        for (int q = 0; q < 8; ++q) sm[kk * LDS + rb + 16 * q] = reg[q];
    } else {
        const int r = tid & 127, kb = tid >> 7;
#pragma unroll
        for (int q = 0; q < 8; ++q) sm[(kb + 2 * q) * LDS + r] = reg[q];
    }
}

// A_K: op(A)(i,kk) is contiguous in kk (transA=1); B_K: op(B)(kk,j) is contiguous in kk (transB=0)
template <bool A_K, bool B_K>
__global__ void __launch_bounds__(GT)
gemm_f32_kernel(int m, int n, int k, const float *__restrict__ A, int lda, const float *__restrict__ B, int ldb,
                float *__restrict__ C, int ldc, int accumulate, int klen, float *__restrict__ partial)
{
    __shared__ __align__(16) float As[2][BK * LDS_A];
    __shared__ __align__(16) float Bs[2][BK * LDS_B];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int i0 = blockIdx.x * BM, j0 = blockIdx.y * BN;
    const int kbeg = blockIdx.z * klen, kend = min(k, kbeg + klen);

    float acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = 0.0f;

    float ra[8], rb[8];
    load_tile<A_K>(ra, A, lda, i0, m, kbeg, kend, tid);
    load_tile<B_K>(rb, B, ldb, j0, n, kbeg, kend, tid);
    store_tile<A_K, LDS_A>(ra, As[0], tid);
    store_tile<B_K, LDS_B>(rb, Bs[0], tid);
    __syncthreads();

    int buf = 0;
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        const bool more = (k0 + BK < kend);
        if (more) {
            load_tile<A_K>(ra, A, lda, i0, m, k0 + BK, kend, tid);
            load_tile<B_K>(rb, B, ldb, j0, n, k0 + BK, kend, tid);
        }
        const float *as = As[buf], *bs = Bs[buf];
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(as + kk * LDS_A + tx * 4);
            const float4 a1 = *reinterpret_cast<const float4 *>(as + kk * LDS_A + 64 + tx * 4);
            const float4 b0 = *reinterpret_cast<const float4 *>(bs + kk * LDS_B + ty * 4);
            const float4 b1 = *reinterpret_cast<const float4 *>(bs + kk * LDS_B + 64 + ty * 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
        }
        if (more) {
            store_tile<A_K, LDS_A>(ra, As[buf ^ 1], tid);
            store_tile<B_K, LDS_B>(rb, Bs[buf ^ 1], tid);
        }
        __syncthreads();
        buf ^= 1;
    }

    // epilogue: rows i contiguous in memory
    float *out = partial ? partial + (size_t)blockIdx.z * m * n : C;
    const int ldo = partial ? m : ldc;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const int j = j0 + (b < 4 ? ty * 4 + b : 64 + ty * 4 + (b - 4));
        if (j >= n) continue;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int i = i0 + (a < 4 ? tx * 4 + a : 64 + tx * 4 + (a - 4));
            if (i >= m) continue;
            float *p = out + (size_t)j * ldo + i;
            *p = (!partial && accumulate) ? *p + acc[a][b] : acc[a][b];
        }
    }
}

__global__ void splitk_reduce_kernel(int m, int n, int nsplit, const float *__restrict__ partial,
                                     float *__restrict__ C, int ldc, int accumulate)
{
    const size_t total = (size_t)m * n;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        float s = 0.0f;
        for (int z = 0; z < nsplit; ++z) s += partial[(size_t)z * total + e];     // fixed order: deterministic
        const int i = (int)(e % m), j = (int)(e / m);
        float *p = C + (size_t)j * ldc + i;
        *p = accumulate ? *p + s : s;
    }
}

bool tc_wanted(const bl_ctx *ctx, int m, int n, int k)
{
    const double macs = (double)m * n * k;
    const bool big = macs >= 64.0 * 1024 * 1024 && k >= 64 && m >= 32 && n >= 32;
    return ctx->gemm_backend == 2 || (ctx->gemm_backend == 0 && big);
}

int gemm_f32_simt(bl_ctx *ctx, int transA, int transB, int m, int n, int k,
                  const float *A, int lda, const float *B, int ldb, float *C, int ldc, int accumulate)
{
    if (m <= 0 || n <= 0) return 0;
    TimedRegion timed(ctx, 0);
    if (k <= 0) {       // empty contraction: C = 0 (or unchanged)
        if (!accumulate) {
            splitk_reduce_kernel<<<cdiv(m * n, 256) > 1024 ? 1024 : cdiv(m * n, 256), 256, 0, ctx->stream>>>(m, n, 0, nullptr, C, ldc, 0);
            BL_LAUNCHED(ctx);
        }
        return 0;
    }
    const int tiles = cdiv(m, BM) * cdiv(n, BN);
    int nsplit = 1;
    if (tiles < ctx->num_sms && k >= 1024) {
        nsplit = cdiv(2 * ctx->num_sms, tiles);
        const int maxsplit = k / 256;
        if (nsplit > maxsplit) nsplit = maxsplit;
        if (nsplit > 128) nsplit = 128;
        if (nsplit < 1) nsplit = 1;
    }
    int klen = cdiv(cdiv(k, nsplit), BK) * BK;
    nsplit = cdiv(k, klen);
    float *partial = nullptr;
    if (nsplit > 1) {
        BL_CHECK(ensure_scratch(ctx, (size_t)nsplit * m * n * sizeof(float)));
        partial = ctx->scratch;
    }
    dim3 grid(cdiv(m, BM), cdiv(n, BN), nsplit);
    if (grid.y > 65535 || grid.z > 65535) return fail(ctx, "gemm: grid too large (n=%d)", n);
#define BL_GEMM_LAUNCH(AK, BKC) gemm_f32_kernel<AK, BKC><<<grid, GT, 0, ctx->stream>>>(m, n, k, A, lda, B, ldb, C, ldc, accumulate, klen, partial)
    if (transA && !transB)       BL_GEMM_LAUNCH(true, true);
    else if (!transA && !transB) BL_GEMM_LAUNCH(false, true);
    else if (!transA && transB)  BL_GEMM_LAUNCH(false, false);
    else return fail(ctx, "gemm: (transA,transB)=(1,1) is not implemented (as in helpers/Matrix.cu:248)");
#undef BL_GEMM_LAUNCH
    BL_LAUNCHED(ctx);
    if (nsplit > 1) {
        int blocks = cdiv(m * n, 256); if (blocks > 2048) blocks = 2048;
        splitk_reduce_kernel<<<blocks, 256, 0, ctx->stream>>>(m, n, nsplit, partial, C, ldc, accumulate);
        BL_LAUNCHED(ctx);
    }
    return 0;
}

} // namespace bl

extern "C" int bl_gemm_f32(bl_ctx *ctx, int transA, int transB, int m, int n, int k,
                           const float *A, int lda, const float *B, int ldb, float *C, int ldc,
                           int accumulate, int mode)
{
    if (!ctx) return bl::fail(nullptr, "bl_gemm_f32: ctx is NULL");
    if (m < 0 || n < 0 || k < 0) return bl::fail(ctx, "bl_gemm_f32: negative dimension");
    const int rowsA = transA ? k : m, rowsB = transB ? n : k;
    if (lda < (rowsA > 1 ? rowsA : 1) || ldb < (rowsB > 1 ? rowsB : 1) || ldc < (m > 1 ? m : 1))
        return bl::fail(ctx, "bl_gemm_f32: leading dimension too small (lda=%d ldb=%d ldc=%d)", lda, ldb, ldc);
    if (transA && transB) return bl::fail(ctx, "gemm: (transA,transB)=(1,1) is not implemented (as in helpers/Matrix.cu:248)");
    if (mode != BL_GEMM_STRICT && mode != BL_GEMM_FAST) return bl::fail(ctx, "bl_gemm_f32: bad mode %d", mode);
    // tensor-core path for the large time-parallel contractions; the tiny ones (a few output tiles, short K) stay on FFMA
    if (bl::tc_wanted(ctx, m, n, k)) {
        // column-major C[m x n] == row-major C'[n x m] = opB^T[n x k] * (opA^T[m x k])^T
        return bl::gemm_tf32_tc(ctx, n, m, k, B, (size_t)ldb, /*a_kmajor=*/!transB, A, (size_t)lda, /*b_kmajor=*/transA != 0,
                                C, ldc, accumulate, mode);
    }
    return bl::gemm_f32_simt(ctx, transA, transB, m, n, k, A, lda, B, ldb, C, ldc, accumulate);
}
