// Probe for the fp16 two-term variant of the tensor-memory recurrent step (successor of tcgen05_ts_step.cu): the 128-row x K=256
// weight slice sits in TMEM as TWO fp16 matrices and each step is 32 kind::f16 MMAs instead of 32 tf32 + 16 bf16:
//     hi = fp16_rn(x), lo' = fp16_rn((x - hi) * 2^11)          (x - hi is exact; the scale keeps lo' out of fp16's subnormals)
//     D[:, 0:N)   = W_hi  * h_hi                                one N = 2N MMA per k-step: B rows [0, N) = h_hi, [N, 2N) = h_lo'
//     D[:, N:2N)  = W_hi  * h_lo'  +  W_lo' * h_hi              second MMA (N = N) accumulates into the upper column half
//     W h ~= D[:, 0:N) + 2^-11 * D[:, N:2N)
// fp16 and tf32 both carry 11 significant bits, so the error budget is the tf32 scheme's (dropped lo*lo term, 2^-22), but an MMA
// covers K = 16 instead of 8, the A operand needs 128 + 128 instead of 256 + 128 TMEM columns and the B tile is half as large.
// Reports accuracy against fp64 and cycles per step next to the tf32 + bf16 scheme of tcgen05_ts_step.cu.
//
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcgen05_f16_step tcgen05_f16_step.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

constexpr int M = 128, N = 16, N2 = 32, K = 256;
constexpr int NT = 256;
constexpr int COL_AHI = 0, COL_ALO = 128, COL_D = 256, TMEM_COLS = 512;
constexpr int KB_BYTES = N2 * 128;                  // one K-block (64 fp16 = 128 B per row) of the [h_hi | h_lo'] tile: 32 rows
constexpr int B_BYTES = (K / 64) * KB_BYTES;
constexpr float LO_SCALE = 2048.0f, LO_UNSCALE = 1.0f / 2048.0f;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool mbar_wait_bounded(uint64_t *bar, uint32_t parity)
{
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// K-major, 128B-swizzled shared-memory matrix descriptor (same encoding as csrc/gemm_tc.cu): SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc(const void *p)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(p) & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// D fp32, A and B fp16 (format 0 of kind::f16), both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t make_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(M >> 4) << 24); }

__device__ __forceinline__ void mma_ts_f16(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t *bar)
{ asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16])
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// x = hi + lo'/2^11 up to 2^-22 |x|
__device__ __forceinline__ void split16(float x, uint32_t &hi, uint32_t &lo)
{
    const __half h = __float2half_rn(x);
    hi = (uint32_t)__half_as_ushort(h);
    lo = (uint32_t)__half_as_ushort(__float2half_rn(__fmul_rn(__fsub_rn(x, __half2float(h)), LO_SCALE)));
}

// byte offset of element (row n, k) of the K-major SWIZZLE_128B tile of N2 rows: 128-byte rows (64 fp16), 8-row atoms of 1024 B,
// 16-byte chunks XORed with the row index; one [N2 x 128 B] box per K-block
__device__ __forceinline__ int off_f16(int n, int k)
{ const int kb = k >> 6, kin = k & 63, c = kin >> 3, e = kin & 7, r = n & 7; return kb * KB_BYTES + (n >> 3) * 1024 + r * 128 + ((c ^ r) << 4) + e * 2; }

struct Params {
    const float *W;        // [M][K]
    const float *h;        // [N][K]
    float *D;              // [M][N]
    long long *cyc;        // [4]
    int *err;
    int steps;
    int terms;             // 1: hi*hi   2: + hi*lo'   3: + lo'*hi
    int nrep;              // timing aid: repeat the MMA batch nrep times (same operands) to separate per-MMA cost from fixed cost
};

__global__ void __launch_bounds__(NT, 1) step_kernel(const Params p)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *B = smem;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + B_BYTES);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot;

    // ---- weights into TMEM, once: lane = row m, column = k pair (even k in the low half)
    if (warp < 4) {
        const int m = warp * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        const float *w = p.W + (size_t)m * K;
        for (int c0 = 0; c0 < K / 2; c0 += 8) {
            uint32_t rh[8], rl[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t h0, l0, h1, l1;
                split16(w[2 * (c0 + i)], h0, l0); split16(w[2 * (c0 + i) + 1], h1, l1);
                rh[i] = h0 | (h1 << 16); rl[i] = l0 | (l1 << 16);
            }
            tmem_st8(lane_base + COL_AHI + c0, rh);
            tmem_st8(lane_base + COL_ALO + c0, rl);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < B_BYTES / 4; i += NT) reinterpret_cast<uint32_t *>(B)[i] = 0u;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const uint32_t idesc32 = make_idesc(N2), idesc16 = make_idesc(N);
    long long c_total = 0, c_write = 0, c_mma = 0, c_ld = 0;
    float keep = 0.0f;
    bool ok = true;
    for (int s = 0; s < p.steps && ok; ++s) {
        const long long t0 = clock64();
        // ---- B operand: previous-step vector of the CTA's sequences, split and written in the UMMA layout (rows n and N + n)
        for (int i = tid; i < N * K / 4; i += NT) {
            const int n = i / (K / 4), k = (i - n * (K / 4)) * 4;
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(p.h + (size_t)n * K + k));
            uint32_t h0, l0, h1, l1, h2, l2, h3, l3;
            split16(v.x, h0, l0); split16(v.y, h1, l1); split16(v.z, h2, l2); split16(v.w, h3, l3);
            *reinterpret_cast<uint2 *>(B + off_f16(n, k)) = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
            *reinterpret_cast<uint2 *>(B + off_f16(N + n, k)) = make_uint2(l0 | (l1 << 16), l2 | (l3 << 16));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        const long long t1 = clock64();
        if (warp == 1) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                for (int rep = 0; rep < p.nrep; ++rep) {
                    // 16 k per MMA = 8 TMEM columns = 32 B inside the swizzle atom
                    for (int ks = 0; ks < K / 16; ++ks) {
                        const uint64_t bd = make_desc(B + (ks >> 2) * KB_BYTES) + (uint64_t)((ks & 3) * 2);
                        if (p.terms >= 2) mma_ts_f16(tmem + COL_D, tmem + COL_AHI + ks * 8, bd, idesc32, (rep | ks) ? 1u : 0u);
                        else              mma_ts_f16(tmem + COL_D, tmem + COL_AHI + ks * 8, bd, idesc16, (rep | ks) ? 1u : 0u);
                    }
                    if (p.terms >= 3)
                        for (int ks = 0; ks < K / 16; ++ks)
                            mma_ts_f16(tmem + COL_D + N, tmem + COL_ALO + ks * 8, make_desc(B + (ks >> 2) * KB_BYTES) + (uint64_t)((ks & 3) * 2), idesc16, 1u);
                }
                commit(bar);
            }
            __syncwarp();
        }
        long long t2 = 0, t3 = 0;
        if (warp < 4) {
            ok = mbar_wait_bounded(bar, (uint32_t)(s & 1));
            t2 = clock64();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float v[16], l[16];
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + COL_D, v);
            tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + COL_D + N, l);
#pragma unroll
            for (int i = 0; i < 16; ++i) { if (p.terms >= 2) v[i] = __fadd_rn(v[i], __fmul_rn(l[i], LO_UNSCALE)); keep += v[i]; }
            if (s == p.steps - 1 && blockIdx.x == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) p.D[(size_t)(warp * 32 + lane) * N + i] = v[i];
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            t3 = clock64();
        }
        ok = __syncthreads_and(ok);
        const long long t4 = clock64();
        if (tid == 0) { c_total += t4 - t0; c_write += t1 - t0; c_mma += t2 - t1; c_ld += t3 - t2; }
    }
    if (!ok && tid == 0) atomicExch(p.err, 1);
    if (tid == 0 && blockIdx.x == 0 && p.steps > 0) {
        p.cyc[0] = c_total / p.steps; p.cyc[1] = c_write / p.steps; p.cyc[2] = c_mma / p.steps; p.cyc[3] = c_ld / p.steps;
    }
    if (keep == 12345.678f) p.D[0] = keep;
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TMEM_COLS) : "memory");
    }
}

int main()
{
    float *dW, *dh, *dD; long long *cyc; int *err;
    cudaMalloc(&dW, (size_t)M * K * 4); cudaMalloc(&dh, (size_t)N * K * 4); cudaMalloc(&dD, (size_t)M * N * 4);
    cudaMallocManaged(&cyc, 4 * sizeof(long long)); cudaMallocManaged(&err, sizeof(int));
    const int smem = B_BYTES + 64 + 1024;
    cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);

    // weight scales: the reference's default U(-0.1, 0.1), trained-size weights U(-3, 3), and tiny weights (fp16 subnormal hi parts)
    for (float wscale : {0.1f, 3.0f, 1e-4f}) {
        std::vector<float> W((size_t)M * K), h((size_t)N * K);
        srand(7);
        for (auto &x : W) x = 2.0f * wscale * ((float)rand() / RAND_MAX - 0.5f);
        for (auto &x : h) x = 2.0f * ((float)rand() / RAND_MAX - 0.5f);       // layer outputs live in (-1, 1)
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; k += 7) h[(size_t)n * K + k] *= 1e-3f;   // some small outputs
        for (int n = 12; n < N; ++n) for (int k = 0; k < K; ++k) h[(size_t)n * K + k] = 0.0f;
        for (int m = 0; m < M; ++m) for (int k = 250; k < K; ++k) W[(size_t)m * K + k] = 0.0f;
        std::vector<double> ref((size_t)M * N);
        double refmax = 0, e32 = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double s = 0; float s32 = 0;
            for (int k = 0; k < K; ++k) { s += (double)W[(size_t)m * K + k] * h[(size_t)n * K + k]; s32 = s32 + W[(size_t)m * K + k] * h[(size_t)n * K + k]; }
            ref[(size_t)m * N + n] = s; refmax = fmax(refmax, fabs(s)); e32 = fmax(e32, fabs((double)s32 - s));
        }
        printf("weights U(-%g, %g): fp32 serial sum (the reference's own order) vs fp64: max err / max|ref| = %.3e\n", wscale, wscale, e32 / refmax);
        cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dh, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
        struct { int terms; const char *name; } cases[] = { {1, "W_hi*h_hi only (fp16)"}, {2, "+ W_hi*h_lo'"}, {3, "+ W_lo'*h_hi"} };
        for (auto &c : cases) {
            Params p{dW, dh, dD, cyc, err, 1, c.terms, 1};
            *err = 0; cudaMemset(dD, 0xff, (size_t)M * N * 4);
            step_kernel<<<1, NT, smem>>>(p);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> D((size_t)M * N);
            cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
            double worst = 0;
            for (size_t i = 0; i < D.size(); ++i) worst = fmax(worst, fabs((double)D[i] - ref[i]));
            printf("  %-28s max err / max|ref| = %.3e   (%s%s)\n", c.name, worst / refmax, cudaGetErrorString(e), *err ? ", mbarrier wait timed out" : "");
            if (e != cudaSuccess) return 1;
        }
    }
    struct { int terms, nrep; const char *name; } tc[] = { {3, 1, "16 x N=32 + 16 x N=16"}, {3, 2, "2 x (16 x N=32 + 16 x N=16)"}, {2, 1, "16 x N=32"}, {2, 2, "32 x N=32"}, {1, 1, "16 x N=16"}, {1, 2, "32 x N=16"} };
    for (auto &c : tc) {
        Params p{dW, dh, dD, cyc, err, 2000, c.terms, c.nrep};
        *err = 0;
        step_kernel<<<nsm, NT, smem>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        printf("%d CTAs x 2000 steps, MMAs/step %-28s: %lld cycles/step = write B %lld + MMA issue..complete %lld + tcgen05.ld %lld + barrier  (%s%s)\n",
               nsm, c.name, cyc[0], cyc[1], cyc[2], cyc[3], cudaGetErrorString(e), *err ? ", mbarrier wait timed out" : "");
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
