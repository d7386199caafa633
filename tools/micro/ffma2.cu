// Microbenchmark: fp32 FMA issue rate on sm_100a -- scalar FFMA (3 register operands) vs packed fma.rn.f32x2 (FFMA2).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k_scalar(float *out, int iters, float b0, float b1)
{
    float acc[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) acc[j] = threadIdx.x + j;
    float a0 = b0 + threadIdx.x, a1 = b1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) acc[j] = fmaf(acc[j], a0, a1);
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < CH; ++j) s += acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH>
__global__ void k_packed(float *out, int iters, float b0, float b1)
{
    unsigned long long acc[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) acc[j] = ((unsigned long long)__float_as_uint(threadIdx.x + j) << 32) | __float_as_uint(1.0f + j);
    unsigned long long a = ((unsigned long long)__float_as_uint(b0 + threadIdx.x) << 32) | __float_as_uint(b0);
    unsigned long long c = ((unsigned long long)__float_as_uint(b1) << 32) | __float_as_uint(b1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[j]) : "l"(a), "l"(c));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int j = 0; j < CH; ++j) s ^= acc[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)(s >> 32)) + __uint_as_float((unsigned)s);
}

int main()
{
    float *out; cudaMalloc(&out, 148 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int mode = 0; mode < 2; ++mode) {
            if (mode == 0) k_scalar<16><<<148, threads>>>(out, 100, 1.0001f, 0.5f); else k_packed<16><<<148, threads>>>(out, 100, 1.0001f, 0.5f);
            cudaEventRecord(e0);
            if (mode == 0) k_scalar<16><<<148, threads>>>(out, iters, 1.0001f, 0.5f); else k_packed<16><<<148, threads>>>(out, iters, 1.0001f, 0.5f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double cycles = ms * 1e-3 * clk * 1e3;
            const double fma_per_sm = (double)iters * 16 * threads * (mode ? 2 : 1);
            printf("%s threads %4d: %.1f FMA/clk/SM  (%.1f TFLOP/s chip)\n", mode ? "fma.rn.f32x2" : "fmaf (FFMA) ", threads, fma_per_sm / cycles,
                   2 * fma_per_sm * 148 / (ms * 1e-3) / 1e12);
        }
    }
    return 0;
}
