// Microbenchmark: issue rate of mma.sync.m16n8k8 TF32 on sm_100a (legacy tensor path), per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_tf32 mma_sync_tf32.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k(float *out, int iters)
{
    float c[CHAINS][4];
    unsigned a[4] = {0x3f800000u + threadIdx.x, 0x3f800000u, 0x3f000000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f000000u + threadIdx.x};
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CHAINS; ++j)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0;
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
void run(int threads)
{
    float *out; cudaMalloc(&out, 148 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<CHAINS><<<148, threads>>>(out, 100);
    cudaEventRecord(e0);
    k<CHAINS><<<148, threads>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double mmas_per_sm = (double)iters * CHAINS * (threads / 32);
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("threads %4d chains %d: %.2f cycles per mma per SM  -> %.0f FMA/clk/SM, %.1f TFLOP/s chip (clock attr %d kHz)\n", threads, CHAINS,
           cycles / mmas_per_sm, 1024.0 * mmas_per_sm / cycles, 2 * 1024.0 * mmas_per_sm * 148 / (ms * 1e-3) / 1e12, clk);
    cudaFree(out);
}

int main()
{
    run<1>(128); run<4>(128); run<8>(128); run<4>(256); run<8>(256); run<4>(512); run<8>(512); run<2>(1024);
    return 0;
}
