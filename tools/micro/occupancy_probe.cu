// Probe: what limits CTAs per SM for a kernel that allocates tensor memory?  Prints the device limits and
// cudaOccupancyMaxActiveBlocksPerMultiprocessor for kernels with and without tcgen05.alloc at several dynamic shared memory sizes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o occupancy_probe occupancy_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __maxnreg__(80) plain_kernel(float *p) { extern __shared__ float s[]; s[threadIdx.x] = p[threadIdx.x]; __syncthreads(); p[threadIdx.x] = s[(threadIdx.x + 1) % blockDim.x]; }

template <int COLS>
__global__ void __maxnreg__(80) tmem_kernel(float *p, long long *cyc)
{
    extern __shared__ float s[];
    __shared__ uint32_t slot;
    const long long t0 = clock64();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&slot)), "n"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    const long long t1 = clock64();
    s[threadIdx.x] = p[threadIdx.x];
    // stay resident for a while so that co-resident CTAs really overlap
    while (clock64() - t1 < 200000) { }
    __syncthreads();
    p[threadIdx.x] = s[(threadIdx.x + 1) % blockDim.x] + (float)slot;
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(slot), "n"(COLS) : "memory");
    if (threadIdx.x == 0) { unsigned sm; asm("mov.u32 %0, %%smid;" : "=r"(sm)); cyc[blockIdx.x * 3] = t0; cyc[blockIdx.x * 3 + 1] = t1; cyc[blockIdx.x * 3 + 2] = sm; }
}

template <typename K> static void report(const char *name, K k, int threads)
{
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncAttributes fa{}; cudaFuncGetAttributes(&fa, k);
    printf("%-24s regs %d static smem %zu:", name, fa.numRegs, fa.sharedSizeBytes);
    for (size_t dyn : {(size_t)0, (size_t)40960, (size_t)81920, (size_t)104448, (size_t)109056, (size_t)122880}) {
        int n = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, threads, dyn);
        printf("  dyn %zu -> %d", dyn, n);
    }
    printf("\n");
}

int main()
{
    cudaDeviceProp pr{}; cudaGetDeviceProperties(&pr, 0);
    printf("%s: SMs %d, smem/SM %zu, smem/block optin %zu, reserved/block %zu, regs/SM %d, max threads/SM %d, max blocks/SM %d\n", pr.name, pr.multiProcessorCount,
           pr.sharedMemPerMultiprocessor, pr.sharedMemPerBlockOptin, pr.reservedSharedMemPerBlock, pr.regsPerMultiprocessor, pr.maxThreadsPerMultiProcessor, pr.maxBlocksPerMultiProcessor);
    for (int threads : {288, 256, 544}) {
        printf("threads %d\n", threads);
        report("plain", plain_kernel, threads);
        report("tcgen05.alloc 256 cols", tmem_kernel<256>, threads);
        report("tcgen05.alloc 512 cols", tmem_kernel<512>, threads);
    }
    // really run 2 x SMs CTAs of the 256-column kernel with 100 KB each: do two CTAs share an SM at the same time?
    float *p; long long *cyc; cudaMalloc(&p, 4096); cudaMallocManaged(&cyc, 296 * 3 * 8);
    cudaMemset(p, 0, 4096);
    tmem_kernel<256><<<296, 288, 100 * 1024>>>(p, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long tmin = 1LL << 62; for (int i = 0; i < 296; ++i) tmin = cyc[i * 3] < tmin ? cyc[i * 3] : tmin;
    int late = 0; for (int i = 0; i < 296; ++i) late += (cyc[i * 3] - tmin > 100000);
    printf("296 CTAs x 288 threads, 256 TMEM columns, 100 KB: %s; %d CTAs started more than 100k cycles after the first (0 = all co-resident)\n", cudaGetErrorString(e), late);
    return 0;
}
