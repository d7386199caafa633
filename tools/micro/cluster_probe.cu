// Probe: how many clusters of CS CTAs (512 threads, 128 regs, `smem` bytes) are co-resident on this GPU, and what a
// DSMEM all-to-all push + cluster barrier costs per step.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_probe cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(512, 1) k(float *out, int steps, int floats_per_peer, long long *cyc)
{
    extern __shared__ float sm[];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned rank = cl.block_rank(), cs = cl.num_blocks();
    float *peer[16];
    for (unsigned r = 0; r < cs; ++r) peer[r] = cl.map_shared_rank(sm, r);
    float acc = 0.f;
    cl.sync();
    const long long t0 = clock64();
    for (int s = 0; s < steps; ++s) {
        float *dstbase = (float *)0;
        const int buf = (s & 1) * 8192;
        for (unsigned r = 0; r < cs; ++r) {
            dstbase = peer[r] + buf + rank * floats_per_peer;
            for (int i = threadIdx.x; i < floats_per_peer; i += blockDim.x) dstbase[i] = (float)(s + i);
        }
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        acc += sm[buf + ((threadIdx.x * 7) % (cs * floats_per_peer))];
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = (t1 - t0) / steps;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main()
{
    float *out; long long *cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMallocManaged(&cyc, 8);
    for (int cs : {2, 4, 8, 16}) {
        for (int smem : {80 * 1024, 120 * 1024, 200 * 1024}) {
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (cs > 8) cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs * 8); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
            printf("cluster %2d smem %3d KB: max active clusters %d (%s) -> %d CTAs\n", cs, smem / 1024, n, cudaGetErrorString(e), n * cs);
        }
    }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    for (int fpp : {32, 416, 1664}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(128); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 80 * 1024;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        *cyc = 0;
        cudaError_t e = cudaLaunchKernelEx(&cfg, k, out, 2000, fpp, cyc);
        cudaError_t e2 = cudaDeviceSynchronize();
        printf("cluster 8, push %5d B to each of 8 CTAs + cluster barrier: %lld cycles/step (%s, %s)\n", fpp * 4, *cyc, cudaGetErrorString(e), cudaGetErrorString(e2));
    }
    return 0;
}
