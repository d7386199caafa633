// Probe: cost of one all-gather per timestep among the 8 CTAs that share a (direction, sequence group) of the persistent
// recurrent kernels, for the candidate exchange protocols.  Every CTA produces 32 cells x 15 sequences of fp32 per step and needs
// the 256 x 15 values of its whole group before it can start the next step (the B operand of the step GEMM), so a step is
//     [ gather -> shared memory -> CTA barrier -> DELAY cycles of stand-in compute -> produce -> publish ].
// Reported: cycles per step minus DELAY = what the protocol costs on the critical path.
//
//   mode 0  counter:   stores, CTA barrier, one thread fence.acq_rel.gpu + atomicAdd; consumers poll ld.acquire.gpu per warp, then
//                      ld.global.cg the data  (what lstm_recurrent_tmem.cu shipped in round 1)
//   mode 1  counter, red.release.gpu instead of fence + atomicAdd
//   mode 2  in-band:   no counter.  Three exchange buffers armed with a sentinel NaN; consumers poll the DATA words themselves
//                      (ld.relaxed.gpu.v4) until none is the sentinel; an otherwise idle warp re-arms the CTA's own words of buffer
//                      (q+1) mod 3 at the start of step q, fences, and signals a named barrier that the producing warps pass
//                      before their stores of step q (so a consumer can never read step q-2 data as step q+1 data)
//   mode 3  cluster of 8, DSMEM push: every thread st.shared::cluster's its value into all 8 CTAs' shared memory, one lane per
//                      warp then arrives (release.cluster) on each CTA's mbarrier; consumers wait on their own mbarrier (acquire.cluster)
//   mode 4  cluster of 8, bulk DSMEM: values staged in own shared memory, one thread issues 8 cp.async.bulk
//                      shared::cta -> shared::cluster copies that complete_tx on the destination CTA's mbarrier
//   EXTRA=1 adds the 9 HBM-only result stores per thread and step of the real forward kernel (they are what a fence has to drain).
//
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o exchange_probe exchange_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

constexpr int NT = 512, C = 8, CL = 32, SG = 15, HP = C * CL;      // 8 slices of 32 cells, 15 sequences per group
constexpr int ROW4 = HP / 4;                                        // float4 per exchange row
constexpr unsigned SENT = 0xFFFFDEADu;                              // a NaN no arithmetic produces

struct Params {
    float *xbuf;            // [groups][3][SG][HP]
    unsigned *flags;        // [groups][32]
    float *sink;            // HBM-only result stores
    long long *cyc;         // per CTA
    int steps, delay, mode, extra;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p)
{ unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint4 ld_relaxed_v4(const void *p)
{ uint4 v; asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void spin(int cycles) { const long long t = clock64(); while (clock64() - t < cycles) { } }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank)
{ uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void extra_stores(const Params &p, int q, int tid, float v)
{
    if (!p.extra) return;
    float *s = p.sink + ((size_t)blockIdx.x * 64 + (q & 63)) * 9 * NT + tid;
#pragma unroll
    for (int i = 0; i < 9; ++i) s[i * NT] = v + (float)i;
}

__global__ void __launch_bounds__(NT, 1) xchg_kernel(const Params p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    float *land = reinterpret_cast<float *>(smem_raw);                       // [2][SG][HP]: gathered values (double buffered for the DSMEM modes)
    float *stagebuf = land + 2 * SG * HP;                                    // [2][SG][CL]: mode 4 staging
    __shared__ uint64_t s_bar[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = blockIdx.x / C, cs = blockIdx.x % C;
    const int seq = tid / CL, cell = tid % CL;
    const bool producer = seq < SG;                                          // warps 0..14; warp 15 is idle (re-arms in mode 2)
    float *xb = p.xbuf + (size_t)grp * 3 * SG * HP;
    unsigned *flag = p.flags + grp * 32;
    const int mode = p.mode;
    if (tid == 0) {
        // mode 3: one arrival per producing warp of every CTA; mode 4: one local arrive.expect_tx per phase
        const int cnt = mode == 3 ? C * SG : 1;
        for (int b = 0; b < 2; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&s_bar[b])), "r"(cnt) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (mode >= 3) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        if (mode == 4 && tid == 0)
            for (int b = 0; b < 2; ++b)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&s_bar[b])), "r"(C * SG * CL * 4) : "memory");
    }

    float val = 0.001f * (float)(tid + blockIdx.x);
    const long long t0 = clock64();
    for (int q = 0; q < p.steps; ++q) {
        float got = 0.0f;
        if (q > 0) {
            if (mode <= 1) {
                if (lane == 0) while (ld_acquire(flag) < (unsigned)(C * q)) { }
                __syncwarp();
                const float4 *src = reinterpret_cast<const float4 *>(xb + (size_t)((q - 1) & 1) * SG * HP);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int i = tid + u * NT;
                    if (i < SG * ROW4) reinterpret_cast<float4 *>(land)[i] = __ldcg(src + i);
                }
            } else if (mode == 2) {
                const uint4 *src = reinterpret_cast<const uint4 *>(xb + (size_t)((q - 1) % 3) * SG * HP);
                const int i0 = tid, i1 = tid + NT;
                const bool h1 = i1 < SG * ROW4;
                uint4 v0, v1 = make_uint4(0, 0, 0, 0);
                bool ok0 = false, ok1 = !h1;
                while (!(ok0 && ok1)) {
                    if (!ok0) { v0 = ld_relaxed_v4(src + i0); ok0 = v0.x != SENT && v0.y != SENT && v0.z != SENT && v0.w != SENT; }
                    if (!ok1) { v1 = ld_relaxed_v4(src + i1); ok1 = v1.x != SENT && v1.y != SENT && v1.z != SENT && v1.w != SENT; }
                }
                reinterpret_cast<uint4 *>(land)[i0] = v0;
                if (h1) reinterpret_cast<uint4 *>(land)[i1] = v1;
            } else {
                mbar_wait_cluster(&s_bar[(q - 1) & 1], (uint32_t)(((q - 1) >> 1) & 1));
                if (mode == 4 && tid == 0)      // re-arm this barrier's next phase (data of step q+1)
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&s_bar[(q - 1) & 1])), "r"(C * SG * CL * 4) : "memory");
            }
        }
        __syncthreads();
        if (mode == 2 && warp == NT / 32 - 1) {
            // re-arm the CTA's own words of buffer (q+1) mod 3, after the barrier: off the critical path
            uint4 *dst = reinterpret_cast<uint4 *>(xb + (size_t)((q + 1) % 3) * SG * HP);
            for (int i = lane; i < SG * (CL / 4); i += 32) dst[(i / (CL / 4)) * ROW4 + cs * (CL / 4) + (i % (CL / 4))] = make_uint4(SENT, SENT, SENT, SENT);
            __threadfence();
            asm volatile("bar.arrive 1, %0;" :: "n"(NT) : "memory");
        }
        if (q > 0) {
            const float *l = (mode >= 3) ? land + (size_t)((q - 1) & 1) * SG * HP : land;
            got = l[(tid * 7) % (SG * HP)];
        }
        spin(p.delay);
        val = got * 0.5f + val * 0.25f + 1.0f;
        if (q + 1 == p.steps) break;
        // ---- produce + publish
        if (mode <= 1) {
            if (producer) xb[(size_t)(q & 1) * SG * HP + seq * HP + cs * CL + cell] = val;
            __syncthreads();
            if (tid == 0) {
                if (mode == 0) { asm volatile("fence.acq_rel.gpu;" ::: "memory"); atomicAdd(flag, 1u); }
                else asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(flag) : "memory");
            }
            extra_stores(p, q, tid, val);
        } else if (mode == 2) {
            if (producer) {
                asm volatile("bar.sync 1, %0;" :: "n"(NT) : "memory");
                xb[(size_t)(q % 3) * SG * HP + seq * HP + cs * CL + cell] = val;
            }
            extra_stores(p, q, tid, val);
        } else if (mode == 3) {
            if (producer) {
                const uint32_t la = smem_u32(land + (size_t)(q & 1) * SG * HP + seq * HP + cs * CL + cell);
#pragma unroll
                for (int r = 0; r < C; ++r) asm volatile("st.shared::cluster.f32 [%0], %1;" :: "r"(mapa(la, r)), "f"(val) : "memory");
                __syncwarp();
                if (lane < C)
                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(mapa(smem_u32(&s_bar[q & 1]), lane)) : "memory");
            }
            extra_stores(p, q, tid, val);
        } else {
            // destination layout per CTA: land[parity][producer slice][SG][CL]
            if (producer) stagebuf[(q & 1) * SG * CL + seq * CL + cell] = val;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid < C) {
                const uint32_t dst = mapa(smem_u32(land + (size_t)(q & 1) * SG * HP + cs * SG * CL), tid);
                const uint32_t bar = mapa(smem_u32(&s_bar[q & 1]), tid);
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(dst), "r"(smem_u32(stagebuf + (q & 1) * SG * CL)), "r"(SG * CL * 4), "r"(bar) : "memory");
            }
            extra_stores(p, q, tid, val);
        }
    }
    const long long t1 = clock64();
    if (mode >= 3) {        // nobody may exit while peers still push into its shared memory
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (tid == 0) p.cyc[blockIdx.x] = (t1 - t0) / p.steps;
    if (val == 123.456f) p.sink[tid] = val;
}

int main(int argc, char **argv)
{
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int steps = 2000;
    const size_t smem = (size_t)(2 * SG * HP + 2 * SG * CL) * 4 + 100 * 1024;    // > half of the SM: one CTA per SM
    cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(xchg_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int max_groups = nsm / C;
    Params p{};
    cudaMalloc(&p.xbuf, (size_t)max_groups * 3 * SG * HP * 4);
    cudaMalloc(&p.flags, (size_t)max_groups * 32 * 4);
    cudaMalloc(&p.sink, (size_t)nsm * 64 * 9 * NT * 4);
    cudaMallocManaged(&p.cyc, nsm * sizeof(long long));
    p.steps = steps;

    // how many 8-CTA clusters of this kernel are co-resident?
    int max_clusters = 0;
    {
        cudaLaunchConfig_t cfg{}; cudaLaunchAttribute at{};
        cfg.gridDim = dim3(nsm / C * C); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem;
        at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = C; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        cfg.attrs = &at; cfg.numAttrs = 1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, xchg_kernel, &cfg);
        printf("co-resident clusters of 8 CTAs (1 CTA per SM): %d  (%s)\n", max_clusters, cudaGetErrorString(e));
    }
    const char *names[] = { "counter: fence.acq_rel + atomicAdd, ld.acquire poll, ld.cg data", "counter: red.release.gpu", "in-band sentinel, 3 buffers, no counter",
                            "cluster DSMEM: st.shared::cluster + remote mbarrier arrive", "cluster DSMEM: cp.async.bulk smem->peer smem, complete_tx" };
    for (int extra = 0; extra <= 1; ++extra)
        for (int delay : {0, 3000})
            for (int mode = 0; mode < 5; ++mode) {
                const bool cluster = mode >= 3;
                int groups = cluster ? std::min(max_clusters, max_groups) : max_groups;
                if (argc > 1) groups = std::min(groups, atoi(argv[1]));
                if (groups < 1) continue;
                p.mode = mode; p.delay = delay; p.extra = extra;
                cudaMemset(p.flags, 0, (size_t)max_groups * 32 * 4);
                cudaMemset(p.xbuf, mode == 2 ? 0xFF : 0, (size_t)max_groups * 3 * SG * HP * 4);
                if (mode == 2) {     // 0xFFFFDEAD
                    std::vector<unsigned> s((size_t)max_groups * 3 * SG * HP, SENT);
                    cudaMemcpy(p.xbuf, s.data(), s.size() * 4, cudaMemcpyHostToDevice);
                }
                cudaLaunchConfig_t cfg{}; cudaLaunchAttribute at[2]{};
                cfg.gridDim = dim3(groups * C); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem;
                int na = 0;
                if (cluster) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = C; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; ++na; }
                else { at[na].id = cudaLaunchAttributeCooperative; at[na].val.cooperative = 1; ++na; }
                cfg.attrs = at; cfg.numAttrs = na;
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                cudaEventRecord(e0);
                cudaError_t e = cudaLaunchKernelEx(&cfg, xchg_kernel, p);
                cudaEventRecord(e1);
                cudaError_t e2 = cudaDeviceSynchronize();
                float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
                std::vector<long long> c(p.cyc, p.cyc + groups * C);
                std::sort(c.begin(), c.end());
                printf("extra=%d delay=%4d mode %d  %-68s groups %2d: %5lld cycles/step beyond the delay (median; max %lld), %.3f us/step by events  (%s / %s)\n",
                       extra, delay, mode, names[mode], groups, c[c.size() / 2] - delay, c.back() - delay, ms * 1e3 / steps, cudaGetErrorString(e), cudaGetErrorString(e2));
                if (e != cudaSuccess || e2 != cudaSuccess) return 1;
            }
    return 0;
}
