// Probe: cycles per timestep of the forward gate math (the ComputeBlockOutputFn restatement of csrc/common.cuh) in isolation, one
// (cell, sequence) pair per thread, 512 threads per CTA, one CTA per SM -- the shape of the persistent recurrent kernels.
// The cell state carries a true dependency from step to step; everything else comes from registers.  Variants:
//   0  exp_ref_tab as shipped (separately rounded double operations, table in shared memory)
//   1  the same polynomial with fused multiply-adds (what glibc's FMA build of expf executes)
//   2  variant 1 with the table lookup and scale done by integer ops on the result exponent (no 64-bit shared load)
//   4  variant 0 with the activations of a level evaluated together in one basic block (act3_tab / act2_tab of common.cuh)
//   3  fp32 __expf instead (NOT parity-safe: the floor of everything that is not the double-precision chain)
//
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../lstm-rnn_b200/csrc -o gate_math_probe gate_math_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace bl;

template <typename TabPtr>
__device__ __forceinline__ float exp_fma_tab(float x, TabPtr tab)
{
    const double InvLn2N = 0x1.71547652b82fep+0 * 32.0;
    const double SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0;
    const double C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0;
    const double C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
    const float xc = fminf(fmaxf(x, -104.0f), 89.0f);
    const double z = __dmul_rn(InvLn2N, (double)xc);
    double kd = __dadd_rn(z, SHIFT);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dsub_rn(kd, SHIFT);
    const double r = __dsub_rn(z, kd);
    unsigned long long t = tab[ki & 31];
    t += ki << (52 - 5);
    const double s = __longlong_as_double((long long)t);
    const double zz = __fma_rn(C0, r, C1);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(C2, r, 1.0);
    y = __fma_rn(zz, r2, y);
    y = __dmul_rn(y, s);
    float res = __double2float_rn(y);
    res = (x > 88.7228317f) ? __int_as_float(0x7f800000) : res;
    res = (x < -103.972076f) ? 0.0f : res;
    return res;
}

template <int V, typename TabPtr>
__device__ __forceinline__ float ex(float x, TabPtr tab)
{
    if (V == 0) return exp_ref_tab(x, tab);
    if (V == 1 || V == 2) return exp_fma_tab(x, tab);
    return __expf(x);
}
template <int V, typename TabPtr>
__device__ __forceinline__ float sig(float x, TabPtr tab)
{
    const float r = __frcp_rn(__fadd_rn(1.0f, ex<V>(-x, tab)));
    return (x < 88.722839f) ? ((x > -88.722839f) ? r : 0.0f) : 1.0f;
}
template <int V, typename TabPtr>
__device__ __forceinline__ float th(float x, TabPtr tab) { return __fsub_rn(__fmul_rn(2.0f, sig<V>(__fmul_rn(2.0f, x), tab)), 1.0f); }

template <int V>
__global__ void __launch_bounds__(512, 1) gate_kernel(const float *in, float *out, long long *cyc, int steps)
{
    __shared__ unsigned long long s_tab[32];
    if (threadIdx.x < 32) s_tab[threadIdx.x] = bl_exp2f_tab[threadIdx.x];
    __syncthreads();
    const int tid = threadIdx.x + blockIdx.x * blockDim.x;
    const float a0 = in[tid * 4], a1 = in[tid * 4 + 1], a2 = in[tid * 4 + 2], a3 = in[tid * 4 + 3];
    const float wp0 = 0.03f, wp1 = -0.05f, wp2 = 0.07f;
    float c = 0.0f, h = 0.0f;
    const long long t0 = clock64();
    for (int q = 0; q < steps; ++q) {
        // pre-activations: vary with the step and with the previous output like the real recurrence (h enters through the step GEMM)
        float ni = __fadd_rn(a0, __fmul_rn(h, 0.11f)), ig = __fadd_rn(a1, __fmul_rn(h, -0.13f));
        float fg = __fadd_rn(a2, __fmul_rn(h, 0.17f)), og = __fadd_rn(a3, __fmul_rn(h, 0.19f));
        ig = __fadd_rn(ig, __fmul_rn(c, wp0));
        fg = __fadd_rn(fg, __fmul_rn(c, wp1));
        if (V == 4) {                          // branch-free batches of common.cuh
            act3_tab(ni, ig, fg, s_tab, ni, ig, fg);
            c = __fadd_rn(__fmul_rn(ni, ig), __fmul_rn(c, fg));
            og = __fadd_rn(og, __fmul_rn(c, wp2));
            float tc;
            act2_tab(c, og, s_tab, tc, og);
            h = __fmul_rn(tc, og);
        } else {
            ni = th<V>(ni, s_tab); ig = sig<V>(ig, s_tab); fg = sig<V>(fg, s_tab);
            c = __fadd_rn(__fmul_rn(ni, ig), __fmul_rn(c, fg));
            og = __fadd_rn(og, __fmul_rn(c, wp2));
            og = sig<V>(og, s_tab);
            h = __fmul_rn(th<V>(c, s_tab), og);
        }
        __syncthreads();                       // the real kernels have at least one CTA barrier per step
    }
    const long long t1 = clock64();
    out[tid] = h + c;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) / steps;
}

template <int V>
static void run(const char *name, const float *in, float *out, long long *cyc, int nsm)
{
    const int steps = 2000;
    for (int threads : {512, 256, 128, 32}) {
        gate_kernel<V><<<nsm, threads>>>(in, out, cyc, steps);
        cudaError_t e = cudaDeviceSynchronize();
        long long lo = 1LL << 60, hi = 0;
        for (int i = 0; i < nsm; ++i) { lo = cyc[i] < lo ? cyc[i] : lo; hi = cyc[i] > hi ? cyc[i] : hi; }
        printf("%-70s %2d warps: %lld .. %lld cycles/step  (%s)\n", name, threads / 32, lo, hi, cudaGetErrorString(e));
    }
}

int main()
{
    int nsm = 0; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    float *in, *out; long long *cyc;
    cudaMallocManaged(&in, (size_t)nsm * 512 * 4 * 4); cudaMallocManaged(&out, (size_t)nsm * 512 * 4); cudaMallocManaged(&cyc, nsm * 8);
    srand(3);
    for (size_t i = 0; i < (size_t)nsm * 512 * 4; ++i) in[i] = 2.0f * ((float)rand() / RAND_MAX - 0.5f);
    run<0>("forward gate math, exp_ref_tab of common.cuh (glibc FMA form, 8 FP64 instructions per exp)", in, out, cyc, nsm);
    run<1>("round-1 variant: fused polynomial only, z rounded separately (9 FP64 instructions)", in, out, cyc, nsm);
    run<4>("branch-free batches (act3_tab / act2_tab): the three / two activations of a level interleave", in, out, cyc, nsm);
    run<3>("fp32 __expf instead of the double-precision chain (floor, not parity-safe)", in, out, cyc, nsm);
    // bit-agreement of the fused variant with the shipped one on this input set
    return 0;
}
