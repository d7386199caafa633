// Persistent recurrent kernels, tensor-memory-resident variant ("v4"): same decomposition, L2 exchange buffers and step-counter
// protocol as lstm_recurrent_reg.cu, but the CTA's slice of the recurrent weights lives in TENSOR MEMORY for the whole pass and the
// per-timestep product runs on the tensor cores (tcgen05.mma with the A operand in TMEM).  Forward:
//
//     pre[128 rows x NB seqs] = Wslice[128 x Hp] * hprev[Hp x NB]          row = gate*32 + cell, NB = 16 or 32
//
// fp32-grade accuracy from three products accumulated in one fp32 TMEM accumulator (hi = tf32_rna(x), lo = x - hi):
//     W h ~= W_hi h_hi + W_hi h_lo      kind::tf32, A = W_hi       TMEM columns [0, Hp)
//                       + W_lo h_hi      kind::f16,  A = bf16(W_lo) TMEM columns [256, 256 + Hp/2), two k per column (even k low)
// (W_lo as tf32 would need Hp more columns and leave no room for the accumulator at Hp = 256; as bf16 its rounding error is 2^-9 of a
// term that is already 2^-11 of the product.  Measured with tools/micro/tcgen05_ts_step.cu: 8.1e-7 of max|W h| against fp64 -- the
// reference's own serial fp32 sum is at 4.4e-7 -- and 1.57 k cycles for the 80 MMAs of a C2 step against 4.8 k for the FFMA GEMM.)
//
// Per step the B operand -- the previous-step vector of the CTA's sequences -- is split while it is copied from the L2 exchange buffer
// into K-major SWIZZLE_128B shared tiles (tf32 hi and lo as the two halves of one tile, bf16), one elected thread issues the MMAs,
// and warps 0..3 (gate = warp, cell = lane) pull the accumulator out of TMEM and stage it [gate][seq][cell] for the gate math, which
// is unchanged.  The BPTT kernel (second half of this file) slices the product the other way round so that its B operand never
// crosses the exchange buffer.  Both are the default whenever pad32(H) <= 256 (bl_lstm_plan_create; BLSTM_REC_V=1 / 2 select the
// register- / shared-memory-resident families of lstm_recurrent_reg.cu / lstm_recurrent.cu instead).
#include "lstm_recurrent.cuh"
#include <cuda_bf16.h>
#include <cmath>
#include <cstdint>
#include <cstdlib>

namespace bl {

constexpr int TM_NT = 512;
constexpr int TM_COL_AHI = 0, TM_COL_ALO = 256, TM_COL_D = 384, TM_COLS = 512;

static int tm_pad32(int x) { return (x + 31) / 32 * 32; }
// RecGeom::K4 (unused by these kernels otherwise) carries the "merged N" switch: BLSTM_TM_MERGE=0 issues the two tf32 products separately
static int tm_merge_default() { const char *e = getenv("BLSTM_TM_MERGE"); return (e && atoi(e) == 0) ? 0 : 1; }

// RecGeom fields used: G, C, CL, SG, NT, npair, Hpad, RS (= Hpad: plain exchange rows), Spad (= NB, the MMA's N), smem
bool choose_geometry_tmem(int H, int S, int ndir, int num_sms, int smem_cap, int forceG, RecGeom *out)
{
    const int Hp = tm_pad32(H);
    if (Hp > 256) return false;
    const int per_dir = num_sms / ndir;
    bool found = false;
    RecGeom best{};
    for (int G = 1; G <= 64 && G <= S; ++G) {
        if (forceG > 0 && G != forceG) continue;
        int C = per_dir / G;
        if (C < 1) break;
        const int CL = cdiv(H, C);
        if (CL > 32) continue;                                // 4 gates x 32 cells = the 128 TMEM lanes
        C = cdiv(H, CL);
        const int SG = cdiv(S, G);
        if ((G - 1) * SG >= S) continue;                      // trailing group would be empty
        if (SG > 32) continue;
        const int NB = SG <= 16 ? 16 : 32;
        if (CL * SG > REC_NPAIR * TM_NT) continue;
        const size_t smem = (size_t)2 * (Hp / 32) * NB * 128 + (size_t)cdiv(Hp, 64) * NB * 128 + (size_t)4 * NB * 32 * sizeof(float) + 1024;
        if ((int)smem > smem_cap) continue;
        const int npair = (CL * SG > TM_NT) ? 2 : 1;
        const double mma = (2.0 * (Hp / 8) + Hp / 16) * (NB == 16 ? 20.0 : 28.0) + 300.0;
        const double gate = 1200.0 + 900.0 * npair;
        const double copy = (double)SG * Hp * 4.0 / 48.0 + 500.0;
        const double cost = mma + gate + copy + 1500.0 + 12.0 * C;
        if (!found || cost < best.cost) {
            found = true;
            best = RecGeom{};
            best.G = G; best.C = C; best.CL = CL; best.SG = SG; best.NT = TM_NT; best.nsub = 1; best.npair = npair;
            best.R = 128; best.Hpad = Hp; best.RS = Hp; best.Spad = NB; best.smem = smem; best.cost = cost;
            best.K4 = tm_merge_default();
        }
    }
    if (found) *out = best;
    return found;
}

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint32_t tm_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ unsigned tm_ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void tm_wait_warp(const unsigned *flag, unsigned target)
{
    if ((threadIdx.x & 31) == 0) { while (tm_ld_acquire(flag) < target) { } }
    __syncwarp();
}
__device__ __forceinline__ void tm_publish(unsigned *flag)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        atomicAdd(flag, 1u);
    }
}

// an MMA batch that never completes would hang the whole cooperative grid: trap instead (the launch then fails loudly)
__device__ __forceinline__ void tm_mbar_wait(uint64_t *bar, uint32_t parity)
{
    for (int spin = 0; spin < (1 << 26); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(tm_smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}

__device__ __forceinline__ bool tm_elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (same encoding as gemm_tc.cu): SBO = 1024 B, version 1, layout type 2
__device__ __forceinline__ uint64_t tm_make_desc(const void *p)
{
    uint64_t d = 0;
    d |= (uint64_t)((tm_smem_u32(p) & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: D fp32, A/B format fmt (kind::tf32: 2 = tf32; kind::f16: 1 = bf16), both K-major, M = 128
__device__ __forceinline__ uint32_t tm_make_idesc(uint32_t fmt, int N)
{ return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

__device__ __forceinline__ void tm_mma_tf32(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tm_mma_bf16(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tm_commit(uint64_t *bar)
{ asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(tm_smem_u32(bar)) : "memory"); }

__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, float (&v)[16])
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

__device__ __forceinline__ float tm_tf32(float x)
{ uint32_t h; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x)); return __uint_as_float(h); }
__device__ __forceinline__ uint32_t tm_bf16(float x) { return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x)); }

// byte offset of (sequence n, k) in a K-major SWIZZLE_128B tile of NB rows: 128-byte rows, 8-row atoms of 1024 B, 16-byte chunks
// XORed with the row index inside the atom; one [NB x 128 B] box per K-block (32 floats / 64 bf16)
__device__ __forceinline__ int tm_off_f32(int NB, int n, int k)
{ const int kb = k >> 5, kin = k & 31, c = kin >> 2, e = kin & 3, r = n & 7; return kb * NB * 128 + (n >> 3) * 1024 + r * 128 + ((c ^ r) << 4) + e * 4; }
__device__ __forceinline__ int tm_off_bf16(int NB, int n, int k)
{ const int kb = k >> 6, kin = k & 63, c = kin >> 3, e = kin & 7, r = n & 7; return kb * NB * 128 + (n >> 3) * 1024 + r * 128 + ((c ^ r) << 4) + e * 2; }

__device__ __forceinline__ void tm_split_store(float *hi, float *lo, size_t idx, float v)
{
    const float h = tm_tf32(v);
    hi[idx] = h;
    lo[idx] = __fsub_rn(v, h);
}

// ------------------------------------------------------------------------------------------------ forward
template <int NPAIR>
__global__ void __launch_bounds__(TM_NT, 1) lstm_fwd_tmem_kernel(const RecFwdParams p)
{
    extern __shared__ uint8_t tm_smem_raw[];
    __shared__ unsigned long long s_tab[32];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_slot;
    const RecGeom &g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Hp = g.Hpad, NB = g.Spad;
    // the tf32 hi and lo copies of the B operand share ONE tile of 2*NB rows (hi rows [0, NB), lo rows [NB, 2*NB)): with `merge` a
    // single N = 2*NB MMA per k-step multiplies W_hi with both (accumulator columns [0, NB) and [NB, 2*NB), added in the epilogue) --
    // 48 instead of 80 MMAs per C2 step; without it two N = NB MMAs address the two halves through their own descriptors
    const bool merge = g.K4 != 0;
    const int NB2 = 2 * NB;
    const int bhi_bytes = (Hp / 32) * NB * 128, bbf_bytes = ((Hp + 63) / 64) * NB * 128;
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(tm_smem_raw) + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B atoms
    uint8_t *Bhl = base, *Bbf = base + 2 * bhi_bytes;
    float *stage = reinterpret_cast<float *>(base + 2 * bhi_bytes + bbf_bytes);          // [4 gates][NB][32 cells]

    const int H = p.H, L = p.L, S = p.S, T = p.T;
    const int d = blockIdx.x / (g.G * g.C);
    const int grp = (blockIdx.x % (g.G * g.C)) / g.C;
    const int cs = blockIdx.x % g.C;
    const int j0 = cs * g.CL, ncell = min(g.CL, H - j0);
    const int s0 = grp * g.SG, nseq = min(g.SG, S - s0);
    unsigned *flag = p.flags + (d * g.G + grp) * 32;

    // rows of the B tiles beyond nseq stay zero for the whole pass
    for (int i = tid; i < (2 * bhi_bytes + bbf_bytes) / 4; i += TM_NT) reinterpret_cast<uint32_t *>(base)[i] = 0u;
    if (tid < 32) s_tab[tid] = bl_exp2f_tab[tid];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(tm_smem_u32(&s_bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tm_smem_u32(&s_slot)), "n"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_slot;

    // the CTA's weight slice into TMEM, once: lane = gate*32 + cell (warp w of the first four owns gate w), column = k (tf32 hi)
    // or k pair (bf16 lo).  W_gate[j, k] = Wi[gate*L*H + d*H*H + j*H + k], k = source cell (LstmLayer.cu:586-596)
    if (warp < 4) {
        const bool ok = lane < ncell;
        const float *w = p.Wi + (size_t)warp * L * H + (size_t)d * H * H + (size_t)(j0 + (ok ? lane : 0)) * H;
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < Hp; c0 += 8) {
            uint32_t r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = (ok && c0 + i < H) ? __float_as_uint(tm_tf32(__ldg(w + c0 + i))) : 0u;
            tm_st8(lane_base + TM_COL_AHI + c0, r);
        }
        for (int c0 = 0; c0 < Hp / 2; c0 += 8) {
            uint32_t r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = 2 * (c0 + i);
                const float x0 = (ok && k < H) ? __ldg(w + k) : 0.0f, x1 = (ok && k + 1 < H) ? __ldg(w + k + 1) : 0.0f;
                r[i] = tm_bf16(__fsub_rn(x0, tm_tf32(x0))) | (tm_bf16(__fsub_rn(x1, tm_tf32(x1))) << 16);
            }
            tm_st8(lane_base + TM_COL_ALO + c0, r);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }

    bool valid[NPAIR]; int cl_[NPAIR], sl_[NPAIR];
    float wb[NPAIR][4], wpe[NPAIR][3], cprev[NPAIR];
#pragma unroll
    for (int u = 0; u < NPAIR; ++u) {
        const int pr = tid + u * TM_NT;
        cl_[u] = pr % g.CL; sl_[u] = pr / g.CL;
        valid[u] = (cl_[u] < ncell) && (sl_[u] < nseq);
        cprev[u] = 0.0f;
        if (valid[u]) {
            const int col = d * H + j0 + cl_[u];
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) wb[u][gi] = __fmul_rn(p.bias, __ldg(p.Wb + gi * L + col));      // bias * w, :97-100
#pragma unroll
            for (int gi = 0; gi < 3; ++gi) wpe[u][gi] = __ldg(p.Wp + gi * L + col);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const uint32_t idesc_tf32 = tm_make_idesc(2u, merge ? NB2 : NB), idesc_bf16 = tm_make_idesc(1u, NB);
    const uint64_t desc_hi = tm_make_desc(Bhl), desc_lo = desc_hi + (uint64_t)(((NB / 8) * 1024) >> 4), desc_bf = tm_make_desc(Bbf);
    const uint64_t kb_step2 = (uint64_t)((NB2 * 128) >> 4), kb_step = (uint64_t)((NB * 128) >> 4);   // next K-block, in descriptor address units
    const int hq4 = Hp / 4;

    for (int q = 0; q < T; ++q) {
        const int t = (d == 0) ? q : T - 1 - q;
        const bool first = (q == 0);
        const bool check = (t >= p.Tmin);
        float *acts_t = p.acts + (size_t)t * S * 4 * L + d * H + j0;
        float *cst_t = p.cst + (size_t)t * S * L + d * H + j0;
        float *y_t = p.Y + (size_t)t * S * p.ldy + d * H + j0;
        const char *pat_t = p.pat + (size_t)t * S;
        float *hx_w = p.hx + (size_t)(d * 2 + (q & 1)) * S * Hp;

        float a[NPAIR][4]; bool dummy[NPAIR];
#pragma unroll
        for (int u = 0; u < NPAIR; ++u) {
            dummy[u] = false;
            if (valid[u]) {
                const int slot = s0 + sl_[u];
                dummy[u] = check && (pat_t[slot] == BL_PATTYPE_NONE);
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) a[u][gi] = acts_t[slot * 4 * L + gi * L + cl_[u]];
            }
        }

        long long *tr = p.trace ? p.trace + ((size_t)blockIdx.x * T + q) * 6 : nullptr;
        if (tr && tid == 0) tr[0] = clock64();
        if (!first) {
            tm_wait_warp(flag, (unsigned)(g.C * q));
            if (tr && tid == 0) tr[1] = clock64();
            // previous-step outputs of this group's sequences: L2 exchange buffer -> split -> the three B tiles
            const float4 *src = reinterpret_cast<const float4 *>(p.hx + ((size_t)(d * 2 + ((q - 1) & 1)) * S + s0) * Hp);
            const int n4 = nseq * hq4;
            float4 v[4]; int idx[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { idx[u] = tid + u * TM_NT; if (idx[u] < n4) v[u] = __ldcg(src + idx[u]); }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (idx[u] >= n4) continue;
                const int n = idx[u] / hq4, k = (idx[u] - n * hq4) * 4;
                const float4 x = v[u];
                const float4 hi = make_float4(tm_tf32(x.x), tm_tf32(x.y), tm_tf32(x.z), tm_tf32(x.w));
                *reinterpret_cast<float4 *>(Bhl + tm_off_f32(NB2, n, k)) = hi;
                *reinterpret_cast<float4 *>(Bhl + tm_off_f32(NB2, NB + n, k)) = make_float4(__fsub_rn(x.x, hi.x), __fsub_rn(x.y, hi.y), __fsub_rn(x.z, hi.z), __fsub_rn(x.w, hi.w));
                uint2 b; b.x = tm_bf16(x.x) | (tm_bf16(x.y) << 16); b.y = tm_bf16(x.z) | (tm_bf16(x.w) << 16);
                *reinterpret_cast<uint2 *>(Bbf + tm_off_bf16(NB, n, k)) = b;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (tr && tid == 0) tr[2] = clock64();
            if (warp == 4) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tm_elect_one()) {
                    // tf32: 8 k per MMA = 8 TMEM columns = 32 B inside the swizzle atom; bf16: 16 k per MMA, likewise 8 columns / 32 B
                    if (merge) {
                        for (int ks = 0; ks < Hp / 8; ++ks)
                            tm_mma_tf32(tmem + TM_COL_D, tmem + TM_COL_AHI + ks * 8, desc_hi + (uint64_t)(ks >> 2) * kb_step2 + (uint64_t)((ks & 3) * 2), idesc_tf32, ks ? 1u : 0u);
                        for (int ks = 0; ks < Hp / 16; ++ks)
                            tm_mma_bf16(tmem + TM_COL_D, tmem + TM_COL_ALO + ks * 8, desc_bf + (uint64_t)(ks >> 2) * kb_step + (uint64_t)((ks & 3) * 2), idesc_bf16, 1u);
                    } else {                                   // small terms first, then the leading one
                        for (int ks = 0; ks < Hp / 16; ++ks)
                            tm_mma_bf16(tmem + TM_COL_D, tmem + TM_COL_ALO + ks * 8, desc_bf + (uint64_t)(ks >> 2) * kb_step + (uint64_t)((ks & 3) * 2), idesc_bf16, ks ? 1u : 0u);
                        for (int ks = 0; ks < Hp / 8; ++ks)
                            tm_mma_tf32(tmem + TM_COL_D, tmem + TM_COL_AHI + ks * 8, desc_lo + (uint64_t)(ks >> 2) * kb_step2 + (uint64_t)((ks & 3) * 2), idesc_tf32, 1u);
                        for (int ks = 0; ks < Hp / 8; ++ks)
                            tm_mma_tf32(tmem + TM_COL_D, tmem + TM_COL_AHI + ks * 8, desc_hi + (uint64_t)(ks >> 2) * kb_step2 + (uint64_t)((ks & 3) * 2), idesc_tf32, 1u);
                    }
                    tm_commit(&s_bar);
                }
                __syncwarp();
            }
            if (warp < 4) {
                tm_mbar_wait(&s_bar, (uint32_t)((q - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int c = 0; c < NB; c += 16) {
                    float dv[16];
                    tm_ld16(tmem + ((uint32_t)(warp * 32) << 16) + TM_COL_D + c, dv);
                    if (merge) {
                        float dl[16];
                        tm_ld16(tmem + ((uint32_t)(warp * 32) << 16) + TM_COL_D + NB + c, dl);
#pragma unroll
                        for (int i = 0; i < 16; ++i) dv[i] = __fadd_rn(dv[i], dl[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) stage[(warp * NB + c + i) * 32 + lane] = dv[i];
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            __syncthreads();
            if (tr && tid == 0) tr[3] = clock64();
        }

        // gate math (identical to lstm_fwd_reg_kernel); only the exchange value h is stored before the publish
        float r_ni[NPAIR], r_ig[NPAIR], r_fg[NPAIR], r_og[NPAIR], r_h[NPAIR];
#pragma unroll
        for (int u = 0; u < NPAIR; ++u) {
            if (!valid[u]) continue;
            const int slot = s0 + sl_[u], cl = cl_[u];
            float h, c;
            if (dummy[u]) {                                       // LstmLayer.cu:78-85
                h = 0.0f; c = 0.0f;
                r_ni[u] = r_ig[u] = r_fg[u] = r_og[u] = 0.0f;
            } else {
                float ni = a[u][0], ig = a[u][1], fg = a[u][2], og = a[u][3];
                if (!first) {                                     // recurrent addProduct, :815-818
                    const float *sp = stage + sl_[u] * 32 + cl;
                    ni = __fadd_rn(ni, sp[0]);
                    ig = __fadd_rn(ig, sp[NB * 32]);
                    fg = __fadd_rn(fg, sp[2 * NB * 32]);
                    og = __fadd_rn(og, sp[3 * NB * 32]);
                }
                ni = __fadd_rn(ni, wb[u][0]); ig = __fadd_rn(ig, wb[u][1]);
                fg = __fadd_rn(fg, wb[u][2]); og = __fadd_rn(og, wb[u][3]);
                if (!first) {                                     // :103-108
                    ig = __fadd_rn(ig, __fmul_rn(cprev[u], wpe[u][0]));
                    fg = __fadd_rn(fg, __fmul_rn(cprev[u], wpe[u][1]));
                }
                act3_tab(ni, ig, fg, s_tab, ni, ig, fg);        // the three first-level activations, interleaved
                c = __fmul_rn(ni, ig);                            // :121-126
                if (!first) c = __fadd_rn(c, __fmul_rn(cprev[u], fg));
                og = __fadd_rn(og, __fmul_rn(c, wpe[u][2]));      // :129-131
                float tc;
                act2_tab(c, og, s_tab, tc, og);
                h = __fmul_rn(tc, og);         // :134
                r_ni[u] = ni; r_ig[u] = ig; r_fg[u] = fg; r_og[u] = og;
            }
            cprev[u] = c; r_h[u] = h;
            hx_w[slot * Hp + j0 + cl] = h;
        }
        if (tr && tid == 0) tr[4] = clock64();
        if (q + 1 < T) tm_publish(flag);
#pragma unroll
        for (int u = 0; u < NPAIR; ++u) {
            if (!valid[u]) continue;
            const int slot = s0 + sl_[u], cl = cl_[u];
            if (!dummy[u]) {
                float *ap = acts_t + slot * 4 * L + cl;
                ap[0] = r_ni[u]; ap[L] = r_ig[u]; ap[2 * L] = r_fg[u]; ap[3 * L] = r_og[u];
            }
            cst_t[slot * L + cl] = cprev[u];
            y_t[slot * p.ldy + cl] = r_h[u];
            if (p.ys_hi) tm_split_store(p.ys_hi, p.ys_lo, ((size_t)t * S + slot) * p.ld_ys + d * ((H + 3) & ~3) + j0 + cl, r_h[u]);
        }
        if (tr && tid == 0) tr[5] = clock64();
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ BPTT
// Output-stationary slicing: the CTA that owns the cells j of a slice has just produced their four gate deltas, so it multiplies
// THOSE (K = 4 gates x 32 cells = 128, straight from registers into the B tiles -- no exchange read for the GEMM operand) with
// the weight columns of ALL source cells k' of its direction (M = pad128(H) rows = 1 or 2 TMEM tiles):
//     Q[k', s] = sum_{gate, j in slice} W_gate[j, k'] * delta_gate[j, s]           A[k'][gate*32 + c] = Wi[gate*L*H + d*H*H + (j0+c)*H + k']
// and publishes its partial Q through the L2 exchange buffer; next step every CTA adds the C partials of its own cells
// (fixed order, so the result is deterministic) to the output error -- the 4 addProducts of LstmLayer.cu:939-942.  Per step and CTA
// that is C*CL*SG floats read (12 KB at C2) instead of the 4*Hp*SG (48 KB) the cell-stationary kernels read, and the same 80 MMAs
// as the forward kernel.  Exchange layout: [dir][parity][group][producer slice][sequence NB][k' R], R = pad128(H).
bool choose_geometry_tmem_bwd(int H, int S, int ndir, int num_sms, int smem_cap, int forceG, RecGeom *out)
{
    if (tm_pad32(H) > 256) return false;
    const int R = (H + 127) / 128 * 128;
    const int per_dir = num_sms / ndir;
    bool found = false;
    RecGeom best{};
    for (int G = 1; G <= 64 && G <= S; ++G) {
        if (forceG > 0 && G != forceG) continue;
        int C = per_dir / G;
        if (C < 1) break;
        const int CL = cdiv(H, C);
        if (CL > 32) continue;
        C = cdiv(H, CL);
        const int SG = cdiv(S, G);
        if ((G - 1) * SG >= S) continue;
        if (SG > 32) continue;
        const int NB = SG <= 16 ? 16 : 32;
        if (CL * SG > REC_NPAIR * TM_NT) continue;
        const size_t smem = (size_t)2 * 4 * NB * 128 + (size_t)2 * NB * 128 + 1024;
        if ((int)smem > smem_cap) continue;
        const int npair = (CL * SG > TM_NT) ? 2 : 1;
        const double mma = (R / 128) * 40.0 * (NB == 16 ? 20.0 : 28.0) + 300.0;
        const double gate = 1000.0 + 700.0 * npair;
        const double xchg = 40.0 * C + (double)R * SG * 4.0 / 64.0 + 500.0;
        const double cost = mma + gate + xchg + 1500.0;
        if (!found || cost < best.cost) {
            found = true;
            best = RecGeom{};
            best.G = G; best.C = C; best.CL = CL; best.SG = SG; best.NT = TM_NT; best.nsub = 1; best.npair = npair;
            best.R = R; best.Hpad = R; best.Spad = NB; best.smem = smem; best.cost = cost;
            best.K4 = tm_merge_default();
            // the plan allocates ndir*2*S*RS floats for the exchange buffer: make that cover [G][C][NB][R] per (dir, parity)
            best.RS = (int)(((size_t)G * C * NB * R + S - 1) / S);
        }
    }
    if (found) *out = best;
    return found;
}

template <int NPAIR>
__global__ void __launch_bounds__(TM_NT, 1) lstm_bwd_tmem_kernel(const RecBwdParams p)
{
    extern __shared__ uint8_t tm_smem_raw[];
    __shared__ unsigned long long s_tab[32];
    __shared__ uint64_t s_bar;
    __shared__ uint32_t s_slot;
    const RecGeom &g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = g.Hpad, MT = R / 128, NB = g.Spad;
    const bool merge = g.K4 != 0;                                          // one N = 2*NB MMA for W_hi*(d_hi | d_lo), see the forward kernel
    const int NB2 = 2 * NB;
    const int bhi_bytes = 4 * NB * 128, bbf_bytes = 2 * NB * 128;          // K = 128: 4 K-blocks of 32 floats, 2 of 64 bf16
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(tm_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *Bhl = base, *Bbf = base + 2 * bhi_bytes;                      // hi rows [0, NB), lo rows [NB, 2*NB) of one tile

    const int H = p.H, L = p.L, S = p.S, T = p.T;
    const int d = blockIdx.x / (g.G * g.C);
    const int grp = (blockIdx.x % (g.G * g.C)) / g.C;
    const int cs = blockIdx.x % g.C;
    const int j0 = cs * g.CL, ncell = min(g.CL, H - j0);
    const int s0 = grp * g.SG, nseq = min(g.SG, S - s0);
    unsigned *flag = p.flags + (d * g.G + grp) * 32;
    const bool inplace = (p.ndir == 1);

    // B entries of cells beyond ncell / sequences beyond nseq are never written: they stay zero for the whole pass
    for (int i = tid; i < (2 * bhi_bytes + bbf_bytes) / 4; i += TM_NT) reinterpret_cast<uint32_t *>(base)[i] = 0u;
    if (tid < 32) s_tab[tid] = bl_exp2f_tab[tid];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(tm_smem_u32(&s_bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tm_smem_u32(&s_slot)), "n"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_slot;

    // weights into TMEM, once.  Tile mt, lane = source cell k' - mt*128, column = gate*32 + c (tf32 hi at mt*128, bf16 lo pairs at 256 + mt*64)
    if (warp < 4) {
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        for (int mt = 0; mt < MT; ++mt) {
            const int kr = mt * 128 + warp * 32 + lane;
            const bool ok = kr < H;
            const float *w = p.Wi + (size_t)d * H * H + (size_t)j0 * H + (ok ? kr : 0);
            for (int c0 = 0; c0 < 128; c0 += 8) {
                uint32_t r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int kk = c0 + i, gi = kk >> 5, c = kk & 31;
                    r[i] = (ok && c < ncell) ? __float_as_uint(tm_tf32(__ldg(w + (size_t)gi * L * H + (size_t)c * H))) : 0u;
                }
                tm_st8(lane_base + TM_COL_AHI + mt * 128 + c0, r);
            }
            for (int c0 = 0; c0 < 64; c0 += 8) {
                uint32_t r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int kk = 2 * (c0 + i), gi = kk >> 5, c = kk & 31;      // kk even: kk and kk+1 share the gate
                    const float x0 = (ok && c < ncell) ? __ldg(w + (size_t)gi * L * H + (size_t)c * H) : 0.0f;
                    const float x1 = (ok && c + 1 < ncell) ? __ldg(w + (size_t)gi * L * H + (size_t)(c + 1) * H) : 0.0f;
                    r[i] = tm_bf16(__fsub_rn(x0, tm_tf32(x0))) | (tm_bf16(__fsub_rn(x1, tm_tf32(x1))) << 16);
                }
                tm_st8(lane_base + TM_COL_ALO + mt * 64 + c0, r);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }

    bool valid[NPAIR]; int cl_[NPAIR], sl_[NPAIR];
    float wpe[NPAIR][3];
    float nfg[NPAIR], ncerr[NPAIR], ndig[NPAIR], ndfg[NPAIR];   // "next step" state, :253-256
#pragma unroll
    for (int u = 0; u < NPAIR; ++u) {
        const int pr = tid + u * TM_NT;
        cl_[u] = pr % g.CL; sl_[u] = pr / g.CL;
        valid[u] = (cl_[u] < ncell) && (sl_[u] < nseq);
        nfg[u] = ncerr[u] = ndig[u] = ndfg[u] = 0.0f;
        if (valid[u]) {
            const int col = d * H + j0 + cl_[u];
#pragma unroll
            for (int gi = 0; gi < 3; ++gi) wpe[u][gi] = __ldg(p.Wp + gi * L + col);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const uint32_t idesc_tf32 = tm_make_idesc(2u, merge ? NB2 : NB), idesc_bf16 = tm_make_idesc(1u, NB);
    const uint64_t desc_hi = tm_make_desc(Bhl), desc_lo = desc_hi + (uint64_t)(((NB / 8) * 1024) >> 4), desc_bf = tm_make_desc(Bbf);
    const uint64_t kb_step2 = (uint64_t)((NB2 * 128) >> 4), kb_step = (uint64_t)((NB * 128) >> 4);
    const int dstride = merge ? NB2 : NB;                        // accumulator columns per 128-row tile
    const size_t ex_slice = (size_t)NB * R;                      // one producer's [NB][R] block

    for (int q = 0; q < T; ++q) {
        const int t = (d == 0) ? T - 1 - q : q;                 // fw walks time backwards, bw forwards (:936, :970)
        const bool firstCall = (q == 0);
        const bool lastCall = (q == T - 1);
        const bool check = (t >= p.Tmin);
        const int tprev = (d == 0) ? t - 1 : t + 1;
        const float *acts_t = p.acts + (size_t)t * S * 4 * L + d * H + j0;
        const float *cst_t = p.cst + (size_t)t * S * L + d * H + j0;
        const float *cst_p = p.cst + (size_t)(lastCall ? t : tprev) * S * L + d * H + j0;
        float *dy_t = p.dY + (size_t)t * S * p.lddy + d * H + j0;
        float *del_t = p.deltas + (size_t)t * S * 4 * L + d * H + j0;
        float *cerr_t = p.cerr + (size_t)t * S * L + d * H + j0;
        const char *pat_t = p.pat + (size_t)t * S;
        float *ex_w = p.dx + (((size_t)(d * 2 + (q & 1)) * g.G + grp) * g.C + cs) * ex_slice;

        float a[NPAIR][4], c[NPAIR], cp[NPAIR], oe[NPAIR], r_dni[NPAIR], r_dog[NPAIR], esum[NPAIR]; bool dummy[NPAIR];
#pragma unroll
        for (int u = 0; u < NPAIR; ++u) {
            dummy[u] = false; cp[u] = 0.0f; r_dni[u] = r_dog[u] = 0.0f; esum[u] = 0.0f;
            if (valid[u]) {
                const int slot = s0 + sl_[u], cl = cl_[u];
                dummy[u] = check && (pat_t[slot] == BL_PATTYPE_NONE);
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) a[u][gi] = acts_t[slot * 4 * L + gi * L + cl];
                c[u] = cst_t[slot * L + cl];
                if (!lastCall) cp[u] = cst_p[slot * L + cl];
                oe[u] = dy_t[slot * p.lddy + cl];
            }
        }

        if (!firstCall) {
            tm_wait_warp(flag, (unsigned)(g.C * q));
            // the partial products of all C producers of this (direction, group) for this thread's cell, in slice order
            const float *ex_r = p.dx + (((size_t)(d * 2 + ((q - 1) & 1)) * g.G + grp) * g.C) * ex_slice;
#pragma unroll
            for (int u = 0; u < NPAIR; ++u) {
                if (!valid[u]) continue;
                const float *rp = ex_r + (size_t)sl_[u] * R + j0 + cl_[u];
                float s = 0.0f;
#pragma unroll 8
                for (int pp = 0; pp < g.C; ++pp) s = __fadd_rn(s, __ldcg(rp + (size_t)pp * ex_slice));
                esum[u] = s;
            }
        }

#pragma unroll
        for (int u = 0; u < NPAIR; ++u) {
            if (!valid[u]) continue;
            const int slot = s0 + sl_[u], cl = cl_[u];
            float e = oe[u];
            if (!firstCall) e = __fadd_rn(e, esum[u]);           // the 4 addProducts of :939-942
            if (inplace) dy_t[slot * p.lddy + cl] = e;           // unidirectional: tmpOutputErrors IS outputErrors (:907-910)
            float dni, dig, dfg, dog, cerr;
            if (dummy[u]) {                                       // :224-234
                dni = dig = dfg = dog = cerr = 0.0f;
                nfg[u] = 0.0f;
            } else {
                const float ni = a[u][0], ig = a[u][1], fg = a[u][2], og = a[u][3];
                const float tc = tanh1_tab(c[u], s_tab);
                dog = __fmul_rn(__fmul_rn(logistic_deriv(og), tc), e);                                   // :246
                cerr = __fadd_rn(__fmul_rn(__fmul_rn(og, tanh_deriv(tc)), e), __fmul_rn(wpe[u][2], dog)); // :250
                if (!firstCall)                                                                            // :252-262
                    cerr = __fadd_rn(cerr, __fadd_rn(__fadd_rn(__fmul_rn(nfg[u], ncerr[u]), __fmul_rn(wpe[u][0], ndig[u])),
                                                     __fmul_rn(wpe[u][1], ndfg[u])));
                dni = __fmul_rn(__fmul_rn(ig, tanh_deriv(ni)), cerr);                                    // :265
                dfg = lastCall ? 0.0f : __fmul_rn(__fmul_rn(logistic_deriv(fg), cp[u]), cerr);           // :268-275
                dig = __fmul_rn(__fmul_rn(logistic_deriv(ig), ni), cerr);                                // :278
                dni = limited_error(dni); dig = limited_error(dig);                                      // :281-284
                dfg = limited_error(dfg); dog = limited_error(dog);
                nfg[u] = fg;
            }
            ncerr[u] = cerr; ndig[u] = dig; ndfg[u] = dfg;
            r_dni[u] = dni; r_dog[u] = dog;
            if (!lastCall) {                                     // B operand of this step's product: k = gate*32 + cell, row = sequence
                const float dv[4] = {dni, dig, dfg, dog};
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                    const float hi = tm_tf32(dv[gi]);
                    *reinterpret_cast<float *>(Bhl + tm_off_f32(NB2, sl_[u], gi * 32 + cl)) = hi;
                    *reinterpret_cast<float *>(Bhl + tm_off_f32(NB2, NB + sl_[u], gi * 32 + cl)) = __fsub_rn(dv[gi], hi);
                    *reinterpret_cast<unsigned short *>(Bbf + tm_off_bf16(NB, sl_[u], gi * 32 + cl)) = (unsigned short)tm_bf16(dv[gi]);
                }
            }
        }
        if (!lastCall) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (warp == 4) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tm_elect_one()) {
                    for (int mt = 0; mt < MT; ++mt) {          // tf32: K = 128 = 16 MMAs of 8; bf16: 8 MMAs of 16
                        const uint32_t dcol = tmem + TM_COL_D + mt * dstride;
                        if (merge) {
                            for (int ks = 0; ks < 16; ++ks)
                                tm_mma_tf32(dcol, tmem + TM_COL_AHI + mt * 128 + ks * 8, desc_hi + (uint64_t)(ks >> 2) * kb_step2 + (uint64_t)((ks & 3) * 2), idesc_tf32, ks ? 1u : 0u);
                            for (int ks = 0; ks < 8; ++ks)
                                tm_mma_bf16(dcol, tmem + TM_COL_ALO + mt * 64 + ks * 8, desc_bf + (uint64_t)(ks >> 2) * kb_step + (uint64_t)((ks & 3) * 2), idesc_bf16, 1u);
                        } else {                               // small terms first, then the leading one
                            for (int ks = 0; ks < 8; ++ks)
                                tm_mma_bf16(dcol, tmem + TM_COL_ALO + mt * 64 + ks * 8, desc_bf + (uint64_t)(ks >> 2) * kb_step + (uint64_t)((ks & 3) * 2), idesc_bf16, ks ? 1u : 0u);
                            for (int ks = 0; ks < 16; ++ks)
                                tm_mma_tf32(dcol, tmem + TM_COL_AHI + mt * 128 + ks * 8, desc_lo + (uint64_t)(ks >> 2) * kb_step2 + (uint64_t)((ks & 3) * 2), idesc_tf32, 1u);
                            for (int ks = 0; ks < 16; ++ks)
                                tm_mma_tf32(dcol, tmem + TM_COL_AHI + mt * 128 + ks * 8, desc_hi + (uint64_t)(ks >> 2) * kb_step2 + (uint64_t)((ks & 3) * 2), idesc_tf32, 1u);
                        }
                    }
                    tm_commit(&s_bar);
                }
                __syncwarp();
            }
            if (warp < 4) {
                tm_mbar_wait(&s_bar, (uint32_t)(q & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int mt = 0; mt < MT; ++mt)
                    for (int cc = 0; cc < NB; cc += 16) {
                        float dv[16];
                        tm_ld16(tmem + ((uint32_t)(warp * 32) << 16) + TM_COL_D + mt * dstride + cc, dv);
                        if (merge) {
                            float dl[16];
                            tm_ld16(tmem + ((uint32_t)(warp * 32) << 16) + TM_COL_D + mt * dstride + NB + cc, dl);
#pragma unroll
                            for (int i = 0; i < 16; ++i) dv[i] = __fadd_rn(dv[i], dl[i]);
                        }
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (cc + i < nseq) ex_w[(size_t)(cc + i) * R + mt * 128 + warp * 32 + lane] = dv[i];
                    }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            tm_publish(flag);
        }
#pragma unroll
        for (int u = 0; u < NPAIR; ++u) {                        // HBM-only results after the publish
            if (!valid[u]) continue;
            const int slot = s0 + sl_[u], cl = cl_[u];
            float *dp = del_t + slot * 4 * L + cl;
            dp[0] = r_dni[u]; dp[L] = ndig[u]; dp[2 * L] = ndfg[u]; dp[3 * L] = r_dog[u];
            cerr_t[slot * L + cl] = ncerr[u];
            if (p.ds_hi) {
                const int Hq = (H + 3) & ~3;
                const size_t bs = ((size_t)t * S + slot) * p.ld_ds + (size_t)d * Hq + j0 + cl, gs = (size_t)p.ndir * Hq;
                tm_split_store(p.ds_hi, p.ds_lo, bs, r_dni[u]); tm_split_store(p.ds_hi, p.ds_lo, bs + gs, ndig[u]);
                tm_split_store(p.ds_hi, p.ds_lo, bs + 2 * gs, ndfg[u]); tm_split_store(p.ds_hi, p.ds_lo, bs + 3 * gs, r_dog[u]);
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ launch
template <typename Params, typename Kernel>
static int launch_tmem(bl_ctx *ctx, Kernel kernel, const Params &p, const char *name)
{
    const RecGeom &g = p.g;
    const int grid = p.ndir * g.G * g.C;
    BL_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    int per_sm = 0;
    BL_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TM_NT, g.smem));
    if (per_sm < 1 || grid > per_sm * ctx->num_sms)
        return fail(ctx, "%s: %d CTAs cannot be co-resident (%d per SM x %d SMs)", name, grid, per_sm, ctx->num_sms);
    BL_CUDA(ctx, cudaMemsetAsync(p.flags, 0, (size_t)p.ndir * g.G * 32 * sizeof(unsigned), ctx->stream));
    void *args[] = { (void *)&p };
    BL_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(TM_NT), args, g.smem, ctx->stream));
    BL_LAUNCHED(ctx);
    return 0;
}

int launch_lstm_fwd_tmem(bl_ctx *ctx, const RecFwdParams &p)
{
    TimedRegion timed(ctx, 1);
    return p.g.npair == 1 ? launch_tmem(ctx, lstm_fwd_tmem_kernel<1>, p, "lstm_fwd_tmem") : launch_tmem(ctx, lstm_fwd_tmem_kernel<2>, p, "lstm_fwd_tmem");
}

int launch_lstm_bwd_tmem(bl_ctx *ctx, const RecBwdParams &p)
{
    TimedRegion timed(ctx, 2);
    return p.g.npair == 1 ? launch_tmem(ctx, lstm_bwd_tmem_kernel<1>, p, "lstm_bwd_tmem") : launch_tmem(ctx, lstm_bwd_tmem_kernel<2>, p, "lstm_bwd_tmem");
}

} // namespace bl
