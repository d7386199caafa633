// Persistent recurrent kernels, second tensor-memory generation ("tm2").  Same decomposition as lstm_recurrent_tmem.cu -- CTA (d, g, c)
// owns CL cells of direction d for the SG sequences of group g, its slice of the recurrent weights stays in TENSOR MEMORY for the whole
// pass and the per-timestep product runs on tcgen05.mma -- but the per-step protocol is rebuilt around what the round-2 probes measured
// (tools/micro/exchange_probe.cu, gate_math_probe.cu, tcgen05_f16_step.cu; profiles/r02_probes.txt, r02_trace_*.txt):
//
//  * in-band exchange: no step counter, no fence + atomic publish, no re-arming.  Every exchanged 32-bit word carries a one-bit step
//    tag and consumers poll the WORDS themselves (ld.relaxed.gpu) until the tag is the one of the step they wait for; two buffers per
//    direction alternate, so the word a consumer polls for step q still holds step q-2 with the opposite tag until the new value
//    lands (same address, same producer thread: coherence orders the two stores).  Forward: the producer splits h once and ships
//    (fp16 hi | fp16 lo') -- |lo'| <= 1/2, so bit 14 of lo' is free for the tag and consumers store the halves straight into the B
//    tile.  BPTT: the fp32 partial products give up their last mantissa bit (rounded to nearest even, then replaced by the tag).
//    2.0 k instead of 4.0 k cycles per all-gather in tools/micro/exchange_probe.cu (sentinel form of the same idea).
//  * fp16 two-term operands: hi = fp16(x), lo' = fp16((x - hi) * 2^11); W h ~= W_hi h_hi + 2^-11 (W_hi h_lo' + W_lo' h_hi), accumulated
//    in two fp32 column halves of one TMEM accumulator by 32 kind::f16 MMAs per 256-wide step (the tf32 + bf16 scheme needed 48) -- fp16
//    and tf32 carry the same 11 significant bits; the B tile is half as large.  Weights whose magnitude does not fit fp16 are scaled by
//    a per-CTA power of two (exact) and the accumulator is scaled back.
//  * warp specialisation, no CTA barrier inside the step: the 16 gate-math warps hand each 64-wide K-block of the B tile to the control
//    warp through its own mbarrier as soon as its values have arrived, the control warp issues that K-block's MMAs at once (the tensor
//    pipe overlaps the arrival skew of the 8 producers), and tcgen05.commit wakes the gate-math warps.
//  * forward: the weight rows sit in TMEM as row = cell * 4 + gate, so the TMEM quadrant a warp may read holds all four gates of 8
//    cells; four shuffles transpose (gate x sequence) among the 4 lanes of a cell and every thread owns one (cell, sequence) pair with
//    its four pre-activations in registers -- no shared-memory staging, no barrier between the MMA and the gate math.
//  * BPTT keeps the output-stationary slicing of lstm_recurrent_tmem.cu (own deltas x all source cells, partial products
//    reduce-scattered through the exchange buffer, summed in slice order).
//  * layers too wide for TMEM alone (H = 512: C5) keep W_lo' in SHARED memory (tcgen05.mma with a shared-memory A descriptor for that
//    product) and W_hi in TMEM.
// Reference: the two loops of LstmLayer.cu:812-886 (forward) and :936-985 (BPTT) with their functors (:54-137, :190-290).
#include "lstm_recurrent.cuh"
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdlib>

namespace bl {

constexpr int T2_GW = 16;                        // gate-math warps
constexpr int T2_NT = (T2_GW + 1) * 32;          // + the control warp
constexpr int T2_KB_MAX = 8;                     // K-blocks of 64 fp16: forward K = Hp <= 512
constexpr int T2_MT_MAX = 4;                     // BPTT: 128-row tiles of source cells, R <= 512
constexpr int T2_CMAX = 16;                      // BPTT: producers per (direction, group)
constexpr unsigned T2_FTAG = 1u << 30;           // forward exchange word = fp16 hi | fp16 lo' << 16: bit 14 of lo' is the step tag
constexpr int T2_FWD_INIT = 0x40, T2_BWD_INIT = 0x01;   // cudaMemset bytes that give every word tag 1 (the first two steps carry tag 0)
constexpr int T2_SPIN = 1 << 22;                 // polls before a kernel gives up and traps (a protocol bug must not hang the GPU)
constexpr float T2_LO_SCALE = 2048.0f, T2_LO_UNSCALE = 1.0f / 2048.0f;
constexpr size_t T2_MIN_SMEM = 120 * 1024;       // more than half an SM: exactly one CTA (one 512-column TMEM allocation) per SM

static int t2_pad(int x, int m) { return (x + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------------ geometry
// RecGeom fields used: G, C, CL, SG, NT, Hpad (forward: K padded to 64; BPTT: R = source cells padded to 128), Spad (= 16), smem,
// K4 (1: W_lo' lives in shared memory), xelems (words of the exchange buffer, all directions and both parities)
bool choose_geometry_tm2(bool bwd, int H, int S, int ndir, int num_sms, int smem_cap, int forceG, RecGeom *out)
{
    const int Hp = t2_pad(H, 64), R = t2_pad(H, 128);
    const int per_dir = num_sms / ndir;
    bool found = false;
    RecGeom best{};
    for (int C0 = 1; C0 <= per_dir && C0 <= 64; ++C0) {
        int CL = t2_pad(cdiv(H, C0), 4);
        if (CL > 32) continue;
        const int C = cdiv(H, CL);
        if (C != C0) continue;                                   // each (C, CL) once
        if (bwd && C > T2_CMAX) continue;
        for (int G = 1; G <= 64 && G <= S; ++G) {
            if (forceG > 0 && G != forceG) continue;
            if (G * C > per_dir) break;
            const int SG = cdiv(S, G);
            if (SG > 16) continue;
            if ((G - 1) * SG >= S) continue;                     // trailing group would be empty
            int lo_smem = 0;
            size_t smem;
            if (!bwd) {
                const int KB = Hp / 64;
                if (KB > T2_KB_MAX) continue;
                if (Hp + 32 > 512) lo_smem = 1;                  // TMEM columns: W_hi Hp/2 + W_lo' Hp/2 + accumulator 32
                if (lo_smem && Hp / 2 + 32 > 512) continue;
                smem = (size_t)KB * 4096 + (lo_smem ? (size_t)KB * 16384 : 0) + (size_t)128 * 65 * 4 + 2048;
            } else {
                const int MT = R / 128;
                if (MT > T2_MT_MAX) continue;
                if (MT * 160 > 512) lo_smem = 1;                 // per tile: W_hi 64 + W_lo' 64 + accumulator 32 columns
                if (lo_smem && MT * 96 > 512) continue;
                smem = (size_t)2 * 4096 + (lo_smem ? (size_t)MT * 2 * 16384 : 0) + (size_t)64 * 128 * 4 + 2048;
            }
            if (smem < T2_MIN_SMEM) smem = T2_MIN_SMEM;
            if ((int)smem > smem_cap) continue;
            // per step: the all-gather and the MMAs cost the same for every split; the gate math is bound by the FP64 pipe
            // (~3 cycles per (cell, sequence) pair forward, ~1 BPTT) and the polled bytes grow with the group's sequences
            const double cost = (bwd ? 1.0 : 3.0) * CL * SG + (bwd ? (double)C * 12.0 : 0.0) + (double)SG * (bwd ? R : Hp) * 4.0 / 40.0 + 8.0 * C;
            if (!found || cost < best.cost) {
                found = true;
                best = RecGeom{};
                best.G = G; best.C = C; best.CL = CL; best.SG = SG; best.NT = T2_NT; best.nsub = 1; best.npair = 1;
                best.R = bwd ? R : 128; best.Hpad = bwd ? R : Hp; best.RS = bwd ? R : Hp; best.Spad = 16; best.smem = smem; best.cost = cost;
                best.K4 = lo_smem;
                best.xelems = bwd ? (size_t)ndir * 2 * G * C * 16 * R : (size_t)ndir * 2 * S * Hp;
            }
        }
    }
    if (found) *out = best;
    return found;
}

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint32_t t2_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint4 t2_ld_relaxed_v4(const void *p)
{
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned t2_ld_relaxed(const void *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// true while any of the four forward exchange words still carries the other step's tag
__device__ __forceinline__ bool t2_stale(const uint4 &v, unsigned tag) { return (((v.x ^ tag) | (v.y ^ tag) | (v.z ^ tag) | (v.w ^ tag)) & T2_FTAG) != 0u; }

__device__ __forceinline__ void t2_mbar_init(uint64_t *bar, int count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(t2_smem_u32(bar)), "r"(count) : "memory"); }
// single-thread wait (the elected MMA thread)
__device__ __forceinline__ void t2_mbar_wait_thread(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    for (int spin = 0; spin < T2_SPIN && !done; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(t2_smem_u32(bar)), "r"(parity) : "memory");
    if (!done) __trap();
}
// elect.sync: ptxas only emits tcgen05.mma without a per-instruction uniformity loop (ELECT / R2UR.BROADCAST / BRA.U.ANY around every
// UTCHMMA: 90-150 cycles per MMA in the second tm2 trace) when the issuing branch is guarded by the elect predicate itself
__device__ __forceinline__ bool t2_elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void t2_mbar_arrive(uint64_t *bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(t2_smem_u32(bar)) : "memory"); }
// Whole-warp wait: ONE lane polls (512 threads spinning on try_wait slowed the tensor pipe's shared-memory operand reads 2.5x in the
// first tm2 trace), the warp re-converges on it.  A wait that never completes would hang the whole cooperative grid: trap instead
// (the launch then fails loudly).
__device__ __forceinline__ void t2_mbar_wait(uint64_t *bar, uint32_t parity)
{
    if ((threadIdx.x & 31) == 0) {
        uint32_t done = 0;
        for (int spin = 0; spin < T2_SPIN && !done; ++spin)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(t2_smem_u32(bar)), "r"(parity) : "memory");
        if (!done) __trap();
    }
    __syncwarp();
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (same encoding as gemm_tc.cu): SBO = 1024 B, version 1, layout type 2
__device__ __forceinline__ uint64_t t2_make_desc(const void *p)
{
    uint64_t d = 0;
    d |= (uint64_t)((t2_smem_u32(p) & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor of kind::f16: D fp32, A and B fp16 (format 0), both K-major, M = 128
__device__ __forceinline__ uint32_t t2_make_idesc(int N) { return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

__device__ __forceinline__ void t2_mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc)       // A in tensor memory
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void t2_mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc)       // A in shared memory
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void t2_commit(uint64_t *bar)
{ asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(t2_smem_u32(bar)) : "memory"); }

__device__ __forceinline__ void t2_st8(uint32_t taddr, const uint32_t (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void t2_ld4(uint32_t taddr, uint32_t (&r)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}

// x ~= hi + lo' / 2^11 to 2^-22 |x| (x - hi is exact in fp32; the scale keeps lo' out of fp16's subnormal range)
__device__ __forceinline__ void t2_split(float x, uint32_t &hi, uint32_t &lo)
{
    const __half h = __float2half_rn(x);
    hi = (uint32_t)__half_as_ushort(h);
    lo = (uint32_t)__half_as_ushort(__float2half_rn(__fmul_rn(__fsub_rn(x, __half2float(h)), T2_LO_SCALE)));
}

// byte offset of (row r, k) inside one K-block (64 fp16 = 128-byte rows, 8-row atoms of 1024 B, 16-byte chunks XORed with the row)
__device__ __forceinline__ int t2_off(int r, int k)
{ return (r >> 3) * 1024 + (r & 7) * 128 + ((((k & 63) >> 3) ^ (r & 7)) << 4) + (k & 7) * 2; }

// power-of-two scale that brings max|w| of the CTA's slice below 2^15 (1.0 for every sane weight): block-wide, all threads call it
__device__ __forceinline__ float t2_slice_scale(float local_max, float *s_red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local_max;
    __syncthreads();
    float m = 0.0f;
    for (int i = 0; i < T2_NT / 32; ++i) m = fmaxf(m, s_red[i]);
    __syncthreads();
    if (!(m < 3.0e38f)) return 1.0f;                             // inf / NaN weights: nothing to save
    int e = 0;
    while (m >= 32768.0f && e < 120) { m *= 0.5f; ++e; }
    return __int_as_float((127 - e) << 23);                      // 2^-e
}

// ------------------------------------------------------------------------------------------------ forward
// trace row (BLSTM_REC_TRACE): [step start, values polled (warp 0), MMAs complete, gate math done, h stored, results stored,
//                               control warp: first K-block ready, MMAs issued]
template <bool LO_SMEM>
__global__ void __launch_bounds__(T2_NT, 1) lstm_fwd_tm2_kernel(const RecFwdParams p)
{
    extern __shared__ uint8_t t2_smem_raw[];
    __shared__ unsigned long long s_tab[32];
    __shared__ uint64_t s_bar_b[T2_KB_MAX], s_bar_mma, s_bar_dfree;
    __shared__ uint32_t s_slot;
    __shared__ float s_red[T2_NT / 32];
    const RecGeom &g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Hp = g.Hpad, KB = Hp / 64;
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(t2_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *Bt = base;                                           // [KB][32 rows: h_hi 0..15 | h_lo' 16..31][128 B]
    uint8_t *Alo = base + (size_t)KB * 4096;                      // LO_SMEM: [KB][128 rows][128 B]
    float *stage = reinterpret_cast<float *>(Alo + (LO_SMEM ? (size_t)KB * 16384 : 0));     // [128][65] prologue staging

    const int H = p.H, L = p.L, S = p.S, T = p.T;
    const int H4 = (H + 3) & ~3;
    const int d = blockIdx.x / (g.G * g.C);
    const int grp = (blockIdx.x % (g.G * g.C)) / g.C;
    const int cs = blockIdx.x % g.C;
    const int j0 = cs * g.CL, ncell = min(g.CL, H - j0), ncell4 = min(g.CL, H4 - j0);
    const int s0 = grp * g.SG, nseq = min(g.SG, S - s0);
    const uint32_t col_ahi = 0, col_alo = Hp / 2, col_d = LO_SMEM ? Hp / 2 : Hp;

    for (int i = tid; i < KB * 4096 / 4; i += T2_NT) reinterpret_cast<uint32_t *>(Bt)[i] = 0u;      // rows beyond nseq stay zero
    if (tid < 32) s_tab[tid] = bl_exp2f_tab[tid];
    if (tid == 0) {
        for (int kb = 0; kb < T2_KB_MAX; ++kb) t2_mbar_init(&s_bar_b[kb], T2_GW / 2);
        t2_mbar_init(&s_bar_mma, 1);
        t2_mbar_init(&s_bar_dfree, T2_GW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == T2_GW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(t2_smem_u32(&s_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_slot;

    // ---- the CTA's weight slice, once: row r = cell * 4 + gate (lane of TMEM), W_gate[j, k] = Wi[gate*L*H + d*H*H + j*H + k]
    // (k = source cell, LstmLayer.cu:586-596), staged through shared memory in 64-wide k chunks so that the global reads coalesce
    const float *Wd = p.Wi + (size_t)d * H * H;
    float wmax = 0.0f;
    for (int idx = tid; idx < 4 * ncell * H; idx += T2_NT) {
        const int gate = idx / (ncell * H), rem = idx - gate * ncell * H;
        wmax = fmaxf(wmax, fabsf(__ldg(Wd + (size_t)gate * L * H + (size_t)j0 * H + rem)));
    }
    const float wscale = t2_slice_scale(wmax, s_red), wunscale = __frcp_rn(wscale);
    for (int kc = 0; kc < KB; ++kc) {
        for (int idx = tid; idx < 128 * 64; idx += T2_NT) {
            const int r = idx >> 6, kk = idx & 63, cell = r >> 2, gate = r & 3, k = kc * 64 + kk;
            stage[r * 65 + kk] = (cell < ncell && k < H) ? __fmul_rn(__ldg(Wd + (size_t)gate * L * H + (size_t)(j0 + cell) * H + k), wscale) : 0.0f;
        }
        __syncthreads();
        if (warp < 4) {
            const int r = warp * 32 + lane;
            const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
            for (int c0 = 0; c0 < 32; c0 += 8) {
                uint32_t rh[8], rl[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uint32_t h0, l0, h1, l1;
                    t2_split(stage[r * 65 + 2 * (c0 + i)], h0, l0);
                    t2_split(stage[r * 65 + 2 * (c0 + i) + 1], h1, l1);
                    rh[i] = h0 | (h1 << 16); rl[i] = l0 | (l1 << 16);                    // even k in the low half of the column
                }
                t2_st8(lane_base + col_ahi + kc * 32 + c0, rh);
                if (!LO_SMEM) t2_st8(lane_base + col_alo + kc * 32 + c0, rl);
                else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        *reinterpret_cast<uint32_t *>(Alo + (size_t)kc * 16384 + t2_off(r, 2 * (c0 + i))) = rl[i];
                }
            }
        }
        __syncthreads();
    }
    if (warp < 4) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (LO_SMEM) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const size_t xbuf = (size_t)S * Hp;                             // one parity of one direction
    unsigned *xd = reinterpret_cast<unsigned *>(p.hx) + (size_t)d * 2 * xbuf;
    long long *trb = p.trace ? p.trace + (size_t)blockIdx.x * T * 8 : nullptr;

    if (warp == T2_GW) {
        // ================================================================ control warp: MMA issue, K-block by K-block
        const uint32_t idesc32 = t2_make_idesc(32), idesc16 = t2_make_idesc(16);
        const uint64_t desc_b = t2_make_desc(Bt), desc_a = t2_make_desc(Alo);
        if (t2_elect_one()) {                                       // ONE thread runs the whole control loop
            for (int q = 1; q < T; ++q) {
                if (q >= 2) t2_mbar_wait_thread(&s_bar_dfree, (uint32_t)(q & 1));      // every warp has read step q-1's accumulator (phase q-2)
                for (int kb = 0; kb < KB; ++kb) {
                    t2_mbar_wait_thread(&s_bar_b[kb], (uint32_t)((q - 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (trb && kb == 0) trb[q * 8 + 6] = clock64();
                    // 16 k per MMA = 8 TMEM columns = 32 B inside the swizzle atom
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        t2_mma_ts(tmem + col_d, tmem + col_ahi + kb * 32 + ks * 8, desc_b + (uint64_t)(kb * 256 + ks * 2), idesc32, (kb | ks) ? 1u : 0u);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        if (!LO_SMEM) t2_mma_ts(tmem + col_d + 16, tmem + col_alo + kb * 32 + ks * 8, desc_b + (uint64_t)(kb * 256 + ks * 2), idesc16, 1u);
                        else t2_mma_ss(tmem + col_d + 16, desc_a + (uint64_t)(kb * 1024 + ks * 2), desc_b + (uint64_t)(kb * 256 + ks * 2), idesc16, 1u);
                    }
                }
                t2_commit(&s_bar_mma);
                if (trb) trb[q * 8 + 7] = clock64();
            }
        }
        __syncwarp();
    } else {
        // ================================================================ gate-math warps
        const int qd = warp & 3, sgp = warp >> 2;
        const int cell = 8 * qd + (lane >> 2), me = lane & 3, seq = 4 * sgp + me;
        const bool valid = cell < ncell && seq < nseq;
        const bool xvalid = cell < ncell4 && seq < nseq;             // the last slice also owns the zero pad cells up to H4
        const int slot = s0 + seq;
        float wb[4] = {0, 0, 0, 0}, wpe[3] = {0, 0, 0}, cprev = 0.0f;
        if (valid) {
            const int col = d * H + j0 + cell;
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) wb[gi] = __fmul_rn(p.bias, __ldg(p.Wb + gi * L + col));          // bias * w, :97-100
#pragma unroll
            for (int gi = 0; gi < 3; ++gi) wpe[gi] = __ldg(p.Wp + gi * L + col);
        }
        // poll assignment: warps 0..7 take the even K-blocks, 8..15 the odd ones; 256 threads x one float4 = 16 rows x 64 k
        const int ph = warp >> 3, prow = (tid & 255) >> 4, pf4 = tid & 15;
        const uint32_t st_off = (uint32_t)((prow >> 3) * 1024 + (prow & 7) * 128 + (((pf4 >> 1) ^ (prow & 7)) << 4) + (pf4 & 1) * 8);
        const uint32_t tm_lane = tmem + ((uint32_t)(qd * 32) << 16) + col_d + 4 * sgp;
        const bool b0 = me & 1, b1 = me & 2;

        for (int q = 0; q < T; ++q) {
            const int t = (d == 0) ? q : T - 1 - q;
            const bool first = (q == 0);
            const bool check = (t >= p.Tmin);
            long long *tr = (trb && tid == 0) ? trb + q * 8 : nullptr;
            if (tr) tr[0] = clock64();
            float a[4] = {0, 0, 0, 0};
            bool dummy = false;
            float *acts_p = p.acts + ((size_t)t * S + slot) * 4 * L + d * H + j0 + cell;
            if (valid) {
                dummy = check && (p.pat[(size_t)t * S + slot] == BL_PATTYPE_NONE);
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) a[gi] = acts_p[gi * L];
            }
            float rec[4] = {0, 0, 0, 0};
            if (!first) {
                // ---- previous-step outputs of this group's sequences: poll the exchange buffer, split, write the B tile K-block by K-block
                const unsigned *xs = xd + (size_t)((q - 1) & 1) * xbuf + (size_t)(s0 + prow) * Hp + pf4 * 4;
                const unsigned want = (((q - 1) >> 1) & 1) ? T2_FTAG : 0u;        // tag of step q-1
                uint4 v[T2_KB_MAX / 2];
#pragma unroll
                for (int u = 0; u < T2_KB_MAX / 2; ++u) {
                    const int kb = ph + 2 * u;
                    if (kb < KB && prow < nseq && kb * 64 + pf4 * 4 < H4) v[u] = t2_ld_relaxed_v4(xs + kb * 64);
                }
#pragma unroll
                for (int u = 0; u < T2_KB_MAX / 2; ++u) {
                    const int kb = ph + 2 * u;
                    if (kb < KB) {
                        if (prow < nseq && kb * 64 + pf4 * 4 < H4) {
                            int spin = 0;
                            while (t2_stale(v[u], want)) {
                                v[u] = t2_ld_relaxed_v4(xs + kb * 64);
                                if (++spin > T2_SPIN) __trap();
                            }
                            // word = hi | lo' << 16 (tag in bit 30): four hi halves into row prow, four lo' halves into row 16 + prow
                            uint8_t *dst = Bt + (size_t)kb * 4096 + st_off;
                            *reinterpret_cast<uint2 *>(dst) = make_uint2(__byte_perm(v[u].x, v[u].y, 0x5410), __byte_perm(v[u].z, v[u].w, 0x5410));
                            *reinterpret_cast<uint2 *>(dst + 2048) = make_uint2(__byte_perm(v[u].x, v[u].y, 0x7632) & 0xBFFFBFFFu,
                                                                                __byte_perm(v[u].z, v[u].w, 0x7632) & 0xBFFFBFFFu);
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) t2_mbar_arrive(&s_bar_b[kb]);
                    }
                }
                if (tr) tr[1] = clock64();
                // ---- the step product: all four gates of 8 cells x 4 sequences per warp out of tensor memory
                t2_mbar_wait(&s_bar_mma, (uint32_t)((q - 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tr) tr[2] = clock64();
                uint32_t r0[4], r1[4];
                t2_ld4(tm_lane, r0);
                t2_ld4(tm_lane + 16, r1);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) t2_mbar_arrive(&s_bar_dfree);
                float x[4];
#pragma unroll
                for (int i = 0; i < 4; ++i)                          // column i = sequence 4*sgp + i, this lane's gate = me
                    x[i] = __fmul_rn(__fmaf_rn(__uint_as_float(r1[i]), T2_LO_UNSCALE, __uint_as_float(r0[i])), wunscale);
                // 4 x 4 transpose among the lanes of a cell: afterwards this lane holds the four gates of sequence 4*sgp + me
                const float k0 = b0 ? x[1] : x[0], k1 = b0 ? x[3] : x[2];               // columns b0, 2 + b0 of gate me
                const float y0 = __shfl_xor_sync(0xffffffffu, b0 ? x[0] : x[1], 1);     // ... of gate me ^ 1
                const float y1 = __shfl_xor_sync(0xffffffffu, b0 ? x[2] : x[3], 1);
                const float keepA = b1 ? k1 : k0, keepB = b1 ? y1 : y0;                 // column me of gates me, me ^ 1
                const float ra = __shfl_xor_sync(0xffffffffu, b1 ? k0 : k1, 2);         // column me of gate me ^ 2
                const float rb = __shfl_xor_sync(0xffffffffu, b1 ? y0 : y1, 2);         // column me of gate me ^ 3
                const float pe = b0 ? keepB : keepA, po = b0 ? keepA : keepB;           // gates 2*b1, 2*b1 + 1
                const float oe = b0 ? rb : ra, oo = b0 ? ra : rb;                       // gates 2*!b1, 2*!b1 + 1
                rec[0] = b1 ? oe : pe; rec[1] = b1 ? oo : po; rec[2] = b1 ? pe : oe; rec[3] = b1 ? po : oo;
            }

            // ---- gate math (ComputeBlockOutputFn, LstmLayer.cu:54-137)
            float ni = 0, ig = 0, fg = 0, og = 0, h = 0.0f, c = 0.0f;
            if (valid && !dummy) {                                    // padded slots: h = c = 0, activations untouched (:78-85)
                ni = a[0]; ig = a[1]; fg = a[2]; og = a[3];
                if (!first) {                                         // recurrent addProduct, :815-818
                    ni = __fadd_rn(ni, rec[0]); ig = __fadd_rn(ig, rec[1]); fg = __fadd_rn(fg, rec[2]); og = __fadd_rn(og, rec[3]);
                }
                ni = __fadd_rn(ni, wb[0]); ig = __fadd_rn(ig, wb[1]); fg = __fadd_rn(fg, wb[2]); og = __fadd_rn(og, wb[3]);
                if (!first) {                                         // :103-108
                    ig = __fadd_rn(ig, __fmul_rn(cprev, wpe[0]));
                    fg = __fadd_rn(fg, __fmul_rn(cprev, wpe[1]));
                }
                ni = tanh_fn_tab(ni, s_tab); ig = logistic_fn_tab(ig, s_tab); fg = logistic_fn_tab(fg, s_tab);
                c = __fmul_rn(ni, ig);                                // :121-126
                if (!first) c = __fadd_rn(c, __fmul_rn(cprev, fg));
                og = __fadd_rn(og, __fmul_rn(c, wpe[2]));             // :129-131
                og = logistic_fn_tab(og, s_tab);
                h = __fmul_rn(tanh_fn_tab(c, s_tab), og);             // :134
            }
            cprev = c;
            if (tr) tr[3] = clock64();
            // ---- the exchange word first: h split once here, (fp16 hi | fp16 lo' << 16), step tag in the free bit 14 of lo'
            if (q + 1 < T && xvalid) {
                uint32_t hh, hl;
                t2_split(h, hh, hl);
                xd[(size_t)(q & 1) * xbuf + (size_t)slot * Hp + j0 + cell] = hh | ((hl & 0xBFFFu) << 16) | (((q >> 1) & 1) ? T2_FTAG : 0u);
            }
            if (tr) tr[4] = clock64();
            if (valid) {
                if (!dummy) { acts_p[0] = ni; acts_p[L] = ig; acts_p[2 * L] = fg; acts_p[3 * L] = og; }
                p.cst[((size_t)t * S + slot) * L + d * H + j0 + cell] = c;
                p.Y[((size_t)t * S + slot) * p.ldy + d * H + j0 + cell] = h;
                if (p.ys_hi) {                                        // TF32 split of the layer output for the backward GEMMs
                    const size_t idx = ((size_t)t * S + slot) * p.ld_ys + d * H4 + j0 + cell;
                    uint32_t hh; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hh) : "f"(h));
                    p.ys_hi[idx] = __uint_as_float(hh);
                    p.ys_lo[idx] = __fsub_rn(h, __uint_as_float(hh));
                }
            }
            if (tr) tr[5] = clock64();
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == T2_GW) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ BPTT
// Output-stationary slicing: the CTA that owns the cells j of a slice has just produced their four gate deltas, so it multiplies THOSE
// (K = 4 gates x 32 cells = 128, straight from registers into the B tile) with the weight columns of ALL source cells k' of its
// direction (M = R = pad128(H): MT tiles of 128 TMEM lanes):
//     Q[k', s] = sum_{gate, j in slice} W_gate[j, k'] * delta_gate[j, s]          A[k'][gate*32 + c] = Wi[gate*L*H + d*H*H + (j0+c)*H + k']
// and publishes its partial Q through the exchange buffer [dir][parity 2][group][producer slice][sequence 16][k' R]; next step every CTA
// adds the C partials of its own cells in slice order (deterministic) to the output error -- the 4 addProducts of LstmLayer.cu:939-942.
// A partial travels as fp32 with its last mantissa bit rounded away (to nearest even) and replaced by the step tag.
// trace row: [step start, partials polled, deltas in the B tile, MMAs complete, partials stored, -, control: B tile ready, MMAs issued]
template <bool LO_SMEM, int CMAX, int MTMAX>
__global__ void __launch_bounds__(T2_NT, 1) lstm_bwd_tm2_kernel(const RecBwdParams p)
{
    extern __shared__ uint8_t t2_smem_raw[];
    __shared__ unsigned long long s_tab[32];
    __shared__ uint64_t s_bar_b, s_bar_mma;
    __shared__ uint32_t s_slot;
    __shared__ float s_red[T2_NT / 32];
    const RecGeom &g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = g.Hpad, MT = R / 128;
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(t2_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *Bt = base;                                           // [2 K-blocks][32 rows: d_hi | d_lo'][128 B], k = gate*32 + cell
    uint8_t *Alo = base + 2 * 4096;                               // LO_SMEM: [MT][2][128 rows][128 B]
    float *stage = reinterpret_cast<float *>(Alo + (LO_SMEM ? (size_t)MT * 2 * 16384 : 0));      // [64 kk][128 k'] prologue staging

    const int H = p.H, L = p.L, S = p.S, T = p.T;
    const int H4 = (H + 3) & ~3;
    const int d = blockIdx.x / (g.G * g.C);
    const int grp = (blockIdx.x % (g.G * g.C)) / g.C;
    const int cs = blockIdx.x % g.C;
    const int j0 = cs * g.CL, ncell = min(g.CL, H - j0);
    const int s0 = grp * g.SG, nseq = min(g.SG, S - s0);
    const bool inplace = (p.ndir == 1);
    const uint32_t col_ahi = 0, col_alo = MT * 64, col_d = LO_SMEM ? MT * 64 : MT * 128;

    for (int i = tid; i < 2 * 4096 / 4; i += T2_NT) reinterpret_cast<uint32_t *>(Bt)[i] = 0u;       // never-written entries stay zero
    if (tid < 32) s_tab[tid] = bl_exp2f_tab[tid];
    if (tid == 0) {
        t2_mbar_init(&s_bar_b, T2_GW);
        t2_mbar_init(&s_bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == T2_GW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(t2_smem_u32(&s_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_slot;

    // ---- weights, once.  Tile mt, lane = source cell k' - mt*128, K index kk = gate*32 + c; staged [kk][k'] in halves of 64 kk
    const float *Wd = p.Wi + (size_t)d * H * H;
    float wmax = 0.0f;
    for (int idx = tid; idx < 4 * ncell * H; idx += T2_NT) {
        const int gate = idx / (ncell * H), rem = idx - gate * ncell * H;
        wmax = fmaxf(wmax, fabsf(__ldg(Wd + (size_t)gate * L * H + (size_t)j0 * H + rem)));
    }
    const float wscale = t2_slice_scale(wmax, s_red), wunscale = __frcp_rn(wscale);
    for (int mt = 0; mt < MT; ++mt)
        for (int half = 0; half < 2; ++half) {
            for (int idx = tid; idx < 64 * 128; idx += T2_NT) {
                const int kl = idx >> 7, kr = idx & 127, kk = half * 64 + kl, gate = kk >> 5, c = kk & 31, ksrc = mt * 128 + kr;
                stage[idx] = (c < ncell && ksrc < H) ? __fmul_rn(__ldg(Wd + (size_t)gate * L * H + (size_t)(j0 + c) * H + ksrc), wscale) : 0.0f;
            }
            __syncthreads();
            if (warp < 4) {
                const int r = warp * 32 + lane;
                const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
                for (int c0 = 0; c0 < 32; c0 += 8) {
                    uint32_t rh[8], rl[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        uint32_t h0, l0, h1, l1;
                        t2_split(stage[(2 * (c0 + i)) * 128 + r], h0, l0);
                        t2_split(stage[(2 * (c0 + i) + 1) * 128 + r], h1, l1);
                        rh[i] = h0 | (h1 << 16); rl[i] = l0 | (l1 << 16);
                    }
                    t2_st8(lane_base + col_ahi + mt * 64 + half * 32 + c0, rh);
                    if (!LO_SMEM) t2_st8(lane_base + col_alo + mt * 64 + half * 32 + c0, rl);
                    else {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            *reinterpret_cast<uint32_t *>(Alo + (size_t)(mt * 2 + half) * 16384 + t2_off(r, 2 * (c0 + i))) = rl[i];
                    }
                }
            }
            __syncthreads();
        }
    if (warp < 4) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (LO_SMEM) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    const size_t ex_slice = (size_t)16 * R;                         // one producer's [16][R] block
    const size_t ex_par = (size_t)g.G * g.C * ex_slice;             // one parity of one direction
    unsigned *exd = reinterpret_cast<unsigned *>(p.dx) + (size_t)d * 2 * ex_par + (size_t)grp * g.C * ex_slice;      // this group's C blocks, parity 0
    long long *trb = p.trace ? p.trace + (size_t)blockIdx.x * T * 8 : nullptr;

    if (warp == T2_GW) {
        // ================================================================ control warp
        const uint32_t idesc32 = t2_make_idesc(32), idesc16 = t2_make_idesc(16);
        const uint64_t desc_b = t2_make_desc(Bt), desc_a = t2_make_desc(Alo);
        if (t2_elect_one()) {                                       // ONE thread runs the whole control loop
            for (int q = 0; q + 1 < T; ++q) {
                t2_mbar_wait_thread(&s_bar_b, (uint32_t)(q & 1));  // all deltas of step q are in the B tile; step q-1's accumulators were read
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (trb) trb[q * 8 + 6] = clock64();
                for (int mt = 0; mt < MT; ++mt) {
                    const uint32_t dcol = tmem + col_d + mt * 32;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)                  // K = 128 = 8 MMAs of 16
                        t2_mma_ts(dcol, tmem + col_ahi + mt * 64 + ks * 8, desc_b + (uint64_t)((ks >> 2) * 256 + (ks & 3) * 2), idesc32, ks ? 1u : 0u);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        if (!LO_SMEM) t2_mma_ts(dcol + 16, tmem + col_alo + mt * 64 + ks * 8, desc_b + (uint64_t)((ks >> 2) * 256 + (ks & 3) * 2), idesc16, 1u);
                        else t2_mma_ss(dcol + 16, desc_a + (uint64_t)((mt * 2 + (ks >> 2)) * 1024 + (ks & 3) * 2), desc_b + (uint64_t)((ks >> 2) * 256 + (ks & 3) * 2), idesc16, 1u);
                    }
                }
                t2_commit(&s_bar_mma);
                if (trb) trb[q * 8 + 7] = clock64();
            }
        }
        __syncwarp();
    } else {
        // ================================================================ gate-math warps: warp = sequence, lane = cell
        const int seq = warp, cell = lane;
        const bool valid = cell < ncell && seq < nseq;
        const int slot = s0 + seq;
        float wpe[3] = {0, 0, 0};
        float nfg = 0.0f, ncerr = 0.0f, ndig = 0.0f, ndfg = 0.0f;  // "next step" state, :253-256
        if (valid) {
            const int col = d * H + j0 + cell;
#pragma unroll
            for (int gi = 0; gi < 3; ++gi) wpe[gi] = __ldg(p.Wp + gi * L + col);
        }
        const int qd = warp & 3, sgp = warp >> 2;                   // epilogue: TMEM quadrant and sequence quad of this warp
        const uint32_t tm_lane = tmem + ((uint32_t)(qd * 32) << 16) + col_d + 4 * sgp;

        for (int q = 0; q < T; ++q) {
            const int t = (d == 0) ? T - 1 - q : q;                 // fw walks time backwards, bw forwards (:936, :970)
            const bool firstCall = (q == 0), lastCall = (q == T - 1);
            const bool check = (t >= p.Tmin);
            const int tprev = (d == 0) ? t - 1 : t + 1;
            long long *tr = (trb && tid == 0) ? trb + q * 8 : nullptr;
            if (tr) tr[0] = clock64();
            const size_t row = (size_t)t * S + slot;
            float a[4] = {0, 0, 0, 0}, c = 0.0f, cp = 0.0f, oe = 0.0f;
            bool dummy = false;
            if (valid) {
                dummy = check && (p.pat[row] == BL_PATTYPE_NONE);
                const float *ap = p.acts + row * 4 * L + d * H + j0 + cell;
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) a[gi] = ap[gi * L];
                c = p.cst[row * L + d * H + j0 + cell];
                if (!lastCall) cp = p.cst[((size_t)tprev * S + slot) * L + d * H + j0 + cell];
                oe = p.dY[row * p.lddy + d * H + j0 + cell];
            }
            float e = oe;
            if (!firstCall && valid) {
                // the partial products of all C producers of this (direction, group) for this thread's cell, added in slice order
                const unsigned *ex_r = exd + (size_t)((q - 1) & 1) * ex_par + (size_t)seq * R + j0 + cell;
                const unsigned want = ((q - 1) >> 1) & 1;            // tag of step q-1
                unsigned pv[CMAX];
#pragma unroll
                for (int pp = 0; pp < CMAX; ++pp)
                    if (pp < g.C) pv[pp] = t2_ld_relaxed(ex_r + (size_t)pp * ex_slice);
                float s = 0.0f;
#pragma unroll
                for (int pp = 0; pp < CMAX; ++pp)
                    if (pp < g.C) {
                        int spin = 0;
                        while ((pv[pp] & 1u) != want) {
                            pv[pp] = t2_ld_relaxed(ex_r + (size_t)pp * ex_slice);
                            if (++spin > T2_SPIN) __trap();
                        }
                        s = __fadd_rn(s, __uint_as_float(pv[pp] & ~1u));
                    }
                e = __fadd_rn(oe, s);                                // the 4 addProducts of :939-942
            }
            if (tr) tr[1] = clock64();

            float dni = 0, dig = 0, dfg = 0, dog = 0, cerr = 0;
            if (valid) {
                if (dummy) {                                         // :224-234
                    nfg = 0.0f;
                } else {
                    const float ni = a[0], ig = a[1], fg = a[2], og = a[3];
                    const float tc = tanh_fn_tab(c, s_tab);
                    dog = __fmul_rn(__fmul_rn(logistic_deriv(og), tc), e);                                   // :246
                    cerr = __fadd_rn(__fmul_rn(__fmul_rn(og, tanh_deriv(tc)), e), __fmul_rn(wpe[2], dog));   // :250
                    if (!firstCall)                                                                          // :252-262
                        cerr = __fadd_rn(cerr, __fadd_rn(__fadd_rn(__fmul_rn(nfg, ncerr), __fmul_rn(wpe[0], ndig)), __fmul_rn(wpe[1], ndfg)));
                    dni = __fmul_rn(__fmul_rn(ig, tanh_deriv(ni)), cerr);                                    // :265
                    dfg = lastCall ? 0.0f : __fmul_rn(__fmul_rn(logistic_deriv(fg), cp), cerr);             // :268-275
                    dig = __fmul_rn(__fmul_rn(logistic_deriv(ig), ni), cerr);                                // :278
                    dni = limited_error(dni); dig = limited_error(dig);                                      // :281-284
                    dfg = limited_error(dfg); dog = limited_error(dog);
                    nfg = fg;
                }
                ncerr = cerr; ndig = dig; ndfg = dfg;
                if (!lastCall) {                                     // B operand of this step's product: k = gate*32 + cell, row = sequence
                    const float dv[4] = {dni, dig, dfg, dog};
#pragma unroll
                    for (int gi = 0; gi < 4; ++gi) {
                        uint32_t hi, lo;
                        t2_split(dv[gi], hi, lo);
                        uint8_t *dst = Bt + (gi >> 1) * 4096 + t2_off(seq, (gi & 1) * 32 + cell);
                        *reinterpret_cast<unsigned short *>(dst) = (unsigned short)hi;
                        *reinterpret_cast<unsigned short *>(dst + 2048) = (unsigned short)lo;          // row + 16
                    }
                }
            }
            if (!lastCall) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) t2_mbar_arrive(&s_bar_b);
            }
            if (tr) tr[2] = clock64();
            // ---- HBM-only results while the tensor pipe works
            if (valid) {
                if (inplace) p.dY[row * p.lddy + d * H + j0 + cell] = e;    // unidirectional: tmpOutputErrors IS outputErrors (:907-910)
                float *dp = p.deltas + row * 4 * L + d * H + j0 + cell;
                dp[0] = dni; dp[L] = dig; dp[2 * L] = dfg; dp[3 * L] = dog;
                p.cerr[row * L + d * H + j0 + cell] = cerr;
                if (p.ds_hi) {
                    const size_t bs = row * p.ld_ds + (size_t)d * H4 + j0 + cell, gs = (size_t)p.ndir * H4;
                    const float dv[4] = {dni, dig, dfg, dog};
#pragma unroll
                    for (int gi = 0; gi < 4; ++gi) {
                        uint32_t hh; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hh) : "f"(dv[gi]));
                        p.ds_hi[bs + gi * gs] = __uint_as_float(hh);
                        p.ds_lo[bs + gi * gs] = __fsub_rn(dv[gi], __uint_as_float(hh));
                    }
                }
            }
            if (!lastCall) {
                // ---- partial products out of tensor memory into the exchange buffer: lane = source cell, 4 sequences per warp
                t2_mbar_wait(&s_bar_mma, (uint32_t)(q & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tr) tr[3] = clock64();
                unsigned *ex_w = exd + (size_t)(q & 1) * ex_par + (size_t)cs * ex_slice;
                const unsigned tag = (q >> 1) & 1;
#pragma unroll
                for (int mt = 0; mt < MTMAX; ++mt)
                    if (mt < MT) {
                        uint32_t r0[4], r1[4];
                        t2_ld4(tm_lane + mt * 32, r0);
                        t2_ld4(tm_lane + mt * 32 + 16, r1);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        const int ksrc = mt * 128 + qd * 32 + lane;
                        if (ksrc < H4) {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (4 * sgp + i < nseq) {
                                    const float o = __fmul_rn(__fmaf_rn(__uint_as_float(r1[i]), T2_LO_UNSCALE, __uint_as_float(r0[i])), wunscale);
                                    // last mantissa bit rounded away (nearest even; inf / NaN keep their class), then the tag
                                    const unsigned bb = __float_as_uint(o);
                                    const unsigned rr = ((bb & 0x7F800000u) == 0x7F800000u) ? (bb & ~1u) : ((bb + ((bb >> 1) & 1u)) & ~1u);
                                    ex_w[(size_t)(4 * sgp + i) * R + ksrc] = rr | tag;
                                }
                        }
                    }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (tr) tr[4] = clock64();
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == T2_GW) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ launch
template <typename Params, typename Kernel>
static int launch_tm2(bl_ctx *ctx, Kernel kernel, const Params &p, float *xbuf, int init_byte, const char *name)
{
    const RecGeom &g = p.g;
    const int grid = p.ndir * g.G * g.C;
    BL_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    int per_sm = 0;
    BL_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, T2_NT, g.smem));
    if (per_sm != 1 || grid > ctx->num_sms)
        return fail(ctx, "%s: %d CTAs cannot be co-resident one per SM (%d per SM x %d SMs)", name, grid, per_sm, ctx->num_sms);
    // every word of both parities starts with step tag 1: the first two steps of a pass carry tag 0
    BL_CUDA(ctx, cudaMemsetAsync(xbuf, init_byte, g.xelems * sizeof(float), ctx->stream));
    void *args[] = { (void *)&p };
    BL_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(T2_NT), args, g.smem, ctx->stream));
    BL_LAUNCHED(ctx);
    return 0;
}

int launch_lstm_fwd_tm2(bl_ctx *ctx, const RecFwdParams &p)
{
    TimedRegion timed(ctx, 1);
    return p.g.K4 ? launch_tm2(ctx, lstm_fwd_tm2_kernel<true>, p, p.hx, T2_FWD_INIT, "lstm_fwd_tm2")
                  : launch_tm2(ctx, lstm_fwd_tm2_kernel<false>, p, p.hx, T2_FWD_INIT, "lstm_fwd_tm2");
}

int launch_lstm_bwd_tm2(bl_ctx *ctx, const RecBwdParams &p)
{
    TimedRegion timed(ctx, 2);
    if (p.g.K4) return launch_tm2(ctx, lstm_bwd_tm2_kernel<true, T2_CMAX, T2_MT_MAX>, p, p.dx, T2_BWD_INIT, "lstm_bwd_tm2");
    if (p.g.C <= 8 && p.g.Hpad <= 256) return launch_tm2(ctx, lstm_bwd_tm2_kernel<false, 8, 2>, p, p.dx, T2_BWD_INIT, "lstm_bwd_tm2");
    return launch_tm2(ctx, lstm_bwd_tm2_kernel<false, T2_CMAX, 3>, p, p.dx, T2_BWD_INIT, "lstm_bwd_tm2");
}

} // namespace bl
