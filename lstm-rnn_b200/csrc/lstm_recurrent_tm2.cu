// Persistent recurrent kernels, second tensor-memory generation ("tm2").  Same decomposition as lstm_recurrent_tmem.cu -- CTA (d, g, c)
// owns CL cells of direction d for the SG sequences of group g, its slice of the recurrent weights stays in TENSOR MEMORY for the whole
// pass and the per-timestep product runs on tcgen05.mma -- but the per-step protocol is rebuilt around what the round-2 probes measured
// (tools/micro/exchange_probe.cu, gate_math_probe.cu, tcgen05_f16_step.cu; profiles/r02_probes.txt, r02_recurrent_trace.txt):
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
//  * thread mapping: warp = sequence, lane = cell in both kernels, so every global access of a warp is one 128-byte line (a first
//    version read the accumulator as 8 cells x 4 gates x 4 sequences per warp and transposed with shuffles: no shared-memory hop, but
//    4 lines per access -- 830 wavefronts per step in the SM's single load / store queue, ahead of the exchange words -- and 32
//    tcgen05.ld per step).  Forward: four warps per sub-group (warp = gate = TMEM quadrant) pull the accumulator out with wide
//    tcgen05.ld (the port moves 64 B per cycle plus ~8 cycles per instruction) and stage it [gate][sequence][cell] in shared memory.
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

// A CTA runs SUB = 1 or 2 independent SUB-GROUPS on the same weight slice: 16 / SUB gate-math warps and one control warp each, every
// sub-group with its own sequence group (up to 16 / SUB sequences), B tile, accumulator columns and barriers -- an attempt to let one
// sub-group's MMAs and gate math fill the other's exchange wait (the driver admits only ONE CTA per SM for a kernel that allocates
// tensor memory, tools/micro/occupancy_probe.cu, so any interleaving has to happen inside the CTA).  Measured: no gain, the two drift to
// a random relative phase; SUB = 2 stays as a tuning option (BLSTM_T2_SUB=2).  Threads = (16 + SUB) * 32.
constexpr int T2_KB_MAX = 8;                     // K-blocks of 64 fp16: forward K = Hp <= 512
constexpr int T2_MT_MAX = 4;                     // BPTT: 128-row tiles of source cells, R <= 512
constexpr int T2_CMAX = 16;                      // BPTT: producers per (direction, group)
constexpr unsigned T2_FTAG = 1u << 30;           // forward exchange word = fp16 hi | fp16 lo' << 16: bit 14 of lo' is the step tag
constexpr int T2_FWD_INIT = 0x40, T2_BWD_INIT = 0x01;   // cudaMemset bytes that give every word tag 1 (the first two steps carry tag 0)
constexpr int T2_SPIN = 1 << 22;                 // polls before a kernel gives up and traps (a protocol bug must not hang the GPU)
constexpr float T2_LO_SCALE = 2048.0f, T2_LO_UNSCALE = 1.0f / 2048.0f;
constexpr float T2_DELTA_SCALE = 8192.0f;        // BPTT B operand: power-of-two scale of the gate deltas (undone with the weight scale)
constexpr size_t T2_MIN_SMEM = 120 * 1024;       // more than half an SM: exactly one CTA (one 512-column TMEM allocation) per SM

static int t2_pad(int x, int m) { return (x + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------------ geometry
// RecGeom fields used: G (sequence groups = sub-groups), C, CL, SG, NT (= (16 + SUB) * 32), nsub (SUB: sub-groups per CTA), Hpad
// (forward: K padded to 64; BPTT: R = source cells padded to 128), Spad (= 16 / SUB, the sequences a sub-group can hold), smem,
// K4 (1: W_lo' lives in shared memory), RP / npair (tuning switches of the exchange poll), xelems (words of the exchange buffer, all
// directions and both parities)
bool choose_geometry_tm2(bool bwd, int H, int S, int ndir, int num_sms, int smem_cap, int forceG, RecGeom *out)
{
    const int Hp = t2_pad(H, 64), R = t2_pad(H, 128);
    const char *es = getenv("BLSTM_T2_SUB");                     // tuning: sub-groups per CTA, 1 or 2 (default: whichever the model prefers)
    const int force_sub = es ? atoi(es) : 0;
    const char *eg = getenv("BLSTM_T2_POLL_DELAY");              // tuning: ns a warp sleeps between its exchange store and its first poll
    const int poll_delay = eg ? atoi(eg) : 0;
    const char *ep = getenv("BLSTM_T2_POLL");                    // tuning: bit 0 first poll ahead of the loop edge, bit 1 re-poll all stale words at once, >> 2 = back-off ns
    const int poll = ep ? atoi(ep) : 2;
    const int per_dir = num_sms / ndir;
    bool found = false;
    RecGeom best{};
    for (int sub = 2; sub >= 1; --sub) {
        if (force_sub && sub != force_sub) continue;
        const int NS = 16 / sub;                                 // sequences per sub-group
        for (int C0 = 1; C0 <= per_dir && C0 <= 64; ++C0) {
            int CL = t2_pad(cdiv(H, C0), 4);
            if (CL > 32) continue;
            const int C = cdiv(H, CL);
            if (C != C0) continue;                               // each (C, CL) once
            if (bwd && C > T2_CMAX) continue;
            for (int G = 1; G <= 128 && G <= S; ++G) {           // G sequence groups = sub-groups; cdiv(G, sub) CTAs per cell slice
                if (forceG > 0 && G != forceG) continue;
                if (cdiv(G, sub) * C > per_dir) break;
                const int SG = cdiv(S, G);
                if (SG > NS) continue;
                if ((G - 1) * SG >= S) continue;                 // trailing group would be empty
                int lo_smem = 0;
                size_t smem;
                if (!bwd) {
                    const int KB = Hp / 64;
                    if (KB > T2_KB_MAX) continue;
                    if (Hp + 32 > 512) lo_smem = 1;              // TMEM columns: W_hi Hp/2 + W_lo' Hp/2 + accumulators 32
                    if (lo_smem && Hp / 2 + 32 > 512) continue;
                    smem = (size_t)KB * 4096 + (lo_smem ? (size_t)KB * 16384 : 0) + (size_t)128 * 65 * 4 + (size_t)4 * 16 * 32 * 4 + 2048;
                } else {
                    const int MT = R / 128;
                    if (MT > T2_MT_MAX) continue;
                    if (MT * 160 > 512) lo_smem = 1;             // per tile: W_hi 64 + W_lo' 64 + accumulators 32 columns
                    if (lo_smem && MT * 96 > 512) continue;
                    smem = (size_t)2 * 4096 + (lo_smem ? (size_t)MT * 2 * 16384 : 0) + (size_t)64 * 128 * 4 + 2048;
                }
                if (smem < T2_MIN_SMEM) smem = T2_MIN_SMEM;
                if ((int)smem > smem_cap) continue;
                // per step: the all-gather latency is the same for every split and a second sub-group hides most of it; the gate math
                // is bound by the FP64 pipe, i.e. by the warps of the CTA that carry at least one (cell, sequence) pair
                const int warps = SG * (G >= sub ? sub : 1);             // warp = sequence
                // (measured: two sub-groups drift to a random relative phase and gain nothing on C2 / C3 / C5 -- 5.13 k against 5.33 k cycles
                // forward, 5.44 k against 5.32 k BPTT -- so a second sub-group is only used when asked for)
                const double chain = bwd ? 4200.0 : 4200.0, math = (bwd ? 45.0 : 70.0) * warps;
                const double cost = chain + (sub == 2 ? 100.0 : 0.0) + math + 2.0 * C;
                if (!found || cost < best.cost) {
                    found = true;
                    best = RecGeom{};
                    best.G = G; best.C = C; best.CL = CL; best.SG = SG; best.NT = (16 + sub) * 32; best.nsub = sub; best.npair = poll;
                    best.R = bwd ? R : 128; best.Hpad = bwd ? R : Hp; best.RS = bwd ? R : Hp; best.Spad = NS; best.smem = smem; best.cost = cost;
                    best.K4 = lo_smem; best.RP = poll_delay;
                    best.xelems = bwd ? (size_t)ndir * 2 * G * C * NS * R : (size_t)ndir * 2 * S * Hp;
                }
            }
        }
    }
    if (found) *out = best;
    return found;
}

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ int cdiv_dev(int a, int b) { return (a + b - 1) / b; }
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
// single-thread wait (the elected MMA thread).  A wait that never completes would hang the whole cooperative grid: trap instead (the
// launch then fails loudly).
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
// Whole-warp wait on an mbarrier.  Measured in this kernel: 300-650 cycles from the last arrival until 16 waiting warps have all passed
// (every lane polling is the fastest form; one polling lane or a suspend-time hint are slower) against ~30 for a named barrier, so
// mbarriers are used only where the tensor pipe or the single control thread is on the other side; warp-to-warp hand-overs go
// through bar.sync.
__device__ __forceinline__ void t2_mbar_wait(uint64_t *bar, uint32_t parity)
{
    t2_mbar_wait_thread(bar, parity);
    __syncwarp();
}
__device__ __forceinline__ void t2_named_barrier(int id, int threads)
{ asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory"); }

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

__device__ __forceinline__ void t2_ld8(uint32_t taddr, uint32_t (&r)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void t2_ld16(uint32_t taddr, uint32_t (&r)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
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
__device__ __forceinline__ float t2_slice_scale(float local_max, float *s_red, int nwarps)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local_max;
    __syncthreads();
    float m = 0.0f;
    for (int i = 0; i < nwarps; ++i) m = fmaxf(m, s_red[i]);
    __syncthreads();
    if (!(m < 3.0e38f)) return 1.0f;                             // inf / NaN weights: nothing to save
    int e = 0;
    while (m >= 32768.0f && e < 120) { m *= 0.5f; ++e; }
    return __int_as_float((127 - e) << 23);                      // 2^-e
}

// ------------------------------------------------------------------------------------------------ forward
// trace row (BLSTM_REC_TRACE): [step start, values polled (warp 0), MMAs complete, accumulator staged and read back, gate math done,
//                               h and results stored, control warp: first K-block ready, MMAs issued]
// registers: the hardware allocates warps in fours, so 17 or 18 warps count as 20: 96 registers per thread
template <int SUB, bool LO_SMEM>
__global__ void __maxnreg__(96) lstm_fwd_tm2_kernel(const RecFwdParams p)
{
    constexpr int GW = 16 / SUB, NT = (16 + SUB) * 32, NS = GW;   // gate-math warps and sequences (= B tile rows per half) of a sub-group
    constexpr int KBB = 2 * NS * 128;                             // bytes of one K-block of a B tile: rows [0, NS) h_hi, [NS, 2 NS) h_lo'
    extern __shared__ uint8_t t2_smem_raw[];
    __shared__ unsigned long long s_tab[32];
    __shared__ uint64_t s_bar_b[SUB][T2_KB_MAX], s_bar_mma[SUB], s_bar_dfree[SUB];
    __shared__ uint32_t s_slot;
    __shared__ float s_red[NT / 32];
    const RecGeom &g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Hp = g.Hpad, KB = Hp / 64;
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(t2_smem_raw) + 1023) & ~(uintptr_t)1023);
    const bool ctrl = warp >= 16;
    const int sub = ctrl ? warp - 16 : warp / GW;                 // sub-group of this warp
    const int lw = warp % GW, ltid = tid - sub * GW * 32;         // gate-math warps: warp / thread index inside the sub-group
    uint8_t *Bt = base + (size_t)sub * KB * KBB;                  // per sub-group [KB][2 NS rows: h_hi | h_lo'][128 B]
    uint8_t *Alo = base + (size_t)SUB * KB * KBB;                 // LO_SMEM: [KB][128 rows][128 B]
    float *stage = reinterpret_cast<float *>(Alo + (LO_SMEM ? (size_t)KB * 16384 : 0));     // [128][65] prologue staging
    float *dstage = stage + 128 * 65 + (size_t)sub * 4 * NS * 32; // per sub-group [gate][sequence][cell]: the step product, transposed

    const int H = p.H, L = p.L, S = p.S, T = p.T;
    const int H4 = (H + 3) & ~3;
    const int GB = cdiv_dev(g.G, SUB);                            // CTAs per (direction, cell slice): SUB sequence groups each
    const int d = blockIdx.x / (GB * g.C);
    const int grp = ((blockIdx.x % (GB * g.C)) / g.C) * SUB + sub;
    const int cs = blockIdx.x % g.C;
    const int j0 = cs * g.CL, ncell = min(g.CL, H - j0), ncell4 = min(g.CL, H4 - j0);
    const bool active = grp < g.G;                                // an odd number of groups leaves the last CTA's second sub-group idle
    const int s0 = grp * g.SG, nseq = active ? min(g.SG, S - s0) : 0;
    // accumulator of a sub-group: [0, NS) hi*hi, [NS, 2 NS) the lo' terms
    const uint32_t col_ahi = 0, col_alo = Hp / 2, col_d = (LO_SMEM ? Hp / 2 : Hp) + sub * 2 * NS;

    for (int i = tid; i < SUB * KB * KBB / 4; i += NT) reinterpret_cast<uint32_t *>(base)[i] = 0u;      // rows beyond nseq stay zero
    if (tid < 32) s_tab[tid] = bl_exp2f_tab[tid];
    if (tid == 0) {
        for (int sb = 0; sb < SUB; ++sb) {
            for (int kb = 0; kb < T2_KB_MAX; ++kb) t2_mbar_init(&s_bar_b[sb][kb], GW / 2);
            t2_mbar_init(&s_bar_mma[sb], 1);
            t2_mbar_init(&s_bar_dfree[sb], 4);                      // the four warps that read the accumulator
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(t2_smem_u32(&s_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_slot;

    // ---- the CTA's weight slice, once: row r = gate * 32 + cell (lane of TMEM), W_gate[j, k] = Wi[gate*L*H + d*H*H + j*H + k]
    // (k = source cell, LstmLayer.cu:586-596), staged through shared memory in 64-wide k chunks so that the global reads coalesce
    const float *Wd = p.Wi + (size_t)d * H * H;
    float wmax = 0.0f;
    for (int idx = tid; idx < 4 * ncell * H; idx += NT) {
        const int gate = idx / (ncell * H), rem = idx - gate * ncell * H;
        wmax = fmaxf(wmax, fabsf(__ldg(Wd + (size_t)gate * L * H + (size_t)j0 * H + rem)));
    }
    const float wscale = t2_slice_scale(wmax, s_red, NT / 32), wunscale = __frcp_rn(wscale);
    for (int kc = 0; kc < KB; ++kc) {
        for (int idx = tid; idx < 128 * 64; idx += NT) {
            const int r = idx >> 6, kk = idx & 63, gate = r >> 5, cell = r & 31, k = kc * 64 + kk;
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
    long long *trb = p.trace ? p.trace + ((size_t)blockIdx.x * SUB + sub) * T * 8 : nullptr;      // one row per sub-group

    if (!active) {
        // idle sub-group: nothing to do until the teardown barrier
    } else if (ctrl) {
        // ================================================================ control warp of the sub-group: MMA issue, K-block by K-block
        const uint32_t idesc32 = t2_make_idesc(2 * NS), idesc16 = t2_make_idesc(NS);
        constexpr int KBD = KBB >> 4;                               // K-block stride in descriptor address units
        const uint64_t desc_b = t2_make_desc(Bt), desc_a = t2_make_desc(Alo);
        if (t2_elect_one()) {                                       // ONE thread runs the whole control loop
            for (int q = 1; q < T; ++q) {
                if (q >= 2) t2_mbar_wait_thread(&s_bar_dfree[sub], (uint32_t)(q & 1));      // every warp has read step q-1's accumulator (phase q-2)
                for (int kb = 0; kb < KB; ++kb) {
                    t2_mbar_wait_thread(&s_bar_b[sub][kb], (uint32_t)((q - 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (trb && kb == 0) trb[q * 8 + 6] = clock64();
                    // 16 k per MMA = 8 TMEM columns = 32 B inside the swizzle atom
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        t2_mma_ts(tmem + col_d, tmem + col_ahi + kb * 32 + ks * 8, desc_b + (uint64_t)(kb * KBD + ks * 2), idesc32, (kb | ks) ? 1u : 0u);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        if (!LO_SMEM) t2_mma_ts(tmem + col_d + NS, tmem + col_alo + kb * 32 + ks * 8, desc_b + (uint64_t)(kb * KBD + ks * 2), idesc16, 1u);
                        else t2_mma_ss(tmem + col_d + NS, desc_a + (uint64_t)(kb * 1024 + ks * 2), desc_b + (uint64_t)(kb * KBD + ks * 2), idesc16, 1u);
                    }
                }
                t2_commit(&s_bar_mma[sub]);
                if (trb) trb[q * 8 + 7] = clock64();
            }
        }
        __syncwarp();
    } else {
        // ================================================================ gate-math warps
        const int seq = lw, cell = lane;
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
        // poll assignment: the lower half of the warps takes the even K-blocks, the upper half the odd ones; GW * 16 threads x one
        // uint4 = NS rows x 64 k
        const int ph = lw / (GW / 2), prow = (ltid % (GW * 16)) >> 4, pf4 = ltid & 15;
        const uint32_t st_off = (uint32_t)((prow >> 3) * 1024 + (prow & 7) * 128 + (((pf4 >> 1) ^ (prow & 7)) << 4) + (pf4 & 1) * 8);
        // warps 0..3 of the sub-group (GW is a multiple of 4: lw == warp & 3, the TMEM quadrant this warp may read) fetch gate lw
        const uint32_t tm_lane = tmem + ((uint32_t)((lw & 3) * 32) << 16) + col_d;
        const int px = g.npair & 3, backoff = g.npair >> 2;          // tuning switches of the exchange poll (BLSTM_T2_POLL)
        uint4 v[T2_KB_MAX / 2];                                      // exchange words of this thread's B tile entries (polled across the loop edge)
#pragma unroll
        for (int u = 0; u < T2_KB_MAX / 2; ++u) v[u] = make_uint4(T2_FTAG, T2_FTAG, T2_FTAG, T2_FTAG);

        // results of the previous step, stored one step late: a step's HBM-only stores (4 wavefronts each: a warp covers 4 sequences)
        // would otherwise sit in the SM's load / store queue AHEAD of the other warps' exchange words and of the next poll
        float d_ni = 0, d_ig = 0, d_fg = 0, d_og = 0, d_c = 0, d_h = 0;
        bool d_acts = false;
        int d_t = 0;
        auto flush_results = [&]() {
            if (!valid) return;
            float *ap = p.acts + ((size_t)d_t * S + slot) * 4 * L + d * H + j0 + cell;
            if (d_acts) { ap[0] = d_ni; ap[L] = d_ig; ap[2 * L] = d_fg; ap[3 * L] = d_og; }
            p.cst[((size_t)d_t * S + slot) * L + d * H + j0 + cell] = d_c;
            p.Y[((size_t)d_t * S + slot) * p.ldy + d * H + j0 + cell] = d_h;
            if (p.ys_hi) {                                            // TF32 split of the layer output for the backward GEMMs
                const size_t idx = ((size_t)d_t * S + slot) * p.ld_ys + d * H4 + j0 + cell;
                uint32_t hh; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hh) : "f"(d_h));
                p.ys_hi[idx] = __uint_as_float(hh);
                p.ys_lo[idx] = __fsub_rn(d_h, __uint_as_float(hh));
            }
        };

        for (int q = 0; q < T; ++q) {
            const int t = (d == 0) ? q : T - 1 - q;
            const bool first = (q == 0);
            const bool check = (t >= p.Tmin);
            long long *tr = (trb && ltid == 0) ? trb + q * 8 : nullptr;
            if (tr) tr[0] = clock64();
            float a[4] = {0, 0, 0, 0};
            bool dummy = false;
            const float *acts_p = p.acts + ((size_t)t * S + slot) * 4 * L + d * H + j0 + cell;
            auto prefetch = [&]() {
                if (valid) {
                    dummy = check && (p.pat[(size_t)t * S + slot] == BL_PATTYPE_NONE);
#pragma unroll
                    for (int gi = 0; gi < 4; ++gi) a[gi] = acts_p[gi * L];
                }
            };
            if (first) prefetch();
            float rec[4] = {0, 0, 0, 0};
            if (!first) {
                // ---- previous-step outputs of this group's sequences: the first poll of every word was issued right after this warp's
                // own exchange store of the previous step (end of the loop body); re-poll ALL words that still carry the old tag at
                // once until none does, then hand the B tile over K-block by K-block
                const unsigned *xs = xd + (size_t)((q - 1) & 1) * xbuf + (size_t)(s0 + prow) * Hp + pf4 * 4;
                const unsigned want = (((q - 1) >> 1) & 1) ? T2_FTAG : 0u;        // tag of step q-1
                if (!(px & 1)) {                                     // tuning: first poll here instead of at the end of the previous step
                    if (g.RP > 0) __nanosleep(g.RP);                  // tuning: give the other producers' words time to land before the first poll
#pragma unroll
                    for (int u = 0; u < T2_KB_MAX / 2; ++u) {
                        const int kb = ph + 2 * u;
                        if (kb < KB && prow < nseq && kb * 64 + pf4 * 4 < H4) v[u] = t2_ld_relaxed_v4(xs + kb * 64);
                    }
                }
                for (int spin = 0;; ++spin) {
                    bool pend = false;
#pragma unroll
                    for (int u = 0; u < T2_KB_MAX / 2; ++u) {
                        const int kb = ph + 2 * u;
                        if (kb < KB && prow < nseq && kb * 64 + pf4 * 4 < H4 && t2_stale(v[u], want)) {
                            if (pend && !(px & 2)) continue;         // tuning: one outstanding re-poll per thread instead of all
                            if (backoff) __nanosleep(backoff);
                            v[u] = t2_ld_relaxed_v4(xs + kb * 64); pend = true;
                        }
                    }
                    if (!pend) break;
                    if (spin > T2_SPIN) __trap();
                }
#pragma unroll
                for (int u = 0; u < T2_KB_MAX / 2; ++u) {
                    const int kb = ph + 2 * u;
                    if (kb < KB) {
                        if (prow < nseq && kb * 64 + pf4 * 4 < H4) {
                            // word = hi | lo' << 16 (tag in bit 30): four hi halves into row prow, four lo' halves into row NS + prow
                            uint8_t *dst = Bt + (size_t)kb * KBB + st_off;
                            *reinterpret_cast<uint2 *>(dst) = make_uint2(__byte_perm(v[u].x, v[u].y, 0x5410), __byte_perm(v[u].z, v[u].w, 0x5410));
                            *reinterpret_cast<uint2 *>(dst + NS * 128) = make_uint2(__byte_perm(v[u].x, v[u].y, 0x7632) & 0xBFFFBFFFu,
                                                                                __byte_perm(v[u].z, v[u].w, 0x7632) & 0xBFFFBFFFu);
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) t2_mbar_arrive(&s_bar_b[sub][kb]);
                    }
                }
                if (tr) tr[1] = clock64();
                // ---- while the tensor pipe works: the previous step's results out, this step's pre-activations in
                flush_results();
                prefetch();
                // ---- the step product out of tensor memory, transposed through shared memory to (warp = sequence, lane = cell): warps
                // 0..3 of the sub-group wait for the tensor pipe and stage gate lw; everybody meets at a named barrier
                if (lw < 4) {
                    t2_mbar_wait(&s_bar_mma[sub], (uint32_t)((q - 1) & 1));
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (tr) tr[2] = clock64();
                    // gate lw of all 32 cells x NS sequences: columns [0, NS) hi*hi, [NS, 2 NS) the lo' terms
                    uint32_t dh[NS], dl[NS];
                    if (NS == 16) { t2_ld16(tm_lane, *reinterpret_cast<uint32_t (*)[16]>(dh)); t2_ld16(tm_lane + NS, *reinterpret_cast<uint32_t (*)[16]>(dl)); }
                    else { t2_ld8(tm_lane, *reinterpret_cast<uint32_t (*)[8]>(dh)); t2_ld8(tm_lane + NS, *reinterpret_cast<uint32_t (*)[8]>(dl)); }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) t2_mbar_arrive(&s_bar_dfree[sub]);
#pragma unroll
                    for (int i = 0; i < NS; ++i)
                        dstage[(lw * NS + i) * 32 + lane] = __fmul_rn(__fmaf_rn(__uint_as_float(dl[i]), T2_LO_UNSCALE, __uint_as_float(dh[i])), wunscale);
                }
                t2_named_barrier(1 + sub, GW * 32);
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) rec[gi] = dstage[(gi * NS + seq) * 32 + cell];
            }
            if (tr) tr[3] = clock64();

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
                act3_tab(ni, ig, fg, s_tab, ni, ig, fg);        // the three first-level activations, interleaved
                c = __fmul_rn(ni, ig);                                // :121-126
                if (!first) c = __fadd_rn(c, __fmul_rn(cprev, fg));
                og = __fadd_rn(og, __fmul_rn(c, wpe[2]));             // :129-131
                float tc;
                act2_tab(c, og, s_tab, tc, og);
                h = __fmul_rn(tc, og);             // :134
            }
            cprev = c;
            if (tr) tr[4] = clock64();
            // ---- the exchange word first: h split once here, (fp16 hi | fp16 lo' << 16), step tag in the free bit 14 of lo'
            if (q + 1 < T && xvalid) {
                uint32_t hh, hl;
                t2_split(h, hh, hl);
                xd[(size_t)(q & 1) * xbuf + (size_t)slot * Hp + j0 + cell] = hh | ((hl & 0xBFFFu) << 16) | (((q >> 1) & 1) ? T2_FTAG : 0u);
            }
            // ---- first poll of the next step's operand, ahead of this step's result stores in the load / store queue
            if (q + 1 < T && (px & 1)) {
                const unsigned *xn = xd + (size_t)(q & 1) * xbuf + (size_t)(s0 + prow) * Hp + pf4 * 4;
#pragma unroll
                for (int u = 0; u < T2_KB_MAX / 2; ++u) {
                    const int kb = ph + 2 * u;
                    if (kb < KB && prow < nseq && kb * 64 + pf4 * 4 < H4) v[u] = t2_ld_relaxed_v4(xn + kb * 64);
                }
            }
            d_ni = ni; d_ig = ig; d_fg = fg; d_og = og; d_c = c; d_h = h; d_acts = valid && !dummy; d_t = t;
            if (tr) tr[5] = clock64();
        }
        flush_results();                                              // the last step's
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 16) {
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
template <int SUB, bool LO_SMEM, int CMAX, int MTMAX>
__global__ void __maxnreg__(96) lstm_bwd_tm2_kernel(const RecBwdParams p)
{
    constexpr int GW = 16 / SUB, NT = (16 + SUB) * 32, NS = GW;
    constexpr int KBB = 2 * NS * 128;                             // bytes of one K-block of a B tile: rows [0, NS) d_hi, [NS, 2 NS) d_lo'
    extern __shared__ uint8_t t2_smem_raw[];
    __shared__ unsigned long long s_tab[32];
    __shared__ uint64_t s_bar_b[SUB], s_bar_mma[SUB];
    __shared__ uint32_t s_slot;
    __shared__ float s_red[NT / 32];
    const RecGeom &g = p.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = g.Hpad, MT = R / 128;
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(t2_smem_raw) + 1023) & ~(uintptr_t)1023);
    const bool ctrl = warp >= 16;
    const int sub = ctrl ? warp - 16 : warp / GW;
    const int lw = warp % GW, ltid = tid - sub * GW * 32;
    uint8_t *Bt = base + (size_t)sub * 2 * KBB;                   // per sub-group [2 K-blocks][2 NS rows: d_hi | d_lo'][128 B], k = gate*32 + cell
    uint8_t *Alo = base + (size_t)SUB * 2 * KBB;                  // LO_SMEM: [MT][2][128 rows][128 B]
    float *stage = reinterpret_cast<float *>(Alo + (LO_SMEM ? (size_t)MT * 2 * 16384 : 0));      // [64 kk][128 k'] prologue staging

    const int H = p.H, L = p.L, S = p.S, T = p.T;
    const int H4 = (H + 3) & ~3;
    const int GB = cdiv_dev(g.G, SUB);
    const int d = blockIdx.x / (GB * g.C);
    const int grp = ((blockIdx.x % (GB * g.C)) / g.C) * SUB + sub;
    const int cs = blockIdx.x % g.C;
    const int j0 = cs * g.CL, ncell = min(g.CL, H - j0);
    const bool active = grp < g.G;
    const int s0 = grp * g.SG, nseq = active ? min(g.SG, S - s0) : 0;
    const bool inplace = (p.ndir == 1);
    // accumulator of tile mt of a sub-group at col_d + mt * 2 NS
    const uint32_t col_ahi = 0, col_alo = MT * 64, col_d = (LO_SMEM ? MT * 64 : MT * 128) + sub * MT * 2 * NS;

    for (int i = tid; i < SUB * 2 * KBB / 4; i += NT) reinterpret_cast<uint32_t *>(base)[i] = 0u;       // never-written entries stay zero
    if (tid < 32) s_tab[tid] = bl_exp2f_tab[tid];
    if (tid == 0) {
        for (int sb = 0; sb < SUB; ++sb) { t2_mbar_init(&s_bar_b[sb], GW); t2_mbar_init(&s_bar_mma[sb], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 16) {
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
    for (int idx = tid; idx < 4 * ncell * H; idx += NT) {
        const int gate = idx / (ncell * H), rem = idx - gate * ncell * H;
        wmax = fmaxf(wmax, fabsf(__ldg(Wd + (size_t)gate * L * H + (size_t)j0 * H + rem)));
    }
    const float wscale = t2_slice_scale(wmax, s_red, NT / 32), wunscale = __fmul_rn(__frcp_rn(wscale), 1.0f / T2_DELTA_SCALE);
    for (int mt = 0; mt < MT; ++mt)
        for (int half = 0; half < 2; ++half) {
            for (int idx = tid; idx < 64 * 128; idx += NT) {
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

    const size_t ex_slice = (size_t)NS * R;                         // one producer's [NS][R] block
    const size_t ex_par = (size_t)g.G * g.C * ex_slice;             // one parity of one direction
    unsigned *exd = reinterpret_cast<unsigned *>(p.dx) + (size_t)d * 2 * ex_par + (size_t)grp * g.C * ex_slice;      // this group's C blocks, parity 0
    long long *trb = p.trace ? p.trace + ((size_t)blockIdx.x * SUB + sub) * T * 8 : nullptr;

    if (!active) {
        // idle sub-group
    } else if (ctrl) {
        // ================================================================ control warp of the sub-group
        const uint32_t idesc32 = t2_make_idesc(2 * NS), idesc16 = t2_make_idesc(NS);
        constexpr int KBD = KBB >> 4;
        const uint64_t desc_b = t2_make_desc(Bt), desc_a = t2_make_desc(Alo);
        if (t2_elect_one()) {                                       // ONE thread runs the whole control loop
            for (int q = 0; q + 1 < T; ++q) {
                t2_mbar_wait_thread(&s_bar_b[sub], (uint32_t)(q & 1));  // all deltas of step q are in the B tile; step q-1's accumulators were read
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (trb) trb[q * 8 + 6] = clock64();
                for (int mt = 0; mt < MT; ++mt) {
                    const uint32_t dcol = tmem + col_d + mt * 2 * NS;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks)                  // K = 128 = 8 MMAs of 16
                        t2_mma_ts(dcol, tmem + col_ahi + mt * 64 + ks * 8, desc_b + (uint64_t)((ks >> 2) * KBD + (ks & 3) * 2), idesc32, ks ? 1u : 0u);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        if (!LO_SMEM) t2_mma_ts(dcol + NS, tmem + col_alo + mt * 64 + ks * 8, desc_b + (uint64_t)((ks >> 2) * KBD + (ks & 3) * 2), idesc16, 1u);
                        else t2_mma_ss(dcol + NS, desc_a + (uint64_t)((mt * 2 + (ks >> 2)) * 1024 + (ks & 3) * 2), desc_b + (uint64_t)((ks >> 2) * KBD + (ks & 3) * 2), idesc16, 1u);
                    }
                }
                t2_commit(&s_bar_mma[sub]);
                if (trb) trb[q * 8 + 7] = clock64();
            }
        }
        __syncwarp();
    } else {
        // ================================================================ gate-math warps: warp = sequence, lane = cell
        const int seq = lw, cell = lane;
        const bool valid = cell < ncell && seq < nseq;
        const int slot = s0 + seq;
        float wpe[3] = {0, 0, 0};
        float nfg = 0.0f, ncerr = 0.0f, ndig = 0.0f, ndfg = 0.0f;  // "next step" state, :253-256
        float gacc[7] = {0, 0, 0, 0, 0, 0, 0};                      // bias (4 gates) and peephole (ig, fg, og) gradient sums of this (cell, sequence)
        if (valid) {
            const int col = d * H + j0 + cell;
#pragma unroll
            for (int gi = 0; gi < 3; ++gi) wpe[gi] = __ldg(p.Wp + gi * L + col);
        }
        // epilogue: TMEM quadrant qd (== warp & 3), sequence octet soct, first tile tile0 of this warp: the sub-group's warps cover two
        // tiles at a time with x8 loads (the tensor-memory port moves 64 B per cycle plus ~8 cycles per tcgen05.ld: 32 loads per step
        // here against 64 with x4 loads)
        constexpr int NOCT = NS / 8;
        const int qd = lw & 3, soct = (lw >> 2) % NOCT, tile0 = lw / (4 * NOCT);
        const uint32_t tm_lane = tmem + ((uint32_t)(qd * 32) << 16) + col_d + 8 * soct;

        for (int q = 0; q < T; ++q) {
            const int t = (d == 0) ? T - 1 - q : q;                 // fw walks time backwards, bw forwards (:936, :970)
            const bool firstCall = (q == 0), lastCall = (q == T - 1);
            const bool check = (t >= p.Tmin);
            const int tprev = (d == 0) ? t - 1 : t + 1;
            long long *tr = (trb && ltid == 0) ? trb + q * 8 : nullptr;
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
                for (int spin = 0;; ++spin) {                        // re-poll ALL partials that still carry the old tag at once
                    bool pend = false;
#pragma unroll
                    for (int pp = 0; pp < CMAX; ++pp)
                        if (pp < g.C && (pv[pp] & 1u) != want) {
                            if (pend && !(g.npair & 2)) continue;
                            if (g.npair >> 2) __nanosleep(g.npair >> 2);
                            pv[pp] = t2_ld_relaxed(ex_r + (size_t)pp * ex_slice); pend = true;
                        }
                    if (!pend) break;
                    if (spin > T2_SPIN) __trap();
                }
                float s = 0.0f;
#pragma unroll
                for (int pp = 0; pp < CMAX; ++pp)
                    if (pp < g.C) s = __fadd_rn(s, __uint_as_float(pv[pp] & ~1u));
                e = __fadd_rn(oe, s);                                // the 4 addProducts of :939-942
            }
            if (tr) tr[1] = clock64();

            float dni = 0, dig = 0, dfg = 0, dog = 0, cerr = 0;
            if (valid) {
                if (dummy) {                                         // :224-234
                    nfg = 0.0f;
                } else {
                    const float ni = a[0], ig = a[1], fg = a[2], og = a[3];
                    const float tc = tanh1_tab(c, s_tab);
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
                        // |delta| <= 1 (limitedError): 2^13 keeps the fp16 halves of small deltas (deep layers: 1e-7) in the normal range;
                        // unscaled, fp16's 5-bit exponent would put an ABSOLUTE floor of 1.5e-11 under every delta
                        uint32_t hi, lo;
                        t2_split(__fmul_rn(dv[gi], T2_DELTA_SCALE), hi, lo);
                        uint8_t *dst = Bt + (gi >> 1) * KBB + t2_off(seq, (gi & 1) * 32 + cell);
                        *reinterpret_cast<unsigned short *>(dst) = (unsigned short)hi;
                        *reinterpret_cast<unsigned short *>(dst + NS * 128) = (unsigned short)lo;      // row + NS
                    }
                }
            }
            if (!lastCall) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) t2_mbar_arrive(&s_bar_b[sub]);
            }
            if (tr) tr[2] = clock64();
            // ---- HBM-only results while the tensor pipe works
            if (valid) {
                // the non-GEMM part of ComputeWeightUpdateFn (:392-408, 440-475): bias sums, peephole sums (ig / fg pair the delta with
                // the cell state of the previous timestep of the direction, og with this timestep's)
                gacc[0] = __fadd_rn(gacc[0], dni); gacc[1] = __fadd_rn(gacc[1], dig); gacc[2] = __fadd_rn(gacc[2], dfg); gacc[3] = __fadd_rn(gacc[3], dog);
                if (!lastCall) { gacc[4] = __fmaf_rn(cp, dig, gacc[4]); gacc[5] = __fmaf_rn(cp, dfg, gacc[5]); }
                gacc[6] = __fmaf_rn(c, dog, gacc[6]);
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
                // ---- partial products out of tensor memory into the exchange buffer: lane = source cell, 4 sequences per warp.  Four warps
                // wait for the tensor pipe, the rest joins them at a named barrier (an mbarrier wakes 16 warps in 300-650 cycles)
                if (lw < 4) t2_mbar_wait(&s_bar_mma[sub], (uint32_t)(q & 1));
                t2_named_barrier(1 + sub, GW * 32);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tr) tr[3] = clock64();
                unsigned *ex_w = exd + (size_t)(q & 1) * ex_par + (size_t)cs * ex_slice;
                const unsigned tag = (q >> 1) & 1;
#pragma unroll
                for (int k = 0; k < (MTMAX + 1) / 2; ++k) {
                    const int mt = tile0 + 2 * k;
                    if (mt < MT) {
                        uint32_t r0[8], r1[8];
                        t2_ld8(tm_lane + mt * 2 * NS, r0);
                        t2_ld8(tm_lane + mt * 2 * NS + NS, r1);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        const int ksrc = mt * 128 + qd * 32 + lane;
                        if (ksrc < H4) {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                if (8 * soct + i < nseq) {
                                    const float o = __fmul_rn(__fmaf_rn(__uint_as_float(r1[i]), T2_LO_UNSCALE, __uint_as_float(r0[i])), wunscale);
                                    // last mantissa bit rounded away (nearest even; inf / NaN keep their class), then the tag
                                    const unsigned bb = __float_as_uint(o);
                                    const unsigned rr = ((bb & 0x7F800000u) == 0x7F800000u) ? (bb & ~1u) : ((bb + ((bb >> 1) & 1u)) & ~1u);
                                    ex_w[(size_t)(8 * soct + i) * R + ksrc] = rr | tag;
                                }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (tr) tr[4] = clock64();
            }
        }
        if (p.gpart) {
            // sum the sub-group's sequences in order (deterministic) through the prologue's staging area, one block per sequence group
            float *red = stage + (size_t)sub * 7 * NS * 32;
#pragma unroll
            for (int i = 0; i < 7; ++i) red[(i * NS + seq) * 32 + cell] = valid ? gacc[i] : 0.0f;
            t2_named_barrier(1 + sub, GW * 32);
            if (lw < 7 && cell < ncell) {
                float sum = 0.0f;
                for (int n = 0; n < nseq; ++n) sum = __fadd_rn(sum, red[(lw * NS + n) * 32 + cell]);
                p.gpart[((size_t)grp * 7 + lw) * L + d * H + j0 + cell] = sum;
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 16) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ launch
template <typename Params, typename Kernel>
static int launch_tm2(bl_ctx *ctx, Kernel kernel, const Params &p, float *xbuf, int init_byte, const char *name)
{
    const RecGeom &g = p.g;
    const int grid = p.ndir * cdiv(g.G, g.nsub) * g.C;
    BL_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    int per_sm = 0;
    BL_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, g.NT, g.smem));
    if (per_sm != 1 || grid > ctx->num_sms) {
        cudaFuncAttributes fa{};
        cudaFuncGetAttributes(&fa, kernel);
        return fail(ctx, "%s: %d CTAs cannot be co-resident one per SM (occupancy %d per SM x %d SMs; %d threads, %d registers, %zu + %zu B shared memory)",
                    name, grid, per_sm, ctx->num_sms, g.NT, fa.numRegs, g.smem, fa.sharedSizeBytes);
    }
    // every word of both parities starts with step tag 1: the first two steps of a pass carry tag 0
    BL_CUDA(ctx, cudaMemsetAsync(xbuf, init_byte, g.xelems * sizeof(float), ctx->stream));
    void *args[] = { (void *)&p };
    BL_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(g.NT), args, g.smem, ctx->stream));
    BL_LAUNCHED(ctx);
    return 0;
}

int launch_lstm_fwd_tm2(bl_ctx *ctx, const RecFwdParams &p)
{
    TimedRegion timed(ctx, 1);
    if (p.g.nsub == 2)
        return p.g.K4 ? launch_tm2(ctx, lstm_fwd_tm2_kernel<2, true>, p, p.hx, T2_FWD_INIT, "lstm_fwd_tm2")
                      : launch_tm2(ctx, lstm_fwd_tm2_kernel<2, false>, p, p.hx, T2_FWD_INIT, "lstm_fwd_tm2");
    return p.g.K4 ? launch_tm2(ctx, lstm_fwd_tm2_kernel<1, true>, p, p.hx, T2_FWD_INIT, "lstm_fwd_tm2")
                  : launch_tm2(ctx, lstm_fwd_tm2_kernel<1, false>, p, p.hx, T2_FWD_INIT, "lstm_fwd_tm2");
}

template <int SUB>
static int launch_bwd_tm2_sub(bl_ctx *ctx, const RecBwdParams &p)
{
    if (p.g.K4) return launch_tm2(ctx, lstm_bwd_tm2_kernel<SUB, true, T2_CMAX, T2_MT_MAX>, p, p.dx, T2_BWD_INIT, "lstm_bwd_tm2");
    if (p.g.C <= 8 && p.g.Hpad <= 256) return launch_tm2(ctx, lstm_bwd_tm2_kernel<SUB, false, 8, 2>, p, p.dx, T2_BWD_INIT, "lstm_bwd_tm2");
    return launch_tm2(ctx, lstm_bwd_tm2_kernel<SUB, false, T2_CMAX, 3>, p, p.dx, T2_BWD_INIT, "lstm_bwd_tm2");
}

int launch_lstm_bwd_tm2(bl_ctx *ctx, const RecBwdParams &p)
{
    TimedRegion timed(ctx, 2);
    return p.g.nsub == 2 ? launch_bwd_tm2_sub<2>(ctx, p) : launch_bwd_tm2_sub<1>(ctx, p);
}

} // namespace bl
