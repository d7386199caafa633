// tcgen05 / TMEM / TMA GEMM for the time-parallel contractions of the LSTM layer (north_star item 1):
// input-to-gate projection, input-error and weight-gradient GEMMs (reference: the cublasSgemm calls behind
// helpers::Matrix, layers/LstmLayer.cu:774-784, 996-1006, 1038-1043; layers/FeedForwardLayer.cu:152, 196, 206).
//
// Kernel: C[M x N] (row-major, ldc) = A[M x K] * B[N x K]^T on fp32 operands in global memory.  Either operand may be
// K-major ([MN][K], K contiguous) or MN-major ([K][MN], MN contiguous -- the "transposed" view of the same row-major
// matrix): tcgen05 reads both from 128B-swizzled shared tiles, so no operand is ever transposed in memory.
//   warp 0      TMA producer: cp.async.bulk.tensor 2D boxes (32 floats = 128 B along the contiguous dimension) into
//               128B-swizzled shared tiles; K-major: one [BM x 32] box, MN-major: BM/32 boxes of [32 k x 32 mn]
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8 per instruction),
//               accumulating in TMEM (fp32); tcgen05.commit releases shared stages and signals the epilogue
//   warp 2      TMEM allocation / deallocation
//   warps 4-7   epilogue: tcgen05.ld (32 lanes x 32 columns per call) -> registers -> global (optionally += C)
// Precision modes (include/blstm_b200.h):
//   BL_GEMM_FAST    one TF32 MMA per k-step (10-bit mantissa products, fp32 accumulate)          -> <= 2e-3 class
//   BL_GEMM_STRICT  error-compensated 3xTF32: every operand is split in global memory into hi = tf32(x) and
//                   lo = x - hi (exact), and the kernel accumulates hi*hi + hi*lo + lo*hi in TMEM       -> fp32 class
// The only preparation pass is elementwise (split + re-pitch to 16-byte aligned rows); see tc_prepare.
// Split-K (grid.z) with ordered partial-sum reduction keeps the result deterministic.
#include "common.cuh"
#include "gemm_tc.cuh"
#include <cuda.h>
#include <cstdint>
#include <algorithm>
#include <cstdlib>

namespace bl {

constexpr int TC_BM = 128, TC_BK = 32, TC_UMMA_K = 8, TC_THREADS = 256;
constexpr int EPI_LD = 36;                     // row pitch (floats) of the epilogue's 32 x 32 transpose tiles: 16-byte aligned rows, conflict-free float4 access

struct GemmTcParams {
    CUtensorMap tmA, tmAlo, tmB, tmBlo;     // lo maps unused in fast mode
    float *C; int ldc;
    int M, N, K;
    int a_mn, b_mn;                          // 1: operand is MN-major (tensor map dims {MN, K}), 0: K-major (dims {K, MN})
    int batches, mt_per_batch;               // grid.y = batches * mt_per_batch: batch b multiplies A rows [b*a_batch_rows, +M) by the same B
    int a_batch_rows; long long c_batch_stride;   // and writes its [M x N] block at C + b*c_batch_stride
    int kblocks_per_split;
    int tiles_x, tiles_y, tiles_z;           // persistent 2-CTA kernel: tile grid (n tiles, batches * m tiles, k splits)
    int accumulate;
    float *partial; int ldp;                 // split-K: slice z at partial + z*M*ldp
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// K-major, 128B-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start address >> 4 in [0,14), LBO (unused for one swizzle atom along K) in [16,30), SBO = 1024 B (8 rows x 128 B) >> 4 in [32,46),
// version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64)
//
// MN-major TF32 operands have ONE legal shared layout, SWIZZLE_128B_BASE32B (layout type 1; cutlass sm100_common.inl:92): 32-byte
// chunks swizzled inside 128 B rows, repeating every 4 rows -- TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  Canonical layout in floats
// ((4,8,m),(4,k)):((1,4,LBO),(32,SBO)) (mma_traits_sm100.hpp:72-75,175): an atom is 32 floats along MN x 4 k-rows of 128 B;
// SBO = distance between successive groups of 4 k-rows (512 B: the rows of a box are consecutive), LBO = distance between successive
// 32-float MN chunks (one [32 k x 128 B] box = 4096 B).
__device__ __forceinline__ uint64_t make_smem_desc(const void *smem_ptr, bool mn_major)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem_ptr) & 0x3FFFF) >> 4);
    if (mn_major) {
        d |= (uint64_t)((TC_BK * 128) >> 4) << 16;
        d |= (uint64_t)(512 >> 4) << 32;
        d |= (uint64_t)1 << 46;
        d |= (uint64_t)1 << 61;
    } else {
        d |= (uint64_t)(1024 >> 4) << 32;
        d |= (uint64_t)1 << 46;
        d |= (uint64_t)2 << 61;
    }
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 [4,6)=1, A/B tf32 [7,10)=[10,13)=2, A / B MN-major at bits 15 / 16,
// N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn, int b_mn)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar)
{ asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------------ the GEMM kernel
template <int BN, bool STRICT, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tf32_tcgen05_kernel(const __grid_constant__ GemmTcParams p)
{
    constexpr int A_BYTES = TC_BM * TC_BK * 4, B_BYTES = BN * TC_BK * 4;
    constexpr int STAGE_BYTES = (STRICT ? 2 : 1) * (A_BYTES + B_BYTES);
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B needs 1024 B alignment
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *tmem_full = empty + STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int batch = blockIdx.y / p.mt_per_batch;
    const int n0 = blockIdx.x * BN, m0 = (blockIdx.y - batch * p.mt_per_batch) * TC_BM;      // m0: row inside the batch's block
    const int a_row0 = batch * p.a_batch_rows + m0;                                            // row in the A tensor map
    const int kb_total = (p.K + TC_BK - 1) / TC_BK;
    const int kb_begin = blockIdx.z * p.kblocks_per_split;
    const int kb_end = min(kb_total, kb_begin + p.kblocks_per_split);
    const int nkb = kb_end - kb_begin;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&p.tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&p.tmB) : "memory");
        if (STRICT) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(&p.tmAlo) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(&p.tmBlo) : "memory");
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // TMEM: BN fp32 accumulator columns (power of two >= 32)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: lane 0 owns the barriers, lanes 0..(loads-1) each issue one box =====
        // load list of a stage: A hi chunks | B hi chunks | A lo chunks | B lo chunks  (K-major operand: 1 chunk, MN-major: tile/32 chunks)
        const int nca = p.a_mn ? TC_BM / 32 : 1, ncb = p.b_mn ? BN / 32 : 1;
        const int nload = (STRICT ? 2 : 1) * (nca + ncb);
        int l = lane;
        const bool is_lo = l >= nca + ncb; if (is_lo) l -= nca + ncb;
        const bool is_b = l >= nca; const int chunk = is_b ? l - nca : l;
        const CUtensorMap *map = is_b ? (is_lo ? &p.tmBlo : &p.tmB) : (is_lo ? &p.tmAlo : &p.tmA);
        const int mn_major = is_b ? p.b_mn : p.a_mn;
        const int mn = (is_b ? n0 : a_row0) + (mn_major ? chunk * 32 : 0);
        const int dst_off = (is_lo ? A_BYTES + B_BYTES : 0) + (is_b ? A_BYTES : 0) + (mn_major ? chunk * TC_BK * 128 : 0);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % STAGES, round = i / STAGES;
            if (lane == 0) {
                mbar_wait(&empty[s], (round & 1) ^ 1);                 // first pass over the ring succeeds immediately
                mbar_expect_tx(&full[s], STAGE_BYTES);
            }
            __syncwarp();
            if (lane < nload) {
                const int kc = (kb_begin + i) * TC_BK;
                uint8_t *dst = smem + s * STAGE_BYTES + dst_off;
                if (mn_major) tma_load_2d(dst, map, &full[s], mn, kc);
                else          tma_load_2d(dst, map, &full[s], kc, mn);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        const uint32_t idesc = make_idesc(TC_BM, BN, p.a_mn, p.b_mn);
        const uint64_t a_step = (uint64_t)(((p.a_mn ? 1024 : TC_UMMA_K * 4)) >> 4), b_step = (uint64_t)(((p.b_mn ? 1024 : TC_UMMA_K * 4)) >> 4);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % STAGES, round = i / STAGES;
            mbar_wait(&full[s], round & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint8_t *st = smem + s * STAGE_BYTES;
                const uint64_t a_hi = make_smem_desc(st, p.a_mn), b_hi = make_smem_desc(st + A_BYTES, p.b_mn);
                const uint64_t a_lo = make_smem_desc(st + A_BYTES + B_BYTES, p.a_mn), b_lo = make_smem_desc(st + 2 * A_BYTES + B_BYTES, p.b_mn);
#pragma unroll
                for (int kk = 0; kk < TC_BK / TC_UMMA_K; ++kk) {
                    // next 8 k: 32 B further inside the swizzle atom (K-major) or the next 8-row atom, 1024 B (MN-major)
                    const uint64_t ad = kk * a_step, bd = kk * b_step;
                    const uint32_t acc0 = (i > 0 || kk > 0) ? 1u : 0u;
                    if (STRICT) {
                        // small terms first, then the leading term
                        umma_tf32(tmem_base, a_lo + ad, b_hi + bd, idesc, acc0);
                        umma_tf32(tmem_base, a_hi + ad, b_lo + bd, idesc, 1u);
                        umma_tf32(tmem_base, a_hi + ad, b_hi + bd, idesc, 1u);
                    } else {
                        umma_tf32(tmem_base, a_hi + ad, b_hi + bd, idesc, acc0);
                    }
                }
                umma_commit(&empty[s]);                                   // frees the stage once these MMAs have read it
                if (i == nkb - 1) umma_commit(tmem_full);                 // accumulator complete
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> global =====
        const int ew = warp & 3;                                          // TMEM lane quarter this warp may access
        const int row = m0 + ew * 32 + lane;
        if (nkb > 0) {
            mbar_wait(tmem_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        float *out = p.partial ? p.partial + ((size_t)blockIdx.z * p.batches + batch) * p.M * p.ldp : p.C + (size_t)batch * p.c_batch_stride;
        const int ldo = p.partial ? p.ldp : p.ldc;
        const bool acc = (!p.partial) && p.accumulate;
        const bool vec_ok = ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            float v[32];
            if (nkb > 0) {
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)c, v);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.0f;
            }
            if (row < p.M) {
                float *dst = out + (size_t)row * ldo + n0 + c;
                if (vec_ok && n0 + c + 32 <= p.N) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        if (acc) { const float4 old = *reinterpret_cast<const float4 *>(dst + i); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                        *reinterpret_cast<float4 *>(dst + i) = o;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (n0 + c + i < p.N) dst[i] = acc ? dst[i] + v[i] : v[i];
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(BN) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ the 2-CTA (cta_group::2) kernel
// A TF32 MMA of 128 x 128 x 8 reads 8 KB of shared memory in 64 tensor-core cycles = 128 B/clk, the whole shared-memory
// bandwidth of an SM, and the 3xTF32 scheme reads every operand tile three times: the single-CTA kernel above is bound by
// shared memory (and by L2 -> SM traffic), not by the tensor pipe.  Here two CTAs of a cluster (two SMs) share one
// 256 x 256 output tile: each holds its own 128 rows of A and HALF of the B tile (128 of the 256 columns), the leader CTA
// issues tcgen05.mma.cta_group::2 (M = 256, N = 256) and each SM accumulates its 128 x 256 half in its own TMEM.  Per SM and
// per MMA that is the same 8 KB of shared-memory reads for twice the math, and half the L2 traffic per flop.
//   both CTAs   warp 0 TMA producer (own A rows, own half of B; all complete_tx go to the LEADER's full barrier),
//               warp 2 TMEM alloc/dealloc (cta_group::2), warps 4-7 epilogue of the CTA's own 128 rows
//   leader      warp 1 issues the MMAs; tcgen05.commit ... multicast::cluster releases the stage in both CTAs
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank)
{ uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank)); return r; }

__device__ __forceinline__ void tma_load_2d_2cta(void *smem_dst, const CUtensorMap *map, uint32_t leader_bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :: "r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}

__device__ __forceinline__ void umma_tf32_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit_2cta(uint64_t *bar)       // arrives on the barrier at this offset in BOTH CTAs
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t cta)        // arrive on the same-offset barrier of CTA `cta` of the cluster
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(map_to_cta(smem_u32(bar), cta)) : "memory");
}

// Persistent: one CTA pair per SM pair walks the output tiles (tile = pair + i * pairs).  TMEM holds TWO 256-column accumulators,
// so the epilogue of tile i (TMEM -> registers -> shared transpose -> global) overlaps the main loop of tile i+1; the TMA/MMA
// ring keeps running across tile boundaries.  acc_full[b] (MMA -> both epilogues, multicast commit) and acc_empty[b] (the 8
// epilogue warps of the pair -> the leader's MMA warp) hand the accumulators back and forth.
template <bool STRICT, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tf32_tcgen05_2cta_kernel(const __grid_constant__ GemmTcParams p)
{
    constexpr int BNH = 128, BN2 = 256;                                   // per-CTA half of B, full tile width
    constexpr int A_BYTES = TC_BM * TC_BK * 4, B_BYTES = BNH * TC_BK * 4;
    constexpr int STAGE_BYTES = (STRICT ? 2 : 1) * (A_BYTES + B_BYTES);
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES);
    uint64_t *empty = full + STAGES;
    uint64_t *acc_full = empty + STAGES;                                  // [2]
    uint64_t *acc_empty = acc_full + 2;                                   // [2], used in the leader CTA
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int kb_total = (p.K + TC_BK - 1) / TC_BK;
    const int tiles_xy = p.tiles_x * p.tiles_y, ntiles = tiles_xy * p.tiles_z;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" :: "l"(&p.tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" :: "l"(&p.tmB) : "memory");
        if (STRICT) {
            asm volatile("prefetch.tensormap [%0];" :: "l"(&p.tmAlo) : "memory");
            asm volatile("prefetch.tensormap [%0];" :: "l"(&p.tmBlo) : "memory");
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(2 * BN2) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                                   // the peer's barriers are initialised before anything signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // tile -> (n tile, (batch, m tile), k split), the same enumeration the one-tile-per-cluster launch used as its grid
    auto tile_coords = [&](int tile, int &n0, int &m0, int &batch, int &z, int &kb_begin, int &nkb) {
        z = tile / tiles_xy;
        const int r = tile - z * tiles_xy, y = r / p.tiles_x, x = r - y * p.tiles_x;
        batch = y / p.mt_per_batch;
        n0 = x * BN2;
        m0 = (y - batch * p.mt_per_batch) * 2 * TC_BM + (int)rank * TC_BM;
        kb_begin = z * p.kblocks_per_split;
        nkb = min(kb_total, kb_begin + p.kblocks_per_split) - kb_begin;
    };

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        const int nca = p.a_mn ? TC_BM / 32 : 1, ncb = p.b_mn ? BNH / 32 : 1;
        const int nload = (STRICT ? 2 : 1) * (nca + ncb);
        int l = lane;
        const bool is_lo = l >= nca + ncb; if (is_lo) l -= nca + ncb;
        const bool is_b = l >= nca; const int chunk = is_b ? l - nca : l;
        const CUtensorMap *map = is_b ? (is_lo ? &p.tmBlo : &p.tmB) : (is_lo ? &p.tmAlo : &p.tmA);
        const int mn_major = is_b ? p.b_mn : p.a_mn;
        const int dst_off = (is_lo ? A_BYTES + B_BYTES : 0) + (is_b ? A_BYTES : 0) + (mn_major ? chunk * TC_BK * 128 : 0);
        int it = 0;                                                       // ring position, runs on across tiles
        for (int tile = pair; tile < ntiles; tile += npairs) {
            int n0, m0, batch, z, kb_begin, nkb;
            tile_coords(tile, n0, m0, batch, z, kb_begin, nkb);
            const int mn = (is_b ? n0 + (int)rank * BNH : batch * p.a_batch_rows + m0) + (mn_major ? chunk * 32 : 0);
            for (int i = 0; i < nkb; ++i, ++it) {
                const int s = it % STAGES, round = it / STAGES;
                if (lane == 0) {
                    mbar_wait(&empty[s], (round & 1) ^ 1);
                    if (rank == 0) mbar_expect_tx(&full[s], 2 * STAGE_BYTES);      // both CTAs' boxes land on the leader's barrier
                }
                __syncwarp();
                if (lane < nload) {
                    const int kc = (kb_begin + i) * TC_BK;
                    uint8_t *dst = smem + s * STAGE_BYTES + dst_off;
                    const uint32_t bar = map_to_cta(smem_u32(&full[s]), 0);
                    if (mn_major) tma_load_2d_2cta(dst, map, bar, mn, kc);
                    else          tma_load_2d_2cta(dst, map, bar, kc, mn);
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===== MMA issuer (leader CTA only) =====
        const uint32_t idesc = make_idesc(2 * TC_BM, BN2, p.a_mn, p.b_mn);
        const uint64_t a_step = (uint64_t)(((p.a_mn ? 1024 : TC_UMMA_K * 4)) >> 4), b_step = (uint64_t)(((p.b_mn ? 1024 : TC_UMMA_K * 4)) >> 4);
        int it = 0, lt = 0;                                               // ring position; local tile counter
        for (int tile = pair; tile < ntiles; tile += npairs, ++lt) {
            int n0, m0, batch, z, kb_begin, nkb;
            tile_coords(tile, n0, m0, batch, z, kb_begin, nkb);
            const int buf = lt & 1;
            mbar_wait(&acc_empty[buf], ((lt >> 1) & 1) ^ 1);              // both epilogues have drained this accumulator (first use: free)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN2);
            for (int i = 0; i < nkb; ++i, ++it) {
                const int s = it % STAGES, round = it / STAGES;
                mbar_wait(&full[s], round & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint8_t *st = smem + s * STAGE_BYTES;
                    const uint64_t a_hi = make_smem_desc(st, p.a_mn), b_hi = make_smem_desc(st + A_BYTES, p.b_mn);
                    const uint64_t a_lo = make_smem_desc(st + A_BYTES + B_BYTES, p.a_mn), b_lo = make_smem_desc(st + 2 * A_BYTES + B_BYTES, p.b_mn);
#pragma unroll
                    for (int kk = 0; kk < TC_BK / TC_UMMA_K; ++kk) {
                        const uint64_t ad = kk * a_step, bd = kk * b_step;
                        const uint32_t acc0 = (i > 0 || kk > 0) ? 1u : 0u;
                        if (STRICT) {
                            umma_tf32_2cta(tmem_d, a_lo + ad, b_hi + bd, idesc, acc0);
                            umma_tf32_2cta(tmem_d, a_hi + ad, b_lo + bd, idesc, 1u);
                            umma_tf32_2cta(tmem_d, a_hi + ad, b_hi + bd, idesc, 1u);
                        } else {
                            umma_tf32_2cta(tmem_d, a_hi + ad, b_hi + bd, idesc, acc0);
                        }
                    }
                    umma_commit_2cta(&empty[s]);
                    if (i == nkb - 1) umma_commit_2cta(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs): own 128 rows x 256 columns of every tile of the pair =====
        // tcgen05.ld hands lane r the 32 consecutive columns of row r; storing that directly makes every warp store touch 32
        // different 128 B lines with 16 B each.  Each warp instead transposes its 32 x 32 block through a private shared
        // tile so that one store instruction writes 4 complete 128 B lines.
        const int ew = warp & 3;
        float *stg = reinterpret_cast<float *>(smem + STAGES * STAGE_BYTES + 256) + ew * (32 * EPI_LD);
        const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;              // this lane's row (within a group of 4) and first column when storing
        int lt = 0;
        for (int tile = pair; tile < ntiles; tile += npairs, ++lt) {
            int n0, m0, batch, z, kb_begin, nkb;
            tile_coords(tile, n0, m0, batch, z, kb_begin, nkb);
            const int buf = lt & 1;
            const int row_base = m0 + ew * 32;
            mbar_wait(&acc_full[buf], (lt >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float *out = p.partial ? p.partial + ((size_t)z * p.batches + batch) * p.M * p.ldp : p.C + (size_t)batch * p.c_batch_stride;
            const int ldo = p.partial ? p.ldp : p.ldc;
            const bool acc = (!p.partial) && p.accumulate;
            const bool vec_ok = ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll 1
            for (int c = 0; c < BN2; c += 32) {
                if (n0 + c >= p.N) break;
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(buf * BN2 + c), v);
#pragma unroll
                for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4 *>(stg + lane * EPI_LD + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                __syncwarp();
                const int col = n0 + c + sub_c;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int r = it * 4 + sub_r, row = row_base + r;
                    if (row >= p.M) continue;
                    float4 o = *reinterpret_cast<const float4 *>(stg + r * EPI_LD + sub_c);
                    float *dst = out + (size_t)row * ldo + col;
                    if (vec_ok && col + 4 <= p.N) {
                        if (acc) { const float4 old = *reinterpret_cast<const float4 *>(dst); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                        *reinterpret_cast<float4 *>(dst) = o;
                    } else {
                        const float e[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (col + i < p.N) dst[i] = acc ? dst[i] + e[i] : e[i];
                    }
                }
                __syncwarp();
            }
            // this warp has read its lanes of the accumulator: hand it back to the leader's MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&acc_empty[buf], 0);
        }
    }
    __syncthreads();
    cluster_sync_all();                                                   // neither CTA may free TMEM or exit while the pair is still running
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(2 * BN2) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ operand preparation
// One elementwise, bandwidth-bound pass per operand: re-pitch the row-major matrix to 16-byte aligned rows and, in strict
// mode, split it into hi = tf32_rna(x) (low 13 mantissa bits zero) and lo = x - hi (exact in fp32).  Optionally the rows
// and/or columns are re-blocked: source blocks of `bw` rows (columns) land at multiples of `bwp` >= bw in the destination,
// the padding is zero.  The LSTM layer uses this to start every (gate, direction) block of H cells on a 16-byte boundary,
// which TMA needs for sub-views along the contiguous dimension.
__device__ __forceinline__ void split_tf32(float v, float &hi, float &lo)
{
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi = __uint_as_float(h);
    lo = __fsub_rn(v, hi);
}

// grid (dst_rows, ceil(ldd/512)), 128 threads x 4 consecutive destination columns
template <bool STRICT>
__global__ void prep_rows_kernel(int cols, const float *__restrict__ src, size_t lds, float *__restrict__ hi, float *__restrict__ lo, size_t ldd,
                                 int rbw, int rbwp, int cbw, int cbwp, int dst_cols)
{
    const int rd = blockIdx.x;
    const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
    if ((size_t)c >= ldd) return;
    int rs = rd; bool row_ok = true;
    if (rbw) { const int b = rd / rbwp, j = rd - b * rbwp; row_ok = j < rbw; rs = b * rbw + j; }
    float v[4], h[4], l[4];
    const float *s = src + (size_t)rs * lds;
    if (!row_ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = 0.0f;
    } else if (!cbw && ((lds & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (c + 4 <= cols)) {
        const float4 q = *reinterpret_cast<const float4 *>(s + c); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int cs = c + i; bool ok = cs < dst_cols;
            if (cbw) { const int b = cs / cbwp, j = cs - b * cbwp; ok = ok && j < cbw; cs = b * cbw + j; }
            v[i] = (ok && cs < cols) ? s[cs] : 0.0f;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { if (STRICT) split_tf32(v[i], h[i], l[i]); else { h[i] = v[i]; l[i] = 0.0f; } }
    *reinterpret_cast<float4 *>(hi + (size_t)rd * ldd + c) = make_float4(h[0], h[1], h[2], h[3]);
    if (STRICT) *reinterpret_cast<float4 *>(lo + (size_t)rd * ldd + c) = make_float4(l[0], l[1], l[2], l[3]);
}

__global__ void sum_slices_kernel(int M, int N, int nsplit, int batches, const float *__restrict__ partial, int ldp,
                                  float *__restrict__ C, int ldc, long long c_batch_stride, int accumulate)
{
    const size_t total = (size_t)M * N;
    const int b = blockIdx.y;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const size_t r = e / N, c = e % N;
        float s = 0.0f;
        for (int z = 0; z < nsplit; ++z) s += partial[(((size_t)z * batches + b) * M + r) * ldp + c];      // fixed order: deterministic
        float *o = C + (size_t)b * c_batch_stride + r * ldc + c;
        *o = accumulate ? *o + s : s;
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn(bl_ctx *ctx)
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
            fail(ctx, "cuTensorMapEncodeTiled is not available from the driver");
            return nullptr;
        }
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2D map over a row-major fp32 matrix: `inner` contiguous floats per row, `outer` rows of pitch ld; box = 32 floats (128 B, one
// swizzle span) x box_outer rows; K-major views use the 16-byte-chunk 128B swizzle, MN-major views the 32-byte-chunk one.
// Boxes running past either extent are zero-filled.
static int make_map(bl_ctx *ctx, CUtensorMap *map, const float *base, int inner, int outer, size_t ld, int box_outer, bool mn_major)
{
    EncodeTiledFn fn = encode_fn(ctx);
    if (!fn) return 1;
    const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_outer};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, "cuTensorMapEncodeTiled failed (%d) inner=%d outer=%d ld=%zu", (int)r, inner, outer, ld);
    return 0;
}

template <int BN, bool STRICT, int STAGES>
static int launch_tc(bl_ctx *ctx, const GemmTcParams &p, dim3 grid)
{
    constexpr int STAGE_BYTES = (STRICT ? 2 : 1) * (TC_BM * TC_BK * 4 + BN * TC_BK * 4);
    const int smem = STAGES * STAGE_BYTES + 1024 + 256;
    auto kernel = gemm_tf32_tcgen05_kernel<BN, STRICT, STAGES>;
    BL_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kernel<<<grid, TC_THREADS, smem, ctx->stream>>>(p);
    BL_LAUNCHED(ctx);
    return 0;
}

template <bool STRICT, int STAGES>
static int launch_tc_2cta(bl_ctx *ctx, const GemmTcParams &p, dim3 grid)
{
    constexpr int STAGE_BYTES = (STRICT ? 2 : 1) * (TC_BM * TC_BK * 4 + 128 * TC_BK * 4);
    const int smem = STAGES * STAGE_BYTES + 1024 + 256 + 4 * 32 * EPI_LD * 4;       // ring + alignment slack + barriers + epilogue tiles
    auto kernel = gemm_tf32_tcgen05_2cta_kernel<STRICT, STAGES>;
    BL_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    BL_CUDA(ctx, cudaLaunchKernelEx(&cfg, kernel, p));
    BL_LAUNCHED(ctx);
    return 0;
}

size_t tc_operand_ld(int cols) { return ((size_t)cols + 3) & ~(size_t)3; }

static int reblocked(int n, int bw, int bwp) { return bw ? (n / bw) * bwp : n; }

size_t tc_operand_elems(int rows, int cols, int rbw, int rbwp, int cbw, int cbwp)
{
    return (size_t)reblocked(rows, rbw, rbwp) * tc_operand_ld(reblocked(cols, cbw, cbwp));
}

// Fills `out` with the prepared form of the row-major matrix src[rows][cols] (leading dimension ld_src).  hi/lo are caller-owned
// buffers of tc_operand_elems() floats (lo unused in fast mode).  In fast mode an already aligned source without re-blocking is
// used in place.
int tc_prepare(bl_ctx *ctx, const float *src, int rows, int cols, size_t ld_src, bool strict, float *hi, float *lo, TcOperand *out,
               int rbw, int rbwp, int cbw, int cbwp)
{
    TimedRegion timed(ctx, 0);
    if ((rbw && (rows % rbw || rbwp < rbw)) || (cbw && (cols % cbw || cbwp < cbw))) return fail(ctx, "tc_prepare: bad re-blocking");
    const int drows = reblocked(rows, rbw, rbwp), dcols = reblocked(cols, cbw, cbwp);
    out->rows = drows; out->cols = dcols; out->strict = strict;
    const bool aligned = ((ld_src & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (!strict && aligned && !rbw && !cbw) { out->hi = src; out->lo = nullptr; out->ld = ld_src; return 0; }
    const size_t ldd = tc_operand_ld(dcols);
    dim3 grid(drows, cdiv((int)ldd, 512));
    if (strict) prep_rows_kernel<true><<<grid, 128, 0, ctx->stream>>>(cols, src, ld_src, hi, lo, ldd, rbw, rbwp, cbw, cbwp, dcols);
    else        prep_rows_kernel<false><<<grid, 128, 0, ctx->stream>>>(cols, src, ld_src, hi, lo, ldd, rbw, rbwp, cbw, cbwp, dcols);
    BL_LAUNCHED(ctx);
    out->hi = hi; out->lo = strict ? lo : nullptr; out->ld = ldd;
    return 0;
}

// C[M x N] row-major (ldc) (+)= A * B^T with A = M x K and B = N x K views of prepared operands
int tc_gemm(bl_ctx *ctx, int M, int N, int K, const TcView &A, const TcView &B, float *C, int ldc, int accumulate)
{
    return tc_gemm_batched(ctx, M, N, K, A, B, C, ldc, accumulate, 1, 0, 0);
}

static int view_check(bl_ctx *ctx, const TcView &v, int mn_extent, int K, const char *which)
{
    const int MN = v.mn_major ? v.op->cols : v.op->rows, KK = v.mn_major ? v.op->rows : v.op->cols;
    if (v.mn0 < 0 || v.k0 < 0 || v.mn0 + mn_extent > MN || v.k0 + K > KK) return fail(ctx, "tc_gemm: %s sub-view out of range", which);
    // sub-views along the contiguous dimension must start on a 16-byte boundary (TMA global address alignment)
    if ((v.mn_major ? v.mn0 : v.k0) & 3) return fail(ctx, "tc_gemm: %s sub-view offset along the contiguous dimension must be a multiple of 4 floats", which);
    return 0;
}

// `batches` products sharing B: batch b uses the A view shifted by b*a_batch_mn along MN and writes C + b*c_batch_stride
int tc_gemm_batched(bl_ctx *ctx, int M, int N, int K, const TcView &A, const TcView &B, float *C, int ldc, int accumulate,
                    int batches, int a_batch_mn, long long c_batch_stride)
{
    TimedRegion timed(ctx, 0);
    const bool strict = A.op->strict;
    if (A.op->strict != B.op->strict) return fail(ctx, "tc_gemm: operands prepared for different precision modes");
    const int a_mn_total = (batches - 1) * a_batch_mn + M;
    BL_CHECK(view_check(ctx, A, a_mn_total, K, "A"));
    BL_CHECK(view_check(ctx, B, N, K, "B"));
    if (A.mn_major && (a_batch_mn & 3)) return fail(ctx, "tc_gemm: batch stride of an MN-major A must be a multiple of 4 floats");
    // strict mode: 256 x 256 tiles on CTA pairs (cta_group::2) unless BLSTM_TC_2CTA=0; BLSTM_TC_BN=256 selects 128x256 single-CTA tiles
    static const int bn_strict_env = getenv("BLSTM_TC_BN") ? atoi(getenv("BLSTM_TC_BN")) : 128;
    static const bool pair_env = !(getenv("BLSTM_TC_2CTA") && atoi(getenv("BLSTM_TC_2CTA")) == 0);
    const bool pair = strict && pair_env;
    const int BN_STRICT = (bn_strict_env == 256) ? 256 : 128; constexpr int BN_FAST = 256;
    const int BN = pair ? 256 : strict ? BN_STRICT : BN_FAST;
    const int BMT = pair ? 2 * TC_BM : TC_BM;                              // rows of one output tile
    const int tiles = batches * cdiv(M, BMT) * cdiv(N, BN) * (pair ? 2 : 1);      // in CTAs
    const int kb_total = cdiv(K, TC_BK);
    int nsplit = 1;
    if (tiles < ctx->num_sms && kb_total >= 16) {
        nsplit = (2 * ctx->num_sms) / tiles;               // at most two full waves of CTAs: a third, mostly empty wave costs more than it buys
        if (nsplit > kb_total / 8) nsplit = kb_total / 8;
        if (nsplit > 64) nsplit = 64;
        if (nsplit < 1) nsplit = 1;
    }
    const int kbs = cdiv(kb_total, nsplit);
    nsplit = cdiv(kb_total, kbs);
    const size_t ldp = tc_operand_ld(N);
    GemmTcParams p;
    p.partial = nullptr;
    if (nsplit > 1) {
        BL_CHECK(ensure_scratch2(ctx, (size_t)nsplit * batches * M * ldp * sizeof(float) + 64));
        p.partial = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(ctx->scratch2) + 15) & ~(uintptr_t)15);
    }
    // the maps cover exactly the sub-view: boxes running past its MN or K extent are zero-filled
    auto maps = [&](const TcView &v, int mn_extent, int box_mn, CUtensorMap *hi, CUtensorMap *lo) -> int {
        const size_t off = v.mn_major ? (size_t)v.k0 * v.op->ld + v.mn0 : (size_t)v.mn0 * v.op->ld + v.k0;
        const int inner = v.mn_major ? mn_extent : K, outer = v.mn_major ? K : mn_extent, box_outer = v.mn_major ? TC_BK : box_mn;
        BL_CHECK(make_map(ctx, hi, v.op->hi + off, inner, outer, v.op->ld, box_outer, v.mn_major));
        if (strict) BL_CHECK(make_map(ctx, lo, v.op->lo + off, inner, outer, v.op->ld, box_outer, v.mn_major));
        else *lo = *hi;
        return 0;
    };
    BL_CHECK(maps(A, a_mn_total, TC_BM, &p.tmA, &p.tmAlo));
    BL_CHECK(maps(B, N, pair ? 128 : BN, &p.tmB, &p.tmBlo));
    p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K; p.a_mn = A.mn_major ? 1 : 0; p.b_mn = B.mn_major ? 1 : 0;
    p.kblocks_per_split = kbs; p.accumulate = accumulate; p.ldp = (int)ldp;
    p.batches = batches; p.mt_per_batch = cdiv(M, BMT); p.a_batch_rows = a_batch_mn; p.c_batch_stride = c_batch_stride;
    dim3 grid(cdiv(N, BN) * (pair ? 2 : 1), batches * cdiv(M, BMT), nsplit);
    if (grid.y > 65535) return fail(ctx, "tc_gemm: M too large");
    p.tiles_x = pair ? (int)grid.x / 2 : (int)grid.x; p.tiles_y = (int)grid.y; p.tiles_z = (int)grid.z;
    if (pair) {
        // persistent launch: one CTA pair per SM pair (or per tile when there are fewer tiles)
        const long long ntiles = (long long)p.tiles_x * p.tiles_y * p.tiles_z;
        const int pairs = (int)std::min<long long>(ntiles, ctx->num_sms / 2);
        BL_CHECK((launch_tc_2cta<true, 3>(ctx, p, dim3(2 * pairs, 1, 1))));
    }
    else if (strict && BN_STRICT == 256) BL_CHECK((launch_tc<256, true, 2>(ctx, p, grid)));
    else if (strict) BL_CHECK((launch_tc<128, true, 3>(ctx, p, grid)));
    else        BL_CHECK((launch_tc<BN_FAST, false, 4>(ctx, p, grid)));
    if (nsplit > 1) {
        int blocks = (int)cdivz((size_t)M * N, 256); if (blocks > 2048) blocks = 2048;
        sum_slices_kernel<<<dim3(blocks, batches), 256, 0, ctx->stream>>>(M, N, nsplit, batches, p.partial, (int)ldp, C, ldc, c_batch_stride, accumulate);
        BL_LAUNCHED(ctx);
    }
    return 0;
}

// Generic entry used by bl_gemm_f32: prepares both operands in the context's scratch, then multiplies.
// C[M x N] row-major (ldc) (+)= A[M x K] * B[N x K]^T.  a_kmajor: A is given as [M][K] (lda) else as [K][M] (lda); same for B.
int gemm_tf32_tc(bl_ctx *ctx, int M, int N, int K, const float *A, size_t lda, bool a_kmajor, const float *B, size_t ldb, bool b_kmajor,
                 float *C, int ldc, int accumulate, int mode)
{
    const bool strict = (mode == BL_GEMM_STRICT);
    const int ar = a_kmajor ? M : K, ac = a_kmajor ? K : M, br = b_kmajor ? N : K, bc = b_kmajor ? K : N;
    const size_t a_elems = tc_operand_elems(ar, ac), b_elems = tc_operand_elems(br, bc);
    BL_CHECK(ensure_scratch(ctx, 2 * (a_elems + b_elems) * sizeof(float) + 64));
    float *S = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(ctx->scratch) + 15) & ~(uintptr_t)15);
    TcOperand a, b;
    BL_CHECK(tc_prepare(ctx, A, ar, ac, lda, strict, S, S + a_elems, &a));
    BL_CHECK(tc_prepare(ctx, B, br, bc, ldb, strict, S + 2 * a_elems, S + 2 * a_elems + b_elems, &b));
    return tc_gemm(ctx, M, N, K, TcView{&a, !a_kmajor, 0, 0}, TcView{&b, !b_kmajor, 0, 0}, C, ldc, accumulate);
}

} // namespace bl
