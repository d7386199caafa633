// Persistent recurrent kernels, register-resident variant ("v3"): same decomposition and step protocol as
// lstm_recurrent.cu, but the CTA's slice of the recurrent weights lives in REGISTERS for the whole pass instead of
// shared memory: thread (rp, ks) owns rows {rp, rp + RH} x 32 consecutive k of the slice (64 weights = 64 registers).
// Per timestep a thread therefore only streams the previous-step vector from shared memory (LDS.128, every lane of a
// warp reads the same address -> one broadcast wavefront) and issues 8 FFMA per LDS: half the shared-memory
// instructions of the 4x4-tile variant, a fifth of its wavefronts, no weight address arithmetic.
// Used whenever the slice fits (rows/2 x K/32 <= 512 threads); lstm_recurrent.cu remains the fallback.
//
// Shared / exchange row layout: 32-float chunks padded to 36 floats, so the k-chunks that share a warp (small slices)
// fall into different banks:  koff(k) = (k >> 5) * 36 + (k & 31).
#include "lstm_recurrent.cuh"
#include <cmath>
#include <cstdint>

namespace bl {

constexpr int RG_NT = 512;          // threads per CTA: 128 registers each
constexpr int RG_KC = 32;           // k per thread chunk
constexpr int RG_CS = 36;           // padded chunk stride (floats)

__host__ __device__ __forceinline__ int rg_koff(int k) { return (k >> 5) * RG_CS + (k & 31); }

static int pad32(int x) { return (x + 31) / 32 * 32; }

// Geometry for the register-resident kernels.  RecGeom fields reused: G, C, CL, SG, R (rows), RQt = RH (row halves),
// KS = KSn (k chunks), Spad = SGp, RS (row stride), RP (stage row pitch), Hpad (= pad32(H)), K4 unused.
bool choose_geometry_reg(bool bwd, int H, int S, int ndir, int num_sms, int smem_cap, int forceG, RecGeom *out)
{
    const int Hp = pad32(H);
    const int Kp = bwd ? 4 * Hp : Hp;
    const int KSn = Kp / RG_KC;
    const int RS = KSn * RG_CS;
    const int per_dir = num_sms / ndir;
    bool found = false;
    RecGeom best{};
    for (int G = 1; G <= 32 && G <= S; ++G) {
        if (forceG > 0 && G != forceG) continue;
        int C = per_dir / G;
        if (C < 1) break;
        const int CL = cdiv(H, C);
        C = cdiv(H, CL);
        const int SG = cdiv(S, G);
        if ((G - 1) * SG >= S) continue;                      // trailing group would be empty
        const int R = bwd ? CL : 4 * CL;
        const int RH = (R + 1) / 2;
        if (RH * KSn > RG_NT) continue;                       // slice does not fit the register file
        if (CL * SG > REC_NPAIR * RG_NT) continue;
        const int SGp = (SG + 3) / 4 * 4;
        const int RP = 2 * RH;
        const size_t smem = ((size_t)SGp * RS + (size_t)KSn * SGp * RP) * sizeof(float);
        if ((int)smem > smem_cap) continue;
        // cost model (cycles per step): FFMA issue of the busiest scheduler + exchange copy + counter round trip
        const int warps = cdiv(RH * KSn, 32);
        const double gemm = (double)(SGp / 4) * 300.0 * cdiv(warps, 4);
        const double gate = 1200.0 + 60.0 * KSn + 900.0 * (CL * SG > RG_NT ? 2 : 1);
        const double copy = (double)SG * RS * 4.0 / 48.0 + 300.0;
        const double cost = gemm + gate + copy + 1500.0 + 12.0 * C;
        if (!found || cost < best.cost) {
            found = true;
            best = RecGeom{};
            best.G = G; best.C = C; best.CL = CL; best.SG = SG; best.NT = RG_NT; best.nsub = 1;
            best.npair = (CL * SG > RG_NT) ? 2 : 1;
            best.R = R; best.RQt = RH; best.KS = KSn; best.Spad = SGp; best.RS = RS; best.RP = RP; best.Hpad = Hp;
            best.smem = smem; best.cost = cost;
        }
    }
    if (found) *out = best;
    return found;
}

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ unsigned rg_ld_acquire(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// counter protocol of lstm_recurrent.cu (cooperative-groups grid sync restricted to one sequence group's slices)
__device__ __forceinline__ void rg_wait(const unsigned *flag, unsigned target)
{
    if (threadIdx.x == 0) { while (rg_ld_acquire(flag) < target) { } }
    __syncthreads();
}
// per-warp variant: lane 0 of every warp polls, so the exchange copy can start without a CTA barrier in between (the copy is
// followed by one anyway)
__device__ __forceinline__ void rg_wait_warp(const unsigned *flag, unsigned target)
{
    if ((threadIdx.x & 31) == 0) { while (rg_ld_acquire(flag) < target) { } }
    __syncwarp();
}
__device__ __forceinline__ void rg_publish(unsigned *flag)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        atomicAdd(flag, 1u);
    }
}

// Packed fp32 FMA (Blackwell FFMA2): one instruction performs two IEEE fma.rn -- bit-identical to two fmaf -- and halves the
// issue slots the step GEMM needs (the FMA datapath rate is unchanged; measured with tools/micro/ffma2.cu).
__device__ __forceinline__ unsigned long long rg_pack2(float lo, float hi)
{ unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void rg_unpack2(unsigned long long v, float &lo, float &hi)
{ asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void rg_fma2(unsigned long long &acc, unsigned long long a, unsigned long long b)
{ asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }

// previous-step vector (tile, [SGp][RS]) times this thread's 2 x 32 register-resident weights (w[e] = {row rp, row rp+RH} at
// k = ks*32 + e, packed); partials to stage[ks][s][row]
__device__ __forceinline__ void reg_gemm(const RecGeom &g, const unsigned long long (&w)[RG_KC],
                                         const float *__restrict__ tile, float *__restrict__ stage, int rp, int ks, int nseq)
{
    const float *tp = tile + ks * RG_CS;
    float *sp = stage + (ks * g.Spad) * g.RP + rp;
    for (int s0 = 0; s0 < nseq; s0 += 4) {
        unsigned long long acc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] = 0ull;
        const float *t0 = tp + s0 * g.RS;
#pragma unroll
        for (int k4 = 0; k4 < RG_KC / 4; ++k4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 h = *reinterpret_cast<const float4 *>(t0 + q * g.RS + k4 * 4);
                rg_fma2(acc[q], w[k4 * 4 + 0], rg_pack2(h.x, h.x));
                rg_fma2(acc[q], w[k4 * 4 + 1], rg_pack2(h.y, h.y));
                rg_fma2(acc[q], w[k4 * 4 + 2], rg_pack2(h.z, h.z));
                rg_fma2(acc[q], w[k4 * 4 + 3], rg_pack2(h.w, h.w));
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float a0, a1;
            rg_unpack2(acc[q], a0, a1);
            sp[(s0 + q) * g.RP] = a0;
            sp[(s0 + q) * g.RP + g.RQt] = a1;
        }
    }
}

__device__ __forceinline__ float rg_stage_sum(const float *sp, int KSn, int kstride)
{
    float s0 = 0.0f, s1 = 0.0f;
    int ks = 0;
    for (; ks + 1 < KSn; ks += 2) { s0 += sp[ks * kstride]; s1 += sp[(ks + 1) * kstride]; }
    if (ks < KSn) s0 += sp[ks * kstride];
    return s0 + s1;
}

// hi = tf32_rna(x), lo = x - hi (exact): the operand split of the strict tensor-core GEMMs (gemm_tc.cu), produced here so that the
// backward pass needs no separate split pass over the deltas / layer outputs
__device__ __forceinline__ void rg_split_store(float *hi, float *lo, size_t idx, float v)
{
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi[idx] = __uint_as_float(h);
    lo[idx] = __fsub_rn(v, __uint_as_float(h));
}

// ------------------------------------------------------------------------------------------------ forward
template <int NPAIR>
__global__ void __launch_bounds__(RG_NT, 1) lstm_fwd_reg_kernel(const RecFwdParams p)
{
    extern __shared__ __align__(16) float smem[];
    __shared__ unsigned long long s_tab[32];
    const RecGeom &g = p.g;
    const int tid = threadIdx.x;
    float *tile = smem;                                            // [SGp][RS]
    float *stage = smem + g.Spad * g.RS;                           // [KSn][SGp][RP]
    const int kstride = g.Spad * g.RP;

    const int H = p.H, L = p.L, S = p.S, T = p.T;
    const int d = blockIdx.x / (g.G * g.C);
    const int grp = (blockIdx.x % (g.G * g.C)) / g.C;
    const int cs = blockIdx.x % g.C;
    const int j0 = cs * g.CL, ncell = min(g.CL, H - j0);
    const int s0 = grp * g.SG, nseq = min(g.SG, S - s0);
    unsigned *flag = p.flags + (d * g.G + grp) * 32;

    for (int i = tid; i < (int)(g.smem / sizeof(float)); i += RG_NT) smem[i] = 0.0f;
    if (tid < 32) s_tab[tid] = bl_exp2f_tab[tid];

    // this thread's weights: rows {rp, rp+RH} of the slice (row = gate*CL + cell), k in [ks*32, ks*32+32)
    const int RH = g.RQt, KSn = g.KS;
    const bool gemm_thread = tid < RH * KSn;
    const int ks = tid / RH, rp = tid - ks * RH;
    unsigned long long w[RG_KC];
    {
        const int r0 = rp, r1 = rp + RH;
        const int g0 = r0 / g.CL, c0 = r0 - g0 * g.CL, g1 = r1 / g.CL, c1 = r1 - g1 * g.CL;
        const bool ok0 = gemm_thread && r0 < g.R && c0 < ncell, ok1 = gemm_thread && r1 < g.R && c1 < ncell;
        // weight layout internal: gate*L*H + d*H*H + j*H + k (k = source cell), LstmLayer.cu:586-596
        const float *q0 = p.Wi + (size_t)g0 * L * H + (size_t)d * H * H + (size_t)(j0 + c0) * H;
        const float *q1 = p.Wi + (size_t)g1 * L * H + (size_t)d * H * H + (size_t)(j0 + c1) * H;
#pragma unroll
        for (int e = 0; e < RG_KC; ++e) {
            const int k = ks * RG_KC + e;
            w[e] = rg_pack2((ok0 && k < H) ? __ldg(q0 + k) : 0.0f, (ok1 && k < H) ? __ldg(q1 + k) : 0.0f);
        }
    }

    bool valid[NPAIR]; int cl_[NPAIR], sl_[NPAIR];
    float wb[NPAIR][4], wpe[NPAIR][3], cprev[NPAIR];
#pragma unroll
    for (int u = 0; u < NPAIR; ++u) {
        const int pr = tid + u * RG_NT;
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
    __syncthreads();

    for (int q = 0; q < T; ++q) {
        const int t = (d == 0) ? q : T - 1 - q;
        const bool first = (q == 0);
        const bool check = (t >= p.Tmin);
        float *acts_t = p.acts + (size_t)t * S * 4 * L + d * H + j0;
        float *cst_t = p.cst + (size_t)t * S * L + d * H + j0;
        float *y_t = p.Y + (size_t)t * S * p.ldy + d * H + j0;
        const char *pat_t = p.pat + (size_t)t * S;
        float *hx_w = p.hx + (size_t)(d * 2 + (q & 1)) * S * g.RS;

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
            rg_wait_warp(flag, (unsigned)(g.C * q));
            if (tr && tid == 0) tr[1] = clock64();
            const float4 *src = reinterpret_cast<const float4 *>(p.hx + ((size_t)(d * 2 + ((q - 1) & 1)) * S + s0) * g.RS);
            float4 *dst = reinterpret_cast<float4 *>(tile);
            const int n4 = nseq * g.RS / 4;
            for (int i = tid; i < n4; i += RG_NT) dst[i] = __ldcg(src + i);
            __syncthreads();
            if (tr && tid == 0) tr[2] = clock64();
            if (gemm_thread) reg_gemm(g, w, tile, stage, rp, ks, nseq);
            __syncthreads();
            if (tr && tid == 0) tr[3] = clock64();
        }

        // gate math; only the exchange value h is stored before the publish so that the release fence has as little as
        // possible to wait for -- the HBM-only results (activations, c, layer output) are stored after it
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
                    const float *sp = stage + sl_[u] * g.RP + cl;
                    ni = __fadd_rn(ni, rg_stage_sum(sp, KSn, kstride));
                    ig = __fadd_rn(ig, rg_stage_sum(sp + g.CL, KSn, kstride));
                    fg = __fadd_rn(fg, rg_stage_sum(sp + 2 * g.CL, KSn, kstride));
                    og = __fadd_rn(og, rg_stage_sum(sp + 3 * g.CL, KSn, kstride));
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
            hx_w[slot * g.RS + rg_koff(j0 + cl)] = h;
        }
        if (tr && tid == 0) tr[4] = clock64();
        if (q + 1 < T) rg_publish(flag);
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
            if (p.ys_hi) rg_split_store(p.ys_hi, p.ys_lo, ((size_t)t * S + slot) * p.ld_ys + d * ((H + 3) & ~3) + j0 + cl, r_h[u]);
        }
        if (tr && tid == 0) tr[5] = clock64();
    }
}

// ------------------------------------------------------------------------------------------------ BPTT
template <int NPAIR>
__global__ void __launch_bounds__(RG_NT, 1) lstm_bwd_reg_kernel(const RecBwdParams p)
{
    extern __shared__ __align__(16) float smem[];
    __shared__ unsigned long long s_tab[32];
    const RecGeom &g = p.g;
    const int tid = threadIdx.x;
    float *tile = smem;
    float *stage = smem + g.Spad * g.RS;
    const int kstride = g.Spad * g.RP;

    const int H = p.H, L = p.L, S = p.S, T = p.T, Hp = g.Hpad;
    const int d = blockIdx.x / (g.G * g.C);
    const int grp = (blockIdx.x % (g.G * g.C)) / g.C;
    const int cs = blockIdx.x % g.C;
    const int j0 = cs * g.CL, ncell = min(g.CL, H - j0);
    const int s0 = grp * g.SG, nseq = min(g.SG, S - s0);
    unsigned *flag = p.flags + (d * g.G + grp) * 32;
    const bool inplace = (p.ndir == 1);

    for (int i = tid; i < (int)(g.smem / sizeof(float)); i += RG_NT) smem[i] = 0.0f;
    if (tid < 32) s_tab[tid] = bl_exp2f_tab[tid];

    // this thread's weights: rows (cells) {rp, rp+RH}, contraction index kk = gate*Hp + j in [ks*32, ks*32+32):
    // W_gate[k, j] = Wi[gate*L*H + d*H*H + j*H + k]   (the (N,N) products of LstmLayer.cu:939-942)
    const int RH = g.RQt, KSn = g.KS;
    const bool gemm_thread = tid < RH * KSn;
    const int ks = tid / RH, rp = tid - ks * RH;
    unsigned long long w[RG_KC];
    {
        const int c0 = rp, c1 = rp + RH;
        const bool ok0 = gemm_thread && c0 < ncell, ok1 = gemm_thread && c1 < ncell;
#pragma unroll
        for (int e = 0; e < RG_KC; ++e) {
            const int kk = ks * RG_KC + e;
            const int gi = kk / Hp, j = kk - gi * Hp;
            const float *q = p.Wi + (size_t)gi * L * H + (size_t)d * H * H + (size_t)j * H + j0;
            w[e] = rg_pack2((ok0 && j < H) ? __ldg(q + c0) : 0.0f, (ok1 && j < H) ? __ldg(q + c1) : 0.0f);
        }
    }

    bool valid[NPAIR]; int cl_[NPAIR], sl_[NPAIR];
    float wpe[NPAIR][3];
    float nfg[NPAIR], ncerr[NPAIR], ndig[NPAIR], ndfg[NPAIR];   // "next step" state, :253-256
#pragma unroll
    for (int u = 0; u < NPAIR; ++u) {
        const int pr = tid + u * RG_NT;
        cl_[u] = pr % g.CL; sl_[u] = pr / g.CL;
        valid[u] = (cl_[u] < ncell) && (sl_[u] < nseq);
        nfg[u] = ncerr[u] = ndig[u] = ndfg[u] = 0.0f;
        if (valid[u]) {
            const int col = d * H + j0 + cl_[u];
#pragma unroll
            for (int gi = 0; gi < 3; ++gi) wpe[u][gi] = __ldg(p.Wp + gi * L + col);
        }
    }
    __syncthreads();

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
        float *dx_w = p.dx + (size_t)(d * 2 + (q & 1)) * S * g.RS;

        float a[NPAIR][4], c[NPAIR], cp[NPAIR], oe[NPAIR], r_dni[NPAIR], r_dog[NPAIR]; bool dummy[NPAIR];
#pragma unroll
        for (int u = 0; u < NPAIR; ++u) {
            dummy[u] = false; cp[u] = 0.0f; r_dni[u] = r_dog[u] = 0.0f;
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
            rg_wait_warp(flag, (unsigned)(g.C * q));
            const float4 *src = reinterpret_cast<const float4 *>(p.dx + ((size_t)(d * 2 + ((q - 1) & 1)) * S + s0) * g.RS);
            float4 *dst = reinterpret_cast<float4 *>(tile);
            const int n4 = nseq * g.RS / 4;
            for (int i = tid; i < n4; i += RG_NT) dst[i] = __ldcg(src + i);
            __syncthreads();
            if (gemm_thread) reg_gemm(g, w, tile, stage, rp, ks, nseq);
            __syncthreads();
        }

#pragma unroll
        for (int u = 0; u < NPAIR; ++u) {
            if (!valid[u]) continue;
            const int slot = s0 + sl_[u], cl = cl_[u];
            float e = oe[u];
            if (!firstCall) e = __fadd_rn(e, rg_stage_sum(stage + sl_[u] * g.RP + cl, KSn, kstride));   // the 4 addProducts of :939-942
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
            float *xp = dx_w + slot * g.RS;                      // exchange values first: the release fence only waits for these
            const int jj = j0 + cl;
            xp[rg_koff(jj)] = dni; xp[rg_koff(Hp + jj)] = dig; xp[rg_koff(2 * Hp + jj)] = dfg; xp[rg_koff(3 * Hp + jj)] = dog;
        }
        if (q + 1 < T) rg_publish(flag);
#pragma unroll
        for (int u = 0; u < NPAIR; ++u) {                        // HBM-only results after the publish
            if (!valid[u]) continue;
            const int slot = s0 + sl_[u], cl = cl_[u];
            float *dp = del_t + slot * 4 * L + cl;
            dp[0] = r_dni[u]; dp[L] = ndig[u]; dp[2 * L] = ndfg[u]; dp[3 * L] = r_dog[u];
            cerr_t[slot * L + cl] = ncerr[u];
            if (p.ds_hi) {
                const int Hq = (H + 3) & ~3;
                const size_t base = ((size_t)t * S + slot) * p.ld_ds + (size_t)d * Hq + j0 + cl, gs = (size_t)p.ndir * Hq;
                rg_split_store(p.ds_hi, p.ds_lo, base, r_dni[u]); rg_split_store(p.ds_hi, p.ds_lo, base + gs, ndig[u]);
                rg_split_store(p.ds_hi, p.ds_lo, base + 2 * gs, ndfg[u]); rg_split_store(p.ds_hi, p.ds_lo, base + 3 * gs, r_dog[u]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ launch
template <typename Params, typename Kernel>
static int launch_reg(bl_ctx *ctx, Kernel kernel, const Params &p, const char *name)
{
    const RecGeom &g = p.g;
    const int grid = p.ndir * g.G * g.C;
    BL_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    int per_sm = 0;
    BL_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, RG_NT, g.smem));
    if (per_sm < 1 || grid > per_sm * ctx->num_sms)
        return fail(ctx, "%s: %d CTAs cannot be co-resident (%d per SM x %d SMs)", name, grid, per_sm, ctx->num_sms);
    BL_CUDA(ctx, cudaMemsetAsync(p.flags, 0, (size_t)p.ndir * g.G * 32 * sizeof(unsigned), ctx->stream));
    void *args[] = { (void *)&p };
    BL_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(RG_NT), args, g.smem, ctx->stream));
    BL_LAUNCHED(ctx);
    return 0;
}

int launch_lstm_fwd_reg(bl_ctx *ctx, const RecFwdParams &p)
{
    TimedRegion timed(ctx, 1);
    return p.g.npair == 1 ? launch_reg(ctx, lstm_fwd_reg_kernel<1>, p, "lstm_fwd_reg") : launch_reg(ctx, lstm_fwd_reg_kernel<2>, p, "lstm_fwd_reg");
}
int launch_lstm_bwd_reg(bl_ctx *ctx, const RecBwdParams &p)
{
    TimedRegion timed(ctx, 2);
    return p.g.npair == 1 ? launch_reg(ctx, lstm_bwd_reg_kernel<1>, p, "lstm_bwd_reg") : launch_reg(ctx, lstm_bwd_reg_kernel<2>, p, "lstm_bwd_reg");
}

} // namespace bl
