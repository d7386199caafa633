// Persistent recurrent kernels: the whole time loop of an LSTM layer pass in ONE launch.
//
// Replaces, per layer and pass, the reference's 2*(T-1)*4 cublasSgemm calls + 2*T Thrust launches
// (layers/LstmLayer.cu:812-829, 847-864 forward; :936-951, 970-985 BPTT).
//
// Decomposition (both kernels): grid = ndir x G x C CTAs, all co-resident (cooperative launch), 1 CTA / SM.
//   CTA (d, g, c) owns cells [c*CL, c*CL+CL) of direction d for the sequences of group g.
//   Its slice of the recurrent weights is loaded into shared memory ONCE and reused for all T steps.
// Per timestep:
//   1. wait on the (d,g) step counter until all C slices published the previous step          (acquire)
//   2. copy the group's previous-step vector (h, or the 4 gate deltas) from the L2-resident exchange
//      buffer into shared memory (float4, ld.global.cg)
//   3. shared-memory GEMM rows x sequences x K on the FFMA pipe: 4x4 register tiles, float4 operand
//      loads, K split over warps, partial sums staged through shared memory
//   4. fused gate nonlinearity / cell update (ComputeBlockOutputFn, LstmLayer.cu:47-138) or delta
//      computation (ComputeBlockErrorsFn, :190-287) in registers; cell state and the "next step"
//      backward state never leave registers between steps
//   5. results to HBM (activations, c, h straight into the [N][L] layer output - no resort pass),
//      h / deltas to the exchange buffer, __threadfence, counter += 1                          (release)
// Forward- and backward-in-time directions are different CTAs of the same grid and run concurrently.
#include "lstm_recurrent.cuh"
#include <cooperative_groups.h>
#include <cstdlib>
#include <cmath>

namespace bl {

// ------------------------------------------------------------------------------------------------ geometry
static int hpad_of(int H)
{   // smallest multiple of 4 with Hpad % 8 == 4: rows of consecutive cells/sequences then land on distinct
    // 16-byte bank groups for LDS.128 (stride == 4 mod 8 floats)
    int h = (H + 3) / 4 * 4;
    if (h % 8 != 4) h += 4;
    return h;
}

bool choose_geometry(bool bwd, int H, int S, int ndir, int num_sms, int smem_cap, int forceG, int forceNsub, int forceNT, RecGeom *out)
{
    // 512 threads (128 registers each) measured fastest on B200 for both kernels (tools/sweep_geometry.py, profiles/):
    // 768 / 1024 threads spill inside the GEMM loop and shorten the per-warp K range
    const int REC_NT = (forceNT == 768 || forceNT == 1024) ? forceNT : 512;
    const int Hpad = hpad_of(H);
    bool found = false;
    RecGeom best{};
    const int per_dir = num_sms / ndir;
    for (int nsub = 1; nsub <= 4; nsub *= 2) {
        if (forceNsub > 0 && nsub != forceNsub) continue;
        const int nt = REC_NT / nsub, nw = nt / 32;
        for (int Gc = 1; Gc <= 16 && Gc * nsub <= S; ++Gc) {          // Gc: sequence groups at CTA level
            const int G = Gc * nsub;
            if (forceG > 0 && G != forceG) continue;
            int C = per_dir / Gc;
            if (C < 1) break;
            const int CL = cdiv(H, C);
            C = cdiv(H, CL);
            const int SG = cdiv(S, G);
            if (CL * SG > REC_NPAIR * nt) continue;
            const int npair = (CL * SG > nt) ? 2 : 1;
            const int R = bwd ? CL : 4 * CL;
            const int RQ = cdiv(R, 4), SQ = cdiv(SG, 4);
            const int K = bwd ? 4 * Hpad : Hpad, K4 = K / 4;
            const int RS = bwd ? 4 * Hpad + 4 : Hpad;
            for (int LR = 1; LR <= 32; LR *= 2) {
                const int LS = 32 / LR;
                const int WR = cdiv(RQ, LR), WS = cdiv(SQ, LS);
                const int tasks = WR * WS;
                int KS = nw / tasks; if (KS < 1) KS = 1;
                if (KS > 8) KS = 8;                                             // more splits only grow the staging buffer and its reduction
                while (KS > 1 && K4 / KS < 8) --KS;
                const int KB4 = cdiv(K4, KS);
                const int Rpad = 4 * WR * LR, Spad = 4 * WS * LS;
                const int RP = Rpad | 1;
                const size_t smem = ((size_t)Rpad * RS + (size_t)nsub * ((size_t)Spad * RS + (size_t)KS * Spad * RP)) * sizeof(float);
                if ((int)smem > smem_cap) continue;
                // cost model (cycles per step), calibrated against the in-kernel clock64 traces (tools/trace_recurrent.py):
                // the GEMM phase runs at the larger of its issue-slot and shared-memory-wavefront counts, the exchange copy
                // at ~48 B/clk/SM out of L2, and the counter round trip grows slowly with the number of slices polling it.
                // Sub-CTAs (nsub > 1) measured no faster than one group per CTA on B200, so they only win on ties.
                const double iters = (double)tasks * KS * KB4;                   // warp-iterations, 64 FFMA + 8 LDS.128 each
                const double issue = iters * 76.0 / 4.0;
                const double wave  = iters * 4.0 * ((LR > 8 ? LR / 8 : 1) + (LS > 8 ? LS / 8 : 1));
                const double copy  = (double)SG * K * 4.0 / 48.0;
                const double sync  = 1500.0 + 12.0 * C;
                const double cost  = ((issue > wave ? issue : wave) * nsub + copy + sync) * (1.0 + 0.05 * (nsub - 1));
                if (!found || cost < best.cost) {
                    found = true;
                    best.G = G; best.C = C; best.CL = CL; best.SG = SG; best.R = R; best.nsub = nsub; best.npair = npair; best.NT = REC_NT;
                    best.LR = LR; best.LS = LS; best.LSlog = (int)std::lround(std::log2((double)LS));
                    best.WR = WR; best.WS = WS; best.KS = KS; best.RQt = WR * LR; best.SQt = WS * LS;
                    best.Rpad = Rpad; best.Spad = Spad; best.K4 = K4; best.KB4 = KB4; best.RS = RS; best.RP = RP;
                    best.Hpad = Hpad; best.smem = smem; best.cost = cost;
                }
            }
        }
    }
    if (found) *out = best;
    return found;
}

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// barrier over one sub-CTA (named barrier sub+1; barrier 0 stays the whole-CTA __syncthreads)
__device__ __forceinline__ void bar_sub(int sub, int nt)
{
    asm volatile("bar.sync %0, %1;" :: "r"(sub + 1), "r"(nt) : "memory");
}

// Step-counter protocol (the one cooperative-groups grid sync uses, restricted to the C slices of one sequence group):
// producer: all stores -> sub-CTA barrier -> ONE thread: __threadfence (cumulative over what the barrier ordered) + atomicAdd
// consumer: ONE thread: acquire-load spin -> sub-CTA barrier -> everybody reads the exchange buffer through L2 (ld.global.cg)
__device__ __forceinline__ void wait_counter(const unsigned *flag, unsigned target, int ts, int sub, int nt)
{
    if (ts == 0) {
        while (ld_acquire_u32(flag) < target) { }
    }
    bar_sub(sub, nt);
}

__device__ __forceinline__ void publish(unsigned *flag, int ts, int sub, int nt)
{
    bar_sub(sub, nt);
    if (ts == 0) {
        __threadfence();
        atomicAdd(flag, 1u);
    }
}

// rows x sequences x K product out of shared memory; partial sums (one per K split) into `stage`.
// 4x4 register tile per thread; the 4 weight rows of a k-quad stay in registers while the 4 sequences stream through,
// which keeps the live set at 16 accumulators + 5 float4 (64-register budget of the 1024-thread CTA).
__device__ __forceinline__ void smem_gemm(const RecGeom &g, const float *__restrict__ Wsm, const float *__restrict__ tile,
                                          float *__restrict__ stage, int warp, int lane, int nw)
{
    const int tasks = g.WR * g.WS * g.KS;
    const int lr = lane >> g.LSlog, ls = lane & (g.LS - 1);
    const int rstep = g.RQt * g.RS / 4, sstep = g.SQt * g.RS / 4;      // float4 strides between a thread's rows / sequences
    for (int wt = warp; wt < tasks; wt += nw) {
        const int ks = wt % g.KS, tl = wt / g.KS;
        const int wr = tl % g.WR, ws = tl / g.WR;
        const int rq = wr * g.LR + lr, sq = ws * g.LS + ls;
        const int kb = ks * g.KB4, ke = min(g.K4, kb + g.KB4);
        const float4 *wp = reinterpret_cast<const float4 *>(Wsm) + rq * (g.RS / 4);
        const float4 *hp = reinterpret_cast<const float4 *>(tile) + sq * (g.RS / 4);
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][q] = 0.0f;
#pragma unroll 2
        for (int k4 = kb; k4 < ke; ++k4) {
            const float4 w0 = wp[k4], w1 = wp[k4 + rstep], w2 = wp[k4 + 2 * rstep], w3 = wp[k4 + 3 * rstep];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 h = hp[k4 + q * sstep];
                acc[0][q] = fmaf(w0.x, h.x, acc[0][q]); acc[1][q] = fmaf(w1.x, h.x, acc[1][q]);
                acc[2][q] = fmaf(w2.x, h.x, acc[2][q]); acc[3][q] = fmaf(w3.x, h.x, acc[3][q]);
                acc[0][q] = fmaf(w0.y, h.y, acc[0][q]); acc[1][q] = fmaf(w1.y, h.y, acc[1][q]);
                acc[2][q] = fmaf(w2.y, h.y, acc[2][q]); acc[3][q] = fmaf(w3.y, h.y, acc[3][q]);
                acc[0][q] = fmaf(w0.z, h.z, acc[0][q]); acc[1][q] = fmaf(w1.z, h.z, acc[1][q]);
                acc[2][q] = fmaf(w2.z, h.z, acc[2][q]); acc[3][q] = fmaf(w3.z, h.z, acc[3][q]);
                acc[0][q] = fmaf(w0.w, h.w, acc[0][q]); acc[1][q] = fmaf(w1.w, h.w, acc[1][q]);
                acc[2][q] = fmaf(w2.w, h.w, acc[2][q]); acc[3][q] = fmaf(w3.w, h.w, acc[3][q]);
            }
        }
        float *sp = stage + (ks * g.Spad + sq) * g.RP + rq;
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                sp[q * g.SQt * g.RP + i * g.RQt] = acc[i][q];
    }
}

// sum over the K splits of one staged GEMM result; sp points at split 0, splits are kstride floats apart
__device__ __forceinline__ float stage_sum(const float *sp, int KS, int kstride)
{
    float s = 0.0f;
    for (int ks = 0; ks < KS; ++ks) s += sp[ks * kstride];
    return s;
}

// ------------------------------------------------------------------------------------------------ forward
template <int NPAIR, int REC_NT>
__global__ void __launch_bounds__(REC_NT, 1) lstm_fwd_persistent_kernel(const RecFwdParams p)
{
    extern __shared__ __align__(16) float smem[];
    __shared__ unsigned long long s_tab[32];
    const RecGeom &g = p.g;
    const int tid = threadIdx.x, lane = tid & 31;
    const int nt = REC_NT / g.nsub, nw = nt >> 5;                 // threads / warps per sub-CTA
    const int sub = tid / nt, ts = tid - sub * nt, warp = ts >> 5;
    float *Wsm = smem;
    float *tile = Wsm + g.Rpad * g.RS + sub * g.Spad * g.RS;
    float *stage = Wsm + g.Rpad * g.RS + g.nsub * g.Spad * g.RS + sub * g.KS * g.Spad * g.RP;
    const int kstride = g.Spad * g.RP;

    const int H = p.H, L = p.L, S = p.S, T = p.T;
    const int Gc = g.G / g.nsub;                                   // sequence groups at CTA level
    const int d = blockIdx.x / (Gc * g.C);
    const int grp = ((blockIdx.x % (Gc * g.C)) / g.C) * g.nsub + sub;
    const int cs = blockIdx.x % g.C;
    const int j0 = cs * g.CL, ncell = min(g.CL, H - j0);
    const int s0 = grp * g.SG, nseq = min(g.SG, S - s0);
    unsigned *flag = p.flags + (d * g.G + grp) * 32;

    // one-time: zero all shared memory (padding rows/columns stay zero), then the weight slice and the exp table.
    for (int i = tid; i < (int)(g.smem / sizeof(float)); i += REC_NT) smem[i] = 0.0f;
    if (tid < 32) s_tab[tid] = bl_exp2f_tab[tid];
    __syncthreads();
    // Wsm row (gate*CL + cell) = column (d*H + j) of the gate's internal matrix: contiguous over the source cell k
    // (weight layout internal: g*L*H + d*H*H + j*H + k, LstmLayer.cu:586-596)
    for (int idx = tid; idx < 4 * ncell * H; idx += REC_NT) {
        const int k = idx % H, rc = idx / H;
        const int cl = rc % ncell, gi = rc / ncell;
        Wsm[(gi * g.CL + cl) * g.RS + k] = __ldg(p.Wi + (size_t)gi * L * H + (size_t)d * H * H + (size_t)(j0 + cl) * H + k);
    }

    // elementwise ownership: pair u of this thread = (cell cl, sequence sl), fixed for the whole pass
    bool valid[NPAIR]; int cl_[NPAIR], sl_[NPAIR];
    float wb[NPAIR][4], wpe[NPAIR][3], cprev[NPAIR];
#pragma unroll
    for (int u = 0; u < NPAIR; ++u) {
        const int pr = ts + u * nt;
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
    if (nseq <= 0) return;                                       // trailing empty group: nothing to do, no one waits on it

    for (int q = 0; q < T; ++q) {
        const int t = (d == 0) ? q : T - 1 - q;
        const bool first = (q == 0);
        const bool check = (t >= p.Tmin);
        float *acts_t = p.acts + (size_t)t * S * 4 * L + d * H + j0;           // + slot*4L + gate*L + cell
        float *cst_t = p.cst + (size_t)t * S * L + d * H + j0;
        float *y_t = p.Y + (size_t)t * S * p.ldy + d * H + j0;
        const char *pat_t = p.pat + (size_t)t * S;
        float *hx_w = p.hx + (size_t)(d * 2 + (q & 1)) * S * g.Hpad + j0;

        // prefetch what does not depend on h: the projected pre-activations and the pattern type
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

        long long *tr = p.trace ? p.trace + ((size_t)(blockIdx.x * g.nsub + sub) * T + q) * 6 : nullptr;
        if (tr && ts == 0) tr[0] = clock64();
        if (!first) {
            wait_counter(flag, (unsigned)(g.C * q), ts, sub, nt);
            if (tr && ts == 0) tr[1] = clock64();
            const float4 *src = reinterpret_cast<const float4 *>(p.hx + ((size_t)(d * 2 + ((q - 1) & 1)) * S + s0) * g.Hpad);
            float4 *dst = reinterpret_cast<float4 *>(tile);
            const int n4 = nseq * g.Hpad / 4;
            for (int i = ts; i < n4; i += nt) dst[i] = __ldcg(src + i);
            bar_sub(sub, nt);
            if (tr && ts == 0) tr[2] = clock64();
            smem_gemm(g, Wsm, tile, stage, warp, lane, nw);
            bar_sub(sub, nt);
            if (tr && ts == 0) tr[3] = clock64();
        }

#pragma unroll
        for (int u = 0; u < NPAIR; ++u) {
            if (!valid[u]) continue;
            const int slot = s0 + sl_[u], cl = cl_[u];
            float h, c;
            if (dummy[u]) {                                       // LstmLayer.cu:78-85
                h = 0.0f; c = 0.0f;
            } else {
                float ni = a[u][0], ig = a[u][1], fg = a[u][2], og = a[u][3];
                if (!first) {                                     // recurrent addProduct, :815-818
                    const float *sp = stage + sl_[u] * g.RP + cl;
                    ni = __fadd_rn(ni, stage_sum(sp, g.KS, kstride));
                    ig = __fadd_rn(ig, stage_sum(sp + g.CL, g.KS, kstride));
                    fg = __fadd_rn(fg, stage_sum(sp + 2 * g.CL, g.KS, kstride));
                    og = __fadd_rn(og, stage_sum(sp + 3 * g.CL, g.KS, kstride));
                }
                ni = __fadd_rn(ni, wb[u][0]); ig = __fadd_rn(ig, wb[u][1]);            // :97-100
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
                float *ap = acts_t + slot * 4 * L + cl;
                ap[0] = ni; ap[L] = ig; ap[2 * L] = fg; ap[3 * L] = og;
            }
            cprev[u] = c;
            cst_t[slot * L + cl] = c;
            y_t[slot * p.ldy + cl] = h;
            hx_w[slot * g.Hpad + cl] = h;
        }
        if (tr && ts == 0) tr[4] = clock64();
        if (q + 1 < T) publish(flag, ts, sub, nt);
        if (tr && ts == 0) tr[5] = clock64();
    }
}

// ------------------------------------------------------------------------------------------------ BPTT
template <int NPAIR, int REC_NT>
__global__ void __launch_bounds__(REC_NT, 1) lstm_bwd_persistent_kernel(const RecBwdParams p)
{
    extern __shared__ __align__(16) float smem[];
    __shared__ unsigned long long s_tab[32];
    const RecGeom &g = p.g;
    const int tid = threadIdx.x, lane = tid & 31;
    const int nt = REC_NT / g.nsub, nw = nt >> 5;
    const int sub = tid / nt, ts = tid - sub * nt, warp = ts >> 5;
    float *Wsm = smem;
    float *tile = Wsm + g.Rpad * g.RS + sub * g.Spad * g.RS;
    float *stage = Wsm + g.Rpad * g.RS + g.nsub * g.Spad * g.RS + sub * g.KS * g.Spad * g.RP;
    const int kstride = g.Spad * g.RP;

    const int H = p.H, L = p.L, S = p.S, T = p.T;
    const int Gc = g.G / g.nsub;
    const int d = blockIdx.x / (Gc * g.C);
    const int grp = ((blockIdx.x % (Gc * g.C)) / g.C) * g.nsub + sub;
    const int cs = blockIdx.x % g.C;
    const int j0 = cs * g.CL, ncell = min(g.CL, H - j0);
    const int s0 = grp * g.SG, nseq = min(g.SG, S - s0);
    unsigned *flag = p.flags + (d * g.G + grp) * 32;
    const bool inplace = (p.ndir == 1);

    for (int i = tid; i < (int)(g.smem / sizeof(float)); i += REC_NT) smem[i] = 0.0f;
    if (tid < 32) s_tab[tid] = bl_exp2f_tab[tid];
    __syncthreads();
    // Wsm row (cell k) = [gate][target cell j] : W_gate[k, j] = Wi[gate*L*H + d*H*H + j*H + k]  (the (N,N) products of :939-942)
    for (int idx = tid; idx < 4 * H * ncell; idx += REC_NT) {
        const int cl = idx % ncell, gj = idx / ncell;           // consecutive threads -> consecutive k: coalesced
        const int j = gj % H, gi = gj / H;
        Wsm[cl * g.RS + gi * g.Hpad + j] = __ldg(p.Wi + (size_t)gi * L * H + (size_t)d * H * H + (size_t)j * H + (j0 + cl));
    }

    bool valid[NPAIR]; int cl_[NPAIR], sl_[NPAIR];
    float wpe[NPAIR][3];
    float nfg[NPAIR], ncerr[NPAIR], ndig[NPAIR], ndfg[NPAIR];   // "next step" state, :253-256
#pragma unroll
    for (int u = 0; u < NPAIR; ++u) {
        const int pr = ts + u * nt;
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
    if (nseq <= 0) return;

    for (int q = 0; q < T; ++q) {
        // the fw direction walks time backwards, the bw direction forwards (:936, :970)
        const int t = (d == 0) ? T - 1 - q : q;
        const bool firstCall = (q == 0);
        const bool lastCall = (q == T - 1);                     // the direction's first timestep: no c_prev
        const bool check = (t >= p.Tmin);
        const int tprev = (d == 0) ? t - 1 : t + 1;             // previous step in the direction's own time order
        const float *acts_t = p.acts + (size_t)t * S * 4 * L + d * H + j0;
        const float *cst_t = p.cst + (size_t)t * S * L + d * H + j0;
        const float *cst_p = p.cst + (size_t)(lastCall ? t : tprev) * S * L + d * H + j0;
        float *dy_t = p.dY + (size_t)t * S * p.lddy + d * H + j0;
        float *del_t = p.deltas + (size_t)t * S * 4 * L + d * H + j0;
        float *cerr_t = p.cerr + (size_t)t * S * L + d * H + j0;
        const char *pat_t = p.pat + (size_t)t * S;
        float *dx_w = p.dx + (size_t)(d * 2 + (q & 1)) * S * g.RS + j0;

        float a[NPAIR][4], c[NPAIR], cp[NPAIR], oe[NPAIR]; bool dummy[NPAIR];
#pragma unroll
        for (int u = 0; u < NPAIR; ++u) {
            dummy[u] = false; cp[u] = 0.0f;
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
            wait_counter(flag, (unsigned)(g.C * q), ts, sub, nt);
            const float4 *src = reinterpret_cast<const float4 *>(p.dx + ((size_t)(d * 2 + ((q - 1) & 1)) * S + s0) * g.RS);
            float4 *dst = reinterpret_cast<float4 *>(tile);
            const int n4 = nseq * g.RS / 4;
            for (int i = ts; i < n4; i += nt) dst[i] = __ldcg(src + i);
            bar_sub(sub, nt);
            smem_gemm(g, Wsm, tile, stage, warp, lane, nw);
            bar_sub(sub, nt);
        }

#pragma unroll
        for (int u = 0; u < NPAIR; ++u) {
            if (!valid[u]) continue;
            const int slot = s0 + sl_[u], cl = cl_[u];
            float e = oe[u];
            if (!firstCall) e = __fadd_rn(e, stage_sum(stage + sl_[u] * g.RP + cl, g.KS, kstride));   // the 4 addProducts of :939-942
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
            float *dp = del_t + slot * 4 * L + cl;
            dp[0] = dni; dp[L] = dig; dp[2 * L] = dfg; dp[3 * L] = dog;
            cerr_t[slot * L + cl] = cerr;
            float *xp = dx_w + slot * g.RS + cl;
            xp[0] = dni; xp[g.Hpad] = dig; xp[2 * g.Hpad] = dfg; xp[3 * g.Hpad] = dog;
        }
        if (q + 1 < T) publish(flag, ts, sub, nt);
    }
}

// ------------------------------------------------------------------------------------------------ launch
template <typename Params, typename Kernel>
static int launch_persistent(bl_ctx *ctx, Kernel kernel, const Params &p, const char *name)
{
    const RecGeom &g = p.g;
    const int grid = p.ndir * (g.G / g.nsub) * g.C;
    BL_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem));
    int per_sm = 0;
    const int REC_NT = g.NT;
    BL_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, REC_NT, g.smem));
    if (per_sm < 1 || grid > per_sm * ctx->num_sms)
        return fail(ctx, "%s: %d CTAs cannot be co-resident (%d per SM x %d SMs)", name, grid, per_sm, ctx->num_sms);
    BL_CUDA(ctx, cudaMemsetAsync(p.flags, 0, (size_t)p.ndir * g.G * 32 * sizeof(unsigned), ctx->stream));
    void *args[] = { (void *)&p };
    BL_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(REC_NT), args, g.smem, ctx->stream));
    BL_LAUNCHED(ctx);
    return 0;
}

int launch_lstm_fwd(bl_ctx *ctx, const RecFwdParams &p)
{
    TimedRegion timed(ctx, 1);
    const char *nm = "lstm_fwd_persistent";
    if (p.g.NT == 512) return p.g.npair == 1 ? launch_persistent(ctx, lstm_fwd_persistent_kernel<1, 512>, p, nm) : launch_persistent(ctx, lstm_fwd_persistent_kernel<2, 512>, p, nm);
    if (p.g.NT == 768) return p.g.npair == 1 ? launch_persistent(ctx, lstm_fwd_persistent_kernel<1, 768>, p, nm) : launch_persistent(ctx, lstm_fwd_persistent_kernel<2, 768>, p, nm);
    return p.g.npair == 1 ? launch_persistent(ctx, lstm_fwd_persistent_kernel<1, 1024>, p, nm) : launch_persistent(ctx, lstm_fwd_persistent_kernel<2, 1024>, p, nm);
}
int launch_lstm_bwd(bl_ctx *ctx, const RecBwdParams &p)
{
    TimedRegion timed(ctx, 2);
    const char *nm = "lstm_bwd_persistent";
    if (p.g.NT == 512) return p.g.npair == 1 ? launch_persistent(ctx, lstm_bwd_persistent_kernel<1, 512>, p, nm) : launch_persistent(ctx, lstm_bwd_persistent_kernel<2, 512>, p, nm);
    if (p.g.NT == 768) return p.g.npair == 1 ? launch_persistent(ctx, lstm_bwd_persistent_kernel<1, 768>, p, nm) : launch_persistent(ctx, lstm_bwd_persistent_kernel<2, 768>, p, nm);
    return p.g.npair == 1 ? launch_persistent(ctx, lstm_bwd_persistent_kernel<1, 1024>, p, nm) : launch_persistent(ctx, lstm_bwd_persistent_kernel<2, 1024>, p, nm);
}

} // namespace bl
