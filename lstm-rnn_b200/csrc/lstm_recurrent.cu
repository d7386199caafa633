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

bool choose_geometry(bool bwd, int H, int S, int ndir, int num_sms, int smem_cap, int forceG, int forceNsub, RecGeom *out)
{
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
            if (cdiv(S, SG) != G) continue;                             // trailing groups would be empty
            if (CL * SG > REC_NPAIR * nt) continue;
            const int R = bwd ? CL : 4 * CL;
            const int RQ = cdiv(R, 4), SQ = cdiv(SG, 4);
            const int K = bwd ? 4 * Hpad : Hpad, K4 = K / 4;
            const int RS = bwd ? 4 * Hpad + 4 : Hpad;
            for (int LR = 1; LR <= 32; LR *= 2) {
                const int LS = 32 / LR;
                const int WR = cdiv(RQ, LR), WS = cdiv(SQ, LS);
                const int tasks = WR * WS;
                int KS = nw / tasks; if (KS < 1) KS = 1;
                while (KS > 1 && K4 / KS < 8) --KS;
                const int KB4 = cdiv(K4, KS);
                const int Rpad = 4 * WR * LR, Spad = 4 * WS * LS;
                const int RP = Rpad | 1;
                const size_t smem = ((size_t)Rpad * RS + (size_t)nsub * ((size_t)Spad * RS + (size_t)KS * Spad * RP)) * sizeof(float);
                if ((int)smem > smem_cap) continue;
                // cost model (cycles per step).  The sub-CTAs of an SM share its FFMA issue slots and shared-memory
                // bandwidth, so their GEMM phases add up; the latency chain of one group (counter round trip, exchange
                // copy, its own GEMM, gate math, publish) overlaps the other groups' GEMMs.
                const double passes = (double)cdiv(tasks * KS, nw);
                const double iters = (double)tasks * KS * KB4;                   // warp-iterations of one group, 64 FFMA + 8 LDS.128 each
                const double issue = iters * 76.0 / 4.0;
                const double wave  = iters * 4.0 * ((LR > 8 ? LR / 8 : 1) + (LS > 8 ? LS / 8 : 1));
                const double gemm1 = (issue > wave ? issue : wave);
                const double lat1  = passes * KB4 * 76.0 * 1.6;                  // one warp's serial path through its tasks
                const double copy  = (double)SG * K * 4.0 / 48.0 + 600.0;
                const double chain = 3500.0 + 10.0 * C + copy + (lat1 > gemm1 ? lat1 : gemm1) + 900.0;
                const double busy  = nsub * (gemm1 + 400.0);
                const double cost  = busy > chain ? busy : chain;
                if (!found || cost < best.cost) {
                    found = true;
                    best.G = G; best.C = C; best.CL = CL; best.SG = SG; best.R = R; best.nsub = nsub;
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
__device__ __forceinline__ void smem_gemm(const RecGeom &g, const float *__restrict__ Wsm, const float *__restrict__ tile,
                                          float *__restrict__ stage, int warp, int lane, int nw)
{
    const int tasks = g.WR * g.WS * g.KS;
    const int lr = lane >> g.LSlog, ls = lane & (g.LS - 1);
    for (int wt = warp; wt < tasks; wt += nw) {
        const int ks = wt % g.KS, tl = wt / g.KS;
        const int wr = tl % g.WR, ws = tl / g.WR;
        const int rq = wr * g.LR + lr, sq = ws * g.LS + ls;
        const int kb = ks * g.KB4, ke = min(g.K4, kb + g.KB4);
        const float4 *wp0 = reinterpret_cast<const float4 *>(Wsm + (size_t)(rq + g.RQt * 0) * g.RS);
        const float4 *wp1 = reinterpret_cast<const float4 *>(Wsm + (size_t)(rq + g.RQt * 1) * g.RS);
        const float4 *wp2 = reinterpret_cast<const float4 *>(Wsm + (size_t)(rq + g.RQt * 2) * g.RS);
        const float4 *wp3 = reinterpret_cast<const float4 *>(Wsm + (size_t)(rq + g.RQt * 3) * g.RS);
        const float4 *hp0 = reinterpret_cast<const float4 *>(tile + (size_t)(sq + g.SQt * 0) * g.RS);
        const float4 *hp1 = reinterpret_cast<const float4 *>(tile + (size_t)(sq + g.SQt * 1) * g.RS);
        const float4 *hp2 = reinterpret_cast<const float4 *>(tile + (size_t)(sq + g.SQt * 2) * g.RS);
        const float4 *hp3 = reinterpret_cast<const float4 *>(tile + (size_t)(sq + g.SQt * 3) * g.RS);
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[i][q] = 0.0f;
#pragma unroll 2
        for (int k4 = kb; k4 < ke; ++k4) {
            const float4 w[4] = {wp0[k4], wp1[k4], wp2[k4], wp3[k4]};
            const float4 h[4] = {hp0[k4], hp1[k4], hp2[k4], hp3[k4]};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    acc[i][q] = fmaf(w[i].x, h[q].x, acc[i][q]);
                    acc[i][q] = fmaf(w[i].y, h[q].y, acc[i][q]);
                    acc[i][q] = fmaf(w[i].z, h[q].z, acc[i][q]);
                    acc[i][q] = fmaf(w[i].w, h[q].w, acc[i][q]);
                }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                stage[((size_t)ks * g.Spad + sq + g.SQt * q) * g.RP + rq + g.RQt * i] = acc[i][q];
    }
}

__device__ __forceinline__ float stage_sum(const RecGeom &g, const float *stage, int sl, int row)
{
    float s = 0.0f;
    for (int ks = 0; ks < g.KS; ++ks) s += stage[((size_t)ks * g.Spad + sl) * g.RP + row];
    return s;
}

// ------------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(REC_NT, 1) lstm_fwd_persistent_kernel(const RecFwdParams p)
{
    extern __shared__ __align__(16) float smem[];
    const RecGeom &g = p.g;
    const int tid = threadIdx.x, lane = tid & 31;
    const int nt = REC_NT / g.nsub, nw = nt >> 5;                 // threads / warps per sub-CTA
    const int sub = tid / nt, ts = tid - sub * nt, warp = ts >> 5;
    float *Wsm = smem;
    float *tile = Wsm + (size_t)g.Rpad * g.RS + (size_t)sub * g.Spad * g.RS;
    float *stage = Wsm + (size_t)g.Rpad * g.RS + (size_t)g.nsub * g.Spad * g.RS + (size_t)sub * g.KS * g.Spad * g.RP;

    const int H = p.H, L = p.L, S = p.S, T = p.T;
    const int Gc = g.G / g.nsub;                                   // sequence groups at CTA level
    const int d = blockIdx.x / (Gc * g.C);
    const int grp = ((blockIdx.x % (Gc * g.C)) / g.C) * g.nsub + sub;
    const int cs = blockIdx.x % g.C;
    const int j0 = cs * g.CL, ncell = min(g.CL, H - j0);
    const int s0 = grp * g.SG, nseq = min(g.SG, S - s0);
    unsigned *flag = p.flags + (size_t)(d * g.G + grp) * 32;

    // one-time: zero both operand tiles (padding rows/columns stay zero), then the weight slice.
    for (int i = tid; i < (int)(g.smem / sizeof(float)); i += REC_NT) smem[i] = 0.0f;
    __syncthreads();
    // Wsm row (gate*CL + cell) = column (d*H + j) of the gate's internal matrix: contiguous over the source cell k
    // (weight layout internal: g*L*H + d*H*H + j*H + k, LstmLayer.cu:586-596)
    for (int idx = tid; idx < 4 * ncell * H; idx += REC_NT) {
        const int k = idx % H, rc = idx / H;
        const int cl = rc % ncell, gi = rc / ncell;
        Wsm[(size_t)(gi * g.CL + cl) * g.RS + k] =
            __ldg(p.Wi + (size_t)gi * L * H + (size_t)d * H * H + (size_t)(j0 + cl) * H + k);
    }

    // elementwise ownership: pair u of this thread = (cell cl, sequence sl), fixed for the whole pass
    bool valid[REC_NPAIR]; int cl_[REC_NPAIR], sl_[REC_NPAIR];
    float wb[REC_NPAIR][4], wpe[REC_NPAIR][3], cprev[REC_NPAIR];
#pragma unroll
    for (int u = 0; u < REC_NPAIR; ++u) {
        const int pr = ts + u * nt;
        cl_[u] = pr % g.CL; sl_[u] = pr / g.CL;
        valid[u] = (cl_[u] < ncell) && (sl_[u] < nseq);
        cprev[u] = 0.0f;
        if (valid[u]) {
            const int col = d * H + j0 + cl_[u];
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) wb[u][gi] = __ldg(p.Wb + gi * L + col);
#pragma unroll
            for (int gi = 0; gi < 3; ++gi) wpe[u][gi] = __ldg(p.Wp + gi * L + col);
        }
    }
    __syncthreads();
    if (nseq <= 0) return;                                       // (cannot happen with the planner's G; keeps named barriers consistent)

    for (int q = 0; q < T; ++q) {
        const int t = (d == 0) ? q : T - 1 - q;
        const bool first = (q == 0);
        const bool check = (t >= p.Tmin);

        // prefetch what does not depend on h: the projected pre-activations and the pattern type
        float a[REC_NPAIR][4]; bool dummy[REC_NPAIR];
#pragma unroll
        for (int u = 0; u < REC_NPAIR; ++u) {
            dummy[u] = false;
            if (valid[u]) {
                const size_t n = (size_t)t * S + s0 + sl_[u];
                const int col = d * H + j0 + cl_[u];
                dummy[u] = check && (p.pat[n] == BL_PATTYPE_NONE);
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) a[u][gi] = p.acts[n * 4 * L + gi * L + col];
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
        for (int u = 0; u < REC_NPAIR; ++u) {
            if (!valid[u]) continue;
            const size_t n = (size_t)t * S + s0 + sl_[u];
            const int col = d * H + j0 + cl_[u];
            float h, c;
            if (dummy[u]) {                                       // LstmLayer.cu:78-85
                h = 0.0f; c = 0.0f;
            } else {
                float ni = a[u][0], ig = a[u][1], fg = a[u][2], og = a[u][3];
                if (!first) {                                     // recurrent addProduct, :815-818
                    ni = __fadd_rn(ni, stage_sum(g, stage, sl_[u], 0 * g.CL + cl_[u]));
                    ig = __fadd_rn(ig, stage_sum(g, stage, sl_[u], 1 * g.CL + cl_[u]));
                    fg = __fadd_rn(fg, stage_sum(g, stage, sl_[u], 2 * g.CL + cl_[u]));
                    og = __fadd_rn(og, stage_sum(g, stage, sl_[u], 3 * g.CL + cl_[u]));
                }
                ni = __fadd_rn(ni, __fmul_rn(p.bias, wb[u][0]));  // :97-100
                ig = __fadd_rn(ig, __fmul_rn(p.bias, wb[u][1]));
                fg = __fadd_rn(fg, __fmul_rn(p.bias, wb[u][2]));
                og = __fadd_rn(og, __fmul_rn(p.bias, wb[u][3]));
                if (!first) {                                     // :103-108
                    ig = __fadd_rn(ig, __fmul_rn(cprev[u], wpe[u][0]));
                    fg = __fadd_rn(fg, __fmul_rn(cprev[u], wpe[u][1]));
                }
                ni = tanh_fn(ni); ig = logistic_fn(ig); fg = logistic_fn(fg);
                c = __fmul_rn(ni, ig);                            // :121-126
                if (!first) c = __fadd_rn(c, __fmul_rn(cprev[u], fg));
                og = __fadd_rn(og, __fmul_rn(c, wpe[u][2]));      // :129-131
                og = logistic_fn(og);
                h = __fmul_rn(tanh_fn(c), og);                    // :134
                float *ap = p.acts + n * 4 * L + col;
                ap[0] = ni; ap[L] = ig; ap[2 * L] = fg; ap[3 * L] = og;
            }
            cprev[u] = c;
            p.cst[n * L + col] = c;
            p.Y[n * p.ldy + col] = h;
            p.hx[((size_t)(d * 2 + (q & 1)) * S + s0 + sl_[u]) * g.Hpad + j0 + cl_[u]] = h;
        }
        if (tr && ts == 0) tr[4] = clock64();
        if (q + 1 < T) publish(flag, ts, sub, nt);
        if (tr && ts == 0) tr[5] = clock64();
    }
}

// ------------------------------------------------------------------------------------------------ BPTT
__global__ void __launch_bounds__(REC_NT, 1) lstm_bwd_persistent_kernel(const RecBwdParams p)
{
    extern __shared__ __align__(16) float smem[];
    const RecGeom &g = p.g;
    const int tid = threadIdx.x, lane = tid & 31;
    const int nt = REC_NT / g.nsub, nw = nt >> 5;                 // threads / warps per sub-CTA
    const int sub = tid / nt, ts = tid - sub * nt, warp = ts >> 5;
    float *Wsm = smem;
    float *tile = Wsm + (size_t)g.Rpad * g.RS + (size_t)sub * g.Spad * g.RS;
    float *stage = Wsm + (size_t)g.Rpad * g.RS + (size_t)g.nsub * g.Spad * g.RS + (size_t)sub * g.KS * g.Spad * g.RP;

    const int H = p.H, L = p.L, S = p.S, T = p.T;
    const int Gc = g.G / g.nsub;                                   // sequence groups at CTA level
    const int d = blockIdx.x / (Gc * g.C);
    const int grp = ((blockIdx.x % (Gc * g.C)) / g.C) * g.nsub + sub;
    const int cs = blockIdx.x % g.C;
    const int j0 = cs * g.CL, ncell = min(g.CL, H - j0);
    const int s0 = grp * g.SG, nseq = min(g.SG, S - s0);
    unsigned *flag = p.flags + (size_t)(d * g.G + grp) * 32;
    const bool inplace = (p.ndir == 1);

    for (int i = tid; i < (int)(g.smem / sizeof(float)); i += REC_NT) smem[i] = 0.0f;
    __syncthreads();
    // Wsm row (cell k) = [gate][target cell j] : W_gate[k, j] = Wi[gate*L*H + d*H*H + j*H + k]  (the (N,N) products of :939-942)
    for (int idx = tid; idx < 4 * H * ncell; idx += REC_NT) {
        const int cl = idx % ncell, gj = idx / ncell;           // consecutive threads -> consecutive k: coalesced
        const int j = gj % H, gi = gj / H;
        Wsm[(size_t)cl * g.RS + gi * g.Hpad + j] =
            __ldg(p.Wi + (size_t)gi * L * H + (size_t)d * H * H + (size_t)j * H + (j0 + cl));
    }

    bool valid[REC_NPAIR]; int cl_[REC_NPAIR], sl_[REC_NPAIR];
    float wpe[REC_NPAIR][3];
    float nfg[REC_NPAIR], ncerr[REC_NPAIR], ndig[REC_NPAIR], ndfg[REC_NPAIR];   // "next step" state, :253-256
#pragma unroll
    for (int u = 0; u < REC_NPAIR; ++u) {
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
    if (nseq <= 0) return;                                       // (cannot happen with the planner's G; keeps named barriers consistent)

    for (int q = 0; q < T; ++q) {
        // the fw direction walks time backwards, the bw direction forwards (:936, :970)
        const int t = (d == 0) ? T - 1 - q : q;
        const bool firstCall = (q == 0);
        const bool lastCall = (q == T - 1);                     // the direction's first timestep: no c_prev
        const bool check = (t >= p.Tmin);
        const int tprev = (d == 0) ? t - 1 : t + 1;             // previous step in the direction's own time order

        float a[REC_NPAIR][4], c[REC_NPAIR], cp[REC_NPAIR], oe[REC_NPAIR]; bool dummy[REC_NPAIR];
#pragma unroll
        for (int u = 0; u < REC_NPAIR; ++u) {
            dummy[u] = false; cp[u] = 0.0f;
            if (valid[u]) {
                const size_t n = (size_t)t * S + s0 + sl_[u];
                const int col = d * H + j0 + cl_[u];
                dummy[u] = check && (p.pat[n] == BL_PATTYPE_NONE);
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) a[u][gi] = p.acts[n * 4 * L + gi * L + col];
                c[u] = p.cst[n * L + col];
                if (!lastCall) cp[u] = p.cst[((size_t)tprev * S + s0 + sl_[u]) * L + col];
                oe[u] = p.dY[n * p.lddy + col];
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
        for (int u = 0; u < REC_NPAIR; ++u) {
            if (!valid[u]) continue;
            const size_t n = (size_t)t * S + s0 + sl_[u];
            const int col = d * H + j0 + cl_[u];
            float e = oe[u];
            if (!firstCall) e = __fadd_rn(e, stage_sum(g, stage, sl_[u], cl_[u]));     // the 4 addProducts of :939-942
            if (inplace) p.dY[n * p.lddy + col] = e;            // unidirectional: tmpOutputErrors IS outputErrors (:907-910)
            float dni, dig, dfg, dog, cerr;
            if (dummy[u]) {                                       // :224-234
                dni = dig = dfg = dog = cerr = 0.0f;
                nfg[u] = 0.0f;
            } else {
                const float ni = a[u][0], ig = a[u][1], fg = a[u][2], og = a[u][3];
                const float tc = tanh_fn(c[u]);
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
            float *dp = p.deltas + n * 4 * L + col;
            dp[0] = dni; dp[L] = dig; dp[2 * L] = dfg; dp[3 * L] = dog;
            p.cerr[n * L + col] = cerr;
            float *xp = p.dx + ((size_t)(d * 2 + (q & 1)) * S + s0 + sl_[u]) * g.RS + j0 + cl_[u];
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
    BL_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, REC_NT, g.smem));
    if (per_sm < 1 || grid > per_sm * ctx->num_sms)
        return fail(ctx, "%s: %d CTAs cannot be co-resident (%d per SM x %d SMs)", name, grid, per_sm, ctx->num_sms);
    BL_CUDA(ctx, cudaMemsetAsync(p.flags, 0, (size_t)p.ndir * g.G * 32 * sizeof(unsigned), ctx->stream));
    void *args[] = { (void *)&p };
    BL_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)kernel, dim3(grid), dim3(REC_NT), args, g.smem, ctx->stream));
    BL_LAUNCHED(ctx);
    return 0;
}

int launch_lstm_fwd(bl_ctx *ctx, const RecFwdParams &p) { TimedRegion timed(ctx, 1); return launch_persistent(ctx, lstm_fwd_persistent_kernel, p, "lstm_fwd_persistent"); }
int launch_lstm_bwd(bl_ctx *ctx, const RecBwdParams &p) { TimedRegion timed(ctx, 2); return launch_persistent(ctx, lstm_bwd_persistent_kernel, p, "lstm_bwd_persistent"); }

} // namespace bl
