// LSTM / BLSTM layer passes behind the C ABI (bl_lstm_*): orchestration of the projection GEMM, the persistent
// recurrent kernels and the gradient contractions.  Reference: layers/LstmLayer.cu:763-886 (forward), :888-1051 (backward).
//
// HBM layout owned by the plan (all fp32):
//   acts   [maxT*S][4L]  column g*L + d*H + j : projection output, overwritten in place by the activated gates
//                         (the reference keeps 8 separate [N][H] arrays; one [N][4L] matrix lets ONE GEMM with
//                          M = 4L produce all gates of both directions straight from the unmodified weight vector,
//                          whose input block is exactly the [P x 4L] column-major matrix of these columns)
//   deltas [maxT*S][4L]  same columns: the clipped gate deltas
//   cst    [maxT*S][L]   cell states, column d*H + j;  cerr [maxT*S][L] cell-state errors
//   hx / dx              L2-resident step-parity exchange buffers of the persistent kernels
// The layer output Y [N][L] is written directly by the recurrent kernel (no ResortOutputsFn pass, :140-161),
// and the output errors dY [N][L] are read directly (no ResortOutputErrorsFn pass, :163-188).
#include "lstm_recurrent.cuh"
#include "gemm_tc.cuh"
#include <algorithm>
#include <cstdint>
#include <cstdlib>

namespace bl {
int gemm_f32_simt(bl_ctx *ctx, int transA, int transB, int m, int n, int k,
                  const float *A, int lda, const float *B, int ldb, float *C, int ldc, int accumulate);
}

struct bl_lstm_plan {
    bl_ctx *ctx;
    int P, L, H, ndir, S, maxT;
    float bias;
    bl::RecGeom gf, gb;
    bool reg_f, reg_b;       // register-resident persistent kernels (lstm_recurrent_reg.cu) vs shared-memory ones
    bool tm_f, tm_b;         // tensor-memory-resident weights + tcgen05 step GEMM (lstm_recurrent_tmem.cu)
    bool t2_f, t2_b;         // second tensor-memory generation: in-band exchange, fp16 operands (lstm_recurrent_tm2.cu)
    float *acts, *deltas, *cst, *cerr, *hx, *dx, *gpart;
    long long *trace, *trace_b;
    bool no_fused_split;     // BLSTM_NO_FUSED_SPLIT, read once at plan creation
    float *tcbuf;            // prepared tensor-core operands (hi then lo of each): X, Win (forward); deltas, Y, Win re-blocked (backward)
    size_t tcbuf_elems;
    float *tc_X, *tc_Wf, *tc_D, *tc_Y, *tc_Wb;
    size_t tc_eX, tc_eWf, tc_eD, tc_eY, tc_eWb;
    bl::TcOperand xsplit;    // strict split of the forward pass's X, reused by the backward pass of the same fraction
    const float *xsplit_src; int xsplit_ld, xsplit_T;
    int ysplit_T;            // T of the fraction whose Y split the forward kernel wrote into tc_Y (0 = none)
    unsigned *flags_f, *flags_b;
    int gsplit;
    int lastT;
};

namespace bl {

// Bias and peephole gradients (the non-GEMM part of ComputeWeightUpdateFn, LstmLayer.cu:392-408, 440-475).
// grid (ceil(L/32), nsplit), block (32, 8): column = d*H + j of the layer, rows = a slice of the patterns.
// part[split][7][L]: 4 bias sums (bias*delta), IG / FG peephole (time-shifted cell state), OG peephole (unshifted).
__global__ void lstm_small_grads_kernel(int N, int S, int H, int L, float bias, const float *__restrict__ deltas,
                                        const float *__restrict__ cst, float *__restrict__ part, int rows_per_split)
{
    __shared__ float red[8][7][33];
    const int col = blockIdx.x * 32 + threadIdx.x;
    const int n0 = blockIdx.y * rows_per_split, n1 = min(N, n0 + rows_per_split);
    float acc[7] = {0, 0, 0, 0, 0, 0, 0};
    if (col < L) {
        const int d = col / H;
        for (int n = n0 + threadIdx.y; n < n1; n += 8) {
            const float *dp = deltas + (size_t)n * 4 * L + col;
            const float dni = dp[0], dig = dp[L], dfg = dp[2 * L], dog = dp[3 * L];
            acc[0] += bias * dni; acc[1] += bias * dig; acc[2] += bias * dfg; acc[3] += bias * dog;
            acc[6] += cst[(size_t)n * L + col] * dog;                               // OG: no time shift (:456-458)
            // IG/FG: fw pairs delta[n] with c[n-S] for n >= S; bw pairs delta[n] with c[n+S] for n < N-S (:460-468, 493-500)
            const int ns = (d == 0) ? n - S : n + S;
            if (ns >= 0 && ns < N) {
                const float cs = cst[(size_t)ns * L + col];
                acc[4] += cs * dig; acc[5] += cs * dfg;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 7; ++i) red[threadIdx.y][i][threadIdx.x] = acc[i];
    __syncthreads();
    if (threadIdx.y == 0 && col < L) {
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            float s = 0.0f;
            for (int y = 0; y < 8; ++y) s += red[y][i][threadIdx.x];
            part[((size_t)blockIdx.y * 7 + i) * L + col] = s;
        }
    }
}

// bias_scale: 1 when the partials already carry the bias factor (lstm_small_grads_kernel), the layer's bias when they are plain delta
// sums (accumulated inside lstm_bwd_tm2_kernel)
__global__ void lstm_small_grads_finish_kernel(int L, int nsplit, float bias_scale, const float *__restrict__ part,
                                               float *__restrict__ dWbias, float *__restrict__ dWpeep)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 7 * L) return;
    const int i = idx / L, col = idx % L;
    float s = 0.0f;
    for (int z = 0; z < nsplit; ++z) s += part[((size_t)z * 7 + i) * L + col];
    if (i < 4) dWbias[i * L + col] = bias_scale * s;          // bias block: g*L + d*H + j
    else       dWpeep[(i - 4) * L + col] = s;    // peephole block: q*L + d*H + j
}

__global__ void gather_cols_kernel(int N, int H, const float *__restrict__ src, int ld, int col0, float *__restrict__ dst)
{
    const size_t total = (size_t)N * H;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x)
        dst[e] = src[(e / H) * ld + col0 + (e % H)];
}

} // namespace bl

extern "C" {

size_t bl_lstm_num_weights(int P, int L, int bidirectional)
{
    return (size_t)L * (4 * ((size_t)P + 1) + (bidirectional ? 2 : 4) * (size_t)L + 3);
}

int bl_lstm_plan_create(bl_ctx *ctx, int P, int L, int bidirectional, int S, int maxT, float bias, bl_lstm_plan **out)
{
    if (!ctx) return bl::fail(nullptr, "bl_lstm_plan_create: ctx is NULL");
    if (!out) return bl::fail(ctx, "bl_lstm_plan_create: out is NULL");
    *out = nullptr;
    if (P <= 0 || L <= 0 || S <= 0 || maxT <= 0) return bl::fail(ctx, "bl_lstm_plan_create: bad sizes P=%d L=%d S=%d maxT=%d", P, L, S, maxT);
    if (bidirectional && (L % 2)) return bl::fail(ctx, "Cannot create a bidirectional layer with an odd layer size");   // LstmLayer.cu:528-529
    bl_lstm_plan *pl = new bl_lstm_plan();
    pl->ctx = ctx; pl->P = P; pl->L = L; pl->ndir = bidirectional ? 2 : 1; pl->H = L / pl->ndir;
    pl->S = S; pl->maxT = maxT; pl->bias = bias; pl->lastT = 0;
    pl->acts = pl->deltas = pl->cst = pl->cerr = pl->hx = pl->dx = pl->gpart = nullptr;
    pl->flags_f = pl->flags_b = nullptr;
    pl->trace = pl->trace_b = nullptr;
    pl->no_fused_split = getenv("BLSTM_NO_FUSED_SPLIT") != nullptr;
    pl->tcbuf = nullptr; pl->tcbuf_elems = 0;
    pl->xsplit_src = nullptr; pl->xsplit_ld = 0; pl->xsplit_T = 0; pl->ysplit_T = 0;

    const int cap = ctx->smem_optin - 2048;      // room for the kernels' static shared memory (exp table) and the driver's reserve
    // tuning overrides (tools/sweep_geometry.py): sequence groups / sub-CTAs / threads of either kernel, and the kernel family
    // (default: tensor-memory-resident weights + tcgen05 step GEMM where the slice fits the 512 TMEM columns, else register-resident
    // weights, else shared memory; BLSTM_REC_V=1 skips the tensor-memory kernels, BLSTM_REC_V=2 forces the shared-memory ones)
    const char *ef = getenv("BLSTM_FWD_G"), *eb = getenv("BLSTM_BWD_G"), *nf = getenv("BLSTM_FWD_NSUB"), *nb = getenv("BLSTM_BWD_NSUB"),
               *tf = getenv("BLSTM_FWD_NT"), *tb = getenv("BLSTM_BWD_NT"), *ev = getenv("BLSTM_REC_V");
    // kernel family: 0 (default) = best that fits: tm2, else tensor-memory generation 1, else register-resident, else shared memory;
    // 1 = register-resident first, 2 = shared memory only, 3 = tensor-memory generation 1 first, 4 = tm2 first (same as 0)
    const int family = ev ? atoi(ev) : 0;
    const bool want_reg = family != 2, want_tmem = family == 0 || family == 3 || family == 4, want_t2 = family == 0 || family == 4;
    const int fG = ef ? atoi(ef) : 0, bG = eb ? atoi(eb) : 0;
    pl->t2_f = want_t2 && bl::choose_geometry_tm2(false, pl->H, S, pl->ndir, ctx->num_sms, ctx->smem_optin - 1024, fG, &pl->gf);
    pl->tm_f = !pl->t2_f && want_tmem && bl::choose_geometry_tmem(pl->H, S, pl->ndir, ctx->num_sms, cap, fG, &pl->gf);
    pl->reg_f = !pl->t2_f && !pl->tm_f && want_reg && bl::choose_geometry_reg(false, pl->H, S, pl->ndir, ctx->num_sms, cap, fG, &pl->gf);
    pl->t2_b = want_t2 && bl::choose_geometry_tm2(true, pl->H, S, pl->ndir, ctx->num_sms, ctx->smem_optin - 1024, bG, &pl->gb);
    pl->tm_b = !pl->t2_b && want_tmem && bl::choose_geometry_tmem_bwd(pl->H, S, pl->ndir, ctx->num_sms, cap, bG, &pl->gb);
    pl->reg_b = !pl->t2_b && !pl->tm_b && want_reg && bl::choose_geometry_reg(true, pl->H, S, pl->ndir, ctx->num_sms, cap, bG, &pl->gb);
    if ((!pl->t2_f && !pl->tm_f && !pl->reg_f && !bl::choose_geometry(false, pl->H, S, pl->ndir, ctx->num_sms, cap, fG, nf ? atoi(nf) : 0, tf ? atoi(tf) : 0, &pl->gf)) ||
        (!pl->t2_b && !pl->tm_b && !pl->reg_b && !bl::choose_geometry(true, pl->H, S, pl->ndir, ctx->num_sms, cap, bG, nb ? atoi(nb) : 0, tb ? atoi(tb) : 0, &pl->gb))) {
        delete pl;
        return bl::fail(ctx, "bl_lstm_plan_create: no persistent-kernel geometry fits (H=%d S=%d): recurrent weights do not fit on chip",
                        L / (bidirectional ? 2 : 1), S);
    }
    const size_t N = (size_t)maxT * S;
    const size_t hx_elems = pl->t2_f ? pl->gf.xelems : (size_t)pl->ndir * 2 * S * pl->gf.RS;
    const size_t dx_elems = pl->t2_b ? pl->gb.xelems : (size_t)pl->ndir * 2 * S * pl->gb.RS;
    pl->gsplit = std::max(64, pl->gb.G);
    int rc = 0;
    rc |= bl_malloc(ctx, (void **)&pl->acts, N * 4 * L * sizeof(float));
    rc |= bl_malloc(ctx, (void **)&pl->deltas, N * 4 * L * sizeof(float));
    rc |= bl_malloc(ctx, (void **)&pl->cst, N * L * sizeof(float));
    rc |= bl_malloc(ctx, (void **)&pl->cerr, N * L * sizeof(float));
    rc |= bl_malloc(ctx, (void **)&pl->hx, hx_elems * sizeof(float));
    rc |= bl_malloc(ctx, (void **)&pl->dx, dx_elems * sizeof(float));
    rc |= bl_malloc(ctx, (void **)&pl->gpart, (size_t)pl->gsplit * 7 * L * sizeof(float));
    rc |= bl_malloc(ctx, (void **)&pl->flags_f, (size_t)pl->ndir * pl->gf.G * 32 * sizeof(unsigned));
    rc |= bl_malloc(ctx, (void **)&pl->flags_b, (size_t)pl->ndir * pl->gb.G * 32 * sizeof(unsigned));
    if (getenv("BLSTM_REC_TRACE")) {
        rc |= bl_malloc(ctx, (void **)&pl->trace, (size_t)ctx->num_sms * 4 * maxT * 8 * sizeof(long long));
        rc |= bl_malloc(ctx, (void **)&pl->trace_b, (size_t)ctx->num_sms * 2 * maxT * 8 * sizeof(long long));
    }
    if (rc) { bl_lstm_plan_destroy(pl); return 1; }
    // zero-filled like the reference's buffers (LstmLayer.cu:554); the exchange buffers' padding columns must stay zero
    rc |= bl_memset(ctx, pl->acts, 0, N * 4 * L * sizeof(float));
    rc |= bl_memset(ctx, pl->deltas, 0, N * 4 * L * sizeof(float));
    rc |= bl_memset(ctx, pl->cst, 0, N * L * sizeof(float));
    rc |= bl_memset(ctx, pl->cerr, 0, N * L * sizeof(float));
    rc |= bl_memset(ctx, pl->hx, 0, hx_elems * sizeof(float));
    rc |= bl_memset(ctx, pl->dx, 0, dx_elems * sizeof(float));
    if (rc) { bl_lstm_plan_destroy(pl); return 1; }
    *out = pl;
    return 0;
}

void bl_lstm_plan_destroy(bl_lstm_plan *pl)
{
    if (!pl) return;
    cudaStreamSynchronize(pl->ctx->stream);
    void *bufs[] = { pl->acts, pl->deltas, pl->cst, pl->cerr, pl->hx, pl->dx, pl->gpart, pl->flags_f, pl->flags_b, pl->trace, pl->trace_b, pl->tcbuf };
    for (void *b : bufs) if (b) cudaFree(b);
    delete pl;
}

int bl_lstm_tm2_geometry(int backward, int H, int S, int ndir, int num_sms, int smem_cap, long long *o)
{
    bl::RecGeom g;
    if (H <= 0 || S <= 0 || ndir < 1 || ndir > 2 || !bl::choose_geometry_tm2(backward != 0, H, S, ndir, num_sms, smem_cap, 0, &g)) return 0;
    o[0] = g.G; o[1] = g.C; o[2] = g.CL; o[3] = g.SG; o[4] = g.NT; o[5] = g.nsub; o[6] = g.Hpad; o[7] = g.Spad; o[8] = g.K4;
    o[9] = (long long)g.smem; o[10] = (long long)ndir * bl::cdiv(g.G, g.nsub) * g.C; o[11] = (long long)g.xelems;
    return 1;
}

int bl_lstm_plan_info(const bl_lstm_plan *pl, int *o)
{
    // smem is a multiple of 4: the low two bits carry nsub (1, 2, 4 -> 1, 2, 0), or 3 for the register-resident kernels
    // (the tensor-memory kernels report their smem rounded up to 16 with 11 (generation 1) or 13 / 14 (tm2 with one / two sub-groups
    // per CTA) in the low four bits)
    o[0] = pl->gf.G; o[1] = pl->gf.C; o[2] = pl->gf.CL;
    o[3] = pl->t2_f ? (int)((pl->gf.smem + 15) & ~(size_t)15) + 12 + pl->gf.nsub : pl->tm_f ? (int)((pl->gf.smem + 15) & ~(size_t)15) + 11 : (int)pl->gf.smem + (pl->reg_f ? 3 : pl->gf.nsub);
    o[4] = pl->gb.G; o[5] = pl->gb.C; o[6] = pl->gb.CL;
    o[7] = pl->t2_b ? (int)((pl->gb.smem + 15) & ~(size_t)15) + 12 + pl->gb.nsub : pl->tm_b ? (int)((pl->gb.smem + 15) & ~(size_t)15) + 11 : (int)pl->gb.smem + (pl->reg_b ? 3 : pl->gb.nsub);
    return 0;
}

// lazily allocates the plan's prepared-operand buffers (only layers large enough for the tensor-core path need them)
static int plan_tc_buffers(bl_lstm_plan *pl)
{
    if (pl->tcbuf) return 0;
    bl_ctx *ctx = pl->ctx;
    const int P = pl->P, L = pl->L, H = pl->H, Hp = (H + 3) & ~3, nb = 4 * pl->ndir;
    const int maxN = pl->maxT * pl->S;
    pl->tc_eX = bl::tc_operand_elems(maxN, P);
    pl->tc_eWf = bl::tc_operand_elems(4 * L, P);
    pl->tc_eD = bl::tc_operand_elems(maxN, 4 * L, 0, 0, H, Hp);
    pl->tc_eY = bl::tc_operand_elems(maxN, L, 0, 0, H, Hp);
    pl->tc_eWb = bl::tc_operand_elems(4 * L, P, H, Hp, 0, 0);
    (void)nb;
    const size_t need = 2 * (pl->tc_eX + pl->tc_eWf + pl->tc_eD + pl->tc_eY + pl->tc_eWb) + 16;
    BL_CUDA(ctx, cudaMalloc(&pl->tcbuf, need * sizeof(float)));
    // the re-blocking pad columns of the splits the recurrent kernels write directly are never touched again: zero them once
    BL_CUDA(ctx, cudaMemsetAsync(pl->tcbuf, 0, need * sizeof(float), ctx->stream));
    pl->tcbuf_elems = need;
    float *b = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(pl->tcbuf) + 15) & ~(uintptr_t)15);
    pl->tc_X = b; b += 2 * pl->tc_eX;
    pl->tc_Wf = b; b += 2 * pl->tc_eWf;
    pl->tc_D = b; b += 2 * pl->tc_eD;
    pl->tc_Y = b; b += 2 * pl->tc_eY;
    pl->tc_Wb = b;
    return 0;
}

int bl_lstm_forward(bl_lstm_plan *pl, const float *W, const float *X, int ldx, const char *patTypes,
                    int T, int Tmin, float *Y, int ldy)
{
    bl_ctx *ctx = pl->ctx;
    if (T <= 0 || T > pl->maxT) return bl::fail(ctx, "bl_lstm_forward: T=%d outside (0, maxT=%d]", T, pl->maxT);
    if (ldx < pl->P || ldy < pl->L) return bl::fail(ctx, "bl_lstm_forward: leading dimension too small");
    const int P = pl->P, L = pl->L, H = pl->H, S = pl->S, N = T * S;
    // (1) input projection of all timesteps, all 4 gates, both directions: acts[4L x N] = Win^T[4L x P] * X[P x N]
    //     (the 8 assignProduct calls of LstmLayer.cu:774-784 as one GEMM; Win is the weight vector's first 4LP floats)
    pl->xsplit_src = nullptr;
    if (bl::tc_wanted(ctx, 4 * L, N, P)) {
        // acts[N][4L] = X[N][P] * Win[4L][P]^T, both operands K-major; the strict split of X stays in the plan for the backward pass
        const bool strict = (ctx->gemm_mode == BL_GEMM_STRICT);
        BL_CHECK(plan_tc_buffers(pl));
        bl::TcOperand Xs, Ws;
        BL_CHECK(bl::tc_prepare(ctx, X, N, P, ldx, strict, pl->tc_X, pl->tc_X + pl->tc_eX, &Xs));
        BL_CHECK(bl::tc_prepare(ctx, W, 4 * L, P, P, strict, pl->tc_Wf, pl->tc_Wf + pl->tc_eWf, &Ws));
        BL_CHECK(bl::tc_gemm(ctx, N, 4 * L, P, bl::TcView{&Xs, false, 0, 0}, bl::TcView{&Ws, false, 0, 0}, pl->acts, 4 * L, 0));
        if (strict) { pl->xsplit = Xs; pl->xsplit_src = X; pl->xsplit_ld = ldx; pl->xsplit_T = T; }
    } else {
        BL_CHECK(bl_gemm_f32(ctx, 1, 0, 4 * L, N, P, W, P, X, ldx, pl->acts, 4 * L, 0, ctx->gemm_mode));
    }
    // (2) recurrent sweep, both directions concurrently
    bl::RecFwdParams p;
    p.Wb = W + (size_t)4 * L * P; p.Wi = p.Wb + 4 * L; p.Wp = p.Wi + (size_t)4 * L * H;
    p.acts = pl->acts; p.cst = pl->cst; p.Y = Y; p.ldy = ldy; p.hx = pl->hx; p.flags = pl->flags_f; p.pat = patTypes;
    p.T = T; p.Tmin = Tmin; p.S = S; p.H = H; p.L = L; p.ndir = pl->ndir; p.bias = pl->bias; p.g = pl->gf; p.trace = pl->trace;
    // the tensor-core backward pass wants the TF32 split of Y: the register-resident kernel writes it while it stores Y
    p.ys_hi = p.ys_lo = nullptr; p.ld_ys = 0;
    pl->ysplit_T = 0;
    if ((pl->reg_f || pl->tm_f || pl->t2_f) && bl::tc_wanted(ctx, P, 4 * L, N) && !pl->no_fused_split) {
        BL_CHECK(plan_tc_buffers(pl));
        p.ys_hi = pl->tc_Y; p.ys_lo = pl->tc_Y + pl->tc_eY; p.ld_ys = (int)bl::tc_operand_ld(pl->ndir * ((H + 3) & ~3));
        pl->ysplit_T = T;
    }
    BL_CHECK(pl->t2_f ? bl::launch_lstm_fwd_tm2(ctx, p) : pl->tm_f ? bl::launch_lstm_fwd_tmem(ctx, p) : pl->reg_f ? bl::launch_lstm_fwd_reg(ctx, p) : bl::launch_lstm_fwd(ctx, p));
    pl->lastT = T;
    return 0;
}

int bl_lstm_backward(bl_lstm_plan *pl, const float *W, const float *X, int ldx, const float *Y, int ldy,
                     float *dY, int lddy, const char *patTypes, int T, int Tmin,
                     float *dX, int lddx, float *dW)
{
    bl_ctx *ctx = pl->ctx;
    if (T <= 0 || T > pl->maxT) return bl::fail(ctx, "bl_lstm_backward: T=%d outside (0, maxT=%d]", T, pl->maxT);
    if (T != pl->lastT) return bl::fail(ctx, "bl_lstm_backward: must follow bl_lstm_forward on the same fraction (T=%d, forward saw %d)", T, pl->lastT);
    if (ldx < pl->P || ldy < pl->L || lddy < pl->L || (dX && lddx < pl->P)) return bl::fail(ctx, "bl_lstm_backward: leading dimension too small");
    const int P = pl->P, L = pl->L, H = pl->H, S = pl->S, N = T * S;
    const size_t inW = (size_t)L * P;
    float *dWbias = dW + 4 * inW, *dWint = dWbias + 4 * L, *dWpeep = dWint + (size_t)4 * L * H;

    // (1) BPTT sweep
    bl::RecBwdParams p;
    p.Wi = W + 4 * inW + 4 * L; p.Wp = p.Wi + (size_t)4 * L * H;
    p.acts = pl->acts; p.cst = pl->cst; p.deltas = pl->deltas; p.cerr = pl->cerr; p.dY = dY; p.lddy = lddy;
    p.dx = pl->dx; p.flags = pl->flags_b; p.pat = patTypes;
    p.T = T; p.Tmin = Tmin; p.S = S; p.H = H; p.L = L; p.ndir = pl->ndir; p.g = pl->gb;
    const bool tc = bl::tc_wanted(ctx, P, 4 * L, N);
    const bool fused_dsplit = tc && (pl->reg_b || pl->tm_b || pl->t2_b) && !pl->no_fused_split;
    p.ds_hi = p.ds_lo = nullptr; p.ld_ds = 0; p.trace = pl->trace_b;
    p.gpart = pl->t2_b ? pl->gpart : nullptr;       // the tm2 BPTT kernel accumulates the bias / peephole gradient sums itself
    if (fused_dsplit) {
        BL_CHECK(plan_tc_buffers(pl));
        p.ds_hi = pl->tc_D; p.ds_lo = pl->tc_D + pl->tc_eD; p.ld_ds = (int)bl::tc_operand_ld(4 * pl->ndir * ((H + 3) & ~3));
    }
    BL_CHECK(pl->t2_b ? bl::launch_lstm_bwd_tm2(ctx, p) : pl->tm_b ? bl::launch_lstm_bwd_tmem(ctx, p) : pl->reg_b ? bl::launch_lstm_bwd_reg(ctx, p) : bl::launch_lstm_bwd(ctx, p));

    if (tc) {
        // ---- tensor-core path: every operand is split (hi/lo TF32) ONCE per layer, in its own row-major layout, and read
        // either K-major or MN-major by the GEMMs -- nothing is transposed in memory.  The (gate, direction) blocks of H
        // cells are re-pitched to Hp = roundup(H, 4) so that block sub-views start on 16-byte boundaries:
        //   D' = deltas [N][4*ndir*Hp]     K-major  -> input error;   MN-major -> input- and recurrent-weight gradients
        //   X' = X      [N][P]             MN-major -> input-weight gradient (the forward pass's split is reused)
        //   Y' = Y      [N][ndir*Hp]       MN-major -> recurrent-weight gradients (time shift = row offset of the view)
        //   W' = Win    [4*ndir*Hp][P]     MN-major -> input error
        const bool strict = true;            // gradients always use the strict (3xTF32) mode
        const int nb = 4 * pl->ndir, Hp = (H + 3) & ~3;
        BL_CHECK(plan_tc_buffers(pl));
        bl::TcOperand D, Xs, Ys, Ws;
        if (fused_dsplit) {       // written by the BPTT kernel
            D.hi = pl->tc_D; D.lo = pl->tc_D + pl->tc_eD; D.ld = (size_t)p.ld_ds; D.rows = N; D.cols = nb * Hp; D.strict = true;
        } else
            BL_CHECK(bl::tc_prepare(ctx, pl->deltas, N, 4 * L, 4 * L, strict, pl->tc_D, pl->tc_D + pl->tc_eD, &D, 0, 0, H, Hp));
        if (pl->xsplit_src == X && pl->xsplit_ld == ldx && pl->xsplit_T == T) Xs = pl->xsplit;
        else BL_CHECK(bl::tc_prepare(ctx, X, N, P, ldx, strict, pl->tc_X, pl->tc_X + pl->tc_eX, &Xs));
        if (pl->ysplit_T == T) {  // written by the forward kernel of this fraction
            Ys.hi = pl->tc_Y; Ys.lo = pl->tc_Y + pl->tc_eY; Ys.ld = bl::tc_operand_ld(pl->ndir * Hp); Ys.rows = N; Ys.cols = pl->ndir * Hp; Ys.strict = true;
        } else
            BL_CHECK(bl::tc_prepare(ctx, Y, N, L, ldy, strict, pl->tc_Y, pl->tc_Y + pl->tc_eY, &Ys, 0, 0, H, Hp));
        // (2) error to the preceding layer: dX[N][P] = deltas[N][4L] * Win[4L][P]   (the 8 products of :996-1006)
        if (dX) {
            BL_CHECK(bl::tc_prepare(ctx, W, 4 * L, P, P, strict, pl->tc_Wb, pl->tc_Wb + pl->tc_eWb, &Ws, H, Hp, 0, 0));
            BL_CHECK(bl::tc_gemm(ctx, N, P, nb * Hp, bl::TcView{&D, false, 0, 0}, bl::TcView{&Ws, true, 0, 0}, dX, lddx, 0));
        }
        // (3) input weight gradients: dWin[4L][P] = deltas^T * X, one [H x P] row block per (gate, direction)
        //     (ComputeWeightUpdateFn case 0x0, :372-389)
        BL_CHECK(bl::tc_gemm_batched(ctx, H, P, N, bl::TcView{&D, true, 0, 0}, bl::TcView{&Xs, true, 0, 0}, dW, P, 0,
                                     /*batches=*/nb, /*a_batch_mn=*/Hp, /*c_batch_stride=*/(long long)H * P));
        // (4) recurrent weight gradients, one [H x H] block per (gate, direction) (case 0x8, :411-437, 493-500):
        //     fw: dW[j][k] = sum_{n>=S} delta[n,j] * h[n-S,k];   bw: dW[j][k] = sum_{n<N-S} delta[n,j] * h[n+S,k]
        //     one launch per direction: its 4 gate blocks are column sub-views of D' (ndir*Hp apart) sharing the same shifted Y'
        if (N - S > 0) {
            for (int d = 0; d < pl->ndir; ++d)
                BL_CHECK(bl::tc_gemm_batched(ctx, H, H, N - S, bl::TcView{&D, true, d * Hp, d == 0 ? S : 0}, bl::TcView{&Ys, true, d * Hp, d == 0 ? 0 : S},
                                             dWint + (size_t)d * H * H, H, 0, /*batches=*/4, /*a_batch_mn=*/pl->ndir * Hp, /*c_batch_stride=*/(long long)L * H));
        } else {
            BL_CHECK(bl_memset(ctx, dWint, 0, (size_t)4 * L * H * sizeof(float)));
        }
    } else {
        // (2) error to the preceding layer: dX[P x N] = Win[P x 4L] * deltas[4L x N]   (the 8 products of :996-1006)
        // (the optional fast mode only relaxes the forward projections; gradients always use the strict path)
        if (dX) BL_CHECK(bl_gemm_f32(ctx, 0, 0, P, N, 4 * L, W, P, pl->deltas, 4 * L, dX, lddx, 0, BL_GEMM_STRICT));

        // (3) input weight gradients: dWin[P x 4L] = X[P x N] * deltas^T   (ComputeWeightUpdateFn case 0x0, :372-389)
        BL_CHECK(bl_gemm_f32(ctx, 0, 1, P, 4 * L, N, X, ldx, pl->deltas, 4 * L, dW, P, 0, BL_GEMM_STRICT));

        // (4) recurrent weight gradients, one [H x H] block per (gate, direction) (case 0x8, :411-437, 493-500):
        //     fw: dW[k,j] = sum_{n>=S}  h[n-S,k] * delta[n,j];   bw: dW[k,j] = sum_{n<N-S} h[n+S,k] * delta[n,j]
        for (int g = 0; g < 4; ++g)
            for (int d = 0; d < pl->ndir; ++d) {
                float *blk = dWint + (size_t)g * L * H + (size_t)d * H * H;
                const int col = g * L + d * H;
                if (N - S > 0) {
                    const float *A = (d == 0) ? Y : Y + (size_t)S * ldy + H;
                    const float *B = (d == 0) ? pl->deltas + (size_t)S * 4 * L + col : pl->deltas + col;
                    BL_CHECK(bl_gemm_f32(ctx, 0, 1, H, H, N - S, A, ldy, B, 4 * L, blk, H, 0, BL_GEMM_STRICT));
                } else {
                    BL_CHECK(bl_memset(ctx, blk, 0, (size_t)H * H * sizeof(float)));
                }
            }
    }

    // (5) bias + peephole gradients: one block of partial sums per sequence group from the tm2 BPTT kernel, else a pass over the
    //     deltas and cell states
    {
        bl::TimedRegion timed(ctx, 3);
        int nsplit = pl->gb.G;
        if (!pl->t2_b) {
            nsplit = pl->gsplit;
            int rows = bl::cdiv(N, nsplit);
            if (rows < 64) rows = 64;
            nsplit = bl::cdiv(N, rows);
            dim3 grid(bl::cdiv(L, 32), nsplit), block(32, 8);
            bl::lstm_small_grads_kernel<<<grid, block, 0, ctx->stream>>>(N, S, H, L, pl->bias, pl->deltas, pl->cst, pl->gpart, rows);
            BL_LAUNCHED(ctx);
        }
        bl::lstm_small_grads_finish_kernel<<<bl::cdiv(7 * L, 256), 256, 0, ctx->stream>>>(L, nsplit, pl->t2_b ? pl->bias : 1.0f, pl->gpart, dWbias, dWpeep);
        BL_LAUNCHED(ctx);
    }
    return 0;
}

// tuning aid: copies the forward kernel's clock64 trace ([CTAs*nsub][T][6]) to the host; returns the number of rows via *rows
int bl_lstm_debug_trace(bl_lstm_plan *pl, int T, long long *host_dst, int *rows)
{
    if (!pl->trace) return bl::fail(pl->ctx, "tracing is off (set BLSTM_REC_TRACE before creating the plan)");
    if (pl->t2_f) return bl::fail(pl->ctx, "the tm2 kernels record 8 stamps per step: use bl_lstm_debug_trace2");
    const int n = pl->ndir * (pl->gf.G / pl->gf.nsub) * pl->gf.C * pl->gf.nsub;
    *rows = n;
    BL_CUDA(pl->ctx, cudaMemcpyAsync(host_dst, pl->trace, (size_t)n * T * 6 * sizeof(long long), cudaMemcpyDeviceToHost, pl->ctx->stream));
    BL_CUDA(pl->ctx, cudaStreamSynchronize(pl->ctx->stream));
    return 0;
}

int bl_lstm_debug_trace2(bl_lstm_plan *pl, int backward, int T, long long *host_dst, int *rows)
{
    const long long *src = backward ? pl->trace_b : pl->trace;
    if (!src) return bl::fail(pl->ctx, "tracing is off (set BLSTM_REC_TRACE before creating the plan)");
    if (backward ? !pl->t2_b : !pl->t2_f) return bl::fail(pl->ctx, "bl_lstm_debug_trace2: the plan does not run the tm2 kernel for this pass");
    const bl::RecGeom &g = backward ? pl->gb : pl->gf;
    const int n = pl->ndir * bl::cdiv(g.G, g.nsub) * g.C * g.nsub;       // one row per sub-group, CTA-major
    *rows = n;
    BL_CUDA(pl->ctx, cudaMemcpyAsync(host_dst, src, (size_t)n * T * 8 * sizeof(long long), cudaMemcpyDeviceToHost, pl->ctx->stream));
    BL_CUDA(pl->ctx, cudaStreamSynchronize(pl->ctx->stream));
    return 0;
}

int bl_lstm_get_internal(bl_lstm_plan *pl, int dir, int which, int T, float *dst)
{
    bl_ctx *ctx = pl->ctx;
    if (dir < 0 || dir >= pl->ndir || which < 0 || which > 9 || T <= 0 || T > pl->maxT)
        return bl::fail(ctx, "bl_lstm_get_internal: bad selector");
    const int L = pl->L, H = pl->H, N = T * pl->S;
    const float *src; int ld, col0;
    if (which == 0)      { src = pl->cst;    ld = L;     col0 = dir * H; }
    else if (which == 1) { src = pl->cerr;   ld = L;     col0 = dir * H; }
    else if (which < 6)  { src = pl->acts;   ld = 4 * L; col0 = (which - 2) * L + dir * H; }
    else                 { src = pl->deltas; ld = 4 * L; col0 = (which - 6) * L + dir * H; }
    int blocks = (int)bl::cdivz((size_t)N * H, 256); if (blocks > 4096) blocks = 4096;
    bl::gather_cols_kernel<<<blocks, 256, 0, ctx->stream>>>(N, H, src, ld, col0, dst);
    BL_LAUNCHED(ctx);
    return 0;
}

} // extern "C"
