/*
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (lstm-rnn_b200/, include/) may
 * include, link or call this file; only tests/, __graft_entry__.smoke() and the cpu_baseline
 * leg of bench.py use it, and only as the checker.
 *
 * Plain-C restatement of the arithmetic of CURRENNT's CPU path (`--cuda false`) for the
 * LSTM/BLSTM training hot path.  Every function cites the reference lines it restates
 * (paths relative to /root/reference/currennt_lib/src).
 *
 * PINNING: this restatement is checked bit-for-bit against the reference's own objects
 * compiled here (oracle/_ref/libcurrennt_ref.so, built by oracle/build_ref.sh) in
 * tests/test_oracle.py, and against the golden vectors under tests/golden/ that were
 * generated from that library (tests/golden/make_golden.py).  The reference repository
 * itself ships no numeric golden vectors for this path (tests/test1/expected_network.jsn
 * is byte-identical to its network.jsn), so "the reference run here" is the pin.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared currennt_oracle.c -lm   (see oracle/Makefile)
 * -ffp-contract=off keeps every multiply and add separately rounded, like the reference's
 * x86-64 host build.  All sums run serially in ascending index order like the reference.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

typedef float real_t;

#define PATTYPE_NONE   0   /* Types.hpp:30-33 */
#define PATTYPE_FIRST  1
#define PATTYPE_NORMAL 2
#define PATTYPE_LAST   3

/* helpers/NumericLimits.cuh:34-44 */
#define RL_MIN       1.1754944e-038f
#define RL_MAX       3.4028235e+038f
#define RL_EXPLIMIT  88.722839f
#define RL_LOGZERO   (-1e30f)

/* ------------------------------------------------------------------ scalar functions */

/* activation_functions/Logistic.cuh:33-43 */
static real_t logistic_fn(real_t x)
{
    if (x < RL_EXPLIMIT) {
        if (x > -RL_EXPLIMIT)
            return 1.0f / (1.0f + expf(-x));
        return 0.0f;
    }
    return 1.0f;
}
/* Logistic.cuh:45-48 (takes the output y) */
static real_t logistic_deriv(real_t y) { return y * (1.0f - y); }

/* activation_functions/Tanh.cuh:33-36 via Maxmin1.cuh:33-36: tanh(x) = 2*sigma(2x) - 1 */
static real_t tanh_fn(real_t x) { return 2.0f * logistic_fn(2.0f * x) - 1.0f; }
/* Tanh.cuh:38-41 */
static real_t tanh_deriv(real_t y) { return 1.0f - (y * y); }

/* helpers/boundRange.cuh:32-35, helpers/limitedError.cuh:31-34 */
static real_t bound_range(real_t x, real_t lo, real_t hi) { return x < lo ? lo : (x > hi ? hi : x); }
static real_t limited_error(real_t e) { return bound_range(e, -1.0f, +1.0f); }

/* helpers/safeExp.cuh:32-40 */
static real_t safe_exp(real_t x)
{
    if (x <= RL_LOGZERO)  return 0.0f;
    if (x >= RL_EXPLIMIT) return RL_MAX;
    return expf(x);
}

/* activation selector used by the feed-forward layer: 0 tanh, 1 logistic, 2 identity
 * (LayerFactory.cu:54-59, activation_functions/Identity.cuh:33-41) */
static real_t act_fn(int act, real_t x)    { return act == 0 ? tanh_fn(x) : act == 1 ? logistic_fn(x) : x; }
static real_t act_deriv(int act, real_t y) { return act == 0 ? tanh_deriv(y) : act == 1 ? logistic_deriv(y) : 1.0f; }

/* exported so tests can pin the scalar functions directly */
real_t orc_logistic(real_t x)       { return logistic_fn(x); }
real_t orc_tanh(real_t x)           { return tanh_fn(x); }
real_t orc_safe_exp(real_t x)       { return safe_exp(x); }
real_t orc_limited_error(real_t x)  { return limited_error(x); }

/* ------------------------------------------------------------------ helpers::Matrix products
 * Column-major, ld == rows (helpers/Matrix.cu:201-211).  One output element at a time, serial
 * fp32 dot product over the contraction index ascending, starting from 0; `add` adds the old
 * value AFTER the dot product (Matrix.cu:41-183).  (transA,transB) = (1,0),(0,0),(0,1); (1,1)
 * is "Not implemented" in the reference (Matrix.cu:248, 316).  Returns 0, or 1 on bad shapes.
 */
int orc_matrix_product(real_t *c, int rowsC, int colsC,
                       const real_t *a, int rowsA, int colsA, int transA,
                       const real_t *b, int rowsB, int colsB, int transB, int add)
{
    int i, j, k;
    if (transA && !transB) {                                  /* Matrix.cu:91-107, 221-237 */
        if (rowsC != colsA || colsC != colsB || rowsA != rowsB) return 1;
        for (j = 0; j < colsC; ++j)
            for (i = 0; i < rowsC; ++i) {
                const real_t *ca = a + (size_t)i * rowsA, *cb = b + (size_t)j * rowsB;
                real_t x = 0;
                for (k = 0; k < rowsA; ++k) x += ca[k] * cb[k];
                c[(size_t)j * rowsC + i] = add ? c[(size_t)j * rowsC + i] + x : x;
            }
    } else if (!transA && !transB) {                          /* Matrix.cu:41-57, 239-255 */
        if (rowsC != rowsA || colsC != colsB || colsA != rowsB) return 1;
        for (j = 0; j < colsC; ++j)
            for (i = 0; i < rowsC; ++i) {
                const real_t *cb = b + (size_t)j * rowsB;
                real_t x = 0;
                for (k = 0; k < colsA; ++k) x += a[(size_t)k * rowsA + i] * cb[k];
                c[(size_t)j * rowsC + i] = add ? c[(size_t)j * rowsC + i] + x : x;
            }
    } else if (!transA && transB) {                           /* Matrix.cu:135-156, 260-279 */
        if (rowsC != rowsA || colsC != rowsB || colsA != colsB) return 1;
        for (j = 0; j < colsC; ++j)
            for (i = 0; i < rowsC; ++i) {
                real_t x = 0;
                for (k = 0; k < colsA; ++k) x += a[(size_t)k * rowsA + i] * b[(size_t)k * rowsB + j];
                c[(size_t)j * rowsC + i] = add ? c[(size_t)j * rowsC + i] + x : x;
            }
    } else {
        return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------ LSTM / BLSTM layer
 * Buffers follow the reference exactly (LstmLayer.cu:543-575): per direction d, twelve arrays
 * of maxT*S*H floats indexed [t][s][j]:
 *   0 tmpOutputs 1 tmpOutputErrors 2 cellStates 3 cellStateErrors
 *   4 niActs 5 igActs 6 fgActs 7 ogActs 8 niDeltas 9 igDeltas 10 fgDeltas 11 ogDeltas
 * Weight vector layout (LstmLayer.cu:535-541, 583-596; decode in :339-357):
 *   input    @0            : g*L*P + (d*H+j)*P + p
 *   bias     @4LP          : g*L + d*H + j
 *   internal @4LP+4L       : g*L*H + d*H*H + j*H + k     (k = source cell)
 *   peephole @4LP+4L+4LH   : q*L + d*H + j, q in {IG,FG,OG}
 */
enum { B_OUT, B_OUTERR, B_CELL, B_CELLERR, B_NI, B_IG, B_FG, B_OG, B_DNI, B_DIG, B_DFG, B_DOG, B_COUNT };

typedef struct {
    int P, L, H, bidir, S, maxT, ndir;
    real_t bias;
    real_t *buf[2][B_COUNT];
} orc_lstm_t;

orc_lstm_t *orc_lstm_create(int P, int L, int bidirectional, int S, int maxT, real_t bias)
{
    int d, b;
    orc_lstm_t *h;
    if (bidirectional && (L % 2)) return NULL;            /* LstmLayer.cu:528-529 */
    h = (orc_lstm_t *)calloc(1, sizeof(*h));
    h->P = P; h->L = L; h->bidir = bidirectional; h->ndir = bidirectional ? 2 : 1;
    h->H = L / h->ndir; h->S = S; h->maxT = maxT; h->bias = bias;
    for (d = 0; d < h->ndir; ++d)
        for (b = 0; b < B_COUNT; ++b)
            h->buf[d][b] = (real_t *)calloc((size_t)maxT * S * h->H, sizeof(real_t));   /* :554 zero-filled */
    return h;
}

void orc_lstm_destroy(orc_lstm_t *h)
{
    int d, b;
    if (!h) return;
    for (d = 0; d < h->ndir; ++d) for (b = 0; b < B_COUNT; ++b) free(h->buf[d][b]);
    free(h);
}

real_t *orc_lstm_buffer(orc_lstm_t *h, int dir, int which) { return h->buf[dir][which]; }

long orc_lstm_num_weights(int P, int L, int bidirectional)
{   /* TrainableLayer.cu:103 with inputWeightsPerBlock 4, internalWeightsPerBlock (bidir?2:4)*L+3 (LstmLayer.cu:525) */
    return (long)L * (4L * (P + 1) + (bidirectional ? 2L : 4L) * L + 3);
}

/* One timestep of ComputeBlockOutputFn (LstmLayer.cu:47-138) for all S*H cells of a direction. */
static void lstm_block_outputs(const orc_lstm_t *h, int d, const real_t *W, const char *patTypes,
                               int t, int firstCall, int checkPatType, int prevDist)
{
    const int H = h->H, L = h->L, P = h->P, n = h->S * H;
    const real_t *bw = W + 4 * (size_t)L * P;                                  /* :535-538 */
    const real_t *pw = bw + 4 * (size_t)L + 4 * (size_t)L * H;                /* :539-541 */
    real_t *cs = h->buf[d][B_CELL], *ni = h->buf[d][B_NI], *ig = h->buf[d][B_IG];
    real_t *fg = h->buf[d][B_FG], *og = h->buf[d][B_OG], *out = h->buf[d][B_OUT];
    int e;
    for (e = n * t; e < n * t + n; ++e) {
        int j = e % H;
        real_t a_ni, a_ig, a_fg, a_og, c;
        if (checkPatType && patTypes[e / H] == PATTYPE_NONE) {                 /* :78-85 */
            if (prevDist > 0) cs[e] = 0;
            out[e] = 0;
            continue;
        }
        a_ni = ni[e]; a_ig = ig[e]; a_fg = fg[e]; a_og = og[e];
        a_ni += h->bias * bw[0 * L + d * H + j];                              /* :97-100 */
        a_ig += h->bias * bw[1 * L + d * H + j];
        a_fg += h->bias * bw[2 * L + d * H + j];
        a_og += h->bias * bw[3 * L + d * H + j];
        if (!firstCall) {                                                      /* :103-108 */
            real_t cp = cs[e + prevDist];
            a_ig += cp * pw[0 * L + d * H + j];
            a_fg += cp * pw[1 * L + d * H + j];
        }
        a_ni = tanh_fn(a_ni); a_ig = logistic_fn(a_ig); a_fg = logistic_fn(a_fg);   /* :111-113 */
        ni[e] = a_ni; ig[e] = a_ig; fg[e] = a_fg;
        c = a_ni * a_ig;                                                       /* :121-126 */
        if (!firstCall) c += cs[e + prevDist] * a_fg;
        cs[e] = c;
        a_og += c * pw[2 * L + d * H + j];                                    /* :129-131 */
        a_og = logistic_fn(a_og);
        og[e] = a_og;
        out[e] = tanh_fn(c) * a_og;                                           /* :134 */
    }
}

/* LstmLayer::computeForwardPass (LstmLayer.cu:763-886).
 * X: preceding layer outputs [T*S][P]; Y: layer outputs [T*S][L] (fw | bw per row, :140-161). */
void orc_lstm_forward(orc_lstm_t *h, const real_t *W, const real_t *X, const char *patTypes,
                      int T, int Tmin, real_t *Y)
{
    const int H = h->H, L = h->L, P = h->P, S = h->S, N = T * S, n = S * H;
    int d, g, t;
    const size_t inW = (size_t)L * P, itW = (size_t)L * H;
    const size_t itOff = 4 * (size_t)L * (P + 1);
    /* input projection for all slots, padded ones included (:772-786) */
    for (d = 0; d < h->ndir; ++d)
        for (g = 0; g < 4; ++g)
            orc_matrix_product(h->buf[d][B_NI + g], H, N, W + g * inW + (size_t)d * (inW / 2) * (h->bidir ? 1 : 0), P, H, 1,
                               X, P, N, 0, 0);
    /* forward-in-time direction (:812-829) */
    for (t = 0; t < T; ++t) {
        if (t != 0)
            for (g = 0; g < 4; ++g)
                orc_matrix_product(h->buf[0][B_NI + g] + (size_t)t * n, H, S, W + itOff + g * itW, H, H, 1,
                                   h->buf[0][B_OUT] + (size_t)(t - 1) * n, H, S, 0, 1);
        lstm_block_outputs(h, 0, W, patTypes, t, t == 0, t >= Tmin, -n);
    }
    /* backward-in-time direction (:832-865) */
    if (h->bidir) {
        for (t = T - 1; t >= 0; --t) {
            if (t != T - 1)
                for (g = 0; g < 4; ++g)
                    orc_matrix_product(h->buf[1][B_NI + g] + (size_t)t * n, H, S, W + itOff + g * itW + itW / 2, H, H, 1,
                                       h->buf[1][B_OUT] + (size_t)(t + 1) * n, H, S, 0, 1);
            lstm_block_outputs(h, 1, W, patTypes, t, t == T - 1, t >= Tmin, +n);
        }
    }
    /* ResortOutputsFn (:140-161, 869-885); the unidirectional layer writes Y directly (:766-769, 884) */
    {
        int s, j;
        for (s = 0; s < N; ++s)
            for (d = 0; d < h->ndir; ++d)
                for (j = 0; j < H; ++j)
                    Y[(size_t)s * L + d * H + j] = h->buf[d][B_OUT][(size_t)s * H + j];
    }
}

/* One timestep of ComputeBlockErrorsFn (LstmLayer.cu:190-287). */
static void lstm_block_errors(const orc_lstm_t *h, int d, const real_t *W, const char *patTypes,
                              int t, int firstCall, int lastCall, int checkPatType, int prevDist)
{
    const int H = h->H, L = h->L, P = h->P, n = h->S * H;
    const real_t *pw = W + 4 * (size_t)L * P + 4 * (size_t)L + 4 * (size_t)L * H;
    const real_t *cs = h->buf[d][B_CELL], *ni = h->buf[d][B_NI], *ig = h->buf[d][B_IG];
    const real_t *fg = h->buf[d][B_FG], *og = h->buf[d][B_OG], *oe = h->buf[d][B_OUTERR];
    real_t *ce = h->buf[d][B_CELLERR], *dni = h->buf[d][B_DNI], *dig = h->buf[d][B_DIG];
    real_t *dfg = h->buf[d][B_DFG], *dog = h->buf[d][B_DOG];
    int e;
    for (e = n * t; e < n * t + n; ++e) {
        int j = e % H;
        real_t outErr = oe[e], a_ni, a_ig, a_og, c, d_og, d_ni, d_fg, d_ig, cerr, w_og;
        if (checkPatType && patTypes[e / H] == PATTYPE_NONE) {                 /* :224-234 */
            dni[e] = 0; dig[e] = 0; dfg[e] = 0; dog[e] = 0; ce[e] = 0;
            continue;
        }
        a_ni = ni[e]; a_ig = ig[e]; a_og = og[e]; c = cs[e];
        d_og = logistic_deriv(a_og) * tanh_fn(c) * outErr;                    /* :246 */
        w_og = pw[2 * L + d * H + j];
        cerr = a_og * tanh_deriv(tanh_fn(c)) * outErr + w_og * d_og;          /* :250 */
        if (!firstCall) {                                                      /* :252-262, "next" = e - prevDist */
            real_t nfg = fg[e - prevDist], nce = ce[e - prevDist];
            real_t ndig = dig[e - prevDist], ndfg = dfg[e - prevDist];
            real_t w_ig = pw[0 * L + d * H + j], w_fg = pw[1 * L + d * H + j];
            cerr += nfg * nce + w_ig * ndig + w_fg * ndfg;
        }
        d_ni = a_ig * tanh_deriv(a_ni) * cerr;                                /* :265 */
        d_fg = 0;                                                              /* :268-275 */
        if (!lastCall)
            d_fg = logistic_deriv(fg[e]) * cs[e + prevDist] * cerr;
        d_ig = logistic_deriv(a_ig) * a_ni * cerr;                            /* :278 */
        dni[e] = limited_error(d_ni); dig[e] = limited_error(d_ig);           /* :281-285 */
        dfg[e] = limited_error(d_fg); dog[e] = limited_error(d_og);
        ce[e] = cerr;
    }
}

/* LstmLayer::computeBackwardPass (LstmLayer.cu:888-1051).
 * dY: this layer's outputErrors [T*S][L] -- for the unidirectional layer the reference works IN PLACE on it
 * (vector swap, :907-910, 1047-1050), so on return it holds dY plus the recurrent error terms; dX: preceding layer's outputErrors [T*S][P] or NULL when the
 * preceding layer is not trainable (:991-992); dW: weightUpdates, same layout as W. */
void orc_lstm_backward(orc_lstm_t *h, const real_t *W, const real_t *X, real_t *dY,
                       const char *patTypes, int T, int Tmin, real_t *dX, real_t *dW)
{
    const int H = h->H, L = h->L, P = h->P, S = h->S, N = T * S, n = S * H;
    const size_t inW = (size_t)L * P, itW = (size_t)L * H;
    const size_t biOff = 4 * inW, itOff = biOff + 4 * (size_t)L, peOff = itOff + 4 * itW;
    int d, g, t, s, j;

    /* ResortOutputErrorsFn (:163-188, 892-906) */
    for (s = 0; s < N; ++s)
        for (d = 0; d < h->ndir; ++d)
            for (j = 0; j < H; ++j)
                h->buf[d][B_OUTERR][(size_t)s * H + j] = dY[(size_t)s * L + d * H + j];

    /* fw direction walks t = T-1 .. 0 (:936-951); firstCall = (t==T-1), lastCall = (t==0) */
    for (t = T - 1; t >= 0; --t) {
        if (t != T - 1)
            for (g = 0; g < 4; ++g)
                orc_matrix_product(h->buf[0][B_OUTERR] + (size_t)t * n, H, S, W + itOff + g * itW, H, H, 0,
                                   h->buf[0][B_DNI + g] + (size_t)(t + 1) * n, H, S, 0, 1);
        lstm_block_errors(h, 0, W, patTypes, t, t == T - 1, t == 0, t >= Tmin, -n);
    }
    /* bw direction walks t = 0 .. T-1 (:970-985); firstCall = (t==0), lastCall = (t==T-1) */
    if (h->bidir) {
        for (t = 0; t < T; ++t) {
            if (t != 0)
                for (g = 0; g < 4; ++g)
                    orc_matrix_product(h->buf[1][B_OUTERR] + (size_t)t * n, H, S, W + itOff + g * itW + itW / 2, H, H, 0,
                                       h->buf[1][B_DNI + g] + (size_t)(t - 1) * n, H, S, 0, 1);
            lstm_block_errors(h, 1, W, patTypes, t, t == 0, t == T - 1, t >= Tmin, +n);
        }
    }

    /* unidirectional: tmpOutputErrors IS outputErrors (swap at :907-910 / :1047-1050) */
    if (!h->bidir)
        memcpy(dY, h->buf[0][B_OUTERR], sizeof(real_t) * (size_t)N * H);

    /* error to the preceding layer (:989-1009): assign with fw NI, then add the other seven */
    if (dX) {
        int first = 1;
        for (d = 0; d < h->ndir; ++d)
            for (g = 0; g < 4; ++g) {
                orc_matrix_product(dX, P, N, W + g * inW + (h->bidir ? (size_t)d * (inW / 2) : 0), P, H, 0,
                                   h->buf[d][B_DNI + g], H, N, 0, !first);
                first = 0;
            }
    }

    /* ComputeWeightUpdateFn (:289-512, 1012-1044): one serial sum over patterns per weight */
    for (g = 0; g < 4; ++g)
        for (d = 0; d < h->ndir; ++d)
            for (j = 0; j < H; ++j) {
                const real_t *delta = h->buf[d][B_DNI + g] + j;
                int p, k, i;
                /* input weights (:372-389), all N patterns */
                for (p = 0; p < P; ++p) {
                    real_t wu = 0;
                    for (i = 0; i < N; ++i) wu += X[(size_t)i * P + p] * delta[(size_t)i * H];
                    dW[g * inW + (size_t)(d * H + j) * P + p] = wu;
                }
                /* bias weights (:392-408) */
                {
                    real_t wu = 0;
                    for (i = 0; i < N; ++i) wu += h->bias * delta[(size_t)i * H];
                    dW[biOff + (size_t)g * L + d * H + j] = wu;
                }
                /* internal weights (:411-437, 493-500): fw pairs delta[i] with out[i-S] for i>=S;
                 * bw pairs delta[i] with out[i+S] for i < N-S */
                for (k = 0; k < H; ++k) {
                    const real_t *src = h->buf[d][B_OUT] + k;
                    real_t wu = 0;
                    if (d == 0) { for (i = S; i < N; ++i)     wu += src[(size_t)(i - S) * H] * delta[(size_t)i * H]; }
                    else        { for (i = 0; i < N - S; ++i) wu += src[(size_t)(i + S) * H] * delta[(size_t)i * H]; }
                    dW[itOff + g * itW + (size_t)d * H * H + (size_t)j * H + k] = wu;
                }
                /* peephole weights (:440-475): IG,FG time-shifted like internal; OG unshifted */
                if (g >= 1) {
                    const real_t *src = h->buf[d][B_CELL] + j;
                    real_t wu = 0;
                    if (g == 3)      { for (i = 0; i < N; ++i)     wu += src[(size_t)i * H] * delta[(size_t)i * H]; }
                    else if (d == 0) { for (i = S; i < N; ++i)     wu += src[(size_t)(i - S) * H] * delta[(size_t)i * H]; }
                    else             { for (i = 0; i < N - S; ++i) wu += src[(size_t)(i + S) * H] * delta[(size_t)i * H]; }
                    dW[peOff + (size_t)(g - 1) * L + d * H + j] = wu;
                }
            }
}

/* ------------------------------------------------------------------ feed-forward layer
 * Weights: [ W: j*P + p ][ b @O*P: j ] (FeedForwardLayer.cu:115, 149-151, 165).
 * forward (FeedForwardLayer.cu:143-172): Y = act(W^T X + bias*b) for all N slots. */
void orc_ff_forward(int act, int P, int O, int N, real_t bias, const real_t *W, const real_t *X, real_t *Y)
{
    int i;
    orc_matrix_product(Y, O, N, W, P, O, 1, X, P, N, 0, 0);
    for (i = 0; i < N * O; ++i) {
        real_t a = Y[i];
        a += bias * W[(size_t)O * P + i % O];                                  /* :58-61 */
        Y[i] = act_fn(act, a);
    }
}

/* backward (FeedForwardLayer.cu:174-224): dY is overwritten with the deltas (:69-80), for all slots
 * including padding; dX = W * delta if the preceding layer is trainable (:190-197);
 * dW = X * delta^T (:206); db_j = sum_n bias*delta (:83-100). */
void orc_ff_backward(int act, int P, int O, int N, real_t bias, const real_t *W, const real_t *X,
                     const real_t *Y, real_t *dY, real_t *dX, real_t *dW)
{
    int i, j;
    for (i = 0; i < N * O; ++i) dY[i] = act_deriv(act, Y[i]) * dY[i];
    if (dX) orc_matrix_product(dX, P, N, W, P, O, 0, dY, O, N, 0, 0);
    orc_matrix_product(dW, P, O, X, P, N, 0, dY, O, N, 1, 0);
    for (j = 0; j < O; ++j) {
        real_t wu = 0;
        for (i = 0; i < N; ++i) wu += bias * dY[(size_t)i * O + j];
        dW[(size_t)O * P + j] = wu;
    }
}

/* ------------------------------------------------------------------ softmax layer
 * forward (SoftmaxLayer.cu:250-315, functors :45-155) on top of the identity feed-forward:
 * per valid pattern: off = 0.5*(min+max) with max starting at FLT_MIN and min at FLT_MAX (:62-63),
 * y = safeExp(x-off), y /= serial sum.  Padded patterns keep the raw W^T x + b. */
void orc_softmax_forward(int O, int N, const char *patTypes, real_t *Y)
{
    int n, i;
    for (n = 0; n < N; ++n) {
        real_t *y = Y + (size_t)n * O, mx = RL_MIN, mn = RL_MAX, off, sum = 0;
        if (patTypes[n] == PATTYPE_NONE) continue;
        for (i = 0; i < O; ++i) { real_t x = y[i]; mn = (mn < x ? mn : x); mx = (mx > x ? mx : x); }
        off = 0.5f * (mn + mx);
        for (i = 0; i < O; ++i) y[i] = safe_exp(y[i] - off);
        for (i = 0; i < O; ++i) sum += y[i];
        for (i = 0; i < O; ++i) y[i] = y[i] / sum;
    }
}

/* backward (SoftmaxLayer.cu:317-353, functors :157-219): e_i <- y_i*(e_i - sum_j y_j e_j), padded untouched;
 * the feed-forward backward then runs with the identity activation. */
void orc_softmax_backward(int O, int N, const char *patTypes, const real_t *Y, real_t *dY)
{
    int n, i;
    for (n = 0; n < N; ++n) {
        const real_t *y = Y + (size_t)n * O; real_t *e = dY + (size_t)n * O, off = 0;
        if (patTypes[n] == PATTYPE_NONE) continue;
        for (i = 0; i < O; ++i) off += y[i] * e[i];
        for (i = 0; i < O; ++i) e[i] = y[i] * (e[i] - off);
    }
}

/* ------------------------------------------------------------------ post-output layers */

/* MulticlassClassificationLayer::calculateError (MulticlassClassificationLayer.cu:48-68, 195-214):
 * -(serial sum over patterns of log(max(FLT_MIN, y[target]))) */
real_t orc_multiclass_error(int O, int N, const int *targetClasses, const real_t *Y)
{
    real_t sum = 0; int n;
    for (n = 0; n < N; ++n) {
        real_t v = 0;
        if (targetClasses[n] != -1) {
            real_t p = Y[(size_t)n * O + targetClasses[n]];
            p = (RL_MIN > p ? RL_MIN : p);
            v = logf(p);
        }
        sum = sum + v;
    }
    return -sum;
}

/* countCorrectClassifications (:70-104, 159-177): argmax with strict > from (0, class 0) */
int orc_multiclass_count_correct(int O, int N, const int *targetClasses, const real_t *Y)
{
    int n, i, correct = 0;
    for (n = 0; n < N; ++n) {
        real_t best = 0; int est = 0;
        if (targetClasses[n] == -1) continue;
        for (i = 0; i < O; ++i) { real_t o = Y[(size_t)n * O + i]; if (o > best) { best = o; est = i; } }
        if (est == targetClasses[n]) ++correct;
    }
    return correct;
}

/* computeBackwardPass (:106-135, 221-240): zero, then e[n,target] = -1/max(FLT_MIN,y) */
void orc_multiclass_backward(int O, int N, const int *targetClasses, const real_t *Y, real_t *dY)
{
    int n;
    memset(dY, 0, sizeof(real_t) * (size_t)N * O);
    for (n = 0; n < N; ++n) {
        real_t p;
        if (targetClasses[n] == -1) continue;
        p = Y[(size_t)n * O + targetClasses[n]];
        p = (RL_MIN > p ? RL_MIN : p);
        dY[(size_t)n * O + targetClasses[n]] = -(1 / p);
    }
}

/* CePostOutputLayer::calculateError (CePostOutputLayer.cu:43-70, 125-143) */
real_t orc_ce_error(int O, int N, const char *patTypes, const real_t *targets, const real_t *Y)
{
    real_t sum = 0; size_t i;
    for (i = 0; i < (size_t)N * O; ++i) {
        real_t v = 0;
        if (patTypes[i / O] != PATTYPE_NONE) {
            real_t t = targets[i], ft = (RL_MIN > t ? RL_MIN : t), o = (RL_MIN > Y[i] ? RL_MIN : Y[i]);
            v = t * logf(ft / o);
        }
        sum = sum + v;
    }
    return sum;
}

/* CePostOutputLayer::computeBackwardPass (:72-98, 150-166) */
void orc_ce_backward(int O, int N, const char *patTypes, const real_t *targets, const real_t *Y, real_t *dY)
{
    size_t i;
    for (i = 0; i < (size_t)N * O; ++i) {
        if (patTypes[i / O] == PATTYPE_NONE) { dY[i] = 0; continue; }
        {
            real_t o = (RL_MIN > Y[i] ? RL_MIN : Y[i]);
            dY[i] = bound_range(-targets[i] / o, -100, +100);
        }
    }
}

/* SsePostOutputLayer::calculateError (SsePostOutputLayer.cu:39-60, 114-132) */
real_t orc_sse_error(int O, int N, const char *patTypes, const real_t *targets, const real_t *Y)
{
    real_t sum = 0; size_t i;
    for (i = 0; i < (size_t)N * O; ++i) {
        real_t v = 0;
        if (patTypes[i / O] != PATTYPE_NONE) { real_t diff = targets[i] - Y[i]; v = diff * diff; }
        sum = sum + v;
    }
    return 0.5f * sum;
}

/* SsePostOutputLayer::computeBackwardPass (:62-88, 139-155) */
void orc_sse_backward(int O, int N, const char *patTypes, const real_t *targets, const real_t *Y, real_t *dY)
{
    size_t i;
    for (i = 0; i < (size_t)N * O; ++i)
        dY[i] = (patTypes[i / O] == PATTYPE_NONE) ? 0 : Y[i] - targets[i];
}

/* RmsePostOutputLayer::computeForwardPass (RmsePostOutputLayer.cu:39-67, 137-153): per-pattern RMSE into rmses[N]
 * (0 for padded patterns) */
void orc_rmse_forward(int O, int N, const char *patTypes, const real_t *targets, const real_t *Y, real_t *rmses)
{
    int n, i;
    for (n = 0; n < N; ++n) {
        real_t r = 0;
        if (patTypes[n] != PATTYPE_NONE) {
            real_t sum = 0;
            for (i = 0; i < O; ++i) { real_t diff = Y[(size_t)n * O + i] - targets[(size_t)n * O + i]; sum = sum + diff * diff; }
            r = sqrtf(sum / O);
        }
        rmses[n] = r;
    }
}

/* RmsePostOutputLayer::calculateError (:126-134): thrust::reduce of the per-pattern RMSEs, pattern order */
real_t orc_rmse_error(int N, const real_t *rmses)
{
    real_t total = 0; int n;
    for (n = 0; n < N; ++n) total = total + rmses[n];
    return total;
}

/* RmsePostOutputLayer::computeBackwardPass (:69-93, 155-170): error = rmse[pattern] * (y - t), every entry */
void orc_rmse_backward(int O, int N, const real_t *rmses, const real_t *targets, const real_t *Y, real_t *dY)
{
    size_t i;
    for (i = 0; i < (size_t)N * O; ++i) dY[i] = rmses[i / O] * (Y[i] - targets[i]);
}

/* WeightedSsePostOutputLayer (WeightedSsePostOutputLayer.cu:40-93, 121-164) and SseMaskPostOutputLayer ("wf",
 * SseMaskPostOutputLayer.cu:40-93, 121-164): the post-output layer is twice as wide as the output layer, its targets
 * hold (target, weight | filter input) pairs; O = size of the output layer */
real_t orc_weightedsse_error(int O, int N, const char *patTypes, const real_t *targets, const real_t *Y)
{
    real_t sum = 0; size_t i;
    for (i = 0; i < (size_t)N * O; ++i) {
        real_t v = 0;
        if (patTypes[i / O] != PATTYPE_NONE) { real_t diff = (Y[i] - targets[2 * i]) * targets[2 * i + 1]; v = diff * diff; }
        sum = sum + v;
    }
    return 0.5f * sum;
}

void orc_weightedsse_backward(int O, int N, const char *patTypes, const real_t *targets, const real_t *Y, real_t *dY)
{
    size_t i;
    for (i = 0; i < (size_t)N * O; ++i)
        dY[i] = (patTypes[i / O] == PATTYPE_NONE) ? 0 : (Y[i] - targets[2 * i]) * targets[2 * i + 1];
}

real_t orc_ssemask_error(int O, int N, const char *patTypes, const real_t *targets, const real_t *Y)
{
    real_t sum = 0; size_t i;
    for (i = 0; i < (size_t)N * O; ++i) {
        real_t v = 0;
        if (patTypes[i / O] != PATTYPE_NONE) { real_t diff = Y[i] * targets[2 * i + 1] - targets[2 * i]; v = diff * diff; }
        sum = sum + v;
    }
    return 0.5f * sum;
}

void orc_ssemask_backward(int O, int N, const char *patTypes, const real_t *targets, const real_t *Y, real_t *dY)
{
    size_t i;
    for (i = 0; i < (size_t)N * O; ++i)
        dY[i] = (patTypes[i / O] == PATTYPE_NONE) ? 0 : (Y[i] * targets[2 * i + 1] - targets[2 * i]) * targets[2 * i + 1];
}

/* BinaryClassificationLayer::calculateError (BinaryClassificationLayer.cu:43-66, 165-181); targets are the target classes as reals */
real_t orc_binary_error(int N, const char *patTypes, const real_t *targets, const real_t *Y)
{
    real_t sum = 0; int n;
    for (n = 0; n < N; ++n) {
        real_t v = 0;
        if (patTypes[n] != PATTYPE_NONE) {
            real_t act = (Y[n] > RL_MIN ? Y[n] : RL_MIN), p = (targets[n] > 0 ? act : 1 - act);
            v = -logf(p);
        }
        sum = sum + v;
    }
    return sum;
}

/* BinaryClassificationLayer::countCorrectClassifications (:68-84, 132-147) */
int orc_binary_count_correct(int N, const char *patTypes, const real_t *targets, const real_t *Y)
{
    int n, c = 0;
    for (n = 0; n < N; ++n) c += (patTypes[n] != PATTYPE_NONE) && ((targets[n] > 0.5f) == (Y[n] > 0.5f));
    return c;
}

/* BinaryClassificationLayer::computeBackwardPass (:86-112, 188-203): padded patterns keep whatever outputErrors held */
void orc_binary_backward(int N, const char *patTypes, const real_t *targets, const real_t *Y, real_t *dY)
{
    int n;
    for (n = 0; n < N; ++n) {
        if (patTypes[n] == PATTYPE_NONE) continue;
        {
            real_t act = (Y[n] > RL_MIN ? Y[n] : RL_MIN), p = (targets[n] > 0 ? act : 1 - act);
            dY[n] = (targets[n] > 0 ? -(1 / p) : (1 / p));
        }
    }
}

/* ------------------------------------------------------------------ optimizer step
 * UpdateWeightFn (optimizers/SteepestDescentOptimizer.cu:39-59): delta = momentum*delta - lr*grad; w += delta */
void orc_sgd_update(long n, real_t learningRate, real_t momentum, real_t *weights, const real_t *weightUpdates,
                    real_t *weightDeltas)
{
    long i;
    for (i = 0; i < n; ++i) {
        real_t delta = momentum * weightDeltas[i] - learningRate * weightUpdates[i];
        weightDeltas[i] = delta;
        weights[i] = weights[i] + delta;
    }
}

/* ------------------------------------------------------------------ data set: truncation + fraction packing */

/* Sequence chunking of DataSet::DataSet (data_sets/DataSet.cpp:527-542): while len>0, cut a chunk of
 * truncSeqLength if len > 1.5*truncSeqLength (double arithmetic), else take the rest.
 * Writes chunk lengths to `out` (capacity cap); returns the number of chunks. */
int orc_truncate_sequence(int seqLength, int truncSeqLength, int *out, int cap)
{
    int k = 0;
    while (seqLength > 0) {
        int len;
        if (truncSeqLength > 0 && seqLength > 1.5 * truncSeqLength)
            len = (truncSeqLength < seqLength ? truncSeqLength : seqLength);
        else
            len = seqLength;
        if (k < cap) out[k] = len;
        seqLength -= len;
        ++k;
    }
    return k;
}

/* DataSet::_makeFractionTask (data_sets/DataSet.cpp:300-414) with no context window and no output lag
 * (Configuration defaults).  seqLengths/seqInputs/seqClasses/seqTargets describe the data set's sequences
 * (already truncated and sorted); the fraction takes sequences firstSeq .. firstSeq+S-1 that exist.
 * Outputs are sized T*S (T returned through *pT): inputs [T][S][P] zero padded, patTypes NONE padded,
 * targetClasses -1 padded (or targets, left zero where padded).  Returns the number of sequences placed. */
int orc_make_fraction(int numSeqs, const int *seqLengths, const real_t *const *seqInputs,
                      const int *const *seqClasses, const real_t *const *seqTargets,
                      int P, int O, int S, int firstSeq,
                      int *pT, int *pTmin, real_t *inputs, char *patTypes, int *targetClasses, real_t *targets)
{
    int T = -2147483647 - 1, Tmin = 2147483647, placed = 0, i, t;
    for (i = firstSeq; i < firstSeq + S; ++i)
        if (i < numSeqs) {                                                     /* :315-327 */
            if (seqLengths[i] > T) T = seqLengths[i];
            if (seqLengths[i] < Tmin) Tmin = seqLengths[i];
            ++placed;
        }
    *pT = T; *pTmin = Tmin;
    if (!placed) return 0;
    memset(inputs, 0, sizeof(real_t) * (size_t)T * S * P);                     /* :330-336 */
    memset(patTypes, PATTYPE_NONE, (size_t)T * S);
    if (targetClasses) for (i = 0; i < T * S; ++i) targetClasses[i] = -1;
    if (targets) memset(targets, 0, sizeof(real_t) * (size_t)T * S * O);
    for (i = 0; i < S; ++i) {
        int q = firstSeq + i, len;
        if (q >= numSeqs) continue;                                            /* :340-341 */
        len = seqLengths[q];
        for (t = 0; t < len; ++t) {
            size_t slot = (size_t)t * S + i;                                   /* :358 */
            memcpy(inputs + slot * P, seqInputs[q] + (size_t)t * P, sizeof(real_t) * P);
            if (targetClasses) targetClasses[slot] = seqClasses[q][t];        /* :370-378 */
            if (targets) memcpy(targets + slot * O, seqTargets[q] + (size_t)t * O, sizeof(real_t) * O);   /* :381-393 */
            patTypes[slot] = (t == 0) ? PATTYPE_FIRST : (t == len - 1) ? PATTYPE_LAST : PATTYPE_NORMAL;   /* :397-406 */
        }
    }
    return placed;
}
