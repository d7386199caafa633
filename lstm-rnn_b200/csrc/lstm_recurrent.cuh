// Persistent recurrent kernels of the LSTM layer: geometry + launch interface (lstm_recurrent.cu).
#pragma once
#include "common.cuh"

namespace bl {

constexpr int REC_NT_MAX = 1024;       // threads per CTA: 1024 (64 registers each) or 768 (85 registers), 1 CTA per SM
constexpr int REC_NPAIR = 2;           // (cell, sequence) pairs per thread in the elementwise phase

// Launch geometry of one persistent kernel.  One CTA owns CL cells of one direction for the SG sequences of
// one sequence group; its slice of the recurrent weights stays in shared memory for the whole pass.
struct RecGeom {
    int G, C, CL, SG;       // sequence groups (all sub-CTAs counted), cell slices per (direction, group), cells per CTA, sequences per group
    int NT;                 // threads per CTA (768 or 1024)
    int nsub;               // sub-CTAs per CTA: each owns one sequence group and has its own barrier and step counter, so one
                            // group's counter wait / exchange copy overlaps the other groups' FFMA work on the same SM; all
                            // sub-CTAs share the CTA's weight slice in shared memory.  CTAs per direction = (G / nsub) * C
    int npair;              // (cell, sequence) pairs per thread in the gate-math phase: 1 or 2
    int R;                  // GEMM rows per CTA: 4*CL (forward: gate x cell), CL (BPTT: cell)
    int LR, LS, LSlog;      // lanes along row quads / sequence quads (LR*LS == 32)
    int WR, WS, KS;         // warp tiles along rows / sequences, K splits
    int RQt, SQt;           // padded row quads / sequence quads (WR*LR, WS*LS)
    int Rpad, Spad;         // 4*RQt, 4*SQt
    int K4, KB4;            // contraction length and per-split length in float4 units
    int RS;                 // shared-memory row stride (floats) of both operand tiles, == global exchange row stride
    int RP;                 // row pitch of the staging buffer
    int Hpad;
    size_t smem;
    double cost;
    size_t xelems;          // tm2 kernels: floats of the exchange buffer (0 = the plan's default ndir*2*S*RS)
};

bool choose_geometry(bool bwd, int H, int S, int ndir, int num_sms, int smem_cap, int forceG, int forceNsub, int forceNT, RecGeom *out);

struct RecFwdParams {
    const float *Wb, *Wi, *Wp;      // bias / internal / peephole segments of the layer's weight vector
    float *acts;                    // [N][4L]  pre-activations in, activations out (column g*L + d*H + j)
    float *cst;                     // [N][L]   cell states
    float *Y; int ldy;              // [N][ldy] layer outputs
    float *hx;                      // [ndir][2][S][Hpad] step-parity exchange buffer for h
    unsigned *flags;                // [ndir*G*32] step counters
    const char *pat;
    int T, Tmin, S, H, L, ndir;
    float bias;
    long long *trace;               // optional [CTAs*nsub][T][6] clock64 stamps (BLSTM_REC_TRACE tuning aid), else NULL
    // optional (register-resident kernels only): TF32 split of the layer output written on the fly for the tensor-core backward
    // pass, hi/lo [N][ld_ys], column d*Hq + j with Hq = roundup(H, 4); NULL = not wanted
    float *ys_hi, *ys_lo; int ld_ys;
    RecGeom g;
};

struct RecBwdParams {
    const float *Wi, *Wp;
    const float *acts, *cst;        // from the forward pass
    float *deltas;                  // [N][4L]
    float *cerr;                    // [N][L] cell state errors
    float *dY; int lddy;            // [N][lddy] output errors (updated in place when !bidirectional)
    float *dx;                      // [ndir][2][S][RS] step-parity exchange buffer for the 4 gate deltas
    unsigned *flags;
    const char *pat;
    int T, Tmin, S, H, L, ndir;
    // optional (register-resident kernels only): TF32 split of the deltas, hi/lo [N][ld_ds], column (gate*ndir + d)*Hq + j
    float *ds_hi, *ds_lo; int ld_ds;
    long long *trace;               // optional [CTAs][T][8] clock64 stamps (tm2 kernels, BLSTM_REC_TRACE), else NULL
    // tm2 kernel only: bias / peephole gradient partials accumulated in registers over the pass, one block per sequence group:
    // gpart[group][7][L] = sum delta_ni, delta_ig, delta_fg, delta_og (bias, before the bias factor), sum c_prev*delta_ig,
    // sum c_prev*delta_fg, sum c*delta_og (LstmLayer.cu:392-408, 440-475); NULL = not wanted
    float *gpart;
    RecGeom g;
};

// register-resident variant (lstm_recurrent_reg.cu): geometry in the same struct (RQt = row halves, KS = 32-float k chunks)
bool choose_geometry_reg(bool bwd, int H, int S, int ndir, int num_sms, int smem_cap, int forceG, RecGeom *out);
int launch_lstm_fwd_reg(bl_ctx *ctx, const RecFwdParams &p);
int launch_lstm_bwd_reg(bl_ctx *ctx, const RecBwdParams &p);

// tensor-memory-resident variants (lstm_recurrent_tmem.cu): weights in TMEM, step GEMM on tcgen05; pad32(H) <= 256 only
bool choose_geometry_tmem(int H, int S, int ndir, int num_sms, int smem_cap, int forceG, RecGeom *out);
int launch_lstm_fwd_tmem(bl_ctx *ctx, const RecFwdParams &p);
bool choose_geometry_tmem_bwd(int H, int S, int ndir, int num_sms, int smem_cap, int forceG, RecGeom *out);
int launch_lstm_bwd_tmem(bl_ctx *ctx, const RecBwdParams &p);

// second tensor-memory generation (lstm_recurrent_tm2.cu): in-band exchange, fp16 two-term operands, warp-specialised step;
// forward up to Hp = 512 and BPTT up to R = 512 (W_lo' in shared memory where TMEM alone is too small)
bool choose_geometry_tm2(bool bwd, int H, int S, int ndir, int num_sms, int smem_cap, int forceG, RecGeom *out);
int launch_lstm_fwd_tm2(bl_ctx *ctx, const RecFwdParams &p);
int launch_lstm_bwd_tm2(bl_ctx *ctx, const RecBwdParams &p);

int launch_lstm_fwd(bl_ctx *ctx, const RecFwdParams &p);
int launch_lstm_bwd(bl_ctx *ctx, const RecBwdParams &p);

} // namespace bl
