/*
 * blstm_b200.h -- C ABI of the B200-native (sm_100a) LSTM/BLSTM training hot path.
 *
 * This is the drop-in boundary: the reference (CURRENNT, naxingyu/lstm-rnn) has no plugin or FFI
 * interface, so the seam is the one its layer classes already have -- they call helpers::Matrix
 * (cuBLAS) and Thrust functors.  Each entry point below replaces one of those call groups; the
 * comment on each names the reference lines it replaces (paths relative to currennt_lib/src).
 * INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, opaque handles; no C++/torch types.
 *   - every function returns 0 on success, non-zero on failure and never throws;
 *     bl_last_error(ctx) (ctx may be NULL for creation failures) returns the message.  The host
 *     wrapper turns non-zero into std::runtime_error, the reference's error convention
 *     (helpers/cublas.cu:44-45, main.cpp:492-495).
 *   - all tensor pointers are DEVICE pointers unless the name says host; the caller owns them
 *     (the reference's layers own outputs/outputErrors/weights, layers/Layer.cpp:41-67).
 *   - work is enqueued on the context's stream and is asynchronous; bl_sync() waits.
 *   - pattern-major layouts exactly as the reference: slot n = t*S + s, features fastest
 *     (data_sets/DataSet.cpp:358).  Every [N][size] tensor takes a leading dimension `ld >= size`
 *     (in floats) so callers may pad rows to 16 bytes for the TMA-fed tensor-core path; the
 *     reference layout is ld == size.
 *   - weights / weightUpdates use the reference's flat per-layer layout unchanged
 *     (layers/LstmLayer.cu:535-541,583-596; layers/FeedForwardLayer.cu:149-165).
 *   - there is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef BLSTM_B200_H
#define BLSTM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bl_ctx       bl_ctx;        /* device + stream + scratch               */
typedef struct bl_lstm_plan bl_lstm_plan;  /* per-layer workspace + launch geometry   */
typedef struct bl_comm      bl_comm;       /* NCCL communicator for data parallelism  */

/* GEMM precision modes (north_star: strict fp32 <= 1e-5, optional TF32/bf16 projection <= 2e-3) */
#define BL_GEMM_STRICT   0   /* fp32-exact products: SIMT FFMA, or 3xTF32 error-compensated tcgen05 */
#define BL_GEMM_FAST     1   /* single-pass TF32 tcgen05                                            */

/* activation selectors of layers::FeedForwardLayer (LayerFactory.cu:54-61) */
#define BL_ACT_TANH      0
#define BL_ACT_LOGISTIC  1
#define BL_ACT_IDENTITY  2

/* pattern types, Types.hpp:30-33 */
#define BL_PATTYPE_NONE   0
#define BL_PATTYPE_FIRST  1
#define BL_PATTYPE_NORMAL 2
#define BL_PATTYPE_LAST   3

/* ------------------------------------------------------------------ context, memory */

/* `stream` is a cudaStream_t (or NULL for a stream owned by the context).  Replaces the implicit
 * default-stream / cublasCreate state of helpers/cublas.cu:37-58. */
int  bl_ctx_create(int device, void *stream, bl_ctx **out);
void bl_ctx_destroy(bl_ctx *ctx);
const char *bl_last_error(const bl_ctx *ctx);
int  bl_sync(bl_ctx *ctx);
/* 0 = strict (default), 1 = fast; applies to the forward projections issued by the layer-level calls
 * (W^T X of the LSTM / feed-forward layers); the backward contractions always run strict */
int  bl_ctx_set_gemm_mode(bl_ctx *ctx, int mode);
/* GEMM backend: 0 = automatic (tcgen05 tensor-core path for large contractions, SIMT FFMA for small ones),
 * 1 = SIMT only, 2 = tcgen05 always (tests).  The precision mode applies to either backend. */
int  bl_ctx_set_gemm_backend(bl_ctx *ctx, int backend);
int  bl_ctx_num_sms(const bl_ctx *ctx);
/* number of CUDA devices visible to the process (--list_devices, main.cpp:509-524); 0 when there is no usable driver */
int  bl_device_count(void);
/* number of kernels this library has launched on the context since creation (bench.py gpu_launches) */
long bl_ctx_launch_count(const bl_ctx *ctx);

/* Optional per-kernel-class device timing for bench.py's roofline line: when enabled, the library brackets its
 * launches with CUDA events on the context's stream.  Classes: 0 gemm, 1 lstm recurrent forward (persistent),
 * 2 lstm BPTT (persistent), 3 everything else (elementwise / reductions).  bl_ctx_timing_read synchronises the
 * stream, returns accumulated milliseconds and launch counts per class, and resets the accumulators. */
#define BL_TIMING_CLASSES 4
int  bl_ctx_timing_enable(bl_ctx *ctx, int on);
int  bl_ctx_timing_read(bl_ctx *ctx, double *ms4, long *count4);

/* thrust::device_vector allocation / thrust::copy replacements (Types.hpp:58-67, layers/InputLayer.cpp:59) */
int  bl_malloc(bl_ctx *ctx, void **ptr, size_t bytes);
int  bl_free(bl_ctx *ctx, void *ptr);
int  bl_memset(bl_ctx *ctx, void *ptr, int value, size_t bytes);
int  bl_memcpy_h2d(bl_ctx *ctx, void *dst, const void *host_src, size_t bytes);
int  bl_memcpy_d2h(bl_ctx *ctx, void *host_dst, const void *src, size_t bytes);
int  bl_memcpy_d2d(bl_ctx *ctx, void *dst, const void *src, size_t bytes);
/* strided row copies: `rows` rows of `row_bytes`, pitches in bytes (packed host <-> padded device) */
int  bl_memcpy2d_h2d(bl_ctx *ctx, void *dst, size_t dst_pitch, const void *host_src, size_t src_pitch, size_t row_bytes, size_t rows);
int  bl_memcpy2d_d2h(bl_ctx *ctx, void *host_dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t row_bytes, size_t rows);
/* Upload fence for pinned staging buffers: bl_upload_mark records a point in the context's stream (after the asynchronous H2D copies
 * enqueued so far) and returns a ticket; bl_upload_wait blocks the host until that point has been reached (returns at once if it
 * already has).  The fraction objects use it so that a pinned buffer never goes back to the staging pool while a copy out of it is
 * still queued -- the reference's synchronous thrust::copy (layers/InputLayer.cpp:59) needs no such thing. */
int  bl_upload_mark(bl_ctx *ctx, unsigned long long *ticket);
int  bl_upload_wait(bl_ctx *ctx, unsigned long long ticket);
/* pinned host memory for the fraction staging buffers */
int  bl_malloc_host(bl_ctx *ctx, void **ptr, size_t bytes);
int  bl_free_host(bl_ctx *ctx, void *ptr);

/* ------------------------------------------------------------------ helpers::Matrix products
 * Drop-in for cublas::multiplyMatrices (helpers/cublas.hpp:30-37, helpers/cublas.cu:60-88) and therefore
 * for Matrix<Gpu>::assignProduct/addProduct (helpers/Matrix.cu:351-377): column-major,
 *   C[m x n] (ldc) = op(A) * op(B) (+ C if accumulate),  op(A) is m x k, op(B) is k x n.
 * (transA,transB) in {(1,0),(0,0),(0,1)}; (1,1) fails like the reference's "Not implemented". */
int bl_gemm_f32(bl_ctx *ctx, int transA, int transB, int m, int n, int k,
                const float *A, int lda, const float *B, int ldb, float *C, int ldc,
                int accumulate, int mode);

/* ------------------------------------------------------------------ LSTM / BLSTM layer
 * Replaces LstmLayer<Gpu>'s constructor buffers (LstmLayer.cu:543-619), computeForwardPass (:763-886:
 * 8 projection SGEMMs, 2*(T-1)*4 recurrent SGEMMs, 2*T ComputeBlockOutputFn launches, ResortOutputsFn)
 * and computeBackwardPass (:888-1051: ResortOutputErrorsFn, 2*(T-1)*4 SGEMMs + 2*T ComputeBlockErrorsFn,
 * 8 input-error SGEMMs, ComputeWeightUpdateFn).
 *   P = preceding layer size, L = layer size (both directions), S = parallel sequences. */
int    bl_lstm_plan_create(bl_ctx *ctx, int P, int L, int bidirectional, int S, int maxT, float bias, bl_lstm_plan **out);
void   bl_lstm_plan_destroy(bl_lstm_plan *plan);
size_t bl_lstm_num_weights(int P, int L, int bidirectional);          /* LstmLayer.cu:525 */

/* X [T*S][ldx] preceding outputs, patTypes [T*S] chars, Y [T*S][ldy] layer outputs ([fw H | bw H] per row). */
int bl_lstm_forward(bl_lstm_plan *plan, const float *W, const float *X, int ldx, const char *patTypes,
                    int T, int Tmin, float *Y, int ldy);
/* dY [T*S][lddy] this layer's outputErrors (for the unidirectional layer it is updated in place with the
 * recurrent error terms, as the reference's vector swap does, LstmLayer.cu:907-910); dX [T*S][lddx] =
 * preceding layer's outputErrors or NULL when that layer is not trainable (LstmLayer.cu:991-992);
 * dW = weightUpdates (gradient sums, same layout as W).  Must follow bl_lstm_forward on the same fraction: X, Y and the
 * input weights must still hold what the forward pass read and wrote (the tensor-core path reuses the forward pass's
 * TF32 split of X when the same X pointer, ldx and T are passed). */
int bl_lstm_backward(bl_lstm_plan *plan, const float *W, const float *X, int ldx, const float *Y, int ldy,
                     float *dY, int lddy, const char *patTypes, int T, int Tmin,
                     float *dX, int lddx, float *dW);
/* Gathers an internal tensor into the reference's [T*S][H] layout (LstmLayer.hpp:169-232 accessors):
 * which: 0 cellStates 1 cellStateErrors 2 niActs 3 igActs 4 fgActs 5 ogActs 6 niDeltas 7 igDeltas 8 fgDeltas 9 ogDeltas */
int bl_lstm_get_internal(bl_lstm_plan *plan, int dir, int which, int T, float *dst);
/* Tuning aid (BLSTM_REC_TRACE=1 at plan creation): per-CTA, per-step clock64 stamps of the forward persistent kernel,
 * [rows][T][6] = {step start, counter seen, exchange copied, GEMM done, gate math done, published}. */
int bl_lstm_debug_trace(bl_lstm_plan *plan, int T, long long *host_dst, int *rows);
/* The same for the second-generation tensor-memory kernels, forward (backward = 0) or BPTT (1): [rows][T][8], rows = CTAs x
 * sub-groups.  Forward: {step start, exchange polled, MMAs complete, accumulator staged, gate math done, exchange word stored,
 * control thread: first K-block ready, control thread: MMAs issued}; BPTT: {step start, partials polled, deltas in the B tile,
 * MMAs complete, partials stored, -, control thread: B tile ready, control thread: MMAs issued}. */
int bl_lstm_debug_trace2(bl_lstm_plan *plan, int backward, int T, long long *host_dst, int *rows);
/* Launch geometry chosen for the persistent kernels: out[0..3] = fwd {G seq groups, C cell slices, cells/CTA, smem bytes},
 * out[4..7] = bwd likewise. */
int bl_lstm_plan_info(const bl_lstm_plan *plan, int *out8);
/* The geometry the second-generation tensor-memory kernels would get for a layer of H cells per direction, S parallel sequences and
 * ndir directions on a device with num_sms SMs and smem_cap bytes of opt-in shared memory -- pure host arithmetic, no device needed
 * (the CPU tests sweep it).  Returns 1 and fills out12 = {G, C, CL, SG, threads, sub-groups per CTA, padded K or R, sequences per
 * sub-group, W_lo' in shared memory, shared memory bytes, CTAs, exchange words} or returns 0 when these kernels do not fit the layer. */
int bl_lstm_tm2_geometry(int backward, int H, int S, int ndir, int num_sms, int smem_cap, long long *out12);

/* ------------------------------------------------------------------ feed-forward / softmax layers
 * FeedForwardLayer<Gpu,TActFn>::computeForwardPass (FeedForwardLayer.cu:143-172): Y = act(W^T X + bias*b) for N slots */
int bl_ff_forward(bl_ctx *ctx, int act, int P, int O, int N, float bias, const float *W,
                  const float *X, int ldx, float *Y, int ldy);
/* ...::computeBackwardPass (FeedForwardLayer.cu:174-224): dY <- act'(Y)*dY in place; dX = W*dY (NULL to skip);
 * dW = [X*dY^T | bias*sum_n dY] */
int bl_ff_backward(bl_ctx *ctx, int act, int P, int O, int N, float bias, const float *W,
                   const float *X, int ldx, const float *Y, int ldy, float *dY, int lddy,
                   float *dX, int lddx, float *dW);
/* SoftmaxLayer's own part of the passes (SoftmaxLayer.cu:263-313 and :328-348), in place, padded patterns untouched */
int bl_softmax_forward(bl_ctx *ctx, int O, int N, const char *patTypes, float *Y, int ldy);
int bl_softmax_backward(bl_ctx *ctx, int O, int N, const char *patTypes, const float *Y, int ldy, float *dY, int lddy);

/* ------------------------------------------------------------------ post-output layers
 * MulticlassClassificationLayer: calculateError + countCorrectClassifications in one pass
 * (MulticlassClassificationLayer.cu:159-177, 195-214); results land in DEVICE scalars. */
int bl_multiclass_error(bl_ctx *ctx, int O, int N, const int *targetClasses, const float *Y, int ldy,
                        float *d_error, int *d_correct);
/* ...::computeBackwardPass (:221-240): zero-fill, then dY[n,target] = -1/max(FLT_MIN,y) */
int bl_multiclass_backward(bl_ctx *ctx, int O, int N, const int *targetClasses, const float *Y, int ldy,
                           float *dY, int lddy);
/* CePostOutputLayer (CePostOutputLayer.cu:125-166) and SsePostOutputLayer (SsePostOutputLayer.cu:114-155) */
int bl_ce_error(bl_ctx *ctx, int O, int N, const char *patTypes, const float *targets, int ldt,
                const float *Y, int ldy, float *d_error);
int bl_ce_backward(bl_ctx *ctx, int O, int N, const char *patTypes, const float *targets, int ldt,
                   const float *Y, int ldy, float *dY, int lddy);
int bl_sse_error(bl_ctx *ctx, int O, int N, const char *patTypes, const float *targets, int ldt,
                 const float *Y, int ldy, float *d_error);
int bl_sse_backward(bl_ctx *ctx, int O, int N, const char *patTypes, const float *targets, int ldt,
                    const float *Y, int ldy, float *dY, int lddy);
/* WeightedSsePostOutputLayer (WeightedSsePostOutputLayer.cu:121-164) and SseMaskPostOutputLayer, type "wf"
 * (SseMaskPostOutputLayer.cu:121-164).  O = size of the OUTPUT layer; a target row holds O (target, weight | filter
 * input) pairs, so ldt >= 2*O.  weightedsse: 1/2 sum ((y-t)*w)^2, dY = (y-t)*w; wf: 1/2 sum (y*f-t)^2, dY = (y*f-t)*f */
int bl_weightedsse_error(bl_ctx *ctx, int O, int N, const char *patTypes, const float *targets, int ldt,
                         const float *Y, int ldy, float *d_error);
int bl_weightedsse_backward(bl_ctx *ctx, int O, int N, const char *patTypes, const float *targets, int ldt,
                            const float *Y, int ldy, float *dY, int lddy);
int bl_ssemask_error(bl_ctx *ctx, int O, int N, const char *patTypes, const float *targets, int ldt,
                     const float *Y, int ldy, float *d_error);
int bl_ssemask_backward(bl_ctx *ctx, int O, int N, const char *patTypes, const float *targets, int ldt,
                        const float *Y, int ldy, float *dY, int lddy);
/* RmsePostOutputLayer: computeForwardPass fills rmses[N] = sqrt(mean_i (y-t)^2), 0 for padded patterns
 * (RmsePostOutputLayer.cu:39-67, 137-153); calculateError sums them (:126-134); computeBackwardPass writes
 * dY = rmses[pattern]*(y-t) for every entry (:75-93, 155-170) */
int bl_rmse_forward(bl_ctx *ctx, int O, int N, const char *patTypes, const float *targets, int ldt,
                    const float *Y, int ldy, float *rmses);
int bl_rmse_error(bl_ctx *ctx, int N, const float *rmses, float *d_error);
int bl_rmse_backward(bl_ctx *ctx, int O, int N, const float *rmses, const float *targets, int ldt,
                     const float *Y, int ldy, float *dY, int lddy);
/* BinaryClassificationLayer (one output per pattern, targets = the 0/1 target classes as reals, -1 for padding):
 * calculateError = -sum log(t>0 ? a : 1-a), a = max(y, FLT_MIN), and countCorrectClassifications ((t>0.5)==(y>0.5)) in
 * one pass (BinaryClassificationLayer.cu:43-84, 132-181); computeBackwardPass dY = -/+ 1/p, padded patterns untouched
 * (:86-112, 188-203) */
int bl_binary_error(bl_ctx *ctx, int N, const char *patTypes, const float *targets, int ldt, const float *Y, int ldy,
                    float *d_error, int *d_correct);
int bl_binary_backward(bl_ctx *ctx, int N, const char *patTypes, const float *targets, int ldt, const float *Y, int ldy,
                       float *dY, int lddy);

/* Elementwise evaluation of the scalar functions every kernel shares (parity tests pin them bit-for-bit against
 * the reference's functors): which = 0 Logistic::fn (Logistic.cuh:33-43), 1 Tanh::fn (Tanh.cuh:33-36),
 * 2 safeExp (safeExp.cuh:32-40), 3 limitedError (limitedError.cuh:31-34).  x, y device arrays of n floats. */
int bl_eval_scalar_fn(bl_ctx *ctx, int which, size_t n, const float *x, float *y);

/* ------------------------------------------------------------------ optimizer step
 * UpdateWeightFn (optimizers/SteepestDescentOptimizer.cu:39-59): delta = momentum*delta - lr*grad; w += delta */
int bl_sgd_update(bl_ctx *ctx, size_t n, float learningRate, float momentum, float *W, const float *dW, float *deltas);
/* batch-mode gradient accumulation over fractions, thrust::transform(plus) of optimizers/Optimizer.cu:77-80: y += x */
int bl_vector_add(bl_ctx *ctx, size_t n, const float *x, float *y);
/* TrainableLayer::injectWeightNoise (layers/TrainableLayer.cu:188-209): w[e] += sigma * N(0,1).  The reference draws from a
 * host mt19937 and copies the noise over; here element e of the call gets a counter-based draw from (seed, offset + e), so the
 * same (seed, offset) gives the same noise on every rank.  The stream is NOT the reference's (documented in DESIGN.md). */
int bl_add_gaussian_noise(bl_ctx *ctx, size_t n, float sigma, unsigned long long seed, unsigned long long offset, float *w);

/* ------------------------------------------------------------------ data parallelism (new)
 * One communicator per process/GPU.  `unique_id` is the 128-byte ncclUniqueId produced by rank 0
 * (bl_comm_unique_id) and distributed by the launcher (bench.py uses torch.distributed for that). */
int  bl_comm_unique_id(void *id128);
int  bl_comm_create(bl_ctx *ctx, int rank, int world, const void *id128, bl_comm **out);
void bl_comm_destroy(bl_comm *comm);
/* rank / world size of the communicator (either pointer may be NULL) */
void bl_comm_info(const bl_comm *comm, int *rank, int *world);
/* In-place sum over ranks of `count` floats.  Default schedule (BLSTM_COMM_MODE unset or "overlap"): issued at once on the communicator's
 * side stream, ordered after the work already enqueued on the context's stream, through a communicator capped to BLSTM_COMM_MAX_CTAS
 * CTAs (default 4: the SMs the persistent recurrent kernels leave free), so that it overlaps with the backward pass of the layers
 * below.  BLSTM_COMM_MODE=grouped: the request is queued and all queued buffers are reduced by one grouped NCCL call in
 * bl_comm_join.  Returns immediately. */
int  bl_allreduce_sum_f32(bl_comm *comm, float *buf, size_t count);
/* The same for the LAST request before the join (nothing is left to overlap with): full-width communicator, context's stream. */
int  bl_allreduce_sum_f32_last(bl_comm *comm, float *buf, size_t count);
/* Completes (in stream order) every all-reduce requested so far: the buffers hold the sums for work enqueued afterwards. */
int  bl_comm_join(bl_comm *comm);

#ifdef __cplusplus
}
#endif
#endif /* BLSTM_B200_H */
