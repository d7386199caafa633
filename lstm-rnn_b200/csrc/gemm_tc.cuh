// Internal interface of the tcgen05 GEMM path (gemm_tc.cu): operands are prepared once (K-major, 16-byte aligned rows,
// hi/lo TF32 split in strict mode) and can then feed several GEMMs as row / K sub-views -- the LSTM backward pass reuses
// the transposed deltas and outputs of a layer for its input-weight and all eight recurrent-weight gradient blocks.
#pragma once
#include "common.cuh"

namespace bl {

struct TcOperand {
    const float *hi;        // [rows][ld] K-major; in fast mode the operand itself
    const float *lo;        // strict mode: x - tf32(x); else NULL
    size_t ld;
    int rows, K;
    bool strict;
};

size_t tc_operand_ld(int K);
int tc_prepare(bl_ctx *ctx, const float *src, int rows, int K, size_t ld_src, bool kmajor, bool strict, float *hi, float *lo, TcOperand *out);
int tc_gemm(bl_ctx *ctx, int M, int N, int K, const TcOperand &A, int a_row0, int a_k0, const TcOperand &B, int b_row0, int b_k0,
            float *C, int ldc, int accumulate);
int tc_gemm_batched(bl_ctx *ctx, int M, int N, int K, const TcOperand &A, int a_row0, int a_k0, const TcOperand &B, int b_row0, int b_k0,
                    float *C, int ldc, int accumulate, int batches, int a_batch_rows, long long c_batch_stride);
int gemm_tf32_tc(bl_ctx *ctx, int M, int N, int K, const float *A, size_t lda, bool a_kmajor, const float *B, size_t ldb, bool b_kmajor,
                 float *C, int ldc, int accumulate, int mode);
// true when bl_gemm_f32 would route an m x n x k contraction to the tensor-core path
bool tc_wanted(const bl_ctx *ctx, int m, int n, int k);

} // namespace bl
