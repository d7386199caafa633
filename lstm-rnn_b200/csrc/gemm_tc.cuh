// Internal interface of the tcgen05 GEMM path (gemm_tc.cu).  An operand is a row-major fp32 matrix prepared once (16-byte
// aligned rows, hi/lo TF32 split in strict mode); GEMMs then take VIEWS of it: either K-major (rows are the M/N index, columns
// the contraction index) or MN-major (the transposed reading of the same memory), with a sub-view offset in both dimensions.
// The LSTM backward pass reads one split of the deltas both ways: K-major for the input error, MN-major for the weight gradients.
#pragma once
#include "common.cuh"

namespace bl {

struct TcOperand {
    const float *hi;        // [rows][ld]; in fast mode the operand itself
    const float *lo;        // strict mode: x - tf32(x); else NULL
    size_t ld;
    int rows, cols;
    bool strict;
};

struct TcView {
    const TcOperand *op;
    bool mn_major;          // false: MN = row, K = column;  true: MN = column, K = row
    int mn0, k0;            // sub-view origin; the offset along the contiguous dimension (columns) must be a multiple of 4
};

size_t tc_operand_ld(int cols);
// floats needed for hi (and for lo) of a prepared [rows][cols] matrix with optional row / column re-blocking (bw -> bwp, 0 = none)
size_t tc_operand_elems(int rows, int cols, int rbw = 0, int rbwp = 0, int cbw = 0, int cbwp = 0);
int tc_prepare(bl_ctx *ctx, const float *src, int rows, int cols, size_t ld_src, bool strict, float *hi, float *lo, TcOperand *out,
               int rbw = 0, int rbwp = 0, int cbw = 0, int cbwp = 0);
int tc_gemm(bl_ctx *ctx, int M, int N, int K, const TcView &A, const TcView &B, float *C, int ldc, int accumulate);
int tc_gemm_batched(bl_ctx *ctx, int M, int N, int K, const TcView &A, const TcView &B, float *C, int ldc, int accumulate,
                    int batches, int a_batch_mn, long long c_batch_stride);
int gemm_tf32_tc(bl_ctx *ctx, int M, int N, int K, const float *A, size_t lda, bool a_kmajor, const float *B, size_t ldb, bool b_kmajor,
                 float *C, int ldc, int accumulate, int mode);
// true when bl_gemm_f32 would route an m x n x k contraction to the tensor-core path
bool tc_wanted(const bl_ctx *ctx, int m, int n, int k);

} // namespace bl
