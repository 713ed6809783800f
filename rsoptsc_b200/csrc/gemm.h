// Dense FP64(-equivalent) contractions of the path (gemm.cu).  All operands column-major,
// leading dimension == row count.
#pragma once
#include "common.h"

namespace rs {
int gemm_backend();            // 0 cuBLAS FP64, 1 tcgen05 split-integer (Ozaki)
void set_gemm_backend(int b);
// G (cols x cols, lower triangle valid) = A^T A,  A: rows x cols
void gram_AtA(const double* A, int64_t rows, int64_t cols, double* G);
// G (rows x rows, lower triangle valid) = A A^T
void gram_AAt(const double* A, int64_t rows, int64_t cols, double* G);
// C (M x N) = A (M x K) * B (K x N)
void gemm_nn(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K);
// same with explicit leading dimensions (sub-blocks of larger matrices)
void gemm_nn_ld(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t N,
                int64_t K);
// C (M x N) += A (M x K) * B (K x N), explicit leading dimensions
void gemm_nn_ld_acc(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                    int64_t N, int64_t K);
// C (M x N) = A^T * B, A: K x M, B: K x N, explicit leading dimensions
void gemm_tn_ld(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t N,
                int64_t K);
// C (M x N) = A (M x K) * B^T, B: N x K
void gemm_nt(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K);
// same with explicit leading dimensions
void gemm_nt_ld(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t N,
                int64_t K);
// C (M x N) = A^T * B, A: K x M, B: K x N
void gemm_tn(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K);
// C (nr x nr, lower triangle) -= V W^T + W V^T, V and W nr x k (trailing update of the tridiagonalisation).
// nranks > 1: only the 128-row groups (absolute rows, the block starts at row r0) of rank `rank` and the first 64
// columns are updated -- the rows the fused multi-GPU tridiagonalisation lets that rank read (sytrd_mg.cuh).
void syr2k_lower_sub(const double* V, int64_t ldv, const double* W, int64_t ldw, int64_t nr, int k, double* C, int64_t ldc,
                     int64_t r0 = 0, int nranks = 1, int rank = 0);
}  // namespace rs
