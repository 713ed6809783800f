// Dense contractions.  Round-1 baseline: cuBLAS FP64 (DSYRK/DGEMM, ~35 TFLOP/s measured on
// B200).  The tcgen05 split-integer path (gemm_ozaki.cu) replaces these where enabled.
#include "gemm.h"
#include "internal.h"

namespace rs {

void gram_AtA(const double* A, int64_t rows, int64_t cols, double* G) {
  const double one = 1.0, zero = 0.0;
  RS_CUBLAS(cublasDsyrk(ctx().cublas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, (int)cols, (int)rows, &one, A, (int)rows,
                        &zero, G, (int)cols));
}
void gram_AAt(const double* A, int64_t rows, int64_t cols, double* G) {
  const double one = 1.0, zero = 0.0;
  RS_CUBLAS(cublasDsyrk(ctx().cublas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, (int)rows, (int)cols, &one, A, (int)rows,
                        &zero, G, (int)rows));
}
void gemm_nn(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K) {
  const double one = 1.0, zero = 0.0;
  RS_CUBLAS(cublasDgemm(ctx().cublas, CUBLAS_OP_N, CUBLAS_OP_N, (int)M, (int)N, (int)K, &one, A, (int)M, B, (int)K,
                        &zero, C, (int)M));
}
void gemm_nt(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K) {
  const double one = 1.0, zero = 0.0;
  RS_CUBLAS(cublasDgemm(ctx().cublas, CUBLAS_OP_N, CUBLAS_OP_T, (int)M, (int)N, (int)K, &one, A, (int)M, B, (int)N,
                        &zero, C, (int)M));
}
void gemm_tn(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K) {
  const double one = 1.0, zero = 0.0;
  RS_CUBLAS(cublasDgemm(ctx().cublas, CUBLAS_OP_T, CUBLAS_OP_N, (int)M, (int)N, (int)K, &one, A, (int)K, B, (int)K,
                        &zero, C, (int)M));
}

}  // namespace rs
