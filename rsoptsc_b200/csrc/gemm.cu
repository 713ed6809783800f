// Dense FP64(-equivalent) contractions of the path: Gram A^T A and the two reconstruction
// GEMMs of the SVT (a4), the centred Gram of the PCA (a2).
//   backend "ozaki"  : tcgen05 kind::i8 split-integer GEMM (gemm_ozaki.cu), FP64-equivalent
//   backend "cublas" : cuBLAS DSYRK/DGEMM (library baseline, ~35 TFLOP/s measured on B200)
// RSOPTSC_GEMM=cublas|ozaki selects; small problems always take the library call.
#include "gemm.h"
#include "internal.h"

namespace rs {

void ozaki_gram_AtA(const double* A, int64_t rows, int64_t cols, double* G);
void ozaki_gram_AAt(const double* A, int64_t rows, int64_t cols, double* G);
void ozaki_gemm_nn(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K);
void ozaki_gemm_nt(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K);
void ozaki_gemm_nn_ld(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                      int64_t N, int64_t K);
void ozaki_gemm_nn_ld_acc(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                          int64_t N, int64_t K);
void ozaki_gemm_tn_ld(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                      int64_t N, int64_t K);
void ozaki_gemm_tn(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K);
void ozaki_gemm_nt_ld(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                      int64_t N, int64_t K);

void ozaki_syr2k_lower_sub(const double* V, int64_t ldv, const double* W, int64_t ldw, int64_t nr, int k, double* C, int64_t ldc,
                           int64_t r0, int nranks, int rank);

static int g_backend = -1;  // 0 cublas, 1 ozaki
int gemm_backend() {
  if (g_backend < 0) {
    const char* e = getenv("RSOPTSC_GEMM");
    g_backend = (e && std::string(e) == "cublas") ? 0 : 1;  // default: the tcgen05 path
  }
  return g_backend;
}
void set_gemm_backend(int b) { g_backend = b; }

// backend 1: tcgen05 path where it pays (large contractions); backend 2: wherever it is legal
// (parity tests at small sizes).  Small problems are dominated by slicing + launch overheads.
static bool use_ozaki(int64_t M, int64_t N, int64_t K) {
  const int b = gemm_backend();
  if (b == 0 || M > 65535 || N > 65535) return false;
  const int64_t lim = b == 2 ? 256 : 1536;
  return M >= lim && N >= lim && K >= lim;
}

void gram_AtA(const double* A, int64_t rows, int64_t cols, double* G) {
  if (use_ozaki(cols, cols, rows)) return ozaki_gram_AtA(A, rows, cols, G);
  const double one = 1.0, zero = 0.0;
  RS_CUBLAS(cublasDsyrk(ctx().cublas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, (int)cols, (int)rows, &one, A, (int)rows,
                        &zero, G, (int)cols));
}
void gram_AAt(const double* A, int64_t rows, int64_t cols, double* G) {
  // few output tiles (rows < 4096 -> less than ~3 waves of 128 x 128 tiles) do not amortise the slicing
  if (rows >= 4096 && use_ozaki(rows, rows, cols)) return ozaki_gram_AAt(A, rows, cols, G);
  const double one = 1.0, zero = 0.0;
  RS_CUBLAS(cublasDsyrk(ctx().cublas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, (int)rows, (int)cols, &one, A, (int)rows,
                        &zero, G, (int)rows));
}
void gemm_nn(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K) {
  if (use_ozaki(M, N, K)) return ozaki_gemm_nn(A, B, C, M, N, K);
  const double one = 1.0, zero = 0.0;
  RS_CUBLAS(cublasDgemm(ctx().cublas, CUBLAS_OP_N, CUBLAS_OP_N, (int)M, (int)N, (int)K, &one, A, (int)M, B, (int)K,
                        &zero, C, (int)M));
}
void gemm_nn_ld(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t N,
                int64_t K) {
  if (use_ozaki(M, N, K)) return ozaki_gemm_nn_ld(A, lda, B, ldb, C, ldc, M, N, K);
  const double one = 1.0, zero = 0.0;
  RS_CUBLAS(cublasDgemm(ctx().cublas, CUBLAS_OP_N, CUBLAS_OP_N, (int)M, (int)N, (int)K, &one, A, (int)lda, B, (int)ldb,
                        &zero, C, (int)ldc));
}
// Panel-shaped contractions (one dimension a few hundred, the others thousands): the tensor-core path
// pays as soon as the operand slicing is amortised.
static bool use_ozaki_panel(int64_t M, int64_t N, int64_t K) {
  const int b = gemm_backend();
  if (b == 0 || M > 65535 || N > 65535) return false;
  const int64_t mn = std::min(M, std::min(N, K));
  return mn >= 256 && (double)M * (double)N * (double)K >= 1536.0 * 1536.0 * 1536.0;
}
void gemm_nn_ld_acc(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M,
                    int64_t N, int64_t K) {
  if (use_ozaki_panel(M, N, K)) return ozaki_gemm_nn_ld_acc(A, lda, B, ldb, C, ldc, M, N, K);
  const double one = 1.0;
  RS_CUBLAS(cublasDgemm(ctx().cublas, CUBLAS_OP_N, CUBLAS_OP_N, (int)M, (int)N, (int)K, &one, A, (int)lda, B, (int)ldb,
                        &one, C, (int)ldc));
}
void gemm_tn_ld(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t N,
                int64_t K) {
  if (use_ozaki_panel(M, N, K)) return ozaki_gemm_tn_ld(A, lda, B, ldb, C, ldc, M, N, K);
  const double one = 1.0, zero = 0.0;
  RS_CUBLAS(cublasDgemm(ctx().cublas, CUBLAS_OP_T, CUBLAS_OP_N, (int)M, (int)N, (int)K, &one, A, (int)lda, B, (int)ldb,
                        &zero, C, (int)ldc));
}
void gemm_nt(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K) {
  if (use_ozaki(M, N, K)) return ozaki_gemm_nt(A, B, C, M, N, K);
  const double one = 1.0, zero = 0.0;
  RS_CUBLAS(cublasDgemm(ctx().cublas, CUBLAS_OP_N, CUBLAS_OP_T, (int)M, (int)N, (int)K, &one, A, (int)M, B, (int)N,
                        &zero, C, (int)M));
}
void gemm_nt_ld(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, int64_t M, int64_t N,
                int64_t K) {
  if (use_ozaki_panel(M, N, K)) return ozaki_gemm_nt_ld(A, lda, B, ldb, C, ldc, M, N, K);
  const double one = 1.0, zero = 0.0;
  RS_CUBLAS(cublasDgemm(ctx().cublas, CUBLAS_OP_N, CUBLAS_OP_T, (int)M, (int)N, (int)K, &one, A, (int)lda, B, (int)ldb,
                        &zero, C, (int)ldc));
}
void syr2k_lower_sub(const double* V, int64_t ldv, const double* W, int64_t ldw, int64_t nr, int k, double* C, int64_t ldc,
                     int64_t r0, int nranks, int rank) {
  Phase ph("sytrd_syr2k");
  if (gemm_backend() != 0 && nr <= 65535) return ozaki_syr2k_lower_sub(V, ldv, W, ldw, nr, k, C, ldc, r0, nranks, rank);
  const double mone = -1.0, one = 1.0;
  RS_CUBLAS(cublasDsyr2k(ctx().cublas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, (int)nr, k, &mone, V, (int)ldv, W, (int)ldw, &one, C,
                         (int)ldc));
}
void gemm_tn(const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K) {
  if (use_ozaki(M, N, K)) return ozaki_gemm_tn(A, B, C, M, N, K);
  const double one = 1.0, zero = 0.0;
  RS_CUBLAS(cublasDgemm(ctx().cublas, CUBLAS_OP_T, CUBLAS_OP_N, (int)M, (int)N, (int)K, &one, A, (int)K, B, (int)K,
                        &zero, C, (int)M));
}

}  // namespace rs
