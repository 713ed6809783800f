// Back-transformation of the symmetric eigensolver: Z <- Q Z with Q = H(0) H(1) ... H(n-2), the
// reflectors of sytrd_kernel.cuh (LAPACK dsytrd('L') storage).  Third stage of the eigensolver
// behind the SVT (a4, R/SimilarityM.R:142-155).
//
// Blocked compact-WY form: for a block of nb reflectors, H(j0) ... H(j0+nb-1) = I - V T V^T with
//     T^-1 = striu(V^T V) + diag(1 / tau)            (upper triangular)
// so one block costs a Gram (V^T V), a triangular solve and two panel GEMMs
//     X = V^T Z,   Y = -T X  (trsm),   Z += V Y,
// both on the tcgen05 split-integer path (gemm_ozaki.cu) when the block is large enough.  2 n^3 flop in
// total, all of it in GEMMs; no 32-bit workspace sizes, so blocks above 32,768 rows work too.
#include "gemm.h"
#include "internal.h"

namespace rs {

namespace {

// Vb (m x nb, ld m): column jj = reflector j0 + jj restricted to rows r0 = j0 + 1 .. n - 1:
// zeros above its unit entry (row jj), the stored tail below.
__global__ void k_extract_V(const double* __restrict__ A, long long lda, int j0, int nb, int m, double* __restrict__ Vb) {
  const int jj = blockIdx.y;
  const double* src = A + (long long)(j0 + 1) + (long long)(j0 + jj) * lda;
  double* dst = Vb + (size_t)jj * m;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x)
    dst[i] = (i < jj) ? 0.0 : (i == jj ? 1.0 : src[i]);
}

// U = striu(S) + diag(1 / tau), S = V^T V given in its lower triangle; strictly lower part of U zeroed
__global__ void k_build_U(const double* __restrict__ S, const double* __restrict__ tau, int nb, double* __restrict__ U) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= nb) return;
  double v = 0.0;
  if (i < j) v = S[j + (size_t)i * nb];  // S(i, j) = S(j, i)
  else if (i == j) v = (tau[j] != 0.0) ? 1.0 / tau[j] : 1e300;  // tau = 0: H = I
  U[i + (size_t)j * nb] = v;
}

}  // namespace

void apply_q_left(const double* d_A, int64_t lda, const double* d_tau, int64_t n, double* d_Z, int64_t ncols) {
  Context& c = ctx();
  if (n < 2 || ncols < 1) return;
  const int64_t nref = n - 1;                              // reflectors 0 .. n-2
  const int64_t nb = n >= 6000 ? 1024 : (n >= 1536 ? 512 : 128);
  const int64_t nblk = (nref + nb - 1) / nb;
  DevBuf<double> Vb((size_t)n * nb), S((size_t)nb * nb), U((size_t)nb * nb), X((size_t)nb * ncols);
  for (int64_t k = nblk - 1; k >= 0; --k) {
    const int64_t j0 = k * nb, nbk = std::min(nb, nref - j0), r0 = j0 + 1, m = n - r0;
    dim3 ge((unsigned)std::min<int64_t>(cdiv(m, 256), 256), (unsigned)nbk);
    {
      Phase ph("bt_gram");
      RS_LAUNCH(k_extract_V, ge, 256, 0, d_A, (long long)lda, (int)j0, (int)nbk, (int)m, Vb.p);
      gram_AtA(Vb.p, m, nbk, S.p);  // lower triangle of V^T V
      dim3 gu((unsigned)cdiv(nbk, 128), (unsigned)nbk);
      RS_LAUNCH(k_build_U, gu, 128, 0, S.p, d_tau + j0, (int)nbk, U.p);
    }
    double* Zr = d_Z + r0;
    {
      Phase ph("bt_vtz");
      gemm_tn_ld(Vb.p, m, Zr, n, X.p, nbk, nbk, ncols, m);  // X = V^T Z            (nbk x ncols)
    }
    {
      Phase ph("bt_trsm");
      const double mone = -1.0;
      RS_CUBLAS(cublasDtrsm(c.cublas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, (int)nbk,
                            (int)ncols, &mone, U.p, (int)nbk, X.p, (int)nbk));  // X <- -T X
    }
    Phase ph("bt_update");
    gemm_nn_ld_acc(Vb.p, m, X.p, nbk, Zr, n, m, ncols, nbk);  // Z += V (-T V^T Z)
  }
}

}  // namespace rs
