// a4: singular value thresholding  J = U max(S - tau, 0) V^T  (R/SimilarityM.R:142-155).
//
// The reference takes a full SVD of the n x n matrix A = Z - Y3/mu.  Here the same J is formed
// from the Gram matrix:  G = A^T A = V diag(lambda) V^T,  sigma = sqrt(lambda),
//     J = A V_r diag(1 - tau/sigma_r) V_r^T          (r = #{sigma > tau})
// i.e. one syrk-shaped contraction, one symmetric eigen-decomposition and two GEMMs whose inner
// dimension shrinks with the rank of J.  In FP64 this matches the SVD route to ~1e-15 relative
// in Z (scratch/proto_admm.py); FP32-level Gram accuracy does NOT (7e-7), because squaring
// buries singular values below sqrt(eps)*sigma_1 -- hence the FP64(-equivalent) requirement.
#include <cmath>

#include "internal.h"
#include "gemm.h"

namespace rs {

void SvtWorkspace::ensure(int64_t n_) {
  if (n == n_ && G.p) return;
  n = n_;
  G.alloc((size_t)n * n);
  T.alloc((size_t)n * n);
  lam.alloc(n);
  info.alloc(1);
  RS_CUSOLVER(cusolverDnDsyevd_bufferSize(ctx().cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, G.p,
                                          (int)n, lam.p, &lwork));
  work.alloc((size_t)lwork);
}

int64_t svt_inplace(SvtWorkspace& ws, double* d_A, int64_t n, double tau) {
  RS_REQUIRE(n <= 46340, "svt: n too large for the 32-bit dense solver interface");
  ws.ensure(n);
  Context& c = ctx();
  {  // G = A^T A (lower triangle)
    Phase ph("gram");
    gram_AtA(d_A, n, n, ws.G.p);
  }
  {  // G -> V (columns), lam ascending
    Phase ph("eig");
    RS_CUSOLVER(cusolverDnDsyevd(c.cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)n, ws.G.p, (int)n,
                                 ws.lam.p, ws.work.p, ws.lwork, ws.info.p));
  }
  std::vector<double> lam(n);
  int info = 0;
  RS_CUDA(cudaMemcpyAsync(lam.data(), ws.lam.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
  RS_CUDA(cudaMemcpyAsync(&info, ws.info.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
  RS_CUDA(cudaStreamSynchronize(c.stream));
  RS_REQUIRE(info == 0, "svt: symmetric eigensolver did not converge (info=" + std::to_string(info) + ")");
  // shrink factors; eigenvalues ascending so the kept ones are a suffix
  std::vector<double> h(n, 0.0);
  int64_t r0 = n;
  for (int64_t i = n - 1; i >= 0; --i) {
    const double sig = lam[i] > 0 ? std::sqrt(lam[i]) : 0.0;
    if (sig > tau) { h[i] = 1.0 - tau / sig; r0 = i; } else break;
  }
  const int64_t r = n - r0;
  if (r == 0) {
    RS_CUDA(cudaMemsetAsync(d_A, 0, sizeof(double) * n * n, c.stream));
    return 0;
  }
  RS_CUDA(cudaMemcpyAsync(ws.lam.p + r0, h.data() + r0, sizeof(double) * r, cudaMemcpyHostToDevice, c.stream));
  const double* Vr = ws.G.p + r0 * n;
  {
    Phase ph("recon");
    gemm_nn(d_A, Vr, ws.T.p, n, r, n);                 // T = A V_r          (n x r)
    scale_columns(ws.T.p, n, r, ws.lam.p + r0);        // T *= diag(h_r)
    gemm_nt(ws.T.p, Vr, d_A, n, n, r);                 // J = T V_r^T        (n x n), over A
  }
  return r;
}

}  // namespace rs
