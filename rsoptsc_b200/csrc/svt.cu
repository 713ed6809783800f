// a4: singular value thresholding  J = U max(S - tau, 0) V^T  (R/SimilarityM.R:142-155).
//
// The reference takes a full SVD of the n x n matrix A = Z - Y3/mu.  Here the same J is formed
// from the Gram matrix:  G = A^T A = V diag(lambda) V^T,  sigma = sqrt(lambda),
//     J = A V_r diag(1 - tau/sigma_r) V_r^T          (r = #{sigma > tau})
// i.e. one syrk-shaped contraction, one symmetric eigen-decomposition and two GEMMs whose inner
// dimension shrinks with the rank of J.  In FP64 this matches the SVD route to ~1e-15 relative
// in Z (scratch/proto_admm.py); FP32-level Gram accuracy does NOT (7e-7), because squaring
// buries singular values below sqrt(eps)*sigma_1 -- hence the FP64(-equivalent) requirement, met
// on the tensor cores by the split-integer GEMM of gemm_ozaki.cu.
#include <cmath>

#include "comm.h"
#include "gemm.h"
#include "internal.h"

namespace rs {

void SvtWorkspace::ensure(int64_t n_) {
  if (n == n_ && G.p) return;
  n = n_;
  G.alloc((size_t)n * n);
  T.alloc((size_t)n * n);
  lam.alloc(n);
}

int64_t svt_inplace(SvtWorkspace& ws, double* d_A, int64_t n, double tau) {
  RS_REQUIRE(n <= 65535, "svt: block larger than 65,535 cells");
  ws.ensure(n);
  Context& c = ctx();
  {  // G = A^T A (lower triangle)
    Phase ph("gram");
    gram_AtA(d_A, n, n, ws.G.p);
  }
  {  // G -> V (columns), lam ascending
    Phase ph("eig");
    // only the eigenvectors with sqrt(lambda) > tau enter J (margin: the count below is re-derived from lambda)
    sym_eig(ws.G.p, n, ws.lam.p, /*vectors=*/true, tau * tau * (1.0 - 1e-9));
  }
  std::vector<double> lam(n);
  RS_CUDA(cudaMemcpyAsync(lam.data(), ws.lam.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
  RS_CUDA(cudaStreamSynchronize(c.stream));
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
  const double* Vr = ws.G.p + r0 * n;
  Phase ph("recon");
  if (3 * r > 2 * n && n >= 1536) {
    // nearly full rank (saturated phase): J = A B with the SYMMETRIC B = V_r h V_r^T = W W^T,
    // W = V_r diag(sqrt h): one half-GEMM (n^2 r) + one GEMM (2 n^3) instead of two GEMMs (4 n^2 r)
    for (int64_t i = r0; i < n; ++i) h[i] = std::sqrt(h[i]);
    RS_CUDA(cudaMemcpyAsync(ws.lam.p + r0, h.data() + r0, sizeof(double) * r, cudaMemcpyHostToDevice, c.stream));
    copy_scale_columns(Vr, n, r, ws.lam.p + r0, ws.T.p);   // W                     (n x r)
    ws.B.alloc((size_t)n * n);
    gram_AAt(ws.T.p, n, r, ws.B.p);                        // B = W W^T  (lower)    (n x n)
    mirror_lower(ws.B.p, n);
    gemm_nn(d_A, ws.B.p, ws.G.p, n, n, n);                 // J = A B, into G (V is no longer needed)
    RS_CUDA(cudaMemcpyAsync(d_A, ws.G.p, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, c.stream));
    ws.B.release();
  } else {
    RS_CUDA(cudaMemcpyAsync(ws.lam.p + r0, h.data() + r0, sizeof(double) * r, cudaMemcpyHostToDevice, c.stream));
    gemm_nn(d_A, Vr, ws.T.p, n, r, n);                 // T = A V_r          (n x r)
    scale_columns(ws.T.p, n, r, ws.lam.p + r0);        // T *= diag(h_r)
    gemm_nt(ws.T.p, Vr, d_A, n, n, r);                 // J = T V_r^T        (n x n), over A
  }
  return r;
}

// ---- collective SVT of a block replicated on every rank (multi-GPU) ----------
//   Gram           every rank forms one column panel of the lower triangle of G = A^T A (panels of
//                  equal area), the panels are all-gathered over NVLink
//   eigensolve     sym_eig(collective): fused multi-GPU tridiagonalisation, divide-and-conquer merges
//                  and back-transformation split by columns
//   reconstruction rank q forms its own columns of J only:  J[:, q] = A (V_r h V_r[q, :]^T)
// so J never travels: the column-local ADMM kernels of rank q read exactly these columns.
int64_t svt_sharded(SvtWorkspace& ws, double* d_A, int64_t n, double tau, const std::vector<int64_t>& cb) {
  RS_REQUIRE(n <= 65535, "svt: block larger than 65,535 cells");
  ws.ensure(n);
  Context& c = ctx();
  const Comm& cm = comm();
  const int P = cm.nranks, rank = cm.rank;
  {  // G = A^T A, lower triangle, one panel per rank
    Phase ph("gram");
    const std::vector<int64_t> gb = split_lower_triangle(n, P, 128);
    const int64_t g0 = gb[rank], g1 = gb[rank + 1];
    if (g1 > g0)  // rows g0.. of the panel: G[g0:, g0:g1] = A[:, g0:]^T A[:, g0:g1]
      gemm_tn_ld(d_A + (size_t)g0 * n, n, d_A + (size_t)g0 * n, n, ws.G.p + g0 + (size_t)g0 * n, n, n - g0, g1 - g0, n);
    comm_allgather_cols(ws.G.p, n, gb);
  }
  {
    Phase ph("eig");
    sym_eig(ws.G.p, n, ws.lam.p, /*vectors=*/true, tau * tau * (1.0 - 1e-9), /*collective=*/true);
  }
  std::vector<double> lam(n);
  RS_CUDA(cudaMemcpyAsync(lam.data(), ws.lam.p, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
  RS_CUDA(cudaStreamSynchronize(c.stream));
  std::vector<double> h(n, 0.0);
  int64_t r0 = n;
  for (int64_t i = n - 1; i >= 0; --i) {
    const double sig = lam[i] > 0 ? std::sqrt(lam[i]) : 0.0;
    if (sig > tau) { h[i] = 1.0 - tau / sig; r0 = i; } else break;
  }
  const int64_t r = n - r0;
  const int64_t c0 = cb[rank], c1 = cb[rank + 1], w = c1 - c0;
  if (r == 0) {
    if (w > 0) RS_CUDA(cudaMemsetAsync(d_A + (size_t)c0 * n, 0, sizeof(double) * n * w, c.stream));
    return 0;
  }
  Phase ph("recon");
  if (w > 0) {
    const double* Vr = ws.G.p + (size_t)r0 * n;
    RS_CUDA(cudaMemcpyAsync(ws.lam.p + r0, h.data() + r0, sizeof(double) * r, cudaMemcpyHostToDevice, c.stream));
    copy_scale_columns(Vr, n, r, ws.lam.p + r0, ws.T.p);                  // T = V_r diag(h)          (n x r)
    DevBuf<double> Bq((size_t)n * w), Jq((size_t)n * w);
    gemm_nt_ld(ws.T.p, n, Vr + c0, n, Bq.p, n, n, w, r);                  // B[:, q] = T V_r[q, :]^T  (n x w)
    gemm_nn_ld(d_A, n, Bq.p, n, Jq.p, n, n, w, n);                        // J[:, q] = A B[:, q]
    RS_CUDA(cudaMemcpyAsync(d_A + (size_t)c0 * n, Jq.p, sizeof(double) * n * w, cudaMemcpyDeviceToDevice, c.stream));
  }
  return r;
}

}  // namespace rs
