// a2: PCA spectrum of the cells (stats::prcomp(t(X))$sdev^2, R/SimilarityM.R:41-49), the K rule
// (R/SimilarityM.R:61-72) and, optionally, the first PCA score columns used as the deterministic
// in-library embedding.
//
// prcomp centres every gene (row of X) over the cells; the non-zero eigenvalues of the
// covariance are those of the smaller Gram matrix of the centred matrix (m x m or n x n).
#include <algorithm>
#include <cmath>

#include "gemm.h"
#include "internal.h"
#include "lanczos.h"

namespace rs {

__global__ void k_fill(double* __restrict__ v, int64_t n, double val) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = val;
}
__global__ void k_center_rows(const double* __restrict__ X, int64_t m, int64_t n, const double* __restrict__ mean,
                              double* __restrict__ Xc) {
  const int64_t total = m * n;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    Xc[e] = X[e] - mean[e % m];
}
// emb[:, c] = V[:, src_c] * scale_c     (n x ncomp)
__global__ void k_scaled_copy(const double* __restrict__ V, int64_t n, int64_t ldv, int col, double scale,
                              double* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = V[i + (int64_t)col * ldv] * scale;
}

void pca_spectrum(const double* d_X, int64_t m, int64_t n, std::vector<double>& eig, double* d_emb, int ncomp) {
  Phase ph("pca");
  Context& c = ctx();
  RS_REQUIRE(m >= 1 && n >= 2, "pca: need at least 2 cells");
  const int64_t d = std::min(m, n);
  RS_REQUIRE(d <= 65535, "pca: min(m, n) too large for the dense eigensolver");
  RS_REQUIRE(ncomp >= 0 && ncomp <= d, "pca: ncomp out of range");
  // row means: X * (1/n)
  DevBuf<double> ones(n), mean(m), Xc((size_t)m * n);
  RS_LAUNCH(k_fill, (unsigned)cdiv(n, 256), 256, 0, ones.p, n, 1.0 / (double)n);
  DenseMatvec mv;
  mv.init(d_X, m, n);
  mv.mul_n(ones.p, mean.p);
  RS_LAUNCH(k_center_rows, 148 * 8, 256, 0, d_X, m, n, mean.p, Xc.p);
  DevBuf<double> G((size_t)d * d), lam(d);
  const bool gene_side = m <= n;
  if (gene_side) gram_AAt(Xc.p, m, n, G.p); else gram_AtA(Xc.p, m, n, G.p);
  sym_eig(G.p, d, lam.p, /*vectors=*/ncomp > 0);
  std::vector<double> h(d);
  RS_CUDA(cudaMemcpyAsync(h.data(), lam.p, sizeof(double) * d, cudaMemcpyDeviceToHost, c.stream));
  RS_CUDA(cudaStreamSynchronize(c.stream));
  eig.resize(d);
  const double denom = (double)std::max<int64_t>(1, n - 1);
  for (int64_t i = 0; i < d; ++i) eig[i] = std::max(h[d - 1 - i], 0.0) / denom;  // descending sdev^2
  if (ncomp && d_emb) {
    if (gene_side) {
      // scores = Xc^T U_top  (n x ncomp); top eigenvectors are the last columns of G
      DevBuf<double> Utop((size_t)m * ncomp);
      for (int cidx = 0; cidx < ncomp; ++cidx)
        RS_CUDA(cudaMemcpyAsync(Utop.p + (size_t)cidx * m, G.p + (size_t)(d - 1 - cidx) * d, sizeof(double) * m,
                                cudaMemcpyDeviceToDevice, c.stream));
      gemm_tn(Xc.p, Utop.p, d_emb, n, ncomp, m);
    } else {
      // scores = V_top * sqrt(lambda_top)
      for (int cidx = 0; cidx < ncomp; ++cidx)
        RS_LAUNCH(k_scaled_copy, (unsigned)cdiv(n, 256), 256, 0, G.p, n, d, (int)(d - 1 - cidx),
                  std::sqrt(std::max(h[d - 1 - cidx], 0.0)), d_emb + (size_t)cidx * n);
    }
  }
}

// prcomp(t(X))$rotation[, 1:ncomp]: gene loadings (m x ncomp) of the leading principal components
// (used by SelectData, R/DataSelection.R:33-41).  Signs are arbitrary (the caller takes abs()).
void pca_loadings(const double* d_X, int64_t m, int64_t n, int ncomp, double* d_load) {
  Phase ph("pca");
  Context& c = ctx();
  const int64_t d = std::min(m, n);
  RS_REQUIRE(d <= 32768 && ncomp >= 1 && ncomp <= d, "pca_loadings: bad shape");
  DevBuf<double> ones(n), mean(m), Xc((size_t)m * n);
  RS_LAUNCH(k_fill, (unsigned)cdiv(n, 256), 256, 0, ones.p, n, 1.0 / (double)n);
  DenseMatvec mv;
  mv.init(d_X, m, n);
  mv.mul_n(ones.p, mean.p);
  RS_LAUNCH(k_center_rows, 148 * 8, 256, 0, d_X, m, n, mean.p, Xc.p);
  DevBuf<double> G((size_t)d * d), lam(d);
  const bool gene_side = m <= n;
  if (gene_side) gram_AAt(Xc.p, m, n, G.p); else gram_AtA(Xc.p, m, n, G.p);
  sym_eig(G.p, d, lam.p, true);
  if (gene_side) {  // eigenvectors of Xc Xc^T are the loadings; leading ones are the last columns
    for (int cidx = 0; cidx < ncomp; ++cidx)
      RS_CUDA(cudaMemcpyAsync(d_load + (size_t)cidx * m, G.p + (size_t)(d - 1 - cidx) * d, sizeof(double) * m,
                              cudaMemcpyDeviceToDevice, c.stream));
  } else {          // loadings = Xc v / sigma for the cell-side eigenvectors v
    std::vector<double> h(d);
    RS_CUDA(cudaMemcpyAsync(h.data(), lam.p, sizeof(double) * d, cudaMemcpyDeviceToHost, c.stream));
    RS_CUDA(cudaStreamSynchronize(c.stream));
    DevBuf<double> Vtop((size_t)n * ncomp), sc(ncomp);
    std::vector<double> hs(ncomp);
    for (int cidx = 0; cidx < ncomp; ++cidx) {
      RS_CUDA(cudaMemcpyAsync(Vtop.p + (size_t)cidx * n, G.p + (size_t)(d - 1 - cidx) * d, sizeof(double) * n,
                              cudaMemcpyDeviceToDevice, c.stream));
      const double l = std::max(h[d - 1 - cidx], 0.0);
      hs[cidx] = l > 0 ? 1.0 / std::sqrt(l) : 0.0;
    }
    sc.upload(hs.data(), ncomp, c.stream);
    gemm_nn(Xc.p, Vtop.p, d_load, m, ncomp, n);
    scale_columns(d_load, m, ncomp, sc.p);
  }
  RS_CUDA(cudaStreamSynchronize(c.stream));
}

int choose_K(const std::vector<double>& eig) {
  // R/SimilarityM.R:61-72: cc = cumsum(eig[-1]); dd = cc[-1] / sum(eig[-1]); K1 = #{dd <= 0.3}
  const size_t L = eig.size();
  if (L < 3) return 10;
  double total = 0;
  for (size_t i = 1; i < L; ++i) total += eig[i];
  double cum = eig[1];
  int K1 = 0;
  for (size_t i = 2; i < L; ++i) {
    cum += eig[i];
    if (cum / total <= 0.3) ++K1;
  }
  if (K1 <= 10) return 10;
  if (K1 >= 30) return 30;
  return K1 + 1;
}

}  // namespace rs
