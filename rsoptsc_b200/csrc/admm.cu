// computeM: linearised ADMM for  min ||Z||_* + lambda ||E||_{2,1}  s.t. X = XZ + E, Z^T 1 = 1,
// Z_Omega = 0   (R/SimilarityM.R:115-197), restructured for the GPU:
//
//   * Z lives on its kNN support only (n*K values, CSR + CSC views); the dense n x n state is
//     Y3 and the SVT operand A/J.
//   * the stop test of iteration `it` (R :157-162) only looks at Err[it-1], Err[it-2] and XXZ,
//     none of which the SVT touches, so it is evaluated BEFORE the SVT (saves one SVT).
//   * J == 0 exactly while ||A||_F <= 1/mu (all singular values thresholded): SVT skipped.
//   * spectral norms: Lanczos (lanczos.cu); the two tests that only compare a norm with
//     epsilon = 1e-5 use the Frobenius bounds ||.||_F/sqrt(min dim) <= ||.||_2 <= ||.||_F first.
#include <algorithm>
#include <cmath>

#include "internal.h"

namespace rs {

static void upload_support(Support& s, const std::vector<int>& rowptr, const std::vector<int>& col) {
  const int64_t n = s.n;
  s.nnz = (int64_t)col.size();
  s.h_rowptr = rowptr;
  s.h_col = col;
  std::vector<int> rowof(s.nnz), colptr(n + 1, 0), rowidx(s.nnz), csrpos(s.nnz);
  for (int64_t i = 0; i < n; ++i)
    for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) { rowof[p] = (int)i; ++colptr[col[p] + 1]; }
  s.max_col_nnz = 0;
  for (int64_t j = 0; j < n; ++j) { s.max_col_nnz = std::max(s.max_col_nnz, colptr[j + 1]); colptr[j + 1] += colptr[j]; }
  std::vector<int> fill(colptr.begin(), colptr.end() - 1);
  for (int64_t p = 0; p < s.nnz; ++p) {  // rows ascend with p, so each column's list is row-sorted
    const int q = fill[col[p]]++;
    rowidx[q] = rowof[p];
    csrpos[q] = (int)p;
  }
  cudaStream_t st = ctx().stream;
  s.rowptr.alloc(n + 1); s.rowptr.upload(rowptr.data(), n + 1, st);
  s.colptr.alloc(n + 1); s.colptr.upload(colptr.data(), n + 1, st);
  s.col.alloc(s.nnz);    s.rowof.alloc(s.nnz); s.rowidx.alloc(s.nnz); s.csrpos.alloc(s.nnz);
  if (s.nnz) {
    s.col.upload(col.data(), s.nnz, st);
    s.rowof.upload(rowof.data(), s.nnz, st);
    s.rowidx.upload(rowidx.data(), s.nnz, st);
    s.csrpos.upload(csrpos.data(), s.nnz, st);
  }
  RS_CUDA(cudaStreamSynchronize(st));
}

void build_support_from_idx(Support& s, const int32_t* h_idx, int64_t n, int K) {
  RS_REQUIRE(n > 0 && K > 0, "support: n and K must be positive");
  RS_REQUIRE(n * (int64_t)K < (int64_t)2147483647, "support: too many entries");
  s.n = n;
  std::vector<int> rowptr(n + 1, 0), col;
  col.reserve((size_t)n * K);
  std::vector<int> tmp(K);
  for (int64_t j = 0; j < n; ++j) {
    for (int t = 0; t < K; ++t) {
      const int c = h_idx[j * K + t];
      RS_REQUIRE(c >= 0 && c < n, "support: neighbour index out of range");
      tmp[t] = c;
    }
    std::sort(tmp.begin(), tmp.end());
    const auto last = std::unique(tmp.begin(), tmp.end());
    col.insert(col.end(), tmp.begin(), last);
    rowptr[j + 1] = (int)col.size();
    tmp.resize(K);
  }
  upload_support(s, rowptr, col);
}

void build_support_from_dense_mask(Support& s, const double* h_D, int64_t n) {
  // entries with D > 0 are forbidden (R/SimilarityM.R:181); D is column-major
  s.n = n;
  std::vector<int> cnt(n, 0);
  for (int64_t j = 0; j < n; ++j)
    for (int64_t i = 0; i < n; ++i)
      if (!(h_D[i + j * n] > 0)) ++cnt[i];
  std::vector<int> rowptr(n + 1, 0);
  for (int64_t i = 0; i < n; ++i) {
    RS_REQUIRE((int64_t)rowptr[i] + cnt[i] < (int64_t)2147483647, "support: too many entries");
    rowptr[i + 1] = rowptr[i] + cnt[i];
  }
  std::vector<int> col(rowptr[n]), fill(rowptr.begin(), rowptr.end() - 1);
  for (int64_t j = 0; j < n; ++j)
    for (int64_t i = 0; i < n; ++i)
      if (!(h_D[i + j * n] > 0)) col[fill[i]++] = (int)j;  // j ascends -> sorted rows
  upload_support(s, rowptr, col);
}

void run_admm(const double* d_X, int64_t m, int64_t n, const Support& s, double lambda, AdmmState& st,
              AdmmResult& res) {
  Context& c = ctx();
  const int maxiter = 100;                                   // R :117
  const double rho = 5.0, mumax = 1e6, epsilon = 1e-5;       // R :118-121
  double mu = 1e-6;
  st.m = m; st.n = n; st.X = d_X;
  const size_t mn = (size_t)m * n, nn = (size_t)n * n;
  st.Zs.alloc(std::max<size_t>(s.nnz, 1)); st.Zs.zero(c.stream);
  st.q.alloc(std::max<size_t>(s.nnz, 1));
  st.XXZ.alloc(mn); st.E.alloc(mn); st.Y1.alloc(mn); st.R.alloc(mn); st.M.alloc(mn);
  st.E.zero(c.stream); st.Y1.zero(c.stream);
  st.Y2.alloc(n); st.Y2.zero(c.stream);
  st.colacc.alloc(n + 1);
  DevBuf<double> Y3(nn), A(nn);
  Y3.zero(c.stream);
  SvtWorkspace svt;
  for (int i = 0; i < maxiter; ++i) res.err1[i] = res.err2[i] = 0.0;

  // XXZ = X - X Z with Z = 0 (R :130)
  double xxz_f2 = admm_residual_update(st, s, mu, /*update_duals=*/false);
  // eta = ||X||_2^2 + ||1_n||_2^2 + 1 (R :178) is iteration-invariant (quirk Q8): hoisted
  const double sx = spectral_norm(d_X, m, n);
  const double eta = sx * sx + std::sqrt((double)n) * std::sqrt((double)n) + 1.0;
  const double min_dim = (double)std::min(m, n);
  res.n_svt = 0;
  int it = 0;
  while (true) {
    ++it;
    if (it >= maxiter) { res.stop = 1; break; }              // R :134-137
    mu = std::min(rho * mu, mumax);                          // R :141
    const double tau = 1.0 / mu;
    if (it >= 3) {                                           // R :157-162 (hoisted above the SVT)
      bool stop = res.err1[it - 2] >= res.err1[it - 3];
      if (!stop) {
        const double f = std::sqrt(xxz_f2);
        if (f <= epsilon) stop = true;                       // ||.||_2 <= ||.||_F
        else if (f / std::sqrt(min_dim) <= epsilon) stop = spectral_norm(st.XXZ.p, m, n) <= epsilon;
      }
      if (stop) { res.stop = 2; break; }
    }
    // step 1: J = SVT_{1/mu}(Z - Y3/mu)                      R :142-155
    const double a_f2 = build_A(s, st.Zs.p, Y3.p, mu, A.p);
    // J == 0 exactly when sigma_1(A) <= tau.  ||A||_F/sqrt(n) <= sigma_1 <= ||A||_F settles most
    // iterations; in between, a residual-certified Lanczos value of sigma_1 (1e-10 relative) decides.
    bool j_zero = std::sqrt(a_f2) <= tau;
    if (!j_zero && std::sqrt(a_f2 / (double)n) <= tau) j_zero = spectral_norm(A.p, n, n) <= tau * (1.0 - 1e-8);
    if (!j_zero) { svt_inplace(svt, A.p, n, tau); ++res.n_svt; }
    const double* J = j_zero ? nullptr : A.p;
    gather_support_term(s, st.Zs.p, J, Y3.p, mu, st.q.p);    // (Z - J + Y3/mu) on the support
    admm_E_update(st, lambda, mu);                           // R :164-176
    admm_Z_update(st, s, mu, eta);                           // R :178-181 (+ Y2, :186)
    xxz_f2 = admm_residual_update(st, s, mu, true);          // R :182, :185
    const double zj_f2 = update_Y3(s, st.Zs.p, J, mu, Y3.p); // R :187
    res.err1[it - 1] = spectral_norm(st.M.p, m, n);          // R :189
    // Err[,2] = ||Z - J||_2 is only ever compared with epsilon (R :191): Frobenius bounds first
    const double zj_f = std::sqrt(zj_f2);
    double err2 = zj_f / std::sqrt((double)n);               // lower bound
    if (err2 <= epsilon && res.err1[it - 1] <= epsilon) {
      if (zj_f <= epsilon) err2 = zj_f;
      else {  // inconclusive: form Z - J densely and take its spectral norm
        DevBuf<double> Dm(nn);
        scatter_support_dense(s, st.Zs.p, Dm.p);
        if (J) {
          const double minus1 = -1.0;
          RS_CUBLAS(cublasDaxpy(c.cublas, (int)nn, &minus1, J, 1, Dm.p, 1));
        }
        err2 = spectral_norm(Dm.p, n, n);
      }
    }
    res.err2[it - 1] = err2;
    if (std::max(res.err1[it - 1], err2) <= epsilon) { res.stop = 3; break; }  // R :191-194
  }
  res.iters = it;
  res.E = std::sqrt(xxz_f2);                                 // norm(XXZ, "f")
}

}  // namespace rs
