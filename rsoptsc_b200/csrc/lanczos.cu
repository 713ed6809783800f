// Lanczos machinery (a5: spectral norms of the stop rule, R/SimilarityM.R:158-163,179,190;
// also used for the NNDSVD seed of the NMF and the PCA scores of CountClusters).
//
// The reference computes every ||.||_2 with a full values-only SVD; only the largest singular
// value is used, so a Lanczos iteration on the Gram operator of the smaller side (matvecs are
// HBM-bound streams over the matrix) gives the same number to ~1e-13 relative.
// Device-resident recurrence (alpha/beta stay on the GPU), full reorthogonalisation, the small
// tridiagonal eigenproblem is solved on the host (tql2).
#include <cmath>
#include <functional>

#include "comm.h"
#include "gemm.h"
#include "internal.h"
#include "lanczos.h"

namespace rs {

// ---- host: symmetric tridiagonal QL with implicit shifts (EISPACK tql2 scheme) -----------
void tridiag_eig(std::vector<double>& d, std::vector<double>& e_in, std::vector<double>* Zv, bool last_row_only) {
  const int n = (int)d.size();
  if (n == 0) return;
  std::vector<double> e(n, 0.0);
  for (int i = 0; i + 1 < n; ++i) e[i] = e_in[i];
  std::vector<double> dummy;
  const bool wantv = Zv != nullptr;
  std::vector<double>& V = wantv ? *Zv : dummy;
  // last_row_only: V (length n) carries only the last row of the eigenvector matrix -- all a
  // Lanczos residual estimate needs -- so the whole solve is O(n^2) instead of O(n^3).
  if (wantv && last_row_only) {
    V.assign((size_t)n, 0.0);
    V[n - 1] = 1.0;
  } else if (wantv) {
    V.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0;  // column-major: V[row + col*n]
  }
  double f = 0.0, tst1 = 0.0;
  const double eps = 2.220446049250313e-16;
  for (int l = 0; l < n; ++l) {
    tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
    int mm = l;
    while (mm < n) {
      if (std::fabs(e[mm]) <= eps * tst1) break;
      ++mm;
    }
    if (mm >= n) mm = n - 1;
    if (mm > l) {
      int iter = 0;
      do {
        ++iter;
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = std::hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; ++i) d[i] -= h;
        f += h;
        p = d[mm];
        double c = 1.0, c2 = c, c3 = c;
        const double el1 = e[l + 1];
        double s = 0.0, s2 = 0.0;
        for (int i = mm - 1; i >= l; --i) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i];
          h = c * p;
          r = std::hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          if (wantv && last_row_only) {
            h = V[i + 1];
            V[i + 1] = s * V[i] + c * h;
            V[i] = c * V[i] - s * h;
          } else if (wantv) {
            double* vi = &V[(size_t)i * n];
            double* vi1 = &V[(size_t)(i + 1) * n];
            for (int k = 0; k < n; ++k) {
              h = vi1[k];
              vi1[k] = s * vi[k] + c * h;
              vi[k] = c * vi[k] - s * h;
            }
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (std::fabs(e[l]) > eps * tst1 && iter < 200);
    }
    d[l] += f;
    e[l] = 0.0;
  }
  // ascending selection sort, carrying vectors
  for (int i = 0; i + 1 < n; ++i) {
    int k = i;
    double p = d[i];
    for (int j = i + 1; j < n; ++j)
      if (d[j] < p) { k = j; p = d[j]; }
    if (k != i) {
      d[k] = d[i];
      d[i] = p;
      if (wantv && last_row_only) std::swap(V[i], V[k]);
      else if (wantv)
        for (int r = 0; r < n; ++r) std::swap(V[(size_t)i * n + r], V[(size_t)k * n + r]);
    }
  }
}

// ---- device vector kernels -------------------------------------------------------------
static constexpr int LT = 256;
static constexpr int LB = 148;  // blocks for the vector kernels (one per SM)

__global__ void k_start_vector(double* __restrict__ q, int64_t d) {
  // deterministic pseudo-random positive start vector, normalised later
  for (int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x; i < d; i += (int64_t)gridDim.x * LT) {
    uint32_t x = (uint32_t)(i * 2654435761u) ^ 0x9e3779b9u;
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    q[i] = 0.5 + (double)(x & 0xffffff) / 16777216.0;
  }
}
__global__ void k_dot_partial(const double* __restrict__ a, const double* __restrict__ b, int64_t d,
                              double* __restrict__ partial) {
  __shared__ double sm[LT / 32 + 1];
  double acc = 0;
  for (int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x; i < d; i += (int64_t)gridDim.x * LT) acc += a[i] * b[i];
  acc = block_sum<LT>(acc, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__device__ __forceinline__ double sum_partials(const double* __restrict__ partial, int count, double* sm) {
  double acc = 0;
  for (int i = threadIdx.x; i < count; i += LT) acc += partial[i];
  return block_sum<LT>(acc, sm);
}
// q = q / ||q|| given partial sums of q.q
__global__ void k_scale_by_norm(double* __restrict__ q, int64_t d, const double* __restrict__ partial, int np,
                                double* __restrict__ beta_out) {
  __shared__ double sm[LT / 32 + 1];
  const double nrm = sqrt(sum_partials(partial, np, sm));
  if (blockIdx.x == 0 && threadIdx.x == 0 && beta_out) *beta_out = nrm;
  const double inv = nrm > 0 ? 1.0 / nrm : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x; i < d; i += (int64_t)gridDim.x * LT) q[i] *= inv;
}
// w -= alpha q_k + beta_{k-1} q_{k-1}; alpha = sum(partial) stored to alpha_out
__global__ void k_three_term(double* __restrict__ w, const double* __restrict__ qk, const double* __restrict__ qkm1,
                             const double* __restrict__ beta_prev, int64_t d, const double* __restrict__ partial,
                             int np, double* __restrict__ alpha_out) {
  __shared__ double sm[LT / 32 + 1];
  const double alpha = sum_partials(partial, np, sm);
  if (blockIdx.x == 0 && threadIdx.x == 0) *alpha_out = alpha;
  const double beta = qkm1 ? *beta_prev : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x; i < d; i += (int64_t)gridDim.x * LT) {
    double v = w[i] - alpha * qk[i];
    if (qkm1) v -= beta * qkm1[i];
    w[i] = v;
  }
}
// c[j] = Q[:,j] . w  (one warp per column)
__global__ void k_Qt_w(const double* __restrict__ Q, int64_t d, int ncol, const double* __restrict__ w,
                       double* __restrict__ c) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * LT + threadIdx.x) >> 5, nw = (gridDim.x * LT) >> 5;
  for (int j = gw; j < ncol; j += nw) {
    const double* qj = Q + (int64_t)j * d;
    double acc = 0;
    for (int64_t i = lane; i < d; i += 32) acc += qj[i] * w[i];
    acc = warp_sum(acc);
    if (lane == 0) c[j] = acc;
  }
}
// w[i] -= sum_j Q[i,j] c[j] ; partial sums of w.w
__global__ void k_sub_Qc(const double* __restrict__ Q, int64_t d, int ncol, const double* __restrict__ c,
                         double* __restrict__ w, double* __restrict__ partial) {
  __shared__ double sm[LT / 32 + 1];
  double nrm = 0;
  for (int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x; i < d; i += (int64_t)gridDim.x * LT) {
    double acc = 0;
    for (int j = 0; j < ncol; ++j) acc += Q[(int64_t)j * d + i] * c[j];
    const double v = w[i] - acc;
    w[i] = v;
    nrm += v * v;
  }
  nrm = block_sum<LT>(nrm, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = nrm;
}
// dst = w / beta (beta = sqrt(sum partial)); zero when the Krylov space is exhausted
__global__ void k_next_q(const double* __restrict__ w, int64_t d, const double* __restrict__ partial, int np,
                         double* __restrict__ beta_out, double* __restrict__ dst) {
  __shared__ double sm[LT / 32 + 1];
  const double beta = sqrt(sum_partials(partial, np, sm));
  if (blockIdx.x == 0 && threadIdx.x == 0) *beta_out = beta;
  const double inv = beta > 0 ? 1.0 / beta : 0.0;
  for (int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x; i < d; i += (int64_t)gridDim.x * LT) dst[i] = w[i] * inv;
}

// ---- dense matvec kernels (column-major) --------------------------------------------------
// y[j] = M[:,j] . x    (warp per column, coalesced down the column)
__global__ void k_gemv_t(const double* __restrict__ M, int64_t rows, int64_t cols, const double* __restrict__ x,
                         double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t gw = ((int64_t)blockIdx.x * LT + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * LT) >> 5;
  for (int64_t j = gw; j < cols; j += nw) {
    const double* mj = M + j * rows;
    double acc = 0;
    for (int64_t i = lane; i < rows; i += 32) acc += mj[i] * x[i];
    acc = warp_sum(acc);
    if (lane == 0) y[j] = acc;
  }
}
// partial[chunk][i] = sum_{j in chunk} M[i,j] x[j]   (thread per row, coalesced across rows)
__global__ void k_gemv_n_partial(const double* __restrict__ M, int64_t rows, int64_t cols, int64_t chunk,
                                 const double* __restrict__ x, double* __restrict__ partial) {
  const int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x;
  const int64_t j0 = (int64_t)blockIdx.y * chunk, j1 = min(cols, j0 + chunk);
  if (i >= rows) return;
  double acc = 0;
  for (int64_t j = j0; j < j1; ++j) acc += M[j * rows + i] * x[j];
  partial[(int64_t)blockIdx.y * rows + i] = acc;
}
__global__ void k_gemv_n_reduce(const double* __restrict__ partial, int64_t rows, int nchunk, double* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x;
  if (i >= rows) return;
  double acc = 0;
  for (int c = 0; c < nchunk; ++c) acc += partial[(int64_t)c * rows + i];
  y[i] = acc;
}

void DenseMatvec::init(const double* M_, int64_t rows_, int64_t cols_) {
  M = M_; rows = rows_; cols = cols_;
  const int64_t row_tiles = cdiv(rows, LT);
  int64_t want = std::max<int64_t>(1, (4 * 148) / row_tiles);
  nchunk = (int)std::min<int64_t>(std::min<int64_t>(want, cols), 1024);
  chunk = cdiv(cols, nchunk);
  nchunk = cdiv(cols, chunk);
  partial.alloc((size_t)nchunk * rows);
}
void DenseMatvec::mul_t(const double* x, double* y) {  // y(cols) = M^T x(rows)
  RS_LAUNCH(k_gemv_t, (unsigned)std::min<int64_t>(cdiv(cols * 32, LT), 148 * 8), LT, 0, M, rows, cols, x, y);
}
void DenseMatvec::mul_n(const double* x, double* y) {  // y(rows) = M x(cols)
  dim3 grid((unsigned)cdiv(rows, LT), (unsigned)nchunk);
  RS_LAUNCH(k_gemv_n_partial, grid, LT, 0, M, rows, cols, chunk, x, partial.p);
  RS_LAUNCH(k_gemv_n_reduce, (unsigned)cdiv(rows, LT), LT, 0, partial.p, rows, nchunk, y);
}

// ---- Lanczos engine ---------------------------------------------------------------------
void Lanczos::init(int64_t dim_, int max_steps_, std::function<void(const double*, double*)> op_, const double* d_start) {
  dim = dim_;
  max_steps = (int)std::min<int64_t>(max_steps_, dim_);
  op = std::move(op_);
  Q.alloc((size_t)dim * (max_steps + 1));
  w.alloc(dim);
  coef.alloc(max_steps + 1);
  d_alpha.alloc(max_steps + 1);
  d_beta.alloc(max_steps + 1);
  partial.alloc(LB);
  steps = 0;
  alpha.clear(); beta.clear();
  if (d_start) RS_CUDA(cudaMemcpyAsync(Q.p, d_start, sizeof(double) * dim, cudaMemcpyDeviceToDevice, ctx().stream));
  else RS_LAUNCH(k_start_vector, LB, LT, 0, Q.p, dim);
  RS_LAUNCH(k_dot_partial, LB, LT, 0, Q.p, Q.p, dim, partial.p);
  RS_LAUNCH(k_scale_by_norm, LB, LT, 0, Q.p, dim, partial.p, LB, (double*)nullptr);
}

void Lanczos::advance(int more) {
  const int target = std::min(max_steps, steps + more);
  for (int k = steps; k < target; ++k) {
    double* qk = Q.p + (int64_t)k * dim;
    op(qk, w.p);
    RS_LAUNCH(k_dot_partial, LB, LT, 0, qk, w.p, dim, partial.p);
    RS_LAUNCH(k_three_term, LB, LT, 0, w.p, qk, k > 0 ? qk - dim : (const double*)nullptr,
              k > 0 ? d_beta.p + (k - 1) : (const double*)nullptr, dim, partial.p, LB, d_alpha.p + k);
    // full reorthogonalisation against q_0..q_k (classical Gram-Schmidt pass)
    RS_LAUNCH(k_Qt_w, (unsigned)std::min(cdiv((int64_t)(k + 1) * 32, LT), 148 * 4), LT, 0, Q.p, dim, k + 1, w.p, coef.p);
    RS_LAUNCH(k_sub_Qc, LB, LT, 0, Q.p, dim, k + 1, coef.p, w.p, partial.p);
    RS_LAUNCH(k_next_q, LB, LT, 0, w.p, dim, partial.p, LB, d_beta.p + k, qk + dim);
  }
  steps = target;
  alpha.resize(steps);
  beta.resize(steps);
  RS_CUDA(cudaMemcpyAsync(alpha.data(), d_alpha.p, sizeof(double) * steps, cudaMemcpyDeviceToHost, ctx().stream));
  RS_CUDA(cudaMemcpyAsync(beta.data(), d_beta.p, sizeof(double) * steps, cudaMemcpyDeviceToHost, ctx().stream));
  RS_CUDA(cudaStreamSynchronize(ctx().stream));
  // Krylov space exhausted (beta ~ 0): truncate
  double scale = 0;
  for (int i = 0; i < steps; ++i) scale = std::max(scale, std::fabs(alpha[i]) + (i ? beta[i - 1] : 0.0));
  for (int i = 0; i < steps; ++i)
    if (!std::isfinite(beta[i]) || !std::isfinite(alpha[i])) nonfinite = true;
  for (int i = 0; i < steps; ++i) {
    if (!(beta[i] > 1e-14 * scale) || !std::isfinite(beta[i]) || !std::isfinite(alpha[i])) {
      steps = std::isfinite(alpha[i]) ? i + 1 : i;
      alpha.resize(steps);
      beta.resize(steps);
      max_steps = steps;
      exhausted = true;
      break;
    }
  }
}

void Lanczos::ritz(std::vector<double>& theta, std::vector<double>& S, bool last_row_only) const {
  theta = alpha;
  std::vector<double> e(beta.begin(), beta.begin() + std::max(0, steps - 1));
  tridiag_eig(theta, e, &S, last_row_only);
}

// Y[:, c] = Q[:, 0..steps) * S[:, sel[c]]
__global__ void k_ritz_vectors(const double* __restrict__ Q, int64_t d, int steps, const double* __restrict__ Ssel,
                               int ncol, double* __restrict__ Y) {
  const int64_t i = (int64_t)blockIdx.x * LT + threadIdx.x;
  if (i >= d) return;
  for (int c = 0; c < ncol; ++c) {
    const double* s = Ssel + (int64_t)c * steps;
    double acc = 0;
    for (int j = 0; j < steps; ++j) acc += Q[(int64_t)j * d + i] * s[j];
    Y[(int64_t)c * d + i] = acc;
  }
}
void Lanczos::ritz_vectors(const std::vector<double>& S, const std::vector<int>& sel, double* d_Y) {
  std::vector<double> h((size_t)sel.size() * steps);
  for (size_t c = 0; c < sel.size(); ++c)
    for (int j = 0; j < steps; ++j) h[c * steps + j] = S[(size_t)sel[c] * steps + j];
  DevBuf<double> dS(h.size());
  dS.upload(h.data(), h.size(), ctx().stream);
  RS_LAUNCH(k_ritz_vectors, (unsigned)cdiv(dim, LT), LT, 0, Q.p, dim, steps, dS.p, (int)sel.size(), d_Y);
  RS_CUDA(cudaStreamSynchronize(ctx().stream));
}

// ---- spectral norm ------------------------------------------------------------------------
// Certified top Ritz value of the running Lanczos process (shared by the two entry points below).
// Returns sigma_1^2 >= 0, or NaN when the recurrence met non-finite values (NaN/Inf in the matrix):
// the callers' finiteness checks then fail the run instead of reporting convergence.
static double top_ritz_value(Lanczos& lz, double rel_tol) {
  double prev = -1, cur = 0;
  while (true) {
    lz.advance(lz.steps == 0 ? 32 : 16);
    if (lz.nonfinite) return std::nan("");
    if (lz.steps == 0) return 0.0;
    std::vector<double> theta, S;
    lz.ritz(theta, S, /*last_row_only=*/true);
    const int k = lz.steps;
    cur = theta[k - 1];
    if (std::isnan(cur)) return cur;
    if (!(cur > 0)) return 0.0;
    // residual bound of the top Ritz pair: |beta_k * s_k|
    const double resid = std::fabs(lz.beta[k - 1] * S[k - 1]);
    // Ritz value error <= min(resid, resid^2 / gap).  Ritz values within 0.1*rel_tol of the top are
    // one cluster (their spread is below the requested accuracy); gap = distance to the rest.
    double gap = cur;
    for (int i = k - 2; i >= 0; --i)
      if (cur - theta[i] > 0.1 * rel_tol * cur) { gap = cur - theta[i]; break; }
    const double err = std::min(resid, resid * resid / gap);
    const bool conv = err <= rel_tol * cur ||
                      (prev > 0 && std::fabs(cur - prev) <= 1e-3 * rel_tol * cur && resid <= 1e-5 * cur);
    if (conv || lz.exhausted) break;
    if (lz.steps >= lz.max_steps) {
      // not certified within the step budget: say so instead of returning a silently loose value
      add_counter("lanczos_uncertified", 1.0);
      break;
    }
    prev = cur;
  }
  return cur;
}

double spectral_norm(const double* d_M, int64_t rows, int64_t cols, double rel_tol) {
  Phase ph("lanczos");
  if (rows == 0 || cols == 0) return 0.0;
  const bool small_rows = rows <= cols;
  const int64_t d = small_rows ? rows : cols;
  DenseMatvec mv;
  DevBuf<double> tmp, B;
  Lanczos lz;
  if (d <= 4096) {
    // small side: form the d x d Gram once (one syrk) and iterate on it -- every Lanczos step is
    // then one L2-resident symmetric matvec instead of two HBM streams over the m x n matrix
    B.alloc((size_t)d * d);
    if (small_rows) gram_AAt(d_M, rows, cols, B.p); else gram_AtA(d_M, rows, cols, B.p);
    mirror_lower(B.p, d);
    mv.init(B.p, d, d);
    lz.init(d, 320, [&](const double* v, double* w) { mv.mul_t(v, w); });  // B symmetric: B^T v = B v
  } else {
    mv.init(d_M, rows, cols);
    tmp.alloc(small_rows ? cols : rows);
    lz.init(d, 320, [&](const double* v, double* w) {
      if (small_rows) { mv.mul_t(v, tmp.p); mv.mul_n(tmp.p, w); }   // w = M (M^T v)
      else            { mv.mul_n(v, tmp.p); mv.mul_t(tmp.p, w); }   // w = M^T (M v)
    });
  }
  return std::sqrt(top_ritz_value(lz, rel_tol));
}

double spectral_norm_colsharded(const double* d_M, int64_t rows, int64_t cols_local, double rel_tol) {
  if (!sharded()) return spectral_norm(d_M, rows, cols_local, rel_tol);
  Phase ph("lanczos");
  if (rows == 0) return 0.0;
  // Operator on the row side, M M^T = sum over ranks of M_r M_r^T (the ranks own disjoint columns).
  // The recurrence runs replicated on bitwise identical data, so every rank takes the same decisions.
  DenseMatvec mv;
  DevBuf<double> tmp, B;
  Lanczos lz;
  if (rows <= 4096) {
    B.alloc((size_t)rows * rows);
    if (cols_local > 0) {
      gram_AAt(d_M, rows, cols_local, B.p);
      mirror_lower(B.p, rows);
    } else {
      B.zero(ctx().stream);
    }
    comm_allreduce_sum(B.p, (size_t)rows * rows);
    mv.init(B.p, rows, rows);
    lz.init(rows, 320, [&](const double* v, double* w) { mv.mul_t(v, w); });
  } else {
    if (cols_local > 0) mv.init(d_M, rows, cols_local);
    tmp.alloc(std::max<int64_t>(cols_local, 1));
    lz.init(rows, 320, [&](const double* v, double* w) {
      if (cols_local > 0) { mv.mul_t(v, tmp.p); mv.mul_n(tmp.p, w); }
      else RS_CUDA(cudaMemsetAsync(w, 0, sizeof(double) * rows, ctx().stream));
      comm_allreduce_sum(w, (size_t)rows);  // one n-vector per step
    });
  }
  return std::sqrt(top_ritz_value(lz, rel_tol));
}

}  // namespace rs
