// a12-a13: NMF::nmf(V, rank = k, method = 'lee', seed = 'nndsvd') + labels (R/Clustering.R:29-37).
//
// The similarity matrix handed to the NMF is the symmetrised kNN-sparse W (<= 2 n K non-zeros),
// so the two per-iteration contractions W^T V (k x n) and V H^T (n x k) are SpMMs over the
// non-zeros (2 nnz k flop each) instead of 2 k n^2 dense flops / 16 n^2 B of dense V traffic.
// Factors are stored cell-major (k contiguous values per cell) so each non-zero gathers one
// contiguous k-vector.  Semantics follow the NMF package (third-party, see oracle header):
//   H <- max(H o (W^T V), eps) / ((W^T W) H + eps);  W <- max(W o (V H^T), eps) / (W (H H^T) + eps)
//   columns of W rescaled to sum 1 every iteration; eps = 1e-9; 'connectivity' stop every 10
//   iterations (stop after 41 unchanged checks); maxIter 2000.
#include <algorithm>
#include <cmath>

#include "gemm.h"
#include "internal.h"
#include "lanczos.h"
#include "sparse.h"

namespace rs {

static constexpr int NT = 256;
static constexpr int GP = 64;       // partial blocks for the k x k Grams / column sums
static constexpr int NMF_KMAX = 128;


// ---- dense -> compressed -------------------------------------------------------------------
__global__ void k_count_cols(const double* __restrict__ V, int64_t n, int* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int64_t j = ((int64_t)blockIdx.x * NT + threadIdx.x) >> 5;
  if (j >= n) return;
  const double* col = V + j * n;
  int c = 0;
  for (int64_t i = lane; i < n; i += 32) c += col[i] != 0.0;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) cnt[j] = c;
}
__global__ void k_fill_cols(const double* __restrict__ V, int64_t n, const int* __restrict__ ptr, int* __restrict__ idx,
                            double* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t j = ((int64_t)blockIdx.x * NT + threadIdx.x) >> 5;
  if (j >= n) return;
  const double* col = V + j * n;
  int base = ptr[j];
  for (int64_t i0 = 0; i0 < n; i0 += 32) {
    const int64_t i = i0 + lane;
    const double v = i < n ? col[i] : 0.0;
    const unsigned mask = __ballot_sync(0xffffffffu, v != 0.0);
    if (v != 0.0) {
      const int off = base + __popc(mask & ((1u << lane) - 1));
      idx[off] = (int)i;
      val[off] = v;
    }
    base += __popc(mask);
  }
}
__global__ void k_count_rows(const double* __restrict__ V, int64_t n, int* __restrict__ cnt) {
  const int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x;
  if (i >= n) return;
  int c = 0;
  for (int64_t j = 0; j < n; ++j) c += V[i + j * n] != 0.0;
  cnt[i] = c;
}
__global__ void k_fill_rows(const double* __restrict__ V, int64_t n, const int* __restrict__ ptr, int* __restrict__ idx,
                            double* __restrict__ val) {
  const int64_t i = (int64_t)blockIdx.x * NT + threadIdx.x;
  if (i >= n) return;
  int off = ptr[i];
  for (int64_t j = 0; j < n; ++j) {
    const double v = V[i + j * n];
    if (v != 0.0) { idx[off] = (int)j; val[off] = v; ++off; }
  }
}

void compress(const double* d_V, int64_t n, bool by_row, Compressed& out) {
  cudaStream_t s = ctx().stream;
  out.n = n;
  DevBuf<int> cnt(n);
  if (by_row) RS_LAUNCH(k_count_rows, (unsigned)cdiv(n, NT), NT, 0, d_V, n, cnt.p);
  else        RS_LAUNCH(k_count_cols, (unsigned)cdiv(n * 32, NT), NT, 0, d_V, n, cnt.p);
  std::vector<int> h(n), ptr(n + 1, 0);
  cnt.download(h.data(), n, s);
  RS_CUDA(cudaStreamSynchronize(s));
  int64_t total = 0;
  for (int64_t i = 0; i < n; ++i) { total += h[i]; RS_REQUIRE(total < 2147483647LL, "nmf: too many non-zeros"); ptr[i + 1] = (int)total; }
  out.nnz = total;
  out.ptr.alloc(n + 1);
  out.ptr.upload(ptr.data(), n + 1, s);
  out.idx.alloc(std::max<int64_t>(total, 1));
  out.val.alloc(std::max<int64_t>(total, 1));
  if (by_row) RS_LAUNCH(k_fill_rows, (unsigned)cdiv(n, NT), NT, 0, d_V, n, out.ptr.p, out.idx.p, out.val.p);
  else        RS_LAUNCH(k_fill_cols, (unsigned)cdiv(n * 32, NT), NT, 0, d_V, n, out.ptr.p, out.idx.p, out.val.p);
  RS_CUDA(cudaStreamSynchronize(s));
}

// y[r] = sum_e val[e] x[idx[e]] over the entries of compressed line r (warp per line)
__global__ void k_spmv(const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
                       int64_t n, const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * NT + threadIdx.x) >> 5;
  if (r >= n) return;
  double acc = 0;
  for (int e = ptr[r] + lane; e < ptr[r + 1]; e += 32) acc += val[e] * x[idx[e]];
  acc = warp_sum(acc);
  if (lane == 0) y[r] = acc;
}
void spmv(const Compressed& A, const double* x, double* y) {
  RS_LAUNCH(k_spmv, (unsigned)cdiv(A.n * 32, NT), NT, 0, A.ptr.p, A.idx.p, A.val.p, A.n, x, y);
}

// ---- k x k Gram of a cell-major n x k factor, column sums -----------------------------------
__device__ __forceinline__ void d_gram_partial(const double* __restrict__ O, int64_t n, int k, double* __restrict__ partial,
                                               double* tile /* shared: 32 rows x k */) {
  const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * chunk, r1 = min(n, r0 + chunk);
  const int kk = k * k;
  double acc[(NMF_KMAX * NMF_KMAX + NT - 1) / NT];
#pragma unroll
  for (int u = 0; u < (NMF_KMAX * NMF_KMAX + NT - 1) / NT; ++u) acc[u] = 0;
  for (int64_t rr = r0; rr < r1; rr += 32) {
    const int rows = (int)min((int64_t)32, r1 - rr);
    __syncthreads();
    for (int e = threadIdx.x; e < rows * k; e += NT) tile[e] = O[rr * k + e];
    __syncthreads();
    int u = 0;
    for (int e = threadIdx.x; e < kk; e += NT, ++u) {
      const int a = e / k, b = e % k;
      double s = acc[u];
      for (int r = 0; r < rows; ++r) s += tile[r * k + a] * tile[r * k + b];
      acc[u] = s;
    }
  }
  int u = 0;
  for (int e = threadIdx.x; e < kk; e += NT, ++u) partial[(int64_t)blockIdx.x * kk + e] = acc[u];
}
__global__ void k_gram_partial(const double* __restrict__ O, int64_t n, int k, double* __restrict__ partial) {
  extern __shared__ double tile[];  // 32 rows x k
  d_gram_partial(O, n, k, partial, tile);
}
__device__ __forceinline__ void d_gram_final(const double* __restrict__ partial, int P, int kk, double* __restrict__ G) {
  const int e = blockIdx.x * NT + threadIdx.x;
  if (e >= kk) return;
  double s = 0;
  for (int p = 0; p < P; ++p) s += partial[(int64_t)p * kk + e];
  G[e] = s;
}
__global__ void k_gram_final(const double* __restrict__ partial, int P, int kk, double* __restrict__ G) {
  d_gram_final(partial, P, kk, G);
}
__device__ __forceinline__ void d_colsum_partial(const double* __restrict__ O, int64_t n, int k, double* __restrict__ partial) {
  // thread c < k sums column c over this block's rows (k <= NT)
  const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * chunk, r1 = min(n, r0 + chunk);
  if (threadIdx.x < k) {
    double s = 0;
    for (int64_t r = r0; r < r1; ++r) s += O[r * k + threadIdx.x];
    partial[(int64_t)blockIdx.x * k + threadIdx.x] = s;
  }
}
__global__ void k_colsum_partial(const double* __restrict__ O, int64_t n, int k, double* __restrict__ partial) {
  d_colsum_partial(O, n, k, partial);
}
__device__ __forceinline__ void d_rescale(double* __restrict__ W, int64_t n, int k, const double* __restrict__ partial, int P,
                                          double* cs /* shared: NMF_KMAX */) {
  if (threadIdx.x < k) {
    double s = 0;
    for (int p = 0; p < P; ++p) s += partial[(int64_t)p * k + threadIdx.x];
    cs[threadIdx.x] = s;
  }
  __syncthreads();
  const int64_t total = n * k;
  for (int64_t e = (int64_t)blockIdx.x * NT + threadIdx.x; e < total; e += (int64_t)gridDim.x * NT) W[e] /= cs[e % k];
}
__global__ void k_rescale(double* __restrict__ W, int64_t n, int k, const double* __restrict__ partial, int P) {
  __shared__ double cs[NMF_KMAX];
  d_rescale(W, n, k, partial, P, cs);
}

// ---- multiplicative update of one factor (warp per cell) -------------------------------------
// T[x][c] <- max(T[x][c] * sum_e val[e] O[idx[e]][c], eps) / (sum_c' G[c][c'] T[x][c'] + eps)
__device__ __forceinline__ void d_mult_update(const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
                                              int64_t n, int k, const double* __restrict__ O, const double* __restrict__ G, double eps,
                                              double* __restrict__ T, double* sm /* shared: G (k*k) + per-warp old row (k each) */) {
  double* sG = sm;
  double* sT = sm + k * k + (threadIdx.x >> 5) * k;
  for (int e = threadIdx.x; e < k * k; e += NT) sG[e] = G[e];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t x = ((int64_t)blockIdx.x * NT + threadIdx.x) >> 5;
  if (x >= n) return;
  for (int c = lane; c < k; c += 32) sT[c] = T[x * k + c];
  __syncwarp();
  const int e0 = ptr[x], e1 = ptr[x + 1];
  for (int c = lane; c < k; c += 32) {
    double num = 0;
    for (int e = e0; e < e1; ++e) num += val[e] * O[(int64_t)idx[e] * k + c];
    double den = 0;
    for (int c2 = 0; c2 < k; ++c2) den += sG[c * k + c2] * sT[c2];
    T[x * k + c] = fmax(sT[c] * num, eps) / (den + eps);
  }
}
__global__ void k_mult_update(const int* __restrict__ ptr, const int* __restrict__ idx, const double* __restrict__ val,
                              int64_t n, int k, const double* __restrict__ O, const double* __restrict__ G, double eps,
                              double* __restrict__ T) {
  extern __shared__ double sm[];
  d_mult_update(ptr, idx, val, n, k, O, G, eps, T, sm);
}

// first arg-max of every cell's k-vector; optional comparison with a previous assignment
__device__ __forceinline__ void d_argmax(const double* __restrict__ F, int64_t n, int k, int base, int32_t* __restrict__ out,
                                         const int32_t* __restrict__ prev, int* __restrict__ changed) {
  const int64_t x = (int64_t)blockIdx.x * NT + threadIdx.x;
  if (x >= n) return;
  int best = 0;
  double bv = F[x * k];
  for (int c = 1; c < k; ++c) {
    const double v = F[x * k + c];
    if (v > bv) { bv = v; best = c; }
  }
  out[x] = best + base;
  if (prev && prev[x] != best + base) *changed = 1;
}
__global__ void k_argmax(const double* __restrict__ F, int64_t n, int k, int base, int32_t* __restrict__ out,
                         const int32_t* __restrict__ prev, int* __restrict__ changed) {
  d_argmax(F, n, k, base, out, prev, changed);
}

// ---- batched forms: blockIdx.y = problem of a rank sweep (same arithmetic as the single-problem kernels) ----------
struct NmfProb {
  double *W, *H, *G, *part;
  int32_t *cur, *prev;
  int* changed;
  int k, done, have_prev;
};
__global__ void kb_gram_partial(const NmfProb* __restrict__ pr, int of_H, int64_t n) {
  extern __shared__ double tile[];
  const NmfProb P = pr[blockIdx.y];
  if (P.done) return;
  d_gram_partial(of_H ? P.H : P.W, n, P.k, P.part, tile);
}
__global__ void kb_gram_final(const NmfProb* __restrict__ pr, int nparts) {
  const NmfProb P = pr[blockIdx.y];
  if (P.done) return;
  d_gram_final(P.part, nparts, P.k * P.k, P.G);
}
// update_H != 0: H <- f(W, G = W^T W) over the CSC view; else W <- f(H, G = H H^T) over the CSR view
__global__ void kb_mult_update(const NmfProb* __restrict__ pr, int update_H, const int* __restrict__ ptr, const int* __restrict__ idx,
                               const double* __restrict__ val, int64_t n, double eps) {
  extern __shared__ double sm[];
  const NmfProb P = pr[blockIdx.y];
  if (P.done) return;
  d_mult_update(ptr, idx, val, n, P.k, update_H ? P.W : P.H, P.G, eps, update_H ? P.H : P.W, sm);
}
__global__ void kb_colsum_partial(const NmfProb* __restrict__ pr, int64_t n) {
  const NmfProb P = pr[blockIdx.y];
  if (P.done) return;
  d_colsum_partial(P.W, n, P.k, P.part);
}
__global__ void kb_rescale(const NmfProb* __restrict__ pr, int64_t n, int nparts) {
  __shared__ double cs[NMF_KMAX];
  const NmfProb P = pr[blockIdx.y];
  if (P.done) return;
  d_rescale(P.W, n, P.k, P.part, nparts, cs);
}
// on_basis != 0: final labels from W (1-based) into cur; else the connectivity check on H against prev
__global__ void kb_argmax(const NmfProb* __restrict__ pr, int64_t n, int on_basis) {
  const NmfProb P = pr[blockIdx.y];
  if (on_basis) { d_argmax(P.W, n, P.k, 1, P.cur, nullptr, P.changed); return; }
  if (P.done) return;
  d_argmax(P.H, n, P.k, 0, P.cur, P.have_prev ? P.prev : nullptr, P.changed);
}

// ---- NNDSVD seed ----------------------------------------------------------------------------
static void nndsvd_seed(const double* d_V, const Compressed& csr, const Compressed& csc, int64_t n, int k,
                        std::vector<double>& W0, std::vector<double>& H0) {
  Phase ph("nndsvd");
  cudaStream_t s = ctx().stream;
  // top-k singular triplets of V: Lanczos on V^T V (two SpMVs per step)
  DevBuf<double> tmp(n);
  Lanczos lz;
  const int cap = (int)std::min<int64_t>(n, 1536);
  lz.init(n, cap, [&](const double* v, double* w) {
    spmv(csr, v, tmp.p);   // tmp = V v
    spmv(csc, tmp.p, w);   // w = V^T tmp
  });
  std::vector<double> theta, S;
  int next = std::min(cap, std::max(2 * k + 24, 48));
  bool dense_fallback = false;
  while (true) {
    lz.advance(next - lz.steps);
    lz.ritz(theta, S, /*last_row_only=*/true);
    const int st = lz.steps;
    // A single-vector Krylov space sees each DISTINCT singular value once: with (numerically)
    // repeated singular values it is exhausted before k triplets exist -> dense path below.
    if (st < k) { dense_fallback = true; break; }
    const double top = theta[st - 1];
    bool conv = true;
    for (int i = 0; i < k && conv; ++i) {
      const double resid = std::fabs(lz.beta[st - 1] * S[st - 1 - i]);
      conv = resid <= 1e-13 * top;
    }
    if (conv || lz.exhausted || st >= lz.max_steps) { lz.ritz(theta, S); break; }
    next = std::min(lz.max_steps, st + std::max(32, st / 3));
  }
  std::vector<int> sel(k);
  DevBuf<double> Vr((size_t)n * k), Ur((size_t)n * k);
  if (!dense_fallback) {
    const int st = lz.steps;
    for (int i = 0; i < k; ++i) sel[i] = st - 1 - i;  // descending sigma
    lz.ritz_vectors(S, sel, Vr.p);
  } else {
    // dense eigen-decomposition of V^T V (the reference takes a full SVD of V anyway)
    RS_REQUIRE(n <= 8192, "nmf: repeated singular values need the dense NNDSVD path, limited to n <= 8192");
    DevBuf<double> G((size_t)n * n), lam(n);
    gram_AtA(d_V, n, n, G.p);
    sym_eig(G.p, n, lam.p, /*vectors=*/true);
    theta.assign(n, 0.0);
    RS_CUDA(cudaMemcpyAsync(theta.data(), lam.p, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    for (int i = 0; i < k; ++i) {
      sel[i] = (int)n - 1 - i;
      RS_CUDA(cudaMemcpyAsync(Vr.p + (size_t)i * n, G.p + (size_t)sel[i] * n, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
    }
    RS_CUDA(cudaStreamSynchronize(s));
  }
  for (int i = 0; i < k; ++i) spmv(csr, Vr.p + (size_t)i * n, Ur.p + (size_t)i * n);  // sigma_i u_i = V v_i
  std::vector<double> hv((size_t)n * k), hu((size_t)n * k);
  Vr.download(hv.data(), hv.size(), s);
  Ur.download(hu.data(), hu.size(), s);
  RS_CUDA(cudaStreamSynchronize(s));
  // Boutsidis & Gallopoulos seed (flag 0), assembled on the host: O(n k)
  W0.assign((size_t)n * k, 0.0);
  H0.assign((size_t)n * k, 0.0);
  for (int i = 0; i < k; ++i) {
    const double sig = std::sqrt(std::max(theta[sel[i]], 0.0));
    double* u = hu.data() + (size_t)i * n;
    const double* v = hv.data() + (size_t)i * n;
    double vn = 0;
    for (int64_t r = 0; r < n; ++r) vn += v[r] * v[r];
    vn = std::sqrt(vn);
    if (!(sig > 0) || !(vn > 0)) continue;
    for (int64_t r = 0; r < n; ++r) u[r] /= sig * vn;
    if (i == 0) {
      for (int64_t r = 0; r < n; ++r) {
        W0[r * k] = std::sqrt(sig) * std::fabs(u[r]);
        H0[r * k] = std::sqrt(sig) * std::fabs(v[r] / vn);
      }
      continue;
    }
    double nup = 0, nun = 0, nvp = 0, nvn = 0;
    for (int64_t r = 0; r < n; ++r) {
      const double a = u[r], b = v[r] / vn;
      if (a > 0) nup += a * a; else nun += a * a;
      if (b > 0) nvp += b * b; else nvn += b * b;
    }
    nup = std::sqrt(nup); nun = std::sqrt(nun); nvp = std::sqrt(nvp); nvn = std::sqrt(nvn);
    const double termp = nup * nvp, termn = nun * nvn;
    const bool pos = termp >= termn;
    const double term = pos ? termp : termn, nu = pos ? nup : nun, nv = pos ? nvp : nvn;
    const double sc = std::sqrt(sig * term);
    for (int64_t r = 0; r < n; ++r) {
      const double a = pos ? std::max(u[r], 0.0) : std::max(-u[r], 0.0);
      const double b = pos ? std::max(v[r] / vn, 0.0) : std::max(-v[r] / vn, 0.0);
      W0[r * k + i] = sc * a / nu;
      H0[r * k + i] = sc * b / nv;
    }
  }
  for (auto& x : W0) if (x < 1e-10) x = 0.0;
  for (auto& x : H0) if (x < 1e-10) x = 0.0;
}

void nmf_lee(const double* d_V, int64_t n, int k, int max_iter, double* h_basis, double* h_coef, int32_t* h_labels,
             NmfResult& res) {
  cudaStream_t s = ctx().stream;
  RS_REQUIRE(n > 0, "nmf: empty matrix");
  RS_REQUIRE(k >= 1 && k <= NMF_KMAX && k <= n, "nmf: rank must be in [1, min(n, 128)]");
  if (max_iter <= 0) max_iter = 2000;
  const double eps = 1e-9;
  const int stop_conv = 40, check_interval = 10;
  Compressed csr, csc;
  compress(d_V, n, true, csr);
  compress(d_V, n, false, csc);
  std::vector<double> W0, H0;
  nndsvd_seed(d_V, csr, csc, n, k, W0, H0);
  Phase ph("nmf");
  DevBuf<double> W((size_t)n * k), H((size_t)n * k), G((size_t)k * k), part((size_t)GP * std::max(k * k, k));
  DevBuf<int32_t> idxA(n), idxB(n);
  DevBuf<int> changed(1);
  W.upload(W0.data(), W0.size(), s);
  H.upload(H0.data(), H0.size(), s);
  const unsigned wgrid = (unsigned)cdiv(n * 32, NT);
  const size_t upd_smem = sizeof(double) * ((size_t)k * k + (NT / 32) * k);
  const size_t gram_smem = sizeof(double) * 32 * k;
  const int kk = k * k;
  bool have_prev = false;
  int inc = 0, it = 0;
  int32_t* cur = idxA.p;
  int32_t* prev = idxB.p;
  if (upd_smem > 48 * 1024)
    RS_CUDA(cudaFuncSetAttribute(k_mult_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
  auto one_iteration = [&] {
    // H update: numerator W^T V by columns (CSC), denominator (W^T W) H
    RS_LAUNCH(k_gram_partial, GP, NT, gram_smem, W.p, n, k, part.p);
    RS_LAUNCH(k_gram_final, (unsigned)cdiv(kk, NT), NT, 0, part.p, GP, kk, G.p);
    RS_LAUNCH(k_mult_update, wgrid, NT, upd_smem, csc.ptr.p, csc.idx.p, csc.val.p, n, k, W.p, G.p, eps, H.p);
    // W update with the new H: numerator V H^T by rows (CSR), denominator W (H H^T)
    RS_LAUNCH(k_gram_partial, GP, NT, gram_smem, H.p, n, k, part.p);
    RS_LAUNCH(k_gram_final, (unsigned)cdiv(kk, NT), NT, 0, part.p, GP, kk, G.p);
    RS_LAUNCH(k_mult_update, wgrid, NT, upd_smem, csr.ptr.p, csr.idx.p, csr.val.p, n, k, H.p, G.p, eps, W.p);
    // rescale columns of W to sum 1
    RS_LAUNCH(k_colsum_partial, GP, NT, 0, W.p, n, k, part.p);
    RS_LAUNCH(k_rescale, 148 * 2, NT, 0, W.p, n, k, part.p, GP);
  };
  // The loop is launch-bound (8 small kernels per iteration): the `check_interval` iterations between
  // two stop checks are captured once into a CUDA graph and replayed.
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  const int64_t launches_before = g_launches.load();
  RS_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  try {
    for (int r = 0; r < check_interval; ++r) one_iteration();
  } catch (...) {
    cudaStreamEndCapture(s, &graph);
    if (graph) cudaGraphDestroy(graph);
    throw;
  }
  RS_CUDA(cudaStreamEndCapture(s, &graph));
  const int64_t launches_per_graph = g_launches.load() - launches_before;
  g_launches.fetch_sub(launches_per_graph);  // captured, not executed
  RS_CUDA(cudaGraphInstantiate(&gexec, graph, 0));
  struct GraphGuard {
    cudaGraph_t g; cudaGraphExec_t e;
    ~GraphGuard() { if (e) cudaGraphExecDestroy(e); if (g) cudaGraphDestroy(g); }
  } guard{graph, gexec};
  for (it = 1; it <= max_iter; ++it) {
    if ((it - 1) % check_interval == 0 && it + check_interval - 1 <= max_iter) {
      RS_CUDA(cudaGraphLaunch(gexec, s));          // iterations it .. it + check_interval - 1
      g_launches.fetch_add(launches_per_graph);
      it += check_interval - 1;
    } else {
      one_iteration();
    }
    if (it % check_interval == 0) {  // 'connectivity' stop on the coef arg-max
      changed.zero(s);
      RS_LAUNCH(k_argmax, (unsigned)cdiv(n, NT), NT, 0, H.p, n, k, 0, cur, have_prev ? prev : (const int32_t*)nullptr,
                changed.p);
      int hc = 0;
      RS_CUDA(cudaMemcpyAsync(&hc, changed.p, sizeof(int), cudaMemcpyDeviceToHost, s));
      RS_CUDA(cudaStreamSynchronize(s));
      if (have_prev && !hc) {
        ++inc;
      } else {
        std::swap(cur, prev);  // store the new assignment
        have_prev = true;
        inc = 0;
      }
      report_progress(/*RSOPTSC_STAGE_NMF*/ 1, it, 0.0);  // interrupt point between stop checks
      if (inc > stop_conv) break;
    }
  }
  res.iters = std::min(it, max_iter);
  // outputs: basis n x k column-major, coef k x n column-major (== cell-major), labels 1-based
  std::vector<double> hw((size_t)n * k);
  W.download(hw.data(), hw.size(), s);
  if (h_coef) H.download(h_coef, (size_t)n * k, s);
  if (h_labels) {
    RS_LAUNCH(k_argmax, (unsigned)cdiv(n, NT), NT, 0, W.p, n, k, 1, cur, (const int32_t*)nullptr, changed.p);
    RS_CUDA(cudaMemcpyAsync(h_labels, cur, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  }
  RS_CUDA(cudaStreamSynchronize(s));
  if (h_basis)
    for (int64_t r = 0; r < n; ++r)
      for (int c = 0; c < k; ++c) h_basis[r + (int64_t)c * n] = hw[r * k + c];
}

// ---- rank sweep (BASELINE.json configs[3]: nmf(V, k) for k = k_lo .. k_hi) ---------------------------------------
// The R problems are independent and each is launch-bound (8 small kernels per iteration): they advance together,
// one batched launch set per iteration with blockIdx.y = problem, V compressed once.  Every problem runs exactly the
// arithmetic of nmf_lee (same device functions, same seed routine), so labels and iteration counts are those of R
// separate calls; a problem that meets its stop rule is frozen (its blocks exit) while the others go on.
void nmf_sweep(const double* d_V, int64_t n, int k_lo, int k_hi, int max_iter, int32_t* h_labels, int32_t* h_iters) {
  cudaStream_t s = ctx().stream;
  RS_REQUIRE(n > 0, "nmf_sweep: empty matrix");
  RS_REQUIRE(k_lo >= 1 && k_hi >= k_lo && k_hi <= NMF_KMAX && k_hi <= n, "nmf_sweep: ranks must satisfy 1 <= k_lo <= k_hi <= min(n, 128)");
  if (max_iter <= 0) max_iter = 2000;
  const double eps = 1e-9;
  const int stop_conv = 40, check_interval = 10;
  const int R = k_hi - k_lo + 1, kmax = k_hi;
  Compressed csr, csc;
  compress(d_V, n, true, csr);
  compress(d_V, n, false, csc);
  std::vector<DevBuf<double>> W(R), H(R), G(R), part(R);
  std::vector<DevBuf<int32_t>> ia(R), ib(R);
  DevBuf<int> changed(R);
  std::vector<NmfProb> prob(R);
  for (int p = 0; p < R; ++p) {
    const int k = k_lo + p;
    std::vector<double> W0, H0;
    nndsvd_seed(d_V, csr, csc, n, k, W0, H0);
    W[p].alloc((size_t)n * k); H[p].alloc((size_t)n * k); G[p].alloc((size_t)k * k);
    part[p].alloc((size_t)GP * std::max(k * k, k));
    ia[p].alloc(n); ib[p].alloc(n);
    W[p].upload(W0.data(), W0.size(), s);
    H[p].upload(H0.data(), H0.size(), s);
    RS_CUDA(cudaStreamSynchronize(s));  // W0 / H0 die with this iteration
    prob[p] = NmfProb{W[p].p, H[p].p, G[p].p, part[p].p, ia[p].p, ib[p].p, changed.p + p, k, 0, 0};
  }
  Phase ph("nmf");
  DevBuf<NmfProb> d_prob(R);
  auto push = [&] { d_prob.upload(prob.data(), R, s); RS_CUDA(cudaStreamSynchronize(s)); };
  push();
  const unsigned wgrid = (unsigned)cdiv(n * 32, NT);
  const size_t upd_smem = sizeof(double) * ((size_t)kmax * kmax + (NT / 32) * kmax);
  const size_t gram_smem = sizeof(double) * 32 * kmax;
  if (upd_smem > 48 * 1024)
    RS_CUDA(cudaFuncSetAttribute(kb_mult_update, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
  const unsigned Ry = (unsigned)R;
  auto one_iteration = [&] {
    RS_LAUNCH(kb_gram_partial, dim3(GP, Ry), NT, gram_smem, d_prob.p, 0, n);
    RS_LAUNCH(kb_gram_final, dim3((unsigned)cdiv(kmax * kmax, NT), Ry), NT, 0, d_prob.p, GP);
    RS_LAUNCH(kb_mult_update, dim3(wgrid, Ry), NT, upd_smem, d_prob.p, 1, csc.ptr.p, csc.idx.p, csc.val.p, n, eps);
    RS_LAUNCH(kb_gram_partial, dim3(GP, Ry), NT, gram_smem, d_prob.p, 1, n);
    RS_LAUNCH(kb_gram_final, dim3((unsigned)cdiv(kmax * kmax, NT), Ry), NT, 0, d_prob.p, GP);
    RS_LAUNCH(kb_mult_update, dim3(wgrid, Ry), NT, upd_smem, d_prob.p, 0, csr.ptr.p, csr.idx.p, csr.val.p, n, eps);
    RS_LAUNCH(kb_colsum_partial, dim3(GP, Ry), NT, 0, d_prob.p, n);
    RS_LAUNCH(kb_rescale, dim3(148 * 2, Ry), NT, 0, d_prob.p, n, GP);
  };
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  const int64_t launches_before = g_launches.load();
  RS_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  try {
    for (int r = 0; r < check_interval; ++r) one_iteration();
  } catch (...) {
    cudaStreamEndCapture(s, &graph);
    if (graph) cudaGraphDestroy(graph);
    throw;
  }
  RS_CUDA(cudaStreamEndCapture(s, &graph));
  const int64_t launches_per_graph = g_launches.load() - launches_before;
  g_launches.fetch_sub(launches_per_graph);
  RS_CUDA(cudaGraphInstantiate(&gexec, graph, 0));
  struct GraphGuard {
    cudaGraph_t g; cudaGraphExec_t e;
    ~GraphGuard() { if (e) cudaGraphExecDestroy(e); if (g) cudaGraphDestroy(g); }
  } guard{graph, gexec};
  std::vector<int> inc(R, 0), iters(R, 0), hc(R, 0);
  int live = R, it = 0;
  const int full = max_iter / check_interval * check_interval;  // iterations covered by whole graphs
  while (live > 0 && it < max_iter) {
    if (it < full) {
      RS_CUDA(cudaGraphLaunch(gexec, s));
      g_launches.fetch_add(launches_per_graph);
      it += check_interval;
    } else {
      one_iteration();
      ++it;
    }
    if (it % check_interval == 0) {  // 'connectivity' stop of every live problem
      changed.zero(s);
      RS_LAUNCH(kb_argmax, dim3((unsigned)cdiv(n, NT), Ry), NT, 0, d_prob.p, n, 0);
      RS_CUDA(cudaMemcpyAsync(hc.data(), changed.p, sizeof(int) * R, cudaMemcpyDeviceToHost, s));
      RS_CUDA(cudaStreamSynchronize(s));
      for (int p = 0; p < R; ++p) {
        if (prob[p].done) continue;
        if (prob[p].have_prev && !hc[p]) {
          ++inc[p];
        } else {
          std::swap(prob[p].cur, prob[p].prev);
          prob[p].have_prev = 1;
          inc[p] = 0;
        }
        if (inc[p] > stop_conv) { prob[p].done = 1; iters[p] = it; --live; }
      }
      push();
      report_progress(/*RSOPTSC_STAGE_NMF*/ 1, it, 0.0);
    }
  }
  for (int p = 0; p < R; ++p)
    if (!prob[p].done) iters[p] = std::min(it, max_iter);
  if (h_labels) {
    RS_LAUNCH(kb_argmax, dim3((unsigned)cdiv(n, NT), Ry), NT, 0, d_prob.p, n, 1);
    for (int p = 0; p < R; ++p)
      RS_CUDA(cudaMemcpyAsync(h_labels + (size_t)p * n, prob[p].cur, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, s));
  }
  RS_CUDA(cudaStreamSynchronize(s));
  if (h_iters)
    for (int p = 0; p < R; ++p) h_iters[p] = iters[p];
}

}  // namespace rs
