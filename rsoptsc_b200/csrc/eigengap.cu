// a14-a18: CountClusters (R/Clustering.R:61-175): Laplacian spectrum (GetComponents), PCA + PAM
// ensemble (GetEnsemble), consensus matrix (GetConsensus) and the eigengap rule.
#include <algorithm>
#include <cmath>

#include "internal.h"
#include "lanczos.h"
#include "sparse.h"

namespace rs {

static constexpr int ET = 256;

// ---- a14: GetComponents ----------------------------------------------------------------------
__global__ void k_fill_ones(double* __restrict__ v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * ET + threadIdx.x;
  if (i < n) v[i] = 1.0;
}
__global__ void k_inv_sqrt(double* __restrict__ d, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * ET + threadIdx.x;
  if (i < n) d[i] = 1.0 / sqrt(d[i]);
}
// L = I - (dd_i * S_ij) * dd_j      (R/Clustering.R:99-101; sqrtm/solve of a diagonal matrix, quirk Q13):
// diagonal block of the Laplacian over the cells `members` (ascending)
__global__ void k_laplacian_block(const double* __restrict__ S, const double* __restrict__ dd, int64_t n,
                                  const int* __restrict__ members, int64_t nc, double* __restrict__ Lc) {
  const int64_t b = blockIdx.y;
  const int64_t j = members[b];
  const double dj = dd[j];
  for (int64_t a = (int64_t)blockIdx.x * ET + threadIdx.x; a < nc; a += (int64_t)gridDim.x * ET) {
    const int64_t i = members[a];
    const double v = (dd[i] * S[i + j * n]) * dj;
    Lc[a + b * nc] = (a == b ? 1.0 : 0.0) - v;
  }
}

// All eigenvalues of one dense symmetric block (lower triangle), appended to `out`.
static void block_eigenvalues(double* d_L, int64_t nc, std::vector<double>& out) {
  Context& c = ctx();
  DevBuf<double> lam(nc);
  {
    Phase ph("eig");
    sym_eig(d_L, nc, lam.p, /*vectors=*/false);
  }
  const size_t at = out.size();
  out.resize(at + nc);
  RS_CUDA(cudaMemcpyAsync(out.data() + at, lam.p, sizeof(double) * nc, cudaMemcpyDeviceToHost, c.stream));
  RS_CUDA(cudaStreamSynchronize(c.stream));
}

// Connected components of the non-zero pattern (host union-find over the compressed entries); an effectively
// dense matrix is one block.  members[c] = ascending cells of component c, components numbered by lowest cell.
static void pattern_components(const Compressed& csc, int64_t n, std::vector<std::vector<int>>& members) {
  Context& c = ctx();
  members.clear();
  if (csc.nnz <= 128 * n) {
    std::vector<int> ptr(n + 1), idx(std::max<int64_t>(csc.nnz, 1)), parent(n);
    csc.ptr.download(ptr.data(), n + 1, c.stream);
    if (csc.nnz) csc.idx.download(idx.data(), csc.nnz, c.stream);
    RS_CUDA(cudaStreamSynchronize(c.stream));
    for (int64_t i = 0; i < n; ++i) parent[i] = (int)i;
    auto find = [&](int x) {
      while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
      return x;
    };
    for (int64_t j = 0; j < n; ++j)
      for (int e = ptr[j]; e < ptr[j + 1]; ++e) {
        const int a = find((int)j), b = find(idx[e]);
        if (a != b) parent[std::max(a, b)] = std::min(a, b);
      }
    std::vector<int> id(n, -1);
    for (int64_t i = 0; i < n; ++i) {
      const int r = find((int)i);
      if (id[r] < 0) { id[r] = (int)members.size(); members.emplace_back(); }
      members[id[r]].push_back((int)i);
    }
  } else {  // effectively dense: one block
    members.emplace_back(n);
    for (int64_t i = 0; i < n; ++i) members[0][i] = (int)i;
  }
}

// Spectrum of I - D^-1/2 S D^-1/2.  The Laplacian is block diagonal over the connected components
// of the non-zero pattern of S, so its spectrum is the union of the blocks' spectra (sum n_c^3
// eigensolver work instead of n^3).  The pattern is taken from the CSC view when S is sparse.
static void laplacian_spectrum(const double* d_S, int64_t n, std::vector<double>& vals, const Compressed* csc_in = nullptr) {
  Context& c = ctx();
  RS_REQUIRE(n <= 65535, "get_components: n too large for this launch shape");
  DevBuf<double> ones(n), deg(n);
  RS_LAUNCH(k_fill_ones, (unsigned)cdiv(n, ET), ET, 0, ones.p, n);
  DenseMatvec mv;
  mv.init(d_S, n, n);
  mv.mul_n(ones.p, deg.p);                                   // degree = S 1      (R :98)
  RS_LAUNCH(k_inv_sqrt, (unsigned)cdiv(n, ET), ET, 0, deg.p, n);
  Compressed own;
  const Compressed* csc = csc_in;
  if (!csc) { compress(d_S, n, false, own); csc = &own; }
  std::vector<std::vector<int>> members;
  pattern_components(*csc, n, members);
  vals.clear();
  vals.reserve(n);
  DevBuf<int> d_mem(n);
  for (const auto& mem : members) {
    const int64_t nc = (int64_t)mem.size();
    DevBuf<double> L((size_t)nc * nc);
    d_mem.upload(mem.data(), nc, c.stream);
    dim3 grid((unsigned)std::min<int64_t>(cdiv(nc, ET), 64), (unsigned)nc);
    RS_LAUNCH(k_laplacian_block, grid, ET, 0, d_S, deg.p, n, d_mem.p, nc, L.p);
    RS_CUDA(cudaStreamSynchronize(c.stream));               // `mem` staging must finish before the next upload
    block_eigenvalues(L.p, nc, vals);
  }
  for (auto& v : vals) v = std::fabs(v);                     // abs(Re(.))   (R :102)
  std::sort(vals.begin(), vals.end());                       // sort         (R :104)
}

// ---- a15 on a sparse similarity: PCA scores by Lanczos ---------------------------------------------
// prcomp(t(S), center = TRUE)$x[, 1:ncomp] (R/Clustering.R:133).  T = t(S); centring its columns gives
// Tc = S^T - 1 mu^T with mu = rowMeans(S); the scores are Tc V for the top right singular vectors V of
// Tc, i.e. the top eigenvectors of Tc^T Tc -- two SpMVs and two rank-one corrections per Lanczos step
// instead of an n^3 dense PCA of which only ncomp columns are used.
__global__ void k_dot1(const double* __restrict__ a, const double* __restrict__ b, int64_t n, double* __restrict__ out) {
  __shared__ double sm[1024 / 32 + 1];
  double acc = 0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) acc += a[i] * b[i];
  acc = block_sum<1024>(acc, sm);
  if (threadIdx.x == 0) *out = acc;
}
__global__ void k_sub_scalar(double* __restrict__ y, int64_t n, const double* __restrict__ s) {
  const int64_t i = (int64_t)blockIdx.x * ET + threadIdx.x;
  if (i < n) y[i] -= *s;
}
__global__ void k_sub_scaled(double* __restrict__ w, const double* __restrict__ mu, int64_t n, const double* __restrict__ s) {
  const int64_t i = (int64_t)blockIdx.x * ET + threadIdx.x;
  if (i < n) w[i] -= mu[i] * (*s);
}
__global__ void k_scale_vec(double* __restrict__ v, int64_t n, double f) {
  const int64_t i = (int64_t)blockIdx.x * ET + threadIdx.x;
  if (i < n) v[i] *= f;
}

static bool pca_scores_sparse(const Compressed& csr, const Compressed& csc, int64_t n, int ncomp, double* d_scores) {
  Phase ph("pca");
  const unsigned g = (unsigned)cdiv(n, ET);
  DevBuf<double> ones(n), mu(n), t(n), scal(2);
  RS_LAUNCH(k_fill_ones, g, ET, 0, ones.p, n);
  spmv(csr, ones.p, mu.p);
  RS_LAUNCH(k_scale_vec, g, ET, 0, mu.p, n, 1.0 / (double)n);
  auto tc_mul = [&](const double* v, double* out) {          // out = Tc v = S^T v - (mu.v) 1
    spmv(csc, v, out);
    RS_LAUNCH(k_dot1, 1, 1024, 0, mu.p, v, n, scal.p);
    RS_LAUNCH(k_sub_scalar, g, ET, 0, out, n, scal.p);
  };
  Lanczos lz;
  const int cap = (int)std::min<int64_t>(n, 1024);
  lz.init(n, cap, [&](const double* v, double* w) {
    tc_mul(v, t.p);
    spmv(csr, t.p, w);                                       // w = Tc^T t = S t - mu sum(t)
    RS_LAUNCH(k_dot1, 1, 1024, 0, ones.p, t.p, n, scal.p + 1);
    RS_LAUNCH(k_sub_scaled, g, ET, 0, w, mu.p, n, scal.p + 1);
  });
  std::vector<double> theta, S;
  int next = std::min(cap, std::max(2 * ncomp + 30, 48));
  while (true) {
    lz.advance(next - lz.steps);
    lz.ritz(theta, S, /*last_row_only=*/true);
    const int st = lz.steps;
    if (st < ncomp) return false;  // Krylov space exhausted (repeated singular values): caller goes dense
    bool conv = true;
    for (int i = 0; i < ncomp && conv; ++i)
      conv = std::fabs(lz.beta[st - 1] * S[st - 1 - i]) <= 1e-13 * theta[st - 1];
    if (conv || lz.exhausted || st >= lz.max_steps) { lz.ritz(theta, S); break; }
    next = std::min(lz.max_steps, st + std::max(32, st / 3));
  }
  std::vector<int> sel(ncomp);
  for (int i = 0; i < ncomp; ++i) sel[i] = lz.steps - 1 - i;
  DevBuf<double> V((size_t)n * ncomp);
  lz.ritz_vectors(S, sel, V.p);
  for (int i = 0; i < ncomp; ++i) tc_mul(V.p + (size_t)i * n, d_scores + (size_t)i * n);
  RS_CUDA(cudaStreamSynchronize(ctx().stream));
  return true;
}

// ---- a14 without the spectrum: how many eigenvalues of Lsym are <= tol ------------------------------------
// CountClusters only needs n_eigs of the similarity's Laplacian (R/Clustering.R:63); eigen(only.values) of an
// n x n matrix is O(n^3) for a number that lives at the low end of the spectrum.  Lsym <= tol  <=>  the scaled
// similarity M = D^-1/2 S D^-1/2 has an eigenvalue >= 1 - tol, so per connected component (inside one the
// eigenvalue 1 of M is simple) a Lanczos process with full reorthogonalisation on the sparse M finds the top of the
// spectrum; a second, deflated run certifies that no copy of a multiple eigenvalue was missed.  Components
// too small to bother (or a run that cannot be certified) go through the dense values-only eigensolver.
__global__ void k_scale_by(double* y, const double* x, const double* __restrict__ dd, int64_t n) {  // y may alias x
  const int64_t i = (int64_t)blockIdx.x * ET + threadIdx.x;
  if (i < n) y[i] = x[i] * dd[i];
}
__global__ void k_axpy_dev(double* __restrict__ w, const double* __restrict__ y, int64_t n, const double* __restrict__ s) {
  const int64_t i = (int64_t)blockIdx.x * ET + threadIdx.x;
  if (i < n) w[i] -= y[i] * (*s);
}

// eigenvalues >= T of M restricted to the component `mem`; returns -1 when the count could not be certified
static int lanczos_count_top(const Compressed& csr, const double* dd, int64_t n, const std::vector<int>& mem, double T) {
  Phase ph("lanczos_count");
  cudaStream_t st = ctx().stream;
  const unsigned g = (unsigned)cdiv(n, ET);
  const int64_t nc = (int64_t)mem.size();
  DevBuf<double> t(n), start(n), scal(2);
  std::vector<double> Yh;        // converged eigenvectors (host bookkeeping of the count only)
  DevBuf<double> Y;              // n x found
  int found = 0;
  const int max_steps = (int)std::min<int64_t>(nc, 640);
  for (int round = 0; round < 6; ++round) {
    // start vector: deterministic pseudo-random values on the component, zero elsewhere (the operator keeps
    // that support exactly), orthogonalised against the eigenvectors found so far
    std::vector<double> h(n, 0.0);
    uint32_t x = 0x9e3779b9u * (uint32_t)(round + 1);
    for (int i : mem) {
      x ^= x << 13; x ^= x >> 17; x ^= x << 5;
      h[i] = 0.25 + (double)(x & 0xffffff) / 16777216.0;
    }
    start.upload(h.data(), n, st);
    auto deflate = [&](double* w) {
      for (int c = 0; c < found; ++c) {
        RS_LAUNCH(k_dot1, 1, 1024, 0, Y.p + (size_t)c * n, w, n, scal.p);
        RS_LAUNCH(k_axpy_dev, g, ET, 0, w, Y.p + (size_t)c * n, n, scal.p);
      }
    };
    deflate(start.p);
    RS_CUDA(cudaStreamSynchronize(st));  // `h` is read by the upload until here
    Lanczos lz;
    lz.init(n, max_steps, [&](const double* v, double* w) {
      RS_LAUNCH(k_scale_by, g, ET, 0, t.p, v, dd, n);
      spmv(csr, t.p, w);
      RS_LAUNCH(k_scale_by, g, ET, 0, w, w, dd, n);
      deflate(w);
    }, start.p);
    std::vector<double> theta, S;
    bool done = false;
    std::vector<int> top;  // converged Ritz values >= T of this round
    while (!done) {
      lz.advance(lz.steps == 0 ? 48 : 24);
      if (lz.nonfinite) return -1;
      const int k = lz.steps;
      if (k == 0) { done = true; break; }  // start vector vanished under deflation: nothing left in this component
      lz.ritz(theta, S, /*last_row_only=*/true);
      const double bk = lz.exhausted ? 0.0 : std::fabs(lz.beta[k - 1]);
      // all Ritz values that could be >= T must have converged, and the first one below T must be a converged
      // (hence genuine) eigenvalue approximation: the top of the spectrum is then complete
      bool ok = true, below = lz.exhausted;
      top.clear();
      for (int i = k - 1; i >= 0; --i) {
        const double r = bk * std::fabs(S[i]);
        if (theta[i] + r >= T) {
          if (r > 1e-10) { ok = false; break; }
          if (theta[i] >= T) top.push_back(i);
        } else {
          below = below || r <= 1e-6;
          break;
        }
      }
      if (ok && (below || (int)top.size() == k)) done = true;
      else if (k >= lz.max_steps) return -1;
    }
    if (top.empty()) return found;        // certified: the deflated operator has nothing left above T
    // keep the Ritz vectors of the new eigenvalues for the deflation of the next round
    lz.ritz(theta, S);
    DevBuf<double> Ynew((size_t)n * (found + top.size()));
    if (found) RS_CUDA(cudaMemcpyAsync(Ynew.p, Y.p, sizeof(double) * (size_t)n * found, cudaMemcpyDeviceToDevice, st));
    lz.ritz_vectors(S, top, Ynew.p + (size_t)n * found);
    Y = std::move(Ynew);
    found += (int)top.size();
    if ((int64_t)found >= nc) return found;
  }
  return -1;
}

// #{eigenvalues of Lsym(S) <= tol} for a sparse similarity (CSR + CSC views given)
static int laplacian_count_le(const double* d_S, int64_t n, double tol, const Compressed& csr, const Compressed& csc) {
  Context& c = ctx();
  DevBuf<double> ones(n), dd(n);
  RS_LAUNCH(k_fill_ones, (unsigned)cdiv(n, ET), ET, 0, ones.p, n);
  spmv(csr, ones.p, dd.p);                                   // degree = S 1      (R :98)
  RS_LAUNCH(k_inv_sqrt, (unsigned)cdiv(n, ET), ET, 0, dd.p, n);
  std::vector<std::vector<int>> members;
  pattern_components(csc, n, members);
  int count = 0;
  DevBuf<int> d_mem(n);
  for (const auto& mem : members) {
    const int64_t nc = (int64_t)mem.size();
    int got = -1;
    if (nc >= 1024) got = lanczos_count_top(csr, dd.p, n, mem, 1.0 - tol);
    if (got < 0) {  // small block, or not certified: all eigenvalues of the dense block
      if (nc >= 1024) add_counter("laplacian_count_dense_fallback", 1.0);
      DevBuf<double> L((size_t)nc * nc);
      d_mem.upload(mem.data(), nc, c.stream);
      dim3 grid((unsigned)std::min<int64_t>(cdiv(nc, ET), 64), (unsigned)nc);
      RS_LAUNCH(k_laplacian_block, grid, ET, 0, d_S, dd.p, n, d_mem.p, nc, L.p);
      RS_CUDA(cudaStreamSynchronize(c.stream));
      std::vector<double> v;
      block_eigenvalues(L.p, nc, v);
      got = 0;
      for (double x : v) got += std::fabs(x) <= tol;
    }
    count += got;
  }
  return count;
}

// ---- a14 on the consensus matrix: exact spectrum through the label signatures ------------------------------
// consen[i,j] depends only on the label signatures (columns of the R x n table) of i and j, so with the G distinct
// signatures, their multiplicities s_g and the G x G matrix C^ between them, D^-1/2 C D^-1/2 = P M^ P^T
// (P: n x G membership, M^_gh = C^_gh / sqrt(delta_g delta_h), delta = C^ s).  Its non-zero eigenvalues are those of
// the G x G matrix sqrt(s_g s_h) M^_gh; the other n - G are zero.  Hence the spectrum of Lsym is
// {1 - mu_g} u {1 (n - G times)}: an O(n R log n + G^3) computation instead of a dense n x n eigensolve, and the
// n x n consensus matrix is never formed (R/Clustering.R:72, 96-105, 143-153).
__global__ void k_scale_sym(double* __restrict__ B, const double* __restrict__ w, int64_t G) {
  const int64_t h = blockIdx.y;
  for (int64_t g = (int64_t)blockIdx.x * ET + threadIdx.x; g < G; g += (int64_t)gridDim.x * ET) B[g + h * G] *= w[g] * w[h];
}
static void consensus_laplacian_spectrum(const std::vector<int32_t>& table, int R, int64_t n, double tau,
                                         std::vector<double>& vals) {
  Phase ph("consensus_spectrum");
  cudaStream_t st = ctx().stream;
  // group the cells by signature (lexicographic sort of the n columns)
  std::vector<int> order(n);
  for (int64_t i = 0; i < n; ++i) order[i] = (int)i;
  auto less = [&](int a, int b) {
    for (int r = 0; r < R; ++r) {
      const int32_t x = table[(size_t)r * n + a], y = table[(size_t)r * n + b];
      if (x != y) return x < y;
    }
    return false;
  };
  std::sort(order.begin(), order.end(), less);
  std::vector<int> rep;
  std::vector<double> size;
  for (int64_t i = 0; i < n; ++i) {
    if (i == 0 || less(order[i - 1], order[i])) { rep.push_back(order[i]); size.push_back(0.0); }
    size.back() += 1.0;
  }
  const int64_t G = (int64_t)rep.size();
  add_counter("consensus_signatures", (double)G);
  std::vector<int32_t> sub((size_t)R * G);
  for (int r = 0; r < R; ++r)
    for (int64_t g = 0; g < G; ++g) sub[(size_t)r * G + g] = table[(size_t)r * n + rep[g]];
  DevBuf<int32_t> d_sub(sub.size());
  d_sub.upload(sub.data(), sub.size(), st);
  DevBuf<double> B((size_t)G * G), s(G), delta(G), lam(G);
  consensus_dense(d_sub.p, R, G, tau, B.p);                  // C^ (symmetric, thresholded)
  s.upload(size.data(), G, st);
  DenseMatvec mv;
  mv.init(B.p, G, G);
  mv.mul_n(s.p, delta.p);                                    // degree of any cell of group g
  std::vector<double> hd(G), w(G);
  delta.download(hd.data(), G, st);
  RS_CUDA(cudaStreamSynchronize(st));
  for (int64_t g = 0; g < G; ++g) w[g] = std::sqrt(size[g] / hd[g]);
  s.upload(w.data(), G, st);
  dim3 grid((unsigned)std::min<int64_t>(cdiv(G, ET), 64), (unsigned)G);
  RS_LAUNCH(k_scale_sym, grid, ET, 0, B.p, s.p, G);
  RS_CUDA(cudaStreamSynchronize(st));
  {
    Phase pe("eig");
    sym_eig(B.p, G, lam.p, /*vectors=*/false);
  }
  std::vector<double> mu(G);
  lam.download(mu.data(), G, st);
  RS_CUDA(cudaStreamSynchronize(st));
  vals.assign(n, 1.0);                                       // the n - G exchangeable directions
  for (int64_t g = 0; g < G; ++g) vals[g] = std::fabs(1.0 - mu[g]);   // abs(Re(.))   (R :102)
  std::sort(vals.begin(), vals.end());                       // sort         (R :104)
}

void get_components(const double* d_S, int64_t n, double tol, double* h_vals, int32_t* n_eigs) {
  std::vector<double> vals;
  laplacian_spectrum(d_S, n, vals);
  int cnt = 0;
  for (double v : vals) cnt += v <= tol;
  if (h_vals) std::memcpy(h_vals, vals.data(), sizeof(double) * n);
  if (n_eigs) *n_eigs = cnt;
}

// ---- a17: consensus ---------------------------------------------------------------------------
// consen[i,j] = #{r : lab[r][i] == lab[r][j]}, zeroed when <= R*tau, (c + c^T)/2 == c.
__global__ void k_consensus(const int32_t* __restrict__ lab, int R, int64_t n, double cut, double* __restrict__ C) {
  extern __shared__ int32_t sl[];  // R x 32 (rows i) + R x 32 (cols j)
  int32_t* li = sl;
  int32_t* lj = sl + R * 32;
  const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
  for (int e = threadIdx.y * 32 + threadIdx.x; e < R * 32; e += 32 * blockDim.y) {
    const int r = e / 32, t = e % 32;
    li[e] = i0 + t < n ? lab[(int64_t)r * n + i0 + t] : -1;
    lj[e] = j0 + t < n ? lab[(int64_t)r * n + j0 + t] : -2;
  }
  __syncthreads();
  const int64_t i = i0 + threadIdx.x;
  for (int c = threadIdx.y; c < 32; c += blockDim.y) {
    const int64_t j = j0 + c;
    if (i >= n || j >= n) continue;
    int cnt = 0;
    for (int r = 0; r < R; ++r) cnt += li[r * 32 + threadIdx.x] == lj[r * 32 + c];
    const double v = (double)cnt;
    C[i + j * n] = v <= cut ? 0.0 : v;
  }
}
void consensus_dense(const int32_t* d_labels, int R, int64_t n, double tau, double* d_consen) {
  Phase ph("consensus");
  dim3 grid((unsigned)cdiv(n, 32), (unsigned)cdiv(n, 32)), block(32, 8);
  RS_REQUIRE(cdiv(n, 32) <= 65535, "consensus: n too large");
  RS_LAUNCH(k_consensus, grid, block, sizeof(int32_t) * R * 64, d_labels, R, n, (double)R * tau, d_consen);
}

// ---- a16: PAM (k-medoids BUILD + SWAP, cluster::pam variant 'original') ------------------------
__global__ void k_pairwise_dist(const double* __restrict__ P, int64_t n, int dim, double* __restrict__ D) {
  const int64_t j = blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * ET + threadIdx.x; i < n; i += (int64_t)gridDim.x * ET) {
    double acc = 0;
    for (int c = 0; c < dim; ++c) {
      const double diff = __dsub_rn(P[i + (int64_t)c * n], P[j + (int64_t)c * n]);
      acc = __dadd_rn(acc, __dmul_rn(diff, diff));
    }
    D[i + j * n] = sqrt(acc);
  }
}
__global__ void k_max_partial(const double* __restrict__ a, int64_t count, double* __restrict__ partial) {
  __shared__ double sm[ET / 32];
  double mx = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * ET + threadIdx.x; i < count; i += (int64_t)gridDim.x * ET) mx = fmax(mx, a[i]);
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < ET / 32; ++w) mx = fmax(mx, sm[w]);
    partial[blockIdx.x] = mx;
  }
}
// BUILD gain of candidate i: sum_j max(dysma[j] - d(i,j), 0); medoids get -inf    (one CTA per i)
__global__ void k_build_gain(const double* __restrict__ D, int64_t n, const double* __restrict__ dysma,
                             const int* __restrict__ is_med, double* __restrict__ gain) {
  __shared__ double sm[ET / 32 + 1];
  const int64_t i = blockIdx.x;
  const double* row = D + i * n;  // D symmetric: column i == row i
  double acc = 0;
  for (int64_t j = threadIdx.x; j < n; j += ET) {
    const double cmd = dysma[j] - row[j];
    if (cmd > 0) acc += cmd;
  }
  acc = block_sum<ET>(acc, sm);
  if (threadIdx.x == 0) gain[i] = is_med[i] ? -INFINITY : acc;
}
__global__ void k_min_update(const double* __restrict__ D, int64_t n, int64_t med, double* __restrict__ dysma) {
  const int64_t j = (int64_t)blockIdx.x * ET + threadIdx.x;
  if (j < n) dysma[j] = fmin(dysma[j], D[j + med * n]);
}
// nearest / second-nearest medoid distance per object
__global__ void k_nearest_two(const double* __restrict__ D, int64_t n, const int* __restrict__ med, int k, double s,
                              double* __restrict__ dysma, double* __restrict__ dysmb, int* __restrict__ near,
                              int* __restrict__ ties) {
  const int64_t j = (int64_t)blockIdx.x * ET + threadIdx.x;
  if (j >= n) return;
  double a = s, b = s;
  int na = 0;
  for (int t = 0; t < k; ++t) {
    const double d = D[j + (int64_t)med[t] * n];
    if (a > d) { b = a; a = d; na = t; } else if (b > d) { b = d; }
  }
  int cnt = 0;  // medoids at exactly the closest distance (pam.c tests dys[ij] == dysma[j])
  for (int t = 0; t < k; ++t) cnt += D[j + (int64_t)med[t] * n] == a;
  dysma[j] = a;
  dysmb[j] = b;
  near[j] = na;
  ties[j] = cnt;
}

// Single-pass SWAP cost for k <= KMAX: T[h][t] = sum_j C_{j, med[t], h}.  Every object j adds its
// "not closest" term to all medoids and a correction to the medoid(s) it is closest to.
template <int KMAX>
__global__ void k_swap_cost_fast(const double* __restrict__ D, int64_t n, const int* __restrict__ med, int k,
                                 const int* __restrict__ is_med, const double* __restrict__ dysma,
                                 const double* __restrict__ dysmb, const int* __restrict__ near,
                                 const int* __restrict__ ties, double* __restrict__ T) {
  __shared__ double sm[ET / 32 + 1];
  const int64_t h = blockIdx.x;
  if (is_med[h]) {
    for (int t = threadIdx.x; t < k; t += ET) T[h * k + t] = INFINITY;
    return;
  }
  const double* dh = D + h * n;
  double shared_term = 0;
  double corr[KMAX];
#pragma unroll
  for (int t = 0; t < KMAX; ++t) corr[t] = 0;
  for (int64_t j = threadIdx.x; j < n; j += ET) {
    const double a = dysma[j], b = dysmb[j], d = dh[j];
    const double base = d < a ? d - a : 0.0;
    shared_term += base;
    const double cterm = ((b > d ? d : b) - a) - base;
    if (ties[j] == 1) {
      const int nr = near[j];
#pragma unroll
      for (int t = 0; t < KMAX; ++t) corr[t] += (t == nr) ? cterm : 0.0;
    } else {
      for (int t = 0; t < k; ++t)
        if (D[j + (int64_t)med[t] * n] == a) {
#pragma unroll
          for (int u = 0; u < KMAX; ++u) corr[u] += (u == t) ? cterm : 0.0;
        }
    }
  }
  shared_term = block_sum<ET>(shared_term, sm);
#pragma unroll
  for (int t = 0; t < KMAX; ++t) {
    if (t < k) {
      const double v = block_sum<ET>(corr[t], sm);
      if (threadIdx.x == 0) T[h * k + t] = shared_term + v;
    }
  }
}
// SWAP cost T[h][t] = sum_j C_{j, med[t], h} for every non-medoid h (one CTA per h)   [KR p.104]
static constexpr int PAM_KMAX = 64;
__global__ void k_swap_cost(const double* __restrict__ D, int64_t n, const int* __restrict__ med, int k,
                            const int* __restrict__ is_med, const double* __restrict__ dysma,
                            const double* __restrict__ dysmb, double* __restrict__ T) {
  __shared__ double sm[ET / 32 + 1];
  const int64_t h = blockIdx.x;
  if (is_med[h]) {
    for (int t = threadIdx.x; t < k; t += ET) T[h * k + t] = INFINITY;
    return;
  }
  const double* dh = D + h * n;
  double shared_term = 0;  // contribution common to all medoids: objects not closest to i
  // pass 1: common term
  for (int64_t j = threadIdx.x; j < n; j += ET) {
    const double a = dysma[j], d = dh[j];
    if (d < a) shared_term += d - a;
  }
  shared_term = block_sum<ET>(shared_term, sm);
  // pass 2: per-medoid correction over the objects whose closest-medoid distance equals d(i,j)
  for (int t = 0; t < k; ++t) {
    const double* di = D + (int64_t)med[t] * n;
    double acc = 0;
    for (int64_t j = threadIdx.x; j < n; j += ET) {
      const double a = dysma[j];
      if (di[j] == a) {
        const double d = dh[j], b = dysmb[j];
        const double small = b > d ? d : b;
        acc += (small - a) - (d < a ? d - a : 0.0);
      }
    }
    acc = block_sum<ET>(acc, sm);
    if (threadIdx.x == 0) T[h * k + t] = shared_term + acc;
  }
}
// lexicographic-first minimum of T over (h, t)
__global__ void k_argmin_first(const double* __restrict__ T, int64_t count, double* __restrict__ out_val,
                               long long* __restrict__ out_idx) {
  __shared__ double sv[ET];
  __shared__ long long si[ET];
  double bv = INFINITY;
  long long bi = -1;
  for (int64_t e = threadIdx.x; e < count; e += ET) {
    const double v = T[e];
    if (v < bv) { bv = v; bi = e; }
  }
  sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = ET / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double v2 = sv[threadIdx.x + o];
      const long long i2 = si[threadIdx.x + o];
      if (i2 >= 0 && (v2 < sv[threadIdx.x] || (v2 == sv[threadIdx.x] && (si[threadIdx.x] < 0 || i2 < si[threadIdx.x])))) {
        sv[threadIdx.x] = v2; si[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { *out_val = sv[0]; *out_idx = si[0]; }
}
// last index among the maxima of gain (pam.c: `ammax <= beter[i]`)
__global__ void k_argmax_last(const double* __restrict__ g, int64_t count, long long* __restrict__ out_idx) {
  __shared__ double sv[ET];
  __shared__ long long si[ET];
  double bv = -INFINITY;
  long long bi = -1;
  for (int64_t e = threadIdx.x; e < count; e += ET) {
    const double v = g[e];
    if (v >= bv && v > -INFINITY) { bv = v; bi = e; }
  }
  sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = ET / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const double v2 = sv[threadIdx.x + o];
      const long long i2 = si[threadIdx.x + o];
      if (i2 >= 0 && (v2 > sv[threadIdx.x] || (v2 == sv[threadIdx.x] && i2 > si[threadIdx.x]))) {
        sv[threadIdx.x] = v2; si[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_idx = si[0];
}
__global__ void k_assign(const double* __restrict__ D, int64_t n, const int* __restrict__ med, int k,
                         int32_t* __restrict__ near) {
  const int64_t j = (int64_t)blockIdx.x * ET + threadIdx.x;
  if (j >= n) return;
  int best = 0;
  double bd = D[j + (int64_t)med[0] * n];
  for (int t = 1; t < k; ++t) {
    const double d = D[j + (int64_t)med[t] * n];
    if (d < bd) { bd = d; best = t; }
  }
  for (int t = 0; t < k; ++t)
    if (med[t] == j) best = t;
  near[j] = best;
}

struct PamEngine {
  int64_t n = 0;
  DevBuf<double> D, dysma, dysmb, gain, T, scal;
  DevBuf<int> is_med, med, near_t, ties;
  DevBuf<long long> lidx;
  DevBuf<int32_t> near;
  double s = 0;
  std::vector<int> build_order;  // greedy BUILD sequence (shared by every k)

  void init(const double* d_P, int64_t n_, int dim, int kmax) {
    n = n_;
    cudaStream_t st = ctx().stream;
    RS_REQUIRE(n <= 65535, "pam: more than 65535 observations (cluster::pam has the same limit)");
    D.alloc((size_t)n * n); dysma.alloc(n); dysmb.alloc(n); gain.alloc(n); T.alloc((size_t)n * kmax); scal.alloc(256);
    is_med.alloc(n); med.alloc(kmax); lidx.alloc(1); near.alloc(n); near_t.alloc(n); ties.alloc(n);
    dim3 grid((unsigned)std::min<int64_t>(cdiv(n, ET), 64), (unsigned)n);
    RS_LAUNCH(k_pairwise_dist, grid, ET, 0, d_P, n, dim, D.p);
    RS_LAUNCH(k_max_partial, 148, ET, 0, D.p, n * n, scal.p);
    std::vector<double> part(148);
    scal.download(part.data(), 148, st);
    RS_CUDA(cudaStreamSynchronize(st));
    double mx = 0;
    for (double v : part) mx = std::max(mx, v);
    s = 1.1 * mx + 1.0;
    // BUILD once up to kmax; the medoid set for k is the first k picks
    std::vector<double> init(n, s);
    dysma.upload(init.data(), n, st);
    is_med.zero(st);
    build_order.clear();
    const int one = 1;
    for (int t = 0; t < kmax; ++t) {
      RS_LAUNCH(k_build_gain, (unsigned)n, ET, 0, D.p, n, dysma.p, is_med.p, gain.p);
      RS_LAUNCH(k_argmax_last, 1, ET, 0, gain.p, n, lidx.p);
      long long pick = -1;
      RS_CUDA(cudaMemcpyAsync(&pick, lidx.p, sizeof(long long), cudaMemcpyDeviceToHost, st));
      RS_CUDA(cudaStreamSynchronize(st));
      RS_REQUIRE(pick >= 0, "pam: BUILD found no candidate");
      build_order.push_back((int)pick);
      RS_CUDA(cudaMemcpyAsync(is_med.p + pick, &one, sizeof(int), cudaMemcpyHostToDevice, st));
      RS_LAUNCH(k_min_update, (unsigned)cdiv(n, ET), ET, 0, D.p, n, (int64_t)pick, dysma.p);
    }
    RS_CUDA(cudaStreamSynchronize(st));
  }

  // SWAP from the BUILD prefix of size k; returns sorted medoids and first-appearance labels
  void solve(int k, std::vector<int>& medoids, std::vector<int32_t>& labels) {
    cudaStream_t st = ctx().stream;
    RS_REQUIRE(k >= 1 && k <= (int)build_order.size() && k <= PAM_KMAX, "pam: k out of range");
    std::vector<int> flags(n, 0);
    medoids.assign(build_order.begin(), build_order.begin() + k);
    std::sort(medoids.begin(), medoids.end());
    for (int m : medoids) flags[m] = 1;
    is_med.upload(flags.data(), n, st);
    med.upload(medoids.data(), k, st);
    RS_LAUNCH(k_nearest_two, (unsigned)cdiv(n, ET), ET, 0, D.p, n, med.p, k, s, dysma.p, dysmb.p, near_t.p, ties.p);
    double sky = 0;
    {  // sky = sum dysma
      std::vector<double> h(n);
      dysma.download(h.data(), n, st);
      RS_CUDA(cudaStreamSynchronize(st));
      for (double v : h) sky += v;
    }
    for (int guard = 0; guard < 100000 && k < n; ++guard) {
      if (k <= 32)
        RS_LAUNCH(k_swap_cost_fast<32>, (unsigned)n, ET, 0, D.p, n, med.p, k, is_med.p, dysma.p, dysmb.p, near_t.p, ties.p, T.p);
      else
        RS_LAUNCH(k_swap_cost, (unsigned)n, ET, 0, D.p, n, med.p, k, is_med.p, dysma.p, dysmb.p, T.p);
      RS_LAUNCH(k_argmin_first, 1, ET, 0, T.p, n * k, scal.p, lidx.p);
      double dz = 0;
      long long where = -1;
      RS_CUDA(cudaMemcpyAsync(&dz, scal.p, sizeof(double), cudaMemcpyDeviceToHost, st));
      RS_CUDA(cudaMemcpyAsync(&where, lidx.p, sizeof(long long), cudaMemcpyDeviceToHost, st));
      RS_CUDA(cudaStreamSynchronize(st));
      if (!(where >= 0) || !(dz < -16.0 * 2.220446049250313e-16 * std::fabs(sky))) break;
      const int hbest = (int)(where / k), tbest = (int)(where % k);
      flags[medoids[tbest]] = 0;
      flags[hbest] = 1;
      medoids[tbest] = hbest;
      std::sort(medoids.begin(), medoids.end());
      sky += dz;
      is_med.upload(flags.data(), n, st);
      med.upload(medoids.data(), k, st);
      RS_LAUNCH(k_nearest_two, (unsigned)cdiv(n, ET), ET, 0, D.p, n, med.p, k, s, dysma.p, dysmb.p, near_t.p, ties.p);
    }
    RS_LAUNCH(k_assign, (unsigned)cdiv(n, ET), ET, 0, D.p, n, med.p, k, near.p);
    std::vector<int32_t> nr(n);
    near.download(nr.data(), n, st);
    RS_CUDA(cudaStreamSynchronize(st));
    labels.assign(n, 0);
    std::vector<int> seen(k, 0);
    int nxt = 0;
    for (int64_t j = 0; j < n; ++j) {
      if (!seen[nr[j]]) seen[nr[j]] = ++nxt;
      labels[j] = seen[nr[j]];
    }
  }
};

void pam(const double* d_P, int64_t n, int dim, int k, int32_t* h_labels, int32_t* h_medoids) {
  Phase ph("pam");
  RS_REQUIRE(k >= 1 && k < n, "pam: need 1 <= k < n");
  PamEngine eng;
  eng.init(d_P, n, dim, k);
  std::vector<int> med;
  std::vector<int32_t> lab;
  eng.solve(k, med, lab);
  if (h_labels) std::memcpy(h_labels, lab.data(), sizeof(int32_t) * n);
  if (h_medoids) for (int t = 0; t < k; ++t) h_medoids[t] = med[t];
}

// ---- CountClusters ------------------------------------------------------------------------------
// First half of GetEnsemble (R/Clustering.R:131-138): PCA scores of t(S), cluster::pam for every k in lo..hi.
// table: (hi - lo + 1) x n row-major labels.
static void ensemble_labels(const double* d_S, int64_t n, int n_comp, int lo, int hi, const Compressed& csr,
                            const Compressed& csc, std::vector<int32_t>& table) {
  DevBuf<double> scores((size_t)n * n_comp);
  if (!(csr.nnz <= 128 * n && pca_scores_sparse(csr, csc, n, n_comp, scores.p))) {  // R :133, sparse route first
    std::vector<double> eig;
    pca_spectrum(d_S, n, n, eig, scores.p, n_comp);                       // R :133 (dense similarity)
  }
  const int R = hi - lo + 1;
  table.assign((size_t)R * n, 0);
  Phase ph("pam");
  PamEngine eng;
  eng.init(scores.p, n, n_comp, hi);
  std::vector<int> med;
  std::vector<int32_t> lab;
  for (int k = lo; k <= hi; ++k) {                                        // R :136-138
    eng.solve(k, med, lab);
    std::copy(lab.begin(), lab.end(), table.begin() + (size_t)(k - lo) * n);
  }
}

// GetEnsemble(data, tol, n_comp, tau, range = lo:hi) (R/Clustering.R:125-154): the truncated, symmetrised
// ensemble consensus matrix, n x n on the device.
void get_ensemble(const double* d_S, int64_t n, int n_comp, double tau, int lo, int hi, double* d_consen) {
  RS_REQUIRE(lo >= 1 && hi >= lo && hi < n, "get_ensemble: invalid range");
  RS_REQUIRE(n_comp >= 1 && n_comp <= n, "get_ensemble: invalid n_comp");
  Compressed csr, csc;
  compress(d_S, n, true, csr);
  compress(d_S, n, false, csc);
  std::vector<int32_t> table;
  ensemble_labels(d_S, n, n_comp, lo, hi, csr, csc, table);
  DevBuf<int32_t> d_table(table.size());
  d_table.upload(table.data(), table.size(), ctx().stream);
  consensus_dense(d_table.p, hi - lo + 1, n, tau, d_consen);              // R :143-153
}

void count_clusters(const double* d_S, int64_t n, double tol, int lo, int hi, int n_comp, int32_t* upper,
                    int32_t* lower, double* h_vals, int32_t* n_eigs) {
  cudaStream_t st = ctx().stream;
  RS_REQUIRE(lo >= 1 && hi >= lo && hi < n, "count_clusters: invalid range");
  RS_REQUIRE(n_comp >= 1 && n_comp <= n, "count_clusters: invalid n_comp");
  std::vector<double> vals;
  Compressed csr, csc;
  compress(d_S, n, true, csr);
  compress(d_S, n, false, csc);
  // R :63 -- only n_eigs of the similarity's Laplacian is used: Lanczos count on the sparse similarity
  int solo = 0;
  if (csr.nnz <= 128 * n && !getenv("RSOPTSC_EIGENGAP_DENSE")) {
    solo = laplacian_count_le(d_S, n, tol, csr, csc);
  } else {
    laplacian_spectrum(d_S, n, vals, &csc);
    for (double v : vals) solo += v <= tol;
  }
  const double tau = solo <= 5 ? 0.3 : (solo <= 10 ? 0.4 : 0.5);          // R :64-70
  const int R = hi - lo + 1;
  std::vector<int32_t> table;
  ensemble_labels(d_S, n, n_comp, lo, hi, csr, csc, table);               // GetEnsemble, R :131-138
  if (getenv("RSOPTSC_EIGENGAP_DENSE")) {                                 // the literal route (tests compare the two)
    DevBuf<int32_t> d_table(table.size());
    d_table.upload(table.data(), table.size(), st);
    DevBuf<double> consen((size_t)n * n);
    consensus_dense(d_table.p, R, n, tau, consen.p);                      // R :143-153
    laplacian_spectrum(consen.p, n, vals);                                // R :72
  } else {
    consensus_laplacian_spectrum(table, R, n, tau, vals);                 // R :143-153 + :72, by label signature
  }
  double best = -INFINITY;
  int arg = 0;
  for (int64_t i = 0; i + 1 < n; ++i) {                                   // R :74-75 (first maximum)
    const double g = vals[i + 1] - vals[i];
    if (g > best) { best = g; arg = (int)i + 1; }
  }
  int low = 0;
  for (double v : vals) low += v <= tol;                                  // R :78
  if (upper) *upper = arg;
  if (lower) *lower = low;
  if (h_vals) std::memcpy(h_vals, vals.data(), sizeof(double) * n);
  if (n_eigs) *n_eigs = low;
}

}  // namespace rs
