// Preprocessing just upstream of the path (R/DataSelection.R):
//   LogNormalize         :117-133   per-cell max scaling, log10(scale * x + 1)
//   ScaleCenterData      :63-76     per-gene z-score (sd == 0 -> 1)
//   ApplyCountThreshold  :91-103    per-gene count filters
//   SelectData           :17-51     expression-frequency filter + PCA-loading feature selection
// Matrices are genes x cells, column-major.  Elementwise stages are HBM-bound streams; the PCA of
// SelectData reuses the centred-Gram + symmetric-eigensolver machinery of a2 (pca.cu).
#include <algorithm>
#include <cmath>
#include <numeric>

#include "gemm.h"
#include "internal.h"
#include "lanczos.h"

namespace rs {

static constexpr int PT = 256;

// ---- LogNormalize: one CTA per cell (column) ---------------------------------------------------
__global__ void k_log_normalize(const double* __restrict__ M, int64_t m, double scale, int normalize,
                                double* __restrict__ out) {
  __shared__ double sm[PT / 32];
  const double* src = M + (int64_t)blockIdx.x * m;
  double* dst = out + (int64_t)blockIdx.x * m;
  double mx = -INFINITY;
  if (normalize) {
    for (int64_t i = threadIdx.x; i < m; i += PT) mx = fmax(mx, src[i]);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = sm[0];
    for (int w = 1; w < PT / 32; ++w) mx = fmax(mx, sm[w]);
  }
  for (int64_t i = threadIdx.x; i < m; i += PT) {
    double x = src[i];
    if (normalize && mx > 0) x = x / mx;
    dst[i] = log10(scale * x + 1.0);
  }
}
void log_normalize(const double* d_M, int64_t m, int64_t n, double scale, bool normalize, double* d_out) {
  Phase ph("log_normalize");
  RS_LAUNCH(k_log_normalize, (unsigned)n, PT, 0, d_M, m, scale, normalize ? 1 : 0, d_out);
}

// ---- per-gene (row) statistics: thread per gene, coalesced across genes ----------------------------
// stats[0*m + g] = mean, [1*m + g] = sample sd, [2*m + g] = #non-zero, [3*m + g] = #(x >= countL), [4*m + g] = #(x >= countU)
__global__ void k_gene_stats(const double* __restrict__ M, int64_t m, int64_t n, double countL, double countU,
                             double* __restrict__ stats) {
  const int64_t g = (int64_t)blockIdx.x * PT + threadIdx.x;
  if (g >= m) return;
  double sum = 0, nnz = 0, cl = 0, cu = 0;
  for (int64_t j = 0; j < n; ++j) {
    const double x = M[g + j * m];
    sum += x;
    nnz += x != 0.0;
    cl += x >= countL;
    cu += x >= countU;
  }
  const double mean = sum / (double)n;
  double ss = 0;
  for (int64_t j = 0; j < n; ++j) {  // two-pass variance like R's sd()
    const double d = M[g + j * m] - mean;
    ss += d * d;
  }
  stats[g] = mean;
  stats[m + g] = n > 1 ? sqrt(ss / (double)(n - 1)) : NAN;
  stats[2 * m + g] = nnz;
  stats[3 * m + g] = cl;
  stats[4 * m + g] = cu;
}
void gene_stats(const double* d_M, int64_t m, int64_t n, double countL, double countU, std::vector<double>& h) {
  DevBuf<double> st((size_t)5 * m);
  RS_LAUNCH(k_gene_stats, (unsigned)cdiv(m, PT), PT, 0, d_M, m, n, countL, countU, st.p);
  h.resize((size_t)5 * m);
  st.download(h.data(), h.size(), ctx().stream);
  RS_CUDA(cudaStreamSynchronize(ctx().stream));
}

__global__ void k_scale_center(const double* __restrict__ M, int64_t m, int64_t n, const double* __restrict__ stats,
                               double* __restrict__ out) {
  const int64_t total = m * n;
  for (int64_t e = (int64_t)blockIdx.x * PT + threadIdx.x; e < total; e += (int64_t)gridDim.x * PT) {
    const int64_t g = e % m;
    double sd = stats[m + g];
    if (sd == 0.0) sd = 1.0;  // R :69
    out[e] = (M[e] - stats[g]) / sd;
  }
}
void scale_center(const double* d_M, int64_t m, int64_t n, double* d_out) {
  Phase ph("scale_center");
  DevBuf<double> st((size_t)5 * m);
  RS_LAUNCH(k_gene_stats, (unsigned)cdiv(m, PT), PT, 0, d_M, m, n, 0.0, 0.0, st.p);
  RS_LAUNCH(k_scale_center, 148 * 8, PT, 0, d_M, m, n, st.p, d_out);
}

// ---- SelectData ---------------------------------------------------------------------------------------
__global__ void k_gather_rows(const double* __restrict__ M, int64_t m, int64_t n, const int* __restrict__ rows,
                              int64_t mv, double* __restrict__ out) {
  const int64_t j = blockIdx.y;
  for (int64_t a = (int64_t)blockIdx.x * PT + threadIdx.x; a < mv; a += (int64_t)gridDim.x * PT)
    out[a + j * mv] = M[rows[a] + j * m];
}
// max_c |load[g, c]| over the first ncomp loading columns
__global__ void k_max_abs_rows(const double* __restrict__ L, int64_t mv, int ncomp, double* __restrict__ out) {
  const int64_t g = (int64_t)blockIdx.x * PT + threadIdx.x;
  if (g >= mv) return;
  double mx = 0;
  for (int c = 0; c < ncomp; ++c) mx = fmax(mx, fabs(L[g + (int64_t)c * mv]));
  out[g] = mx;
}

void select_data(const double* d_M, int64_t m, int64_t n, double gene_expression_threshold, int n_features,
                 std::vector<int>& selected /* 0-based rows of M */, int* max_component_out) {
  Phase ph("select_data");
  cudaStream_t s = ctx().stream;
  std::vector<double> st;
  gene_stats(d_M, m, n, 0.0, 0.0, st);
  std::vector<int> gene_use;
  if (gene_expression_threshold > 0) {                                     // R :22-30
    const double alpha = gene_expression_threshold / (double)n;
    for (int64_t g = 0; g < m; ++g) {
      const double f = st[2 * m + g] / (double)n;
      if (f > alpha && f < 1.0 - alpha) gene_use.push_back((int)g);
    }
  } else {
    gene_use.resize(m);
    std::iota(gene_use.begin(), gene_use.end(), 0);
  }
  const int64_t mv = (int64_t)gene_use.size();
  RS_REQUIRE(mv >= 3 && n >= 3, "select_data: fewer than 3 genes pass the expression filter");
  DevBuf<int> d_rows(mv);
  d_rows.upload(gene_use.data(), mv, s);
  DevBuf<double> Mv((size_t)mv * n);
  RS_REQUIRE(n <= 65535, "select_data: too many cells for this launch shape");
  dim3 grid((unsigned)std::min<int64_t>(cdiv(mv, PT), 64), (unsigned)n);
  RS_LAUNCH(k_gather_rows, grid, PT, 0, d_M, m, n, d_rows.p, mv, Mv.p);
  // prcomp(t(M_variable)): sdev (descending) and the loadings of the leading components   (R :33)
  std::vector<double> eig;
  pca_spectrum(Mv.p, mv, n, eig, nullptr, 0);
  const int64_t L = (int64_t)eig.size();
  std::vector<double> sdev(L);
  for (int64_t i = 0; i < L; ++i) sdev[i] = std::sqrt(eig[i]);
  // eigengaps = |sdev[-c(1,L)] - sdev[-(1:2)]| ; max_component = which.max + 1      (R :34-35)
  int64_t arg = 0;
  double best = -1;
  for (int64_t i = 0; i + 2 < L; ++i) {
    const double gap = std::fabs(sdev[i + 1] - sdev[i + 2]);
    if (gap > best) { best = gap; arg = i; }
  }
  const int ncomp = (int)std::min<int64_t>(arg + 2, L);
  if (max_component_out) *max_component_out = ncomp;
  DevBuf<double> load((size_t)mv * ncomp), maxc(mv);
  pca_loadings(Mv.p, mv, n, ncomp, load.p);
  RS_LAUNCH(k_max_abs_rows, (unsigned)cdiv(mv, PT), PT, 0, load.p, mv, ncomp, maxc.p);
  std::vector<double> h(mv);
  maxc.download(h.data(), mv, s);
  RS_CUDA(cudaStreamSynchronize(s));
  // keep every gene whose max coefficient reaches the adj_n-th highest one           (R :44-48)
  const int64_t adj_n = std::min<int64_t>(mv, n_features);
  RS_REQUIRE(adj_n >= 1, "select_data: n_features must be positive");
  std::vector<double> sorted(h);
  std::sort(sorted.begin(), sorted.end(), std::greater<double>());
  const double nth = sorted[adj_n - 1];
  selected.clear();
  for (int64_t g = 0; g < mv; ++g)
    if (h[g] >= nth) selected.push_back(gene_use[g]);
}

}  // namespace rs
