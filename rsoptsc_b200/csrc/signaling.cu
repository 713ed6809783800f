// Cell-cell signalling probability of one ligand-receptor pair:
// the n^2 elementwise core of GetSignalingPartners (R/Signaling.R:187-309), i.e. updateFlatBoth /
// updateFlatUp / updateFlatDown / updateFlatNone (:75-163) followed by the per-ligand-cell
// normalisation (:291-293).  The reference evaluates the n^2 entries with foreach %dopar%.
//   alpha_ij = exp(-1 / (L_i R_j))      (0 when L_i or R_j is 0)         getAlpha :24-30
//   K = alpha / (alpha + beta_j), D = alpha / (alpha + gamma_j)  (0/0 := 0)   getK/getD :1-15
//   both: alpha K D beta gamma | up: alpha K beta | down: alpha D gamma | none: alpha
//   P[i, j] = value_ij / sum_j value_ij   (NaN -> 0)
// HBM write-bound (8 n^2 B out).  Thread = ligand cell i (stores coalesced across i), grid.y = column
// chunk; the row sums are accumulated per chunk and combined in a fixed order (deterministic), and the
// values are recomputed in the write pass instead of being stored (two exp passes, no n^2 scratch).
#include "internal.h"

namespace rs {

static constexpr int ST = 128;
static constexpr int SIG_CHUNKS = 32;

__device__ __forceinline__ double sig_value(double L, double R, double b, double g, int has_b, int has_g) {
  const double a = (L == 0.0 || R == 0.0) ? 0.0 : exp(-1.0 / (L * R));
  double v = a;
  if (has_b) {
    const double k = (a == 0.0 && b == 0.0) ? 0.0 : a / (a + b);
    v = v * k;
  }
  if (has_g) {
    const double d = (a == 0.0 && g == 0.0) ? 0.0 : a / (a + g);
    v = v * d;
  }
  if (has_b) v = v * b;
  if (has_g) v = v * g;
  return v;
}

__global__ void k_signaling_rowsum(const double* __restrict__ lig, const double* __restrict__ rec,
                                   const double* __restrict__ beta, const double* __restrict__ gamma, int64_t n,
                                   int64_t chunk, double* __restrict__ partial /* SIG_CHUNKS x n */) {
  const int64_t i = (int64_t)blockIdx.x * ST + threadIdx.x;
  if (i >= n) return;
  const double L = lig[i];
  const int hb = beta != nullptr, hg = gamma != nullptr;
  const int64_t j0 = (int64_t)blockIdx.y * chunk, j1 = min(n, j0 + chunk);
  double sum = 0;
  for (int64_t j = j0; j < j1; ++j) sum += sig_value(L, rec[j], hb ? beta[j] : 0.0, hg ? gamma[j] : 0.0, hb, hg);
  partial[(int64_t)blockIdx.y * n + i] = sum;
}

__global__ void k_signaling_pair(const double* __restrict__ lig, const double* __restrict__ rec,
                                 const double* __restrict__ beta, const double* __restrict__ gamma, int64_t n,
                                 int64_t chunk, const double* __restrict__ partial, double* __restrict__ P) {
  const int64_t i = (int64_t)blockIdx.x * ST + threadIdx.x;
  if (i >= n) return;
  const double L = lig[i];
  const int hb = beta != nullptr, hg = gamma != nullptr;
  double sum = 0;
  if (partial)
    for (int c = 0; c < SIG_CHUNKS; ++c) sum += partial[(int64_t)c * n + i];  // fixed order: deterministic
  const int64_t j0 = (int64_t)blockIdx.y * chunk, j1 = min(n, j0 + chunk);
  for (int64_t j = j0; j < j1; ++j) {
    double v = sig_value(L, rec[j], hb ? beta[j] : 0.0, hg ? gamma[j] : 0.0, hb, hg);
    if (partial) {
      v = v / sum;
      if (isnan(v)) v = 0.0;  // rows with zero total (R :293)
    }
    P[i + j * n] = v;
  }
}

void signaling_pair(const double* d_lig, const double* d_rec, const double* d_beta, const double* d_gamma, int64_t n,
                    bool normalize, double* d_P) {
  Phase ph("signaling");
  const int64_t chunk = cdiv(n, SIG_CHUNKS);
  dim3 grid((unsigned)cdiv(n, ST), SIG_CHUNKS);
  DevBuf<double> partial;
  if (normalize) {
    partial.alloc((size_t)SIG_CHUNKS * n);
    RS_LAUNCH(k_signaling_rowsum, grid, ST, 0, d_lig, d_rec, d_beta, d_gamma, n, chunk, partial.p);
  }
  RS_LAUNCH(k_signaling_pair, grid, ST, 0, d_lig, d_rec, d_beta, d_gamma, n, chunk, normalize ? partial.p : nullptr,
            d_P);
}

}  // namespace rs
