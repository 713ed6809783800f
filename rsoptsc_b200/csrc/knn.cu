// a3: exact K nearest neighbours in the low-dimensional embedding (R/SimilarityM.R:76-77,83-84;
// FNN::get.knnx on the first 3 embedding columns, self included).
//
// Warp-level selection: one warp owns one query cell and streams all n candidates through
// shared-memory tiles (coalesced loads, one candidate per lane per step).  The running top-K
// (K <= 32) is held one entry per lane, sorted ascending by (distance, index); a candidate that
// beats the current K-th entry is inserted with a ballot + shuffle shift.  After the first few
// tiles insertions are rare (expected K ln(n/K) per query), so the kernel is bound by the FP64
// distance arithmetic: 3 sub, 3 mul, 2 add per pair, evaluated without FMA contraction in the
// reference's left-to-right order so that index sets match the CPU oracle bit for bit.
#include "internal.h"

namespace rs {

static constexpr int KNN_WARPS = 8;     // queries per CTA
static constexpr int KNN_TILE = 512;    // candidates staged per tile (DIM x 512 doubles <= 32 KB)
static constexpr int KNN_MAXDIM = 8;

template <int DIM>
__global__ void __launch_bounds__(KNN_WARPS * 32)
k_knn(const double* __restrict__ emb, int64_t n, int64_t ld, int K, int32_t* __restrict__ idx) {
  __shared__ double tile[DIM][KNN_TILE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t qi = (int64_t)blockIdx.x * KNN_WARPS + warp;
  const bool active = qi < n;
  double qc[DIM];
#pragma unroll
  for (int c = 0; c < DIM; ++c) qc[c] = active ? emb[qi + c * ld] : 0.0;
  // lane t holds the t-th best (distance, index); lanes >= K are kept at +inf and never win
  double best_d = INFINITY;
  int best_i = 0x7fffffff;
  for (int64_t t0 = 0; t0 < n; t0 += KNN_TILE) {
    __syncthreads();
    const int cnt = (int)min((int64_t)KNN_TILE, n - t0);
    for (int e = threadIdx.x; e < cnt; e += KNN_WARPS * 32)
#pragma unroll
      for (int c = 0; c < DIM; ++c) tile[c][e] = emb[t0 + e + c * ld];
    __syncthreads();
    if (!active) continue;
    for (int e0 = 0; e0 < cnt; e0 += 32) {
      const int e = e0 + lane;
      double d = INFINITY;
      const int ci = (int)(t0 + e);
      if (e < cnt) {
        double diff = __dsub_rn(qc[0], tile[0][e]);
        d = __dmul_rn(diff, diff);
#pragma unroll
        for (int c = 1; c < DIM; ++c) {
          diff = __dsub_rn(qc[c], tile[c][e]);
          d = __dadd_rn(d, __dmul_rn(diff, diff));
        }
      }
      // current K-th best (threshold), warp-uniform
      const double thr_d = __shfl_sync(0xffffffffu, best_d, K - 1);
      const int thr_i = __shfl_sync(0xffffffffu, best_i, K - 1);
      unsigned pending = __ballot_sync(0xffffffffu, e < cnt && (d < thr_d || (d == thr_d && ci < thr_i)));
      while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        const double cd = __shfl_sync(0xffffffffu, d, src);
        const int cidx = __shfl_sync(0xffffffffu, ci, src);
        // re-check against the (possibly tightened) threshold
        const double td = __shfl_sync(0xffffffffu, best_d, K - 1);
        const int ti = __shfl_sync(0xffffffffu, best_i, K - 1);
        if (!(cd < td || (cd == td && cidx < ti))) continue;
        const bool less = best_d < cd || (best_d == cd && best_i < cidx);  // my entry stays before the candidate
        const int pos = __popc(__ballot_sync(0xffffffffu, less));
        const double up_d = __shfl_up_sync(0xffffffffu, best_d, 1);
        const int up_i = __shfl_up_sync(0xffffffffu, best_i, 1);
        if (lane == pos) { best_d = cd; best_i = cidx; }
        else if (lane > pos) { best_d = up_d; best_i = up_i; }
        if (lane >= K) { best_d = INFINITY; best_i = 0x7fffffff; }
      }
    }
  }
  if (active && lane < K) idx[qi * K + lane] = best_i;
}

void knn_bruteforce(const double* d_emb, int64_t n, int64_t ld, int dim, int K, int32_t* d_idx) {
  Phase ph("knn");
  RS_REQUIRE(K >= 1 && K <= 32, "knn: K must be in [1, 32]");
  RS_REQUIRE(K <= n, "knn: K exceeds the number of cells");
  RS_REQUIRE(dim >= 1 && dim <= KNN_MAXDIM, "knn: embedding dimension must be in [1, 8]");
  const unsigned grid = (unsigned)cdiv(n, KNN_WARPS);
  switch (dim) {
    case 1: RS_LAUNCH(k_knn<1>, grid, KNN_WARPS * 32, 0, d_emb, n, ld, K, d_idx); break;
    case 2: RS_LAUNCH(k_knn<2>, grid, KNN_WARPS * 32, 0, d_emb, n, ld, K, d_idx); break;
    case 3: RS_LAUNCH(k_knn<3>, grid, KNN_WARPS * 32, 0, d_emb, n, ld, K, d_idx); break;
    case 4: RS_LAUNCH(k_knn<4>, grid, KNN_WARPS * 32, 0, d_emb, n, ld, K, d_idx); break;
    case 5: RS_LAUNCH(k_knn<5>, grid, KNN_WARPS * 32, 0, d_emb, n, ld, K, d_idx); break;
    case 6: RS_LAUNCH(k_knn<6>, grid, KNN_WARPS * 32, 0, d_emb, n, ld, K, d_idx); break;
    case 7: RS_LAUNCH(k_knn<7>, grid, KNN_WARPS * 32, 0, d_emb, n, ld, K, d_idx); break;
    default: RS_LAUNCH(k_knn<8>, grid, KNN_WARPS * 32, 0, d_emb, n, ld, K, d_idx); break;
  }
}

}  // namespace rs
