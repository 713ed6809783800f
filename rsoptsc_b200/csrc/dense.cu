// Dense / support-indexed elementwise kernels of the SimilarityM path (HBM-bound stages):
//   a1  column normalisation            R/SimilarityM.R:34-39
//   A = Z - Y3/mu, Y3 += mu (Z - J)     R/SimilarityM.R:142,187
//   a11 threshold + symmetrise          R/SimilarityM.R:94-95
// All matrices are column-major FP64.  Reductions are two-stage with a fixed order so that
// results are run-to-run deterministic.
#include "comm.h"
#include "internal.h"

namespace rs {

static constexpr int RED_BLOCKS = 592;  // 4 x 148 SMs
static constexpr int RED_THREADS = 256;

// ---- deterministic reductions -------------------------------------------------------
struct RedScratch {
  DevBuf<double> partial, result;
  void ensure() {
    if (!partial.p) partial.alloc(RED_BLOCKS * 2);
    if (!result.p) result.alloc(2);
  }
};
static RedScratch& red_storage() {
  static RedScratch r;
  return r;
}
static RedScratch& red() {
  RedScratch& r = red_storage();
  r.ensure();
  return r;
}
void release_dense_scratch() {  // rsoptsc_finalize / device switch: hand the scratch back without re-creating it
  red_storage().partial.release();
  red_storage().result.release();
}

__global__ void k_final_sum(const double* __restrict__ partial, int count, double* __restrict__ out) {
  __shared__ double sm[RED_THREADS / 32 + 1];
  double v = 0;
  for (int i = threadIdx.x; i < count; i += RED_THREADS) v += partial[i];
  v = block_sum<RED_THREADS>(v, sm);
  if (threadIdx.x == 0) out[0] = v;
}

static double finish_sum(int nblocks) {
  RS_LAUNCH(k_final_sum, 1, RED_THREADS, 0, red().partial.p, nblocks, red().result.p);
  double h = 0;
  RS_CUDA(cudaMemcpyAsync(&h, red().result.p, sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
  RS_CUDA(cudaStreamSynchronize(ctx().stream));
  return h;
}

__global__ void k_frob2(const double* __restrict__ a, int64_t count, double* __restrict__ partial) {
  __shared__ double sm[RED_THREADS / 32 + 1];
  double v = 0;
  const int64_t stride = (int64_t)gridDim.x * RED_THREADS;
  for (int64_t i = (int64_t)blockIdx.x * RED_THREADS + threadIdx.x; i < count; i += stride) {
    const double x = a[i];
    v += x * x;
  }
  v = block_sum<RED_THREADS>(v, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = v;
}

double frob2(const double* d_a, int64_t count) {
  Phase ph("dense_ew");
  RS_LAUNCH(k_frob2, RED_BLOCKS, RED_THREADS, 0, d_a, count, red().partial.p);
  return finish_sum(RED_BLOCKS);
}

// ---- a1: normalisation ---------------------------------------------------------------
__global__ void k_minmax(const double* __restrict__ a, int64_t count, double* __restrict__ pmin,
                         double* __restrict__ pmax) {
  __shared__ double smn[RED_THREADS / 32], smx[RED_THREADS / 32];
  double mn = INFINITY, mx = -INFINITY;
  const int64_t stride = (int64_t)gridDim.x * RED_THREADS;
  for (int64_t i = (int64_t)blockIdx.x * RED_THREADS + threadIdx.x; i < count; i += stride) {
    const double x = a[i];
    mn = fmin(mn, x);
    mx = fmax(mx, x);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < RED_THREADS / 32; ++w) { mn = fmin(mn, smn[w]); mx = fmax(mx, smx[w]); }
    pmin[blockIdx.x] = mn;
    pmax[blockIdx.x] = mx;
  }
}
__global__ void k_minmax_final(const double* __restrict__ pmin, const double* __restrict__ pmax, int count,
                               double* __restrict__ out) {
  if (threadIdx.x == 0) {
    double mn = INFINITY, mx = -INFINITY;
    for (int i = 0; i < count; ++i) { mn = fmin(mn, pmin[i]); mx = fmax(mx, pmax[i]); }
    out[0] = mn;
    out[1] = mx - mn;  // max(data - min)
  }
}
// One CTA per column: x = ((d - mn) / mx) / ||.||_2   (two passes over the column, 2nd from cache)
__global__ void k_normalize_cols(const double* __restrict__ data, int64_t m, const double* __restrict__ mnmx,
                                 double* __restrict__ X) {
  __shared__ double sm[RED_THREADS / 32 + 1];
  const double mn = mnmx[0], mx = mnmx[1];
  const double* src = data + (int64_t)blockIdx.x * m;
  double* dst = X + (int64_t)blockIdx.x * m;
  double acc = 0;
  for (int64_t i = threadIdx.x; i < m; i += RED_THREADS) {
    const double x = (src[i] - mn) / mx;
    acc += x * x;
  }
  const double nrm = sqrt(block_sum<RED_THREADS>(acc, sm));
  for (int64_t i = threadIdx.x; i < m; i += RED_THREADS) dst[i] = ((src[i] - mn) / mx) / nrm;
}

void normalize_columns(const double* d_data, int64_t m, int64_t n, double* d_X) {
  Phase ph("normalize");
  double* part = red().partial.p;
  RS_LAUNCH(k_minmax, RED_BLOCKS, RED_THREADS, 0, d_data, m * n, part, part + RED_BLOCKS);
  RS_LAUNCH(k_minmax_final, 1, 32, 0, part, part + RED_BLOCKS, RED_BLOCKS, red().result.p);
  RS_LAUNCH(k_normalize_cols, (unsigned)n, RED_THREADS, 0, d_data, m, red().result.p, d_X);
}

// ---- A = Z - Y3/mu -------------------------------------------------------------------
// (y3 and a start at the same offset of two 256-byte aligned allocations: one scalar head element
// restores the 16-byte alignment of the double2 body when that offset is odd)
__global__ void k_scale_neg(const double* __restrict__ y3, double mu, int64_t count, double* __restrict__ a) {
  const int64_t head = (reinterpret_cast<uintptr_t>(y3) & 8) ? 1 : 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 2;
  for (int64_t i = head + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i + 1 < count; i += stride) {
    const double2 v = *reinterpret_cast<const double2*>(y3 + i);
    *reinterpret_cast<double2*>(a + i) = make_double2(-(v.x / mu), -(v.y / mu));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (head && count > 0) a[0] = -(y3[0] / mu);
    if (((count - head) & 1) && count > head) a[count - 1] = -(y3[count - 1] / mu);
  }
}
__global__ void k_scatter_add(const int* __restrict__ rowof, const int* __restrict__ col, const double* __restrict__ zs,
                              int64_t nnz, int64_t n, double scale, double* __restrict__ a) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < nnz) a[(int64_t)rowof[p] + (int64_t)col[p] * n] += scale * zs[p];
}

// entry p of the support lives at block-storage offset off[p] (Support::dense_off); `sel` (may be null)
// restricts the pass to a list of entries (the columns a rank owns)
__global__ void k_scatter_add_off(const long long* __restrict__ off, const double* __restrict__ zs, int64_t cnt,
                                  double scale, double* __restrict__ a, const int* __restrict__ sel = nullptr) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt) return;
  const int64_t p = sel ? sel[t] : t;
  a[off[p]] += scale * zs[p];
}

// A = Z - Y3/mu on the block-diagonal storage; fro2[c] = ||A_c||_F^2 per connected component
void build_A(const Support& s, const double* d_Zs, const double* d_Y3, double mu, double* d_A,
             std::vector<double>& fro2, const ShardPlan* plan) {
  Phase ph("build_A");
  const int64_t nn = s.total_dense;
  if (plan && plan->on) {
    // owned columns only; the blocks that are split across the ranks are then gathered (their SVT is collective)
    for (const auto& sg : plan->segs) {
      const int blocks = (int)std::min<int64_t>(RED_BLOCKS * 2, std::max<int64_t>(1, sg.count / 2048));
      RS_LAUNCH(k_scale_neg, blocks, 256, 0, d_Y3 + sg.off, mu, sg.count, d_A + sg.off);
    }
    if (plan->nent())
      RS_LAUNCH(k_scatter_add_off, cdiv(plan->nent(), 256), 256, 0, s.dense_off.p, d_Zs, plan->nent(), 1.0, d_A,
                plan->d_own_entries.p);
    for (size_t c = 0; c < s.comp_n.size(); ++c)
      if (plan->big[c]) comm_allgather_cols(d_A + s.comp_off[c], s.comp_n[c], plan->cb[c]);
  } else {
    RS_LAUNCH(k_scale_neg, RED_BLOCKS * 2, 256, 0, d_Y3, mu, nn, d_A);
    if (s.nnz) RS_LAUNCH(k_scatter_add_off, cdiv(s.nnz, 256), 256, 0, s.dense_off.p, d_Zs, s.nnz, 1.0, d_A);
  }
  fro2.assign(s.comp_n.size(), 0.0);
  for (size_t c = 0; c < s.comp_n.size(); ++c) {
    if (plan && plan->on && !plan->big[c] && plan->owner[c] != plan->rank) continue;
    const int64_t cnt = s.comp_n[c] * s.comp_n[c];
    const int blocks = (int)std::min<int64_t>(RED_BLOCKS, std::max<int64_t>(1, cnt / 4096));
    RS_LAUNCH(k_frob2, blocks, RED_THREADS, 0, d_A + s.comp_off[c], cnt, red().partial.p);
    fro2[c] = finish_sum(blocks);
  }
}

// ---- q[p] = Zs - J + Y3/mu on the support -------------------------------------------
__global__ void k_gather_term(const long long* __restrict__ doff, const double* __restrict__ zs,
                              const double* __restrict__ J, const double* __restrict__ y3, double mu, int64_t cnt,
                              double* __restrict__ q, const int* __restrict__ sel) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cnt) return;
  const int64_t p = sel ? sel[t] : t;
  const int64_t off = doff[p];
  const double j = J ? J[off] : 0.0;
  q[p] = zs[p] - j + y3[off] / mu;
}
void gather_support_term(const Support& s, const double* d_Zs, const double* d_J, const double* d_Y3, double mu,
                         double* d_q, const ShardPlan* plan) {
  Phase ph("gather");
  const bool sh = plan && plan->on;
  const int64_t cnt = sh ? plan->nent() : s.nnz;
  if (cnt)
    RS_LAUNCH(k_gather_term, cdiv(cnt, 256), 256, 0, s.dense_off.p, d_Zs, d_J, d_Y3, mu, cnt, d_q,
              sh ? plan->d_own_entries.p : (const int*)nullptr);
}

// ---- Y3 += mu (Z - J) ------------------------------------------------------------------
__global__ void k_y3_sub_J(const double* __restrict__ J, double mu, int64_t count, double* __restrict__ y3,
                           double* __restrict__ partial) {
  __shared__ double sm[RED_THREADS / 32 + 1];
  double acc = 0;
  const int64_t head = (reinterpret_cast<uintptr_t>(y3) & 8) ? 1 : 0;  // see k_scale_neg
  const int64_t stride = (int64_t)gridDim.x * RED_THREADS * 2;
  for (int64_t i = head + ((int64_t)blockIdx.x * RED_THREADS + threadIdx.x) * 2; i + 1 < count; i += stride) {
    const double2 j = *reinterpret_cast<const double2*>(J + i);
    double2 y = *reinterpret_cast<double2*>(y3 + i);
    y.x -= mu * j.x;
    y.y -= mu * j.y;
    *reinterpret_cast<double2*>(y3 + i) = y;
    acc += j.x * j.x + j.y * j.y;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (head && count > 0) {
      const double j = J[0];
      y3[0] -= mu * j;
      acc += j * j;
    }
    if (((count - head) & 1) && count > head) {
      const double j = J[count - 1];
      y3[count - 1] -= mu * j;
      acc += j * j;
    }
  }
  acc = block_sum<RED_THREADS>(acc, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
// sparse part: Y3[i,j] += mu*Zs ; partial sums of (Zs-J)^2 - J^2 over the support
__global__ void k_y3_add_Z(const long long* __restrict__ doff, const double* __restrict__ zs,
                           const double* __restrict__ J, double mu, int64_t cnt, double* __restrict__ y3,
                           double* __restrict__ partial, const int* __restrict__ sel) {
  __shared__ double sm[RED_THREADS / 32 + 1];
  double acc = 0;
  const int64_t stride = (int64_t)gridDim.x * RED_THREADS;
  for (int64_t t = (int64_t)blockIdx.x * RED_THREADS + threadIdx.x; t < cnt; t += stride) {
    const int64_t p = sel ? sel[t] : t;
    const int64_t off = doff[p];
    const double z = zs[p];
    const double j = J ? J[off] : 0.0;
    y3[off] += mu * z;
    acc += (z - j) * (z - j) - j * j;
  }
  acc = block_sum<RED_THREADS>(acc, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// Returns this rank's share of ||Z - J||_F^2 (the caller sums over ranks; NaN propagates to its finiteness check).
double update_Y3(const Support& s, const double* d_Zs, const double* d_J, double mu, double* d_Y3,
                 const ShardPlan* plan) {
  Phase ph("update_Y3");
  const int64_t nn = s.total_dense;  // block-diagonal storage (sum n_c^2)
  const bool sh = plan && plan->on;
  double total = 0;
  // order matters: the sparse kernel reads J on the support, so it may run after the dense one
  if (d_J && !sh) {
    RS_LAUNCH(k_y3_sub_J, RED_BLOCKS, RED_THREADS, 0, d_J, mu, nn, d_Y3, red().partial.p);
    total += finish_sum(RED_BLOCKS);
  } else if (d_J) {
    for (const auto& sg : plan->segs) {
      const int blocks = (int)std::min<int64_t>(RED_BLOCKS, std::max<int64_t>(1, sg.count / 4096));
      RS_LAUNCH(k_y3_sub_J, blocks, RED_THREADS, 0, d_J + sg.off, mu, sg.count, d_Y3 + sg.off, red().partial.p);
      total += finish_sum(blocks);
    }
  }
  const int blocks = 148;
  const int64_t cnt = sh ? plan->nent() : s.nnz;
  RS_LAUNCH(k_y3_add_Z, blocks, RED_THREADS, 0, s.dense_off.p, d_Zs, d_J, mu, cnt, d_Y3, red().partial.p,
            sh ? plan->d_own_entries.p : (const int*)nullptr);
  total += finish_sum(blocks);
  return total;
}

// ---- upper triangle <- lower triangle (column-major d x d) -----------------------------------
__global__ void k_mirror_lower(double* __restrict__ B, int64_t d, int64_t j0) {
  const int64_t j = j0 + blockIdx.y;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < j; i += (int64_t)gridDim.x * 256) B[i + j * d] = B[j + i * d];
}
void mirror_lower(double* d_B, int64_t d) {
  for (int64_t j0 = 1; j0 < d; j0 += 65535) {  // column 0 has nothing above the diagonal
    const int64_t nc = std::min<int64_t>(65535, d - j0);
    dim3 grid((unsigned)std::min<int64_t>(cdiv(d, 256), 16), (unsigned)nc);
    RS_LAUNCH(k_mirror_lower, grid, 256, 0, d_B, d, j0);
  }
}

// ---- dst[:, c] = src[:, c] * scale[c] ----------------------------------------------------------
__global__ void k_copy_scale_columns(const double* __restrict__ src, int64_t rows, const double* __restrict__ scale,
                                     double* __restrict__ dst) {
  const double sc = scale[blockIdx.y];
  const double* s = src + (int64_t)blockIdx.y * rows;
  double* d = dst + (int64_t)blockIdx.y * rows;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = s[i] * sc;
}
void copy_scale_columns(const double* d_src, int64_t rows, int64_t cols, const double* d_scale, double* d_dst) {
  for (int64_t c0 = 0; c0 < cols; c0 += 65535) {
    const int64_t nc = std::min<int64_t>(65535, cols - c0);
    dim3 grid((unsigned)std::min<int64_t>(cdiv(rows, 256), 8), (unsigned)nc);
    RS_LAUNCH(k_copy_scale_columns, grid, 256, 0, d_src + c0 * rows, rows, d_scale + c0, d_dst + c0 * rows);
  }
}

// ---- T[:, c] *= scale[c] --------------------------------------------------------------
__global__ void k_scale_columns(double* __restrict__ T, int64_t rows, const double* __restrict__ scale) {
  const double sc = scale[blockIdx.y];
  double* colp = T + (int64_t)blockIdx.y * rows;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (int64_t)gridDim.x * blockDim.x)
    colp[i] *= sc;
}
void scale_columns(double* d_T, int64_t rows, int64_t cols, const double* d_scale) {
  if (!cols) return;
  RS_REQUIRE(cols <= 65535 * 64, "scale_columns: too many columns");
  // gridDim.y <= 65535: loop in slabs
  for (int64_t c0 = 0; c0 < cols; c0 += 65535) {
    const int64_t nc = cols - c0 < 65535 ? cols - c0 : 65535;
    dim3 grid((unsigned)std::min<int64_t>(cdiv(rows, 256), 8), (unsigned)nc);
    RS_LAUNCH(k_scale_columns, grid, 256, 0, d_T + c0 * rows, rows, d_scale + c0);
  }
}

// ---- a11: threshold + symmetrise --------------------------------------------------------
__device__ __forceinline__ double thr_abs(double z) { return z <= 2.220446049250313e-16 ? 0.0 : fabs(z); }

// Tiled transpose-add: each 32x32 tile of W reads the tile of Z at (bi,bj) straight and the
// tile at (bj,bi) through shared memory.  HBM: 16 n^2 B read + 8 n^2 B written.
__global__ void k_symmetrise_dense(const double* __restrict__ Z, int64_t n, double* __restrict__ W) {
  __shared__ double tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
  // load transposed partner tile: rows j0.., cols i0..
  for (int c = threadIdx.y; c < 32; c += blockDim.y) {
    const int64_t r = j0 + threadIdx.x, cc = i0 + c;
    tile[c][threadIdx.x] = (r < n && cc < n) ? thr_abs(Z[r + cc * n]) : 0.0;
  }
  __syncthreads();
  for (int c = threadIdx.y; c < 32; c += blockDim.y) {
    const int64_t r = i0 + threadIdx.x, cc = j0 + c;
    if (r < n && cc < n) W[r + cc * n] = 0.5 * (thr_abs(Z[r + cc * n]) + tile[threadIdx.x][c]);
  }
}
void threshold_symmetrise_dense(const double* d_Z, int64_t n, double* d_W) {
  Phase ph("symm");
  dim3 grid((unsigned)cdiv(n, 32), (unsigned)cdiv(n, 32)), block(32, 8);
  RS_LAUNCH(k_symmetrise_dense, grid, block, 0, d_Z, n, d_W);
}

__global__ void k_symmetrise_sparse(const int* __restrict__ rowof, const int* __restrict__ col,
                                    const double* __restrict__ zs, int64_t nnz, int64_t n, double* __restrict__ W) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nnz) return;
  const double h = 0.5 * thr_abs(zs[p]);
  if (h == 0.0) return;
  const int64_t i = rowof[p], j = col[p];
  // each W entry receives at most two addends (from (i,j) and (j,i)): order-independent
  atomicAdd(&W[i + j * n], h);
  atomicAdd(&W[j + i * n], h);
}
void threshold_symmetrise_sparse(const Support& s, const double* d_Zs, double* d_W) {
  Phase ph("symm");
  RS_CUDA(cudaMemsetAsync(d_W, 0, sizeof(double) * s.n * s.n, ctx().stream));
  if (s.nnz) RS_LAUNCH(k_symmetrise_sparse, cdiv(s.nnz, 256), 256, 0, s.rowof.p, s.col.p, d_Zs, s.nnz, s.n, d_W);
}
void scatter_support_blocks(const Support& s, const double* d_Zs, double* d_Zb, const ShardPlan* plan) {
  RS_CUDA(cudaMemsetAsync(d_Zb, 0, sizeof(double) * s.total_dense, ctx().stream));
  const bool sh = plan && plan->on;
  const int64_t cnt = sh ? plan->nent() : s.nnz;
  if (cnt)
    RS_LAUNCH(k_scatter_add_off, cdiv(cnt, 256), 256, 0, s.dense_off.p, d_Zs, cnt, 1.0, d_Zb,
              sh ? plan->d_own_entries.p : (const int*)nullptr);
}
// tmp[p] = Zs[p] for the owned entries, 0 elsewhere; summed over the ranks every entry has exactly one owner
__global__ void k_copy_selected(const double* __restrict__ src, const int* __restrict__ sel, int64_t cnt,
                                double* __restrict__ dst) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < cnt) dst[sel[t]] = src[sel[t]];
}
void assemble_support_values(const Support& s, const ShardPlan& plan, double* d_Zs) {
  if (!plan.on || !s.nnz) return;
  DevBuf<double> tmp((size_t)s.nnz);
  tmp.zero(ctx().stream);
  if (plan.nent()) RS_LAUNCH(k_copy_selected, cdiv(plan.nent(), 256), 256, 0, d_Zs, plan.d_own_entries.p, plan.nent(), tmp.p);
  comm_allreduce_sum(tmp.p, (size_t)s.nnz);
  RS_CUDA(cudaMemcpyAsync(d_Zs, tmp.p, sizeof(double) * s.nnz, cudaMemcpyDeviceToDevice, ctx().stream));
  RS_CUDA(cudaStreamSynchronize(ctx().stream));
}
void scatter_support_dense(const Support& s, const double* d_Zs, double* d_Z) {
  Phase ph("symm");
  RS_CUDA(cudaMemsetAsync(d_Z, 0, sizeof(double) * s.n * s.n, ctx().stream));
  if (s.nnz) RS_LAUNCH(k_scatter_add, cdiv(s.nnz, 256), 256, 0, s.rowof.p, s.col.p, d_Zs, s.nnz, s.n, 1.0, d_Z);
}

}  // namespace rs
