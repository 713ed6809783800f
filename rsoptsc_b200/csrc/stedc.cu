// Divide-and-conquer eigensolver for the symmetric tridiagonal matrix produced by sytrd_kernel.cuh:
// the second stage of the symmetric eigensolver behind the SVT (a4, R/SimilarityM.R:142-155), the PCA
// (a2) and the Laplacian spectrum (a14).  Numerics in stedc_core.h (shared with the CPU harness
// tests/stedc_host_check.cpp); this file is the B200 mapping:
//
//   leaves (<= 64 rows)   one CTA per leaf, implicit QL in shared memory
//   per level             one gather of the coupling rows, merge planning on the host (O(n log n):
//                         sort, deflation), then per merge on the device:
//     secular equation    one warp per root, sums over poles split over the lanes
//     Loewner weights     one warp per pole
//     M = R S             the k x k eigenvectors of D + rho z z^T scattered straight into the n_m x n_m
//                         update matrix (deflated columns = unit vectors, Givens rotations as row
//                         operations), so that the whole merge is Q_out = diag(Q1, Q2) M: two GEMMs
//     GEMM                tcgen05 split-integer FP64-equivalent GEMM (gemm_ozaki.cu) for blocks of
//                         1536 rows and more, cuBLAS below
//
// The O(n^3) work is entirely in the GEMMs; everything else is O(n^2).
#include <cmath>

#include "comm.h"
#include "gemm.h"
#include "internal.h"
#include "stedc_core.h"

namespace rs {

namespace {

using stedc::MergePlan;
using stedc::Rotation;

constexpr int LEAF = 64;

// ---- leaves: implicit QL (tql2 organisation) on a <= 64 x 64 tridiagonal block ---------------------
// Thread 0 runs the scalar recurrences of one sweep and records its plane rotations; all threads then
// apply the sweep to their own row of the eigenvector matrix.
__global__ void __launch_bounds__(LEAF) k_leaf_eig(const double* __restrict__ dmod, const double* __restrict__ e,
                                                   const int* __restrict__ loff, double* Q, long long ldq, double* lam) {
  __shared__ double sd[LEAF], se[LEAF + 1], sc[LEAF], ss[LEAF];
  __shared__ double Z[LEAF * LEAF];
  __shared__ int ctl[3];  // [0] sweep pending, [1] l, [2] m
  const int tid = threadIdx.x;
  const int off = loff[blockIdx.x], s = loff[blockIdx.x + 1] - off;
  if (tid < s) {
    sd[tid] = dmod[off + tid];
    se[tid] = (tid + 1 < s) ? e[off + tid] : 0.0;
  }
  for (int c = 0; c < s; ++c)
    if (tid < s) Z[tid + c * LEAF] = (tid == c) ? 1.0 : 0.0;
  __shared__ double sf, stst;
  if (tid == 0) { sf = 0.0; stst = 0.0; }
  __syncthreads();
  const double eps = stedc::EPS;
  for (int l = 0; l < s; ++l) {
    for (int iter = 0; iter < 200; ++iter) {
      if (tid == 0) {
        if (iter == 0) stst = fmax(stst, fabs(sd[l]) + fabs(se[l]));
        int mm = l;
        while (mm < s) {
          if (fabs(se[mm]) <= eps * stst) break;
          ++mm;
        }
        if (mm >= s) mm = s - 1;
        ctl[0] = 0;
        if (mm > l && (iter == 0 || fabs(se[l]) > eps * stst)) {
          ctl[0] = 1; ctl[1] = l; ctl[2] = mm;
          double g = sd[l];
          double p = (sd[l + 1] - g) / (2.0 * se[l]);
          double r = hypot(p, 1.0);
          if (p < 0) r = -r;
          sd[l] = se[l] / (p + r);
          sd[l + 1] = se[l] * (p + r);
          const double dl1 = sd[l + 1];
          double h = g - sd[l];
          for (int i = l + 2; i < s; ++i) sd[i] -= h;
          sf += h;
          p = sd[mm];
          double c = 1.0, c2 = c, c3 = c;
          const double el1 = se[l + 1];
          double sn = 0.0, s2 = 0.0;
          for (int i = mm - 1; i >= l; --i) {
            c3 = c2; c2 = c; s2 = sn;
            g = c * se[i];
            h = c * p;
            r = hypot(p, se[i]);
            se[i + 1] = sn * r;
            sn = se[i] / r;
            c = p / r;
            p = c * sd[i] - sn * g;
            sd[i + 1] = h + sn * (c * g + sn * sd[i]);
            sc[i] = c;
            ss[i] = sn;
          }
          p = -sn * s2 * c3 * el1 * se[l] / dl1;
          se[l] = sn * p;
          sd[l] = c * p;
        }
      }
      __syncthreads();
      if (!ctl[0]) break;
      const int l0 = ctl[1], mm = ctl[2];
      if (tid < s) {
        for (int i = mm - 1; i >= l0; --i) {
          const double c = sc[i], sn = ss[i];
          const double h = Z[tid + (i + 1) * LEAF], zi = Z[tid + i * LEAF];
          Z[tid + (i + 1) * LEAF] = sn * zi + c * h;
          Z[tid + i * LEAF] = c * zi - sn * h;
        }
      }
      __syncthreads();
    }
    if (tid == 0) { sd[l] += sf; se[l] = 0.0; }
    __syncthreads();
  }
  if (tid < s) lam[off + tid] = sd[tid];
  for (int c = 0; c < s; ++c)
    if (tid < s) Q[(off + tid) + (long long)(off + c) * ldq] = Z[tid + c * LEAF];
}

__global__ void k_gather_z(const double* __restrict__ Q, long long ldq, int n, const int* __restrict__ zrow,
                           const double* __restrict__ zsgn, double* __restrict__ z) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) z[c] = zsgn[c] == 0.0 ? 0.0 : zsgn[c] * Q[zrow[c] + (long long)c * ldq];
}

struct WarpSum {
  __device__ double operator()(double v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
};

__global__ void __launch_bounds__(256) k_secular(int k, const double* __restrict__ dl, const double* __restrict__ zz,
                                                 double rho, int* __restrict__ org, double* __restrict__ tau,
                                                 double* __restrict__ lam) {
  const int j = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (j >= k) return;
  int o;
  double t;
  stedc::secular_root(k, j, dl, zz, rho, lane, 32, WarpSum(), &o, &t);
  if (lane == 0) { org[j] = o; tau[j] = t; lam[j] = dl[o] + t; }
}

__global__ void __launch_bounds__(256) k_zhat(int k, const double* __restrict__ dl, const double* __restrict__ zz,
                                              const int* __restrict__ org, const double* __restrict__ tau,
                                              double* __restrict__ zhat) {
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= k) return;
  double p = stedc::loewner_partial(k, i, dl, org, tau, lane, 32);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) p *= __shfl_xor_sync(0xffffffffu, p, o);
  if (lane == 0) zhat[i] = copysign(sqrt(fabs(p)), zz[i]);
}

// column j of the eigenvector matrix of D + rho z z^T, normalised, scattered into M
__global__ void __launch_bounds__(128) k_build_M(int k, int nm, const double* __restrict__ dl, const int* __restrict__ org,
                                                 const double* __restrict__ tau, const double* __restrict__ zhat,
                                                 const int* __restrict__ idx_nd, const int* __restrict__ pos_nd,
                                                 double* __restrict__ M, int c_lo, int c_hi) {
  __shared__ double red[4];
  const int j = blockIdx.x, tid = threadIdx.x;
  if (pos_nd[j] < c_lo || pos_nd[j] >= c_hi) return;  // multi-GPU: this rank builds the columns [c_lo, c_hi) of M only
  const int oj = org[j];
  const double tj = tau[j], dorg = dl[oj];
  double acc = 0;
  for (int i = tid; i < k; i += 128) {
    const double u = zhat[i] / ((dl[i] - dorg) - tj);
    acc += u * u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) red[tid >> 5] = acc;
  __syncthreads();
  const double inv = 1.0 / sqrt((red[0] + red[1]) + (red[2] + red[3]));
  double* col = M + (size_t)pos_nd[j] * nm;
  for (int i = tid; i < k; i += 128) col[idx_nd[i]] = zhat[i] / ((dl[i] - dorg) - tj) * inv;
}

__global__ void k_set_ones(int nd, int nm, const int* __restrict__ idx_def, const int* __restrict__ pos_def,
                           double* __restrict__ M, int c_lo, int c_hi) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < nd && pos_def[t] >= c_lo && pos_def[t] < c_hi) M[idx_def[t] + (size_t)pos_def[t] * nm] = 1.0;
}

// M <- R_1 (R_2 (... (R_m M))): the column rotations of the deflation step as row operations on M
__global__ void k_rot_rows(int nm, const Rotation* __restrict__ rots, int nrot, double* __restrict__ M, int c_lo, int c_hi) {
  const int c = c_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c_hi) return;
  double* col = M + (size_t)c * nm;
  for (int r = nrot - 1; r >= 0; --r) {
    const Rotation R = rots[r];
    const double xa = col[R.a], xb = col[R.b];
    if (xa == 0.0 && xb == 0.0) continue;
    col[R.a] = R.c * xa - R.s * xb;
    col[R.b] = R.s * xa + R.c * xb;
  }
}

// Qn[:, pos[t]] = Qc[:, idx[t]]  (a merge in which nothing couples)
__global__ void k_permute_cols(const double* __restrict__ Qc, double* __restrict__ Qn, long long ldq, int nm,
                               const int* __restrict__ idx, const int* __restrict__ pos) {
  const int t = blockIdx.y;
  const double* src = Qc + (long long)idx[t] * ldq;
  double* dst = Qn + (long long)pos[t] * ldq;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nm; r += gridDim.x * blockDim.x) dst[r] = src[r];
}

struct Block {
  int off, size;
};

template <class T>
static void append(std::vector<char>& buf, const std::vector<T>& v, size_t& offset_out) {
  const size_t al = 16;
  size_t off = (buf.size() + al - 1) / al * al;
  offset_out = off;
  buf.resize(off + v.size() * sizeof(T));
  if (!v.empty()) memcpy(buf.data() + off, v.data(), v.size() * sizeof(T));
}

}  // namespace

// Eigen-decomposition of the symmetric tridiagonal (d_d: n diagonal, d_e: n-1 off-diagonal, device).
// d_lam (n, device): ascending eigenvalues; d_Z (n x n, ld n): eigenvectors; d_S1, d_S2: n x n scratch.
// collective (multi-GPU, every rank calls with the same d, e): the merges of stedc_split_min() rows and more are
// split by OUTPUT columns -- rank q builds the columns q of M and forms Q_out[:, q] = diag(Q1, Q2) M[:, q]; the
// column panels are all-gathered (Q stays replicated, so the next level's coupling rows are local).  The
// levels below run replicated: they hold ~1/8 of the flops.
static int64_t stedc_split_min() {
  static int64_t v = -1;
  if (v < 0) {
    const char* e = getenv("RSOPTSC_STEDC_SPLIT_MIN");
    v = e ? std::max<int64_t>(atoll(e), 2) : 3072;
  }
  return v;
}

void stedc_device(const double* d_d, const double* d_e, int64_t n64, double* d_lam, double* d_Z, double* d_S1, double* d_S2,
                  bool collective) {
  Context& c = ctx();
  collective = collective && sharded();
  const int n = (int)n64;
  RS_REQUIRE(n64 >= 1 && n64 <= 65535, "stedc: size out of range");
  std::vector<double> d(n), e(std::max(n - 1, 1), 0.0);
  RS_CUDA(cudaMemcpyAsync(d.data(), d_d, sizeof(double) * n, cudaMemcpyDeviceToHost, c.stream));
  if (n > 1) RS_CUDA(cudaMemcpyAsync(e.data(), d_e, sizeof(double) * (n - 1), cudaMemcpyDeviceToHost, c.stream));
  RS_CUDA(cudaStreamSynchronize(c.stream));

  for (int i = 0; i < n; ++i)
    RS_REQUIRE(std::isfinite(d[i]) && (i + 1 >= n || std::isfinite(e[i])), "symmetric eigensolver: non-finite entries in the matrix");
  std::vector<int> loff;
  stedc::make_leaves(n, LEAF, loff);
  const int L = (int)loff.size() - 1;
  for (int b = 1; b < L; ++b) {  // rank-one tears at every leaf boundary
    const int i = loff[b];
    d[i - 1] -= std::fabs(e[i - 1]);
    d[i] -= std::fabs(e[i - 1]);
  }
  int levels = 0;
  for (int b = L; b > 1; b = (b + 1) / 2) ++levels;
  // ping-pong so that the last level writes d_Z
  double* Qc = (levels % 2 == 0) ? d_Z : d_S1;
  double* Qn = (levels % 2 == 0) ? d_S1 : d_Z;
  const long long ldq = n;
  RS_CUDA(cudaMemsetAsync(d_Z, 0, sizeof(double) * (size_t)n * n, c.stream));
  RS_CUDA(cudaMemsetAsync(d_S1, 0, sizeof(double) * (size_t)n * n, c.stream));

  DevBuf<double> dv_d(n), dv_z(n), dv_zsgn(n), dv_lam(n), dv_tau(n), dv_zhat(n), dv_lamnew(n);
  DevBuf<int> dv_loff(L + 1), dv_zrow(n), dv_org(n);
  dv_d.upload(d.data(), n, c.stream);
  dv_loff.upload(loff.data(), L + 1, c.stream);
  std::vector<double> lam(n);
  {
    Phase ph("stedc_leaves");
    RS_LAUNCH(k_leaf_eig, (unsigned)L, LEAF, 0, dv_d.p, d_e, dv_loff.p, Qc, ldq, dv_lam.p);
    dv_lam.download(lam.data(), n, c.stream);
    RS_CUDA(cudaStreamSynchronize(c.stream));
  }

  std::vector<Block> blocks;
  for (int b = 0; b < L; ++b) blocks.push_back({loff[b], loff[b + 1] - loff[b]});
  DevBuf<char> dv_pack;
  std::vector<double> z(n), zsgn(n), lamnew(n);
  std::vector<int> zrow(n);

  while (blocks.size() > 1) {
    const size_t nmerge = blocks.size() / 2;
    // ---- coupling vectors z = [last row of Q1 ; sign(beta) first row of Q2] --------------------
    Phase* ph_plan = new Phase("stedc_plan_secular");
    std::fill(zsgn.begin(), zsgn.end(), 0.0);
    std::fill(zrow.begin(), zrow.end(), 0);
    for (size_t m = 0; m < nmerge; ++m) {
      const Block &B1 = blocks[2 * m], &B2 = blocks[2 * m + 1];
      const double beta = e[B1.off + B1.size - 1];
      for (int cidx = 0; cidx < B1.size; ++cidx) { zrow[B1.off + cidx] = B1.off + B1.size - 1; zsgn[B1.off + cidx] = 1.0; }
      for (int cidx = 0; cidx < B2.size; ++cidx) { zrow[B2.off + cidx] = B2.off; zsgn[B2.off + cidx] = beta < 0 ? -1.0 : 1.0; }
    }
    dv_zrow.upload(zrow.data(), n, c.stream);
    dv_zsgn.upload(zsgn.data(), n, c.stream);
    RS_LAUNCH(k_gather_z, (unsigned)cdiv(n, 256), 256, 0, Qc, ldq, n, dv_zrow.p, dv_zsgn.p, dv_z.p);
    dv_z.download(z.data(), n, c.stream);
    RS_CUDA(cudaStreamSynchronize(c.stream));

    // ---- plans; stage 1 upload: poles and weights of every merge ----------------------------------
    std::vector<MergePlan> plans(nmerge);
    std::vector<char> pack;
    struct Off { size_t dl, zz, idx_nd, pos_nd, idx_def, pos_def, rots; };
    std::vector<Off> offs(nmerge);
    for (size_t m = 0; m < nmerge; ++m) {
      const Block &B1 = blocks[2 * m], &B2 = blocks[2 * m + 1];
      const int off = B1.off, nm = B1.size + B2.size;
      stedc::plan_merge(&lam[off], &z[off], e[off + B1.size - 1], nm, plans[m]);
      append(pack, plans[m].dl, offs[m].dl);
      append(pack, plans[m].zz, offs[m].zz);
      append(pack, plans[m].idx_nd, offs[m].idx_nd);
      append(pack, plans[m].idx_def, offs[m].idx_def);
      append(pack, plans[m].rots, offs[m].rots);
      add_counter("stedc_deflated", (double)plans[m].idx_def.size());
      add_counter("stedc_poles", (double)plans[m].k);
    }
    dv_pack.alloc(pack.size() + 16);
    if (!pack.empty()) RS_CUDA(cudaMemcpyAsync(dv_pack.p, pack.data(), pack.size(), cudaMemcpyHostToDevice, c.stream));
    for (size_t m = 0; m < nmerge; ++m) {
      const MergePlan& P = plans[m];
      if (P.k == 0) continue;
      const int off = blocks[2 * m].off;
      RS_LAUNCH(k_secular, (unsigned)cdiv(P.k, 8), 256, 0, P.k, (const double*)(dv_pack.p + offs[m].dl),
                (const double*)(dv_pack.p + offs[m].zz), P.rho, dv_org.p + off, dv_tau.p + off, dv_lamnew.p + off);
    }
    dv_lamnew.download(lamnew.data(), n, c.stream);
    RS_CUDA(cudaStreamSynchronize(c.stream));

    delete ph_plan;
    // ---- output order; stage 2 upload: positions -----------------------------------------------
    std::vector<char> pack2;
    std::vector<std::vector<int>> pos_nd(nmerge), pos_def(nmerge);
    for (size_t m = 0; m < nmerge; ++m) {
      const int off = blocks[2 * m].off;
      std::vector<double> lamj(lamnew.begin() + off, lamnew.begin() + off + plans[m].k), dnew;
      stedc::order_outputs(plans[m], lamj, pos_nd[m], pos_def[m], dnew);
      for (size_t t = 0; t < dnew.size(); ++t) lam[off + t] = dnew[t];
      append(pack2, pos_nd[m], offs[m].pos_nd);
      append(pack2, pos_def[m], offs[m].pos_def);
    }
    DevBuf<char> dv_pack2(pack2.size() + 16);
    if (!pack2.empty()) RS_CUDA(cudaMemcpyAsync(dv_pack2.p, pack2.data(), pack2.size(), cudaMemcpyHostToDevice, c.stream));

    // ---- per merge: M, then Q_next[block] = Q_cur[block] M -----------------------------------------
    Phase ph_apply("stedc_apply");
    for (size_t m = 0; m < nmerge; ++m) {
      const MergePlan& P = plans[m];
      const int off = blocks[2 * m].off, nm = P.nm, k = P.k, nd = (int)P.idx_def.size(), nrot = (int)P.rots.size();
      const double* Qblk = Qc + off + (long long)off * ldq;
      double* Qout = Qn + off + (long long)off * ldq;
      const int* p_idx_def = (const int*)(dv_pack.p + offs[m].idx_def);
      const int* p_pos_def = (const int*)(dv_pack2.p + offs[m].pos_def);
      if (k == 0 && nrot == 0) {
        dim3 grid((unsigned)std::min(cdiv(nm, 256), 64), (unsigned)nm);
        RS_LAUNCH(k_permute_cols, grid, 256, 0, Qblk, Qout, ldq, nm, p_idx_def, p_pos_def);
        continue;
      }
      double* M = d_S2;
      // output columns this rank forms: all of them, or its panel of a split merge
      const bool split = collective && nm >= stedc_split_min();
      std::vector<int64_t> mb;
      int c_lo = 0, c_hi = nm;
      if (split) {
        mb = split_even(nm, comm().nranks, nm >= 4096 ? 128 : 1);
        c_lo = (int)mb[comm().rank];
        c_hi = (int)mb[comm().rank + 1];
      }
      const int wc = c_hi - c_lo;
      if (wc > 0) {
        RS_CUDA(cudaMemsetAsync(M + (size_t)c_lo * nm, 0, sizeof(double) * (size_t)nm * wc, c.stream));
        if (k > 0) {
          const double* p_dl = (const double*)(dv_pack.p + offs[m].dl);
          const double* p_zz = (const double*)(dv_pack.p + offs[m].zz);
          RS_LAUNCH(k_zhat, (unsigned)cdiv(k, 8), 256, 0, k, p_dl, p_zz, dv_org.p + off, dv_tau.p + off, dv_zhat.p + off);
          RS_LAUNCH(k_build_M, (unsigned)k, 128, 0, k, nm, p_dl, dv_org.p + off, dv_tau.p + off, dv_zhat.p + off,
                    (const int*)(dv_pack.p + offs[m].idx_nd), (const int*)(dv_pack2.p + offs[m].pos_nd), M, c_lo, c_hi);
        }
        if (nd > 0) RS_LAUNCH(k_set_ones, (unsigned)cdiv(nd, 256), 256, 0, nd, nm, p_idx_def, p_pos_def, M, c_lo, c_hi);
        if (nrot > 0)
          RS_LAUNCH(k_rot_rows, (unsigned)cdiv(wc, 128), 128, 0, nm, (const Rotation*)(dv_pack.p + offs[m].rots), nrot, M,
                    c_lo, c_hi);
        // Q_in is block diagonal (the two children): two half-height GEMMs instead of one full one
        const int n1 = blocks[2 * m].size, n2 = nm - n1;
        const double* Mc = M + (size_t)c_lo * nm;
        double* Qo = Qout + (long long)c_lo * ldq;
        gemm_nn_ld(Qblk, ldq, Mc, nm, Qo, ldq, n1, wc, n1);
        gemm_nn_ld(Qblk + n1 + (long long)n1 * ldq, ldq, Mc + n1, nm, Qo + n1, ldq, n2, wc, n2);
      }
      // full-height columns travel: outside the diagonal block they are zero on every rank
      if (split) comm_allgather_cols(Qn + (long long)off * ldq, ldq, mb);
    }
    std::vector<Block> next;
    for (size_t m = 0; m < nmerge; ++m) next.push_back({blocks[2 * m].off, blocks[2 * m].size + blocks[2 * m + 1].size});
    if (blocks.size() % 2 == 1) {
      const Block& B = blocks.back();
      RS_CUDA(cudaMemcpy2DAsync(Qn + B.off + (long long)B.off * ldq, sizeof(double) * ldq, Qc + B.off + (long long)B.off * ldq,
                                sizeof(double) * ldq, sizeof(double) * B.size, B.size, cudaMemcpyDeviceToDevice, c.stream));
      next.push_back(B);
    }
    // the packs are read by kernels still in flight: order their release after this level
    RS_CUDA(cudaStreamSynchronize(c.stream));
    std::swap(Qc, Qn);
    blocks.swap(next);
  }

  if (L == 1) {
    // a single leaf: sort on the host, permute the columns
    std::vector<int> ord(n), pos(n);
    for (int i = 0; i < n; ++i) ord[i] = i;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return lam[a] < lam[b]; });
    std::vector<double> sorted(n);
    for (int p = 0; p < n; ++p) { sorted[p] = lam[ord[p]]; pos[ord[p]] = p; }
    std::vector<int> idx(n);
    for (int i = 0; i < n; ++i) idx[i] = i;
    DevBuf<int> dv_idx(n), dv_pos(n);
    dv_idx.upload(idx.data(), n, c.stream);
    dv_pos.upload(pos.data(), n, c.stream);
    RS_CUDA(cudaMemcpyAsync(d_S1, Qc, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToDevice, c.stream));
    dim3 grid(1, (unsigned)n);
    RS_LAUNCH(k_permute_cols, grid, 64, 0, d_S1, d_Z, ldq, n, dv_idx.p, dv_pos.p);
    lam = sorted;
    RS_CUDA(cudaStreamSynchronize(c.stream));
  }
  RS_CUDA(cudaMemcpyAsync(d_lam, lam.data(), sizeof(double) * n, cudaMemcpyHostToDevice, c.stream));
  RS_CUDA(cudaStreamSynchronize(c.stream));
}

}  // namespace rs
