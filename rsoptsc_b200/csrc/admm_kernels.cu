// Column-local kernels of one ADMM iteration of computeM (R/SimilarityM.R:164-187).
//
// Z is non-zero only on the kNN support (R/SimilarityM.R:181), so
//   * the gradient H is needed only on the support: an SDDMM  X[:,i] . R[:,j]      (a7, a8)
//   * X - X Z is an SpMM over the support                                           (a9)
// and, with X replicated, every one of these is independent per column j of Z.  One CTA owns
// one column j and walks its CSC entry list.  The kernels are HBM/L2-bound: per iteration they
// stream the m x n arrays XXZ, E, Y1, R, M once and gather nnz columns of X (L2-resident).
#include "internal.h"

namespace rs {

static constexpr int CT = 256;  // threads per column CTA

// ---- a6: E update, and R = XXZ - E + Y1/mu ------------------------------------------
// QQ = XXZ + Y1/mu ; E[:,j] = QQ[:,j] * (|q| - lambda/mu)/|q| if |q| > lambda/mu else 0
__global__ void k_E_update(const double* __restrict__ XXZ, const double* __restrict__ Y1, int64_t m, double lambda,
                           double mu, double* __restrict__ E, double* __restrict__ R) {
  __shared__ double sm[CT / 32 + 1];
  const int64_t base = (int64_t)blockIdx.x * m;
  double acc = 0;
  for (int64_t t = threadIdx.x; t < m; t += CT) {
    const double qq = XXZ[base + t] + Y1[base + t] / mu;
    acc += qq * qq;
  }
  const double nq = sqrt(block_sum<CT>(acc, sm));
  const double thr = lambda / mu;
  const double sc = nq <= thr ? 0.0 : (nq - thr) / nq;  // zero columns: NaN overwritten by the <= test (quirk Q10)
  for (int64_t t = threadIdx.x; t < m; t += CT) {
    const double x = XXZ[base + t], y = Y1[base + t] / mu;
    const double e = sc == 0.0 ? 0.0 : sc * (x + y);
    E[base + t] = e;
    R[base + t] = (x - e) + y;
  }
}

void admm_E_update(AdmmState& st, double lambda, double mu) {
  Phase ph("E_update");
  if (st.nloc > 0)
    RS_LAUNCH(k_E_update, (unsigned)st.nloc, CT, 0, st.XXZ.p, st.Y1.p, st.m, lambda, mu, st.E.p, st.R.p);
}

// ---- a7 + a8: gradient step on the support, Y2 update --------------------------------
// H[i,j] = -X[:,i].R[:,j] - r_j + q[p] ;  Z[i,j] -= H/eta ;  r_j = 1 - colsum_j(Z) + Y2_j/mu
// then Y2_j += mu (1 - colsum_j(Z_new))                                (R/SimilarityM.R:179-186)
__global__ void k_Z_update(const double* __restrict__ X, const double* __restrict__ R, int64_t m,
                           const int* __restrict__ colptr, const int* __restrict__ rowidx,
                           const int* __restrict__ csrpos, const double* __restrict__ q, double mu, double eta,
                           double* __restrict__ Zs, double* __restrict__ Y2, const int* __restrict__ cols) {
  __shared__ double sm[CT / 32 + 1];
  const int j = cols ? cols[blockIdx.x] : blockIdx.x;  // cell; the m x nloc arrays are indexed by blockIdx.x
  const int q0 = colptr[j], q1 = colptr[j + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double part = 0;
  for (int e = q0 + threadIdx.x; e < q1; e += CT) part += Zs[csrpos[e]];
  const double colsum_old = block_sum<CT>(part, sm);
  const double rj = 1.0 - colsum_old + Y2[j] / mu;
  const double* Rj = R + (int64_t)blockIdx.x * m;
  double newpart = 0;
  for (int e = q0 + warp; e < q1; e += CT / 32) {
    const double* Xi = X + (int64_t)rowidx[e] * m;
    double dot = 0;
    for (int64_t t = lane; t < m; t += 32) dot += Xi[t] * Rj[t];
    dot = warp_sum(dot);
    if (lane == 0) {
      const int p = csrpos[e];
      const double h = (-dot - rj) + q[p];
      const double z = Zs[p] - h / eta;
      Zs[p] = z;
      newpart += z;
    }
  }
  const double colsum_new = block_sum<CT>(newpart, sm);
  if (threadIdx.x == 0) Y2[j] += mu * (1.0 - colsum_new);
}

void admm_Z_update(AdmmState& st, const Support& s, double mu, double eta) {
  Phase ph("Z_update");
  if (st.nloc > 0)
    RS_LAUNCH(k_Z_update, (unsigned)st.nloc, CT, 0, st.X, st.R.p, st.m, s.colptr.p, s.rowidx.p, s.csrpos.p, st.q.p, mu,
              eta, st.Zs.p, st.Y2.p, st.cols);
}

// ---- a9 + a10: XXZ = X - X Z ; Y1 += mu (XXZ - E) ; M = XXZ - E -----------------------
__global__ void k_residual(const double* __restrict__ X, int64_t m, const int* __restrict__ colptr,
                           const int* __restrict__ rowidx, const int* __restrict__ csrpos,
                           const double* __restrict__ Zs, const double* __restrict__ E, double mu, int update_duals,
                           double* __restrict__ XXZ, double* __restrict__ Y1, double* __restrict__ M,
                           double* __restrict__ colacc, const int* __restrict__ cols) {
  __shared__ double sm[CT / 32 + 1];
  __shared__ int s_row[CT];
  __shared__ double s_z[CT];
  const int j = cols ? cols[blockIdx.x] : blockIdx.x;
  const int q0 = colptr[j], q1 = colptr[j + 1];
  const int64_t base = (int64_t)blockIdx.x * m, xbase = (int64_t)j * m;
  double nrm = 0;
  for (int64_t t0 = 0; t0 < m; t0 += CT) {
    const int64_t t = t0 + threadIdx.x;
    double acc = 0;
    for (int c0 = q0; c0 < q1; c0 += CT) {  // entry list staged through shared memory
      __syncthreads();
      if (c0 + threadIdx.x < q1) {
        s_row[threadIdx.x] = rowidx[c0 + threadIdx.x];
        s_z[threadIdx.x] = Zs[csrpos[c0 + threadIdx.x]];
      }
      __syncthreads();
      const int cnt = min(CT, q1 - c0);
      if (t < m)
        for (int c = 0; c < cnt; ++c) acc += s_z[c] * X[(int64_t)s_row[c] * m + t];
    }
    if (t < m) {
      const double r = X[xbase + t] - acc;
      XXZ[base + t] = r;
      const double d = r - E[base + t];
      if (update_duals) {
        Y1[base + t] += mu * d;
        M[base + t] = d;
      }
      nrm += r * r;
    }
  }
  nrm = block_sum<CT>(nrm, sm);
  if (threadIdx.x == 0) colacc[blockIdx.x] = nrm;
}

__global__ void k_sum_vec(const double* __restrict__ v, int64_t n, double* __restrict__ out) {
  __shared__ double sm[CT / 32 + 1];
  double acc = 0;
  for (int64_t i = threadIdx.x; i < n; i += CT) acc += v[i];
  acc = block_sum<CT>(acc, sm);
  if (threadIdx.x == 0) out[0] = acc;
}

double admm_residual_update(AdmmState& st, const Support& s, double mu, bool update_duals) {
  Phase ph("residual");
  if (st.nloc == 0) return 0.0;
  RS_LAUNCH(k_residual, (unsigned)st.nloc, CT, 0, st.X, st.m, s.colptr.p, s.rowidx.p, s.csrpos.p, st.Zs.p, st.E.p, mu,
            update_duals ? 1 : 0, st.XXZ.p, st.Y1.p, st.M.p, st.colacc.p, st.cols);
  RS_LAUNCH(k_sum_vec, 1, CT, 0, st.colacc.p, st.nloc, st.colacc.p + st.nloc);
  double h = 0;
  RS_CUDA(cudaMemcpyAsync(&h, st.colacc.p + st.nloc, sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
  RS_CUDA(cudaStreamSynchronize(ctx().stream));
  return h;
}

}  // namespace rs
