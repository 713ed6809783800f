// computeM: linearised ADMM for  min ||Z||_* + lambda ||E||_{2,1}  s.t. X = XZ + E, Z^T 1 = 1,
// Z_Omega = 0   (R/SimilarityM.R:115-197), restructured for the GPU:
//
//   * Z lives on its kNN support only (n*K values, CSR + CSC views); the dense state is Y3 and
//     the SVT operand A/J.
//   * Y3, A and J are block diagonal over the connected components of the support graph (the SVD
//     of a block-diagonal matrix is block diagonal, Y3 starts at 0): they are stored as the
//     concatenated n_c x n_c blocks and the SVT runs per block (sum n_c^3 instead of n^3).
//   * the stop test of iteration `it` (R :157-162) only looks at Err[it-1], Err[it-2] and XXZ,
//     none of which the SVT touches, so it is evaluated BEFORE the SVT (saves one SVT).
//   * J_c == 0 exactly while sigma_1(A_c) <= 1/mu (all singular values thresholded): SVT skipped;
//     decided by Frobenius bounds, else by a residual-certified Lanczos value of sigma_1.
//   * spectral norms: Lanczos (lanczos.cu); the two tests that only compare a norm with
//     epsilon = 1e-5 use the Frobenius bounds ||.||_F/sqrt(min dim) <= ||.||_2 <= ||.||_F first.
#include <algorithm>
#include <cmath>

#include "comm.h"
#include "internal.h"

namespace rs {

int64_t shard_split_min() {
  const char* e = getenv("RSOPTSC_SPLIT_MIN_N");  // tests lower it to split small components
  const int64_t v = e ? atoll(e) : 2048;
  return v < 2 ? 2 : v;
}

void build_shard_plan(const Support& s, int P, int rank, int64_t split_min, ShardPlan& plan, bool device) {
  plan = ShardPlan();
  plan.on = P > 1;
  plan.P = P;
  plan.rank = rank;
  const size_t nc = s.comp_n.size();
  plan.big.assign(nc, 0);
  plan.owner.assign(nc, -1);
  plan.cb.assign(nc, {});
  std::vector<size_t> small;
  for (size_t c = 0; c < nc; ++c) {
    if (P > 1 && s.comp_n[c] >= split_min) {
      plan.big[c] = 1;
      plan.cb[c] = split_even(s.comp_n[c], P);
    } else {
      small.push_back(c);
    }
  }
  // small components: largest first onto the least loaded rank (cost ~ n_c^3), ties to the lower rank / index
  std::stable_sort(small.begin(), small.end(), [&](size_t a, size_t b) { return s.comp_n[a] > s.comp_n[b]; });
  std::vector<double> load(P, 0.0);
  for (size_t c : small) {
    int best = 0;
    for (int q = 1; q < P; ++q)
      if (load[q] < load[best]) best = q;
    plan.owner[c] = best;
    const double w = (double)s.comp_n[c];
    load[best] += w * w * w + 1.0;
  }
  for (int64_t j = 0; j < s.n; ++j) {
    const int c = s.h_comp[j];
    const bool mine = plan.big[c] ? (s.h_local[j] >= plan.cb[c][rank] && s.h_local[j] < plan.cb[c][rank + 1])
                                  : plan.owner[c] == rank;
    if (mine) plan.own_cols.push_back((int)j);
  }
  std::vector<char> own(s.n, 0);
  for (int j : plan.own_cols) own[j] = 1;
  for (int64_t p = 0; p < s.nnz; ++p)
    if (own[s.h_col[p]]) plan.own_entries.push_back((int)p);
  for (size_t c = 0; c < nc; ++c) {
    const int64_t ncc = s.comp_n[c];
    if (plan.big[c]) {
      const int64_t lo = plan.cb[c][rank], hi = plan.cb[c][rank + 1];
      if (hi > lo) plan.segs.push_back({s.comp_off[c] + lo * ncc, (hi - lo) * ncc});
    } else if (plan.owner[c] == rank) {
      plan.segs.push_back({s.comp_off[c], ncc * ncc});
    }
  }
  if (!device) return;
  cudaStream_t st = ctx().stream;
  plan.d_own_cols.alloc(std::max<size_t>(plan.own_cols.size(), 1));
  plan.d_own_entries.alloc(std::max<size_t>(plan.own_entries.size(), 1));
  if (!plan.own_cols.empty()) plan.d_own_cols.upload(plan.own_cols.data(), plan.own_cols.size(), st);
  if (!plan.own_entries.empty()) plan.d_own_entries.upload(plan.own_entries.data(), plan.own_entries.size(), st);
  RS_CUDA(cudaStreamSynchronize(st));
}

static void upload_support(Support& s, const std::vector<int>& rowptr, const std::vector<int>& col, bool device = true) {
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
  // connected components of the (symmetrised) support graph: union-find over the entries
  std::vector<int> parent(n);
  for (int64_t i = 0; i < n; ++i) parent[i] = (int)i;
  auto find = [&](int x) {
    while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
    return x;
  };
  for (int64_t p = 0; p < s.nnz; ++p) {
    const int a = find(rowof[p]), b = find(col[p]);
    if (a != b) parent[std::max(a, b)] = std::min(a, b);
  }
  std::vector<int> comp_id(n, -1), local(n, 0);
  s.comp_n.clear();
  for (int64_t i = 0; i < n; ++i) {       // components numbered by their smallest cell, cells keep their order
    const int r = find((int)i);
    if (comp_id[r] < 0) { comp_id[r] = (int)s.comp_n.size(); s.comp_n.push_back(0); }
    comp_id[i] = comp_id[r];
    local[i] = (int)s.comp_n[comp_id[i]]++;
  }
  s.comp_off.assign(s.comp_n.size(), 0);
  s.total_dense = 0;
  for (size_t c = 0; c < s.comp_n.size(); ++c) { s.comp_off[c] = s.total_dense; s.total_dense += s.comp_n[c] * s.comp_n[c]; }
  s.h_comp = comp_id;
  s.h_local = local;
  std::vector<long long> dense_off(s.nnz);
  for (int64_t p = 0; p < s.nnz; ++p) {
    const int c = comp_id[rowof[p]];
    dense_off[p] = s.comp_off[c] + local[rowof[p]] + (long long)local[col[p]] * s.comp_n[c];
  }
  if (!device) return;  // host-only planning (rsoptsc_host_shard_owner)
  cudaStream_t st = ctx().stream;
  s.dense_off.alloc(std::max<int64_t>(s.nnz, 1));
  if (s.nnz) s.dense_off.upload(dense_off.data(), s.nnz, st);
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

void build_support_from_idx(Support& s, const int32_t* h_idx, int64_t n, int K, bool device) {
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
  upload_support(s, rowptr, col, device);
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
  // multi-GPU: every rank holds X and the support; the column-local state is split (ShardPlan)
  ShardPlan plan;
  const bool sh = sharded();
  if (sh) build_shard_plan(s, comm().nranks, comm().rank, shard_split_min(), plan);
  const ShardPlan* pp = sh ? &plan : nullptr;
  st.m = m; st.n = n; st.X = d_X;
  st.nloc = sh ? plan.nloc() : n;
  st.cols = sh ? plan.d_own_cols.p : nullptr;
  const size_t mn = (size_t)m * std::max<int64_t>(st.nloc, 1);
  st.Zs.alloc(std::max<size_t>(s.nnz, 1)); st.Zs.zero(c.stream);
  st.q.alloc(std::max<size_t>(s.nnz, 1));
  st.XXZ.alloc(mn); st.E.alloc(mn); st.Y1.alloc(mn); st.R.alloc(mn); st.M.alloc(mn);
  st.E.zero(c.stream); st.Y1.zero(c.stream);
  st.Y2.alloc(n); st.Y2.zero(c.stream);
  st.colacc.alloc(std::max<int64_t>(st.nloc, 1) + 1);
  const size_t nd = (size_t)std::max<int64_t>(s.total_dense, 1);  // block-diagonal dense state
  DevBuf<double> Y3(nd), A(nd);
  Y3.zero(c.stream);
  SvtWorkspace svt;
  for (int i = 0; i < maxiter; ++i) res.err1[i] = res.err2[i] = 0.0;
  auto norm_cols = [&](const double* M) {  // spectral norm of an m x n matrix held as this rank's columns
    return sh ? spectral_norm_colsharded(M, m, st.nloc) : spectral_norm(M, m, n);
  };
  auto mine = [&](size_t cidx) { return !sh || plan.big[cidx] || plan.owner[cidx] == plan.rank; };

  // XXZ = X - X Z with Z = 0 (R :130)
  double xxz_f2 = comm_sum(admm_residual_update(st, s, mu, /*update_duals=*/false));
  // eta = ||X||_2^2 + ||1_n||_2^2 + 1 (R :178) is iteration-invariant (quirk Q8): hoisted
  const double sx = spectral_norm(d_X, m, n);
  RS_REQUIRE(std::isfinite(sx) && std::isfinite(xxz_f2),
             "computeM: non-finite values in X (NaN/Inf input or an all-zero cell column, R/SimilarityM.R:34-39)");
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
        else if (f / std::sqrt(min_dim) <= epsilon) stop = norm_cols(st.XXZ.p) <= epsilon;
      }
      if (stop) { res.stop = 2; break; }
    }
    // step 1: J = SVT_{1/mu}(Z - Y3/mu)                      R :142-155
    // A and J are block diagonal over the connected components of the support graph: one SVT per
    // block (sum n_c^3 instead of n^3), written over A in the block storage.
    std::vector<double> a_f2;
    build_A(s, st.Zs.p, Y3.p, mu, A.p, a_f2, pp);
    bool any_J = false;
    std::vector<double*> small_blocks;   // small blocks with J != 0: batched below (svt_small_batch)
    std::vector<int64_t> small_sizes;
    const int64_t small_lim = svt_small_limit();
    for (size_t cidx = 0; cidx < s.comp_n.size(); ++cidx) {
      if (!mine(cidx)) continue;
      const int64_t nc = s.comp_n[cidx];
      double* Ac = A.p + s.comp_off[cidx];
      // J_c == 0 exactly when sigma_1(A_c) <= tau.  ||A_c||_F/sqrt(n_c) <= sigma_1 <= ||A_c||_F settles
      // most iterations; in between a residual-certified Lanczos value of sigma_1 (1e-10 rel.) decides.
      // (big components: A_c is replicated, so every rank takes the same decision)
      bool j_zero = std::sqrt(a_f2[cidx]) <= tau;
      if (!j_zero && std::sqrt(a_f2[cidx] / (double)nc) <= tau) j_zero = spectral_norm(Ac, nc, nc) <= tau * (1.0 - 1e-8);
      if (j_zero) {
        RS_CUDA(cudaMemsetAsync(Ac, 0, sizeof(double) * nc * nc, c.stream));
      } else {
        if (sh && plan.big[cidx]) svt_sharded(svt, Ac, nc, tau, plan.cb[cidx]);
        else if (nc >= 2 && nc < small_lim) { small_blocks.push_back(Ac); small_sizes.push_back(nc); }
        else svt_inplace(svt, Ac, nc, tau);
        ++res.n_svt;
        any_J = true;
      }
    }
    if (small_blocks.size() >= 2) svt_small_batch(small_blocks, small_sizes, tau);
    else if (small_blocks.size() == 1) svt_inplace(svt, small_blocks[0], small_sizes[0], tau);
    const double* J = any_J ? A.p : nullptr;
    gather_support_term(s, st.Zs.p, J, Y3.p, mu, st.q.p, pp);  // (Z - J + Y3/mu) on the support
    admm_E_update(st, lambda, mu);                           // R :164-176
    admm_Z_update(st, s, mu, eta);                           // R :178-181 (+ Y2, :186)
    xxz_f2 = comm_sum(admm_residual_update(st, s, mu, true));  // R :182, :185
    const double zj_f2 = std::max(comm_sum(update_Y3(s, st.Zs.p, J, mu, Y3.p, pp)), 0.0);  // R :187
    res.err1[it - 1] = norm_cols(st.M.p);                    // R :189
    RS_REQUIRE(std::isfinite(res.err1[it - 1]) && std::isfinite(zj_f2) && std::isfinite(xxz_f2),
               "computeM: the iteration produced non-finite values (NaN/Inf in the input?)");
    // Err[,2] = ||Z - J||_2 is only ever compared with epsilon (R :191): Frobenius bounds first
    const double zj_f = std::sqrt(zj_f2);
    double err2 = zj_f / std::sqrt((double)n);               // lower bound
    if (err2 <= epsilon && res.err1[it - 1] <= epsilon) {
      if (zj_f <= epsilon) err2 = zj_f;
      else {  // inconclusive: form Z - J block by block; the norm of a block-diagonal matrix is the max
        DevBuf<double> Dm(nd);
        scatter_support_blocks(s, st.Zs.p, Dm.p, pp);
        err2 = 0.0;
        for (size_t cidx = 0; cidx < s.comp_n.size(); ++cidx) {
          if (!mine(cidx)) continue;
          const int64_t nc = s.comp_n[cidx];
          double* Dc = Dm.p + s.comp_off[cidx];
          int64_t lo = 0, hi = nc;
          if (sh && plan.big[cidx]) { lo = plan.cb[cidx][plan.rank]; hi = plan.cb[cidx][plan.rank + 1]; }
          if (J && hi > lo) {
            const double minus1 = -1.0;
            RS_CUBLAS(cublasDaxpy_64(c.cublas, nc * (hi - lo), &minus1, J + s.comp_off[cidx] + lo * nc, 1, Dc + lo * nc, 1));
          }
          if (sh && plan.big[cidx]) comm_allgather_cols(Dc, nc, plan.cb[cidx]);
          err2 = std::max(err2, spectral_norm(Dc, nc, nc));
        }
        err2 = comm_max(err2);
      }
    }
    res.err2[it - 1] = err2;
    report_progress(/*RSOPTSC_STAGE_ADMM*/ 0, it, res.err1[it - 1]);  // print("Err = ...") at R :191; interrupt point
    if (std::max(res.err1[it - 1], err2) <= epsilon) { res.stop = 3; break; }  // R :191-194
  }
  res.iters = it;
  res.E = std::sqrt(xxz_f2);                                 // norm(XXZ, "f")
  if (sh) assemble_support_values(s, plan, st.Zs.p);
}

}  // namespace rs
