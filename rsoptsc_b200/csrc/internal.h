// Internal interfaces between the translation units of librsoptsc_b200.so.
#pragma once
#include <atomic>
#include <map>

#include "common.h"

namespace rs {

// ---- launch accounting + phase timers (capi.cu) ------------------------------------
extern std::atomic<int64_t> g_launches;
void add_counter(const char* name, double v);  // work counters read back by rsoptsc_counter()
struct Phase {  // RAII: device time of a region on the library stream
  int id;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  explicit Phase(const char* name);
  ~Phase();
};

// progress / interrupt hook of the C ABI (rsoptsc_set_progress): throws Error(7) when the caller asks to stop
void report_progress(int stage, int iter, double value);

#define RS_LAUNCH(kernel, grid, block, smem, ...)                                   \
  do {                                                                              \
    kernel<<<(grid), (block), (smem), ::rs::ctx().stream>>>(__VA_ARGS__);           \
    ::rs::g_launches.fetch_add(1, std::memory_order_relaxed);                       \
    RS_KERNEL_CHECK();                                                              \
  } while (0)

// ---- support of Z: CSR over rows (the reference's mask is per row jj, SimilarityM.R:90-92)
// plus the CSC view the column-local ADMM kernels walk. ------------------------------
struct Support {
  int64_t n = 0, nnz = 0;
  DevBuf<int> rowptr, col, rowof;      // CSR (row-major position p); rowof[p] = row of entry p
  DevBuf<int> colptr, rowidx, csrpos;  // CSC: for column j, entries q in [colptr[j], colptr[j+1])
                                       //      rowidx[q] = i, csrpos[q] = p
  std::vector<int> h_rowptr, h_col;
  int max_col_nnz = 0;
  // Connected components of the support graph.  Z, Y3, A = Z - Y3/mu and J = SVT(A) are block
  // diagonal over them (exactly: the SVD of a block-diagonal matrix is block diagonal), so the
  // dense n x n state is stored as the concatenation of the n_c x n_c diagonal blocks and the SVT
  // runs per block (cost sum n_c^3 instead of n^3).
  std::vector<int64_t> comp_n, comp_off;  // block sizes / offsets (in doubles) into the dense state
  int64_t total_dense = 0;                // sum n_c^2
  DevBuf<long long> dense_off;            // per support entry: offset of (i,j) inside the block storage
  std::vector<int> h_comp, h_local;       // per cell: its component and its index inside the component
};

// Multi-GPU split of one computeM problem.  Cells (columns of Z, E, Y1, XXZ, Y3,
// J -- column-local per R/SimilarityM.R:164-187) are owned by exactly one rank:
//   * components with at least split_min cells ("big") are cut into one contiguous range of their
//     local columns per rank; their SVT is collective (svt_sharded);
//   * smaller components belong to one rank each (largest first onto the least loaded rank, cost
//     n_c^3), together with all their column-local state; their SVT is local.
struct ShardPlan {
  bool on = false;
  int P = 1, rank = 0;
  std::vector<char> big;                  // per component
  std::vector<int> owner;                 // per component: owning rank, -1 for big components
  std::vector<std::vector<int64_t>> cb;   // big components: P + 1 local column boundaries
  std::vector<int> own_cols;              // cells owned by this rank, ascending
  std::vector<int> own_entries;           // CSR positions of the support entries in owned columns
  DevBuf<int> d_own_cols, d_own_entries;
  struct Seg { int64_t off, count; };
  std::vector<Seg> segs;                  // owned parts of the block-diagonal dense storage
  int64_t nloc() const { return (int64_t)own_cols.size(); }
  int64_t nent() const { return (int64_t)own_entries.size(); }
};
void build_shard_plan(const Support& s, int P, int rank, int64_t split_min, ShardPlan& plan, bool device = true);
int64_t shard_split_min();  // RSOPTSC_SPLIT_MIN_N, default 2048
void build_support_from_idx(Support& s, const int32_t* h_idx, int64_t n, int K, bool device = true);
void build_support_from_dense_mask(Support& s, const double* h_D, int64_t n);

// ---- dense elementwise (dense.cu) ---------------------------------------------------
void normalize_columns(const double* d_data, int64_t m, int64_t n, double* d_X);
// A = -Y3/mu (+ Z scattered on the support) in block-diagonal storage; fro2[c] = ||A_c||_F^2
void build_A(const Support& s, const double* d_Zs, const double* d_Y3, double mu, double* d_A,
             std::vector<double>& fro2, const ShardPlan* plan = nullptr);
// q[p] = Zs[p] - J[i,j] + Y3[i,j]/mu on the support (J may be null == 0)
void gather_support_term(const Support& s, const double* d_Zs, const double* d_J, const double* d_Y3,
                         double mu, double* d_q, const ShardPlan* plan = nullptr);
// Y3 += mu*(Z - J) ; returns ||Z - J||_F^2   (J may be null == 0)
double update_Y3(const Support& s, const double* d_Zs, const double* d_J, double mu, double* d_Y3,
                 const ShardPlan* plan = nullptr);
// sharded runs: Zs of the owned columns is current on every rank, the rest is stale -> make it whole again
void assemble_support_values(const Support& s, const ShardPlan& plan, double* d_Zs);
void scale_columns(double* d_T, int64_t rows, int64_t cols, const double* d_scale);
void copy_scale_columns(const double* d_src, int64_t rows, int64_t cols, const double* d_scale, double* d_dst);
void mirror_lower(double* d_B, int64_t d);  // upper triangle <- lower triangle
void threshold_symmetrise_dense(const double* d_Z, int64_t n, double* d_W);
void threshold_symmetrise_sparse(const Support& s, const double* d_Zs, double* d_W /* n x n, zeroed here */);
void scatter_support_dense(const Support& s, const double* d_Zs, double* d_Z /* n x n, zeroed here */);
void scatter_support_blocks(const Support& s, const double* d_Zs, double* d_Zb /* block storage, zeroed here */,
                            const ShardPlan* plan = nullptr);
double frob2(const double* d_a, int64_t count);

// ---- column-local ADMM step (admm_kernels.cu) ---------------------------------------
struct AdmmState {
  int64_t m = 0, n = 0;
  int64_t nloc = 0;           // columns held by this rank (== n unless sharded)
  const int* cols = nullptr;  // their global indices (device; null == identity)
  const double* X = nullptr;  // m x n normalised
  DevBuf<double> Zs;          // nnz, CSR order
  DevBuf<double> XXZ, E, Y1, R, M;  // m x n
  DevBuf<double> Y2;          // n
  DevBuf<double> q;           // nnz support term
  DevBuf<double> colacc;      // n  per-column ||XXZ||^2
};
void admm_E_update(AdmmState& st, double lambda, double mu);
void admm_Z_update(AdmmState& st, const Support& s, double mu, double eta);
// XXZ = X - X Z ; Y1 += mu (XXZ - E) ; M = XXZ - E ; returns ||XXZ||_F^2
double admm_residual_update(AdmmState& st, const Support& s, double mu, bool update_duals);

// ---- spectral norm (lanczos.cu) -----------------------------------------------------
// Largest singular value of the column-major rows x cols matrix (Lanczos on the smaller Gram
// operator, full reorthogonalisation).  rel_tol on the Ritz value.
double spectral_norm(const double* d_M, int64_t rows, int64_t cols, double rel_tol = 2e-10);
// same for a matrix whose COLUMNS are split across the ranks (d_M: rows x cols_local on this rank):
// Gram of the row side summed over ranks (NCCL allreduce), identical result on every rank
double spectral_norm_colsharded(const double* d_M, int64_t rows, int64_t cols_local, double rel_tol = 2e-10);
// Host: eigen-decomposition of a symmetric tridiagonal (d: k diag, e: k-1 off-diag).
// On return d holds ascending eigenvalues, Zv (k x k column-major, may be null) eigenvectors.
void tridiag_eig(std::vector<double>& d, std::vector<double>& e, std::vector<double>* Zv, bool last_row_only = false);

// ---- singular value thresholding (svt.cu) -------------------------------------------
// eig.cu: dense symmetric eigensolver (lower triangle of A; eigenvectors overwrite A when requested,
// eigenvalues ascending in d_lam)
// keep_above: only the eigenvectors of eigenvalues >= keep_above are needed (the trailing columns, eigenvalues being
// ascending); the other columns of A may be left undefined.  Default: all of them.
// collective: every rank of the communicator calls with the same (replicated) matrix; the stages are
// split across the ranks and the result is replicated again.
void sym_eig(double* d_A, int64_t n, double* d_lam, bool vectors, double keep_above = -1.0e300, bool collective = false);
// same with eigenvectors, solver forced: 0 library (cuSOLVER), 1 native (sytrd_kernel.cuh + stedc.cu)
void sym_eig_select(double* d_A, int64_t n, double* d_lam, int solver);
// backtransform.cu: Z (n x n) <- Q Z, Q = product of the Householder reflectors stored by the tridiagonalisation
void apply_q_left(const double* d_A, int64_t lda, const double* d_tau, int64_t n, double* d_Z, int64_t ncols);  // Z: n x ncols, ld n

struct SvtWorkspace {
  int64_t n = 0;
  DevBuf<double> G, T, B, lam;
  void ensure(int64_t n);
};
// J = SVT_tau(A) written over A.  Returns rank(J).
int64_t svt_inplace(SvtWorkspace& ws, double* d_A, int64_t n, double tau);
// Several small blocks (every size below svt_small_limit()) with the same tau: Gram + eigensolve of the blocks side by
// side on two extra streams, reconstructions on the library stream.  J over every block.
void svt_small_batch(const std::vector<double*>& blocks, const std::vector<int64_t>& sizes, double tau);
// largest block size the batch route takes (0: route off -- native eigensolver forced, RSOPTSC_SVT_BATCH=0, ...)
int64_t svt_small_limit();
// Collective variant for a block replicated on every rank (comm.h): on return the columns
// [cb[rank], cb[rank+1]) of d_A hold J, the other columns are unspecified.
int64_t svt_sharded(SvtWorkspace& ws, double* d_A, int64_t n, double tau, const std::vector<int64_t>& cb);

// ---- drivers --------------------------------------------------------------------------
struct AdmmResult {
  double E = 0;
  int iters = 0, stop = 0, n_svt = 0;
  double err1[100], err2[100];
};
// Runs computeM on device-resident X with the given support; Zs (nnz) returned in st.Zs.
void run_admm(const double* d_X, int64_t m, int64_t n, const Support& s, double lambda,
              AdmmState& st, AdmmResult& res);

// knn.cu
void knn_bruteforce(const double* d_emb, int64_t n, int64_t ld, int dim, int K, int32_t* d_idx);
// pca.cu : eig (host, min(m,n) descending), optional device emb n x ncomp
void pca_spectrum(const double* d_X, int64_t m, int64_t n, std::vector<double>& eig, double* d_emb,
                  int ncomp);
int choose_K(const std::vector<double>& eig);
void pca_loadings(const double* d_X, int64_t m, int64_t n, int ncomp, double* d_load /* m x ncomp */);

// preprocess.cu (R/DataSelection.R)
void log_normalize(const double* d_M, int64_t m, int64_t n, double scale, bool normalize, double* d_out);
void scale_center(const double* d_M, int64_t m, int64_t n, double* d_out);
void gene_stats(const double* d_M, int64_t m, int64_t n, double countL, double countU, std::vector<double>& h);
void select_data(const double* d_M, int64_t m, int64_t n, double gene_expression_threshold, int n_features,
                 std::vector<int>& selected, int* max_component_out);

// nmf.cu
struct NmfResult { int iters = 0; };
void nmf_lee(const double* d_V, int64_t n, int k, int max_iter, double* h_basis, double* h_coef,
             int32_t* h_labels, NmfResult& res);

// rank sweep k = k_lo .. k_hi on one similarity matrix, all problems advanced together (batched launches);
// h_labels: (k_hi - k_lo + 1) x n row-major, 1-based; h_iters: iterations of every rank
void nmf_sweep(const double* d_V, int64_t n, int k_lo, int k_hi, int max_iter, int32_t* h_labels, int32_t* h_iters);

// graph.cu (RepresentationMap pieces)
void dist_flat(const double* d_emb, int64_t n, int dim, double* d_D);
void adjacency(const double* d_S, int64_t n, double* d_A);
void graph_components(const double* d_adj, int64_t n, int32_t* h_membership, int32_t* n_components);
void graph_distances(const double* d_adj, int64_t n, double* d_D);

// signaling.cu
void signaling_pair(const double* d_lig, const double* d_rec, const double* d_beta, const double* d_gamma, int64_t n,
                    bool normalize, double* d_P);

// eigengap.cu
void get_components(const double* d_S, int64_t n, double tol, double* h_vals, int32_t* n_eigs);
void consensus_dense(const int32_t* d_labels, int R, int64_t n, double tau, double* d_consen);
void get_ensemble(const double* d_S, int64_t n, int n_comp, double tau, int lo, int hi, double* d_consen);
void pam(const double* d_P, int64_t n, int dim, int k, int32_t* h_labels, int32_t* h_medoids);
void count_clusters(const double* d_S, int64_t n, double tol, int lo, int hi, int n_comp,
                    int32_t* upper, int32_t* lower, double* h_vals, int32_t* n_eigs);

}  // namespace rs
