/* rsoptsc_b200 -- C ABI of the B200 (sm_100a) implementation of RSoptSC's cell-by-cell
 * numerical hot path.  This is the drop-in boundary: the entry points are what an R `.Call`
 * shim (r/rsoptsc_shim.c) binds in place of the bodies of the reference's R
 * functions.  The reference has no native interface of its own (NAMESPACE:1-86 has no
 * useDynLib), so every function cites the R function it replaces.
 *
 * Conventions
 *   - plain C, no C++/torch types; all matrices are column-major FP64 exactly as R holds them
 *   - sizes are int64_t, indices int32_t and 0-based at this boundary (the R shim adds 1)
 *   - every function returns 0 on success, non-zero on failure; rsoptsc_last_error() gives
 *     the message (thread-local).  Nothing is thrown across the boundary.
 *   - the caller owns all host buffers, the library owns all device memory
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails
 *   - `*_dev` variants take DEVICE pointers for the large inputs/outputs (used by bench.py to
 *     time the path with inputs resident in HBM); the plain variants take HOST pointers and
 *     perform the host<->device copies themselves.
 */
#ifndef RSOPTSC_B200_H
#define RSOPTSC_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSOPTSC_ABI_VERSION 3

/* stop codes of the ADMM (R/SimilarityM.R:134-137, :157-162, :191-194) */
#define RSOPTSC_STOP_MAXITER 1
#define RSOPTSC_STOP_ERRMIN 2
#define RSOPTSC_STOP_EPSILON 3

/* status returned when the caller's progress callback asked to stop (rsoptsc_set_progress) */
#define RSOPTSC_STATUS_INTERRUPTED 9

/* ---- lifecycle ------------------------------------------------------------------- */
int rsoptsc_abi_version(void);
/* Select the CUDA device for this process (one process per GPU) and create the stream and
 * dense-solver handles.  device < 0 keeps the current device. */
int rsoptsc_init(int device);
int rsoptsc_finalize(void);
const char* rsoptsc_last_error(void);
/* Number of kernels of this library launched since init (bench.py's gpu_launches). */
int64_t rsoptsc_kernel_launches(void);
/* Device-time (ms, CUDA events on the library stream) spent in the named phase since the last
 * rsoptsc_timers_reset(); names: "gram", "eig", "recon", "column_step", "lanczos", "dense_ew",
 * "knn", "normalize", "pca", "symm", "nmf", "nndsvd".  Returns -1 for an unknown name. */
double rsoptsc_phase_ms(const char* name);
int64_t rsoptsc_phase_calls(const char* name);
int rsoptsc_timers_reset(void);
int rsoptsc_timers_enable(int on);
/* Caller-defined device timer (CUDA events on the library stream) accumulated under `name`;
 * read back with rsoptsc_phase_ms(name).  bench.py brackets every timed step with these. */
/* Work counters accumulated since the last rsoptsc_timers_reset():
 *   "ozaki_fp64_flops"  FP64-equivalent flops (2 per output MAC) executed by the tcgen05 GEMM
 *   "ozaki_int8_ops"    int8 tensor ops (2 per slice-pair MAC) the same launches issued      */
double rsoptsc_counter(const char* name);
/* i-th phase name seen so far, NULL past the end (enumeration for reports). */
const char* rsoptsc_phase_name(int32_t i);
int rsoptsc_timer_begin(const char* name);
int rsoptsc_timer_end(const char* name);

/* Progress / interrupt hook.  The reference prints `Err = ...` once per ADMM iteration
 * (R/SimilarityM.R:191) and, being R code, can be interrupted between iterations; NMF::nmf checks its
 * stop rule every 10 iterations (R/Clustering.R:29-33).  fn is called on the calling thread
 *   stage RSOPTSC_STAGE_ADMM  once per computeM iteration, iter = the reference's `iter`, value = Err[iter, 1]
 *   stage RSOPTSC_STAGE_NMF   at every connectivity check of the NMF, iter = iterations done, value = 0
 * A non-zero return aborts the running entry point with RSOPTSC_STATUS_INTERRUPTED after the library has
 * released its device memory (the R shim maps this to R_CheckUserInterrupt + Rprintf).  With a multi-GPU
 * communicator every rank must return the same value.  fn = NULL removes the hook. */
#define RSOPTSC_STAGE_ADMM 0
#define RSOPTSC_STAGE_NMF 1
typedef int (*rsoptsc_progress_fn)(int32_t stage, int32_t iter, double value, void* user);
int rsoptsc_set_progress(rsoptsc_progress_fn fn, void* user);

/* ---- multi-GPU: one process per GPU, one communicator per process ------
 * After rsoptsc_comm_init with nranks > 1 the library is SPMD: every rank calls the same entry point
 * with the same inputs (the genes x cells matrix is replicated, <= 2.4 GB) and receives the same
 * outputs.  Inside, one computeM problem is split: column blocks of Z/E/Y1/XXZ/Y3/J per rank
 * (R/SimilarityM.R:164-187 is column-local), the SVT Gram / eigensolver / reconstruction of the large
 * connected blocks split across all ranks with NCCL all-gathers of column panels and a fused
 * peer-memory tridiagonalisation, the stop-rule norms by NCCL allreduce of the genes x genes Gram.
 * The reference has no counterpart (single R process); R hosts bind it through one R worker per GPU. */
/* 128-byte NCCL unique id, created on rank 0 and handed to every rank by the host's own channel. */
int rsoptsc_comm_unique_id(uint8_t* id128);
/* Collective over the nranks processes; call after rsoptsc_init(device).  nranks == 1 is a no-op. */
int rsoptsc_comm_init(int32_t nranks, int32_t rank, const uint8_t* id128);
int rsoptsc_comm_finalize(void);
/* nranks / rank of the active communicator (1 / 0 without one); p2p = 1 once the peer-mapped exchange
 * regions of the fused tridiagonalisation exist; collectives / bytes: data-path NCCL calls issued and
 * their payload since rsoptsc_comm_init.  Any pointer may be NULL. */
int rsoptsc_comm_info(int32_t* nranks, int32_t* rank, int32_t* p2p, int64_t* collectives, double* bytes);

/* device-memory helpers for the *_dev entry points */
int rsoptsc_dev_alloc(void** dptr, int64_t bytes);
int rsoptsc_dev_free(void* dptr);
int rsoptsc_dev_upload(void* dptr, const void* host, int64_t bytes);
int rsoptsc_dev_download(void* host, const void* dptr, int64_t bytes);
int rsoptsc_dev_sync(void);

/* ---- SimilarityM stages ---------------------------------------------------------- */
/* a1  R/SimilarityM.R:34-39   X = column-L2-normalised (data - min)/max ; data, X: m x n */
int rsoptsc_normalize(const double* data, int64_t m, int64_t n, double* X);
int rsoptsc_normalize_dev(const double* d_data, int64_t m, int64_t n, double* d_X);

/* a2  R/SimilarityM.R:41-72   sdev^2 of prcomp(t(X)) (min(m,n) values, descending) and K.
 * emb (n x ncomp, column-major, may be NULL): first ncomp PCA score columns of t(X), the
 * deterministic stand-in for Rtsne/umap. */
int rsoptsc_pca_spectrum(const double* X, int64_t m, int64_t n, double* eig, int32_t* K,
                         double* emb, int32_t ncomp);

/* a3  R/SimilarityM.R:76-77,83-84  exact K nearest neighbours (Euclid, self included) of each
 * row of emb (n x dim, column-major, leading dimension n) using the first `dim` columns.
 * idx: n x K row-major (idx[j*K + t]), 0-based, ascending (distance, index). */
int rsoptsc_knn(const double* emb, int64_t n, int32_t dim, int32_t K, int32_t* idx);

/* a4-a10  computeM(D, X, lambda)  R/SimilarityM.R:115-197.
 * Mask given as neighbour lists (row j of Z may be non-zero only in columns idx[j*K..]).
 * Z (n x n dense, may be NULL), E = ||X - XZ||_F, iters = value of `iter` at return,
 * stop = RSOPTSC_STOP_*, err_trace (may be NULL): 2*100 doubles, Err[i,1] at [i], Err[i,2]
 * lower bound at [100+i] (the reference prints Err[i,1] at :190). */
int rsoptsc_computeM(const double* X, int64_t m, int64_t n, const int32_t* idx, int32_t K,
                     double lambda, double* Z, double* E, int32_t* iters, int32_t* stop,
                     double* err_trace);
/* Same with the reference's dense mask argument D (n x n, entries > 0 are forbidden). */
int rsoptsc_computeM_dense_mask(const double* D, const double* X, int64_t m, int64_t n,
                                double lambda, double* Z, double* E, int32_t* iters,
                                int32_t* stop, double* err_trace);

/* a11  R/SimilarityM.R:94-95   W = 0.5*(|Z'| + |Z'^T|), Z' = Z with entries <= DBL_EPSILON
 * zeroed.  Z, W: n x n dense. */
int rsoptsc_threshold_symmetrise(const double* Z, int64_t n, double* W);
int rsoptsc_threshold_symmetrise_dev(const double* d_Z, int64_t n, double* d_W);

/* Whole SimilarityM (R/SimilarityM.R:27-97): a1, a2, a3, a4-a10, a11.
 *   data      m x n raw (log-normalised) expression
 *   emb       n x emb_dim embedding (what Rtsne/umap return in the reference), or NULL to use
 *             the in-library PCA embedding; the first 3 columns are used (quirk Q1)
 *   K_override  > 0 forces K (config 5), 0 applies the reference rule
 *   W         n x n dense similarity or NULL
 *   sparse COO output (optional, all NULL to skip): *nnz_inout holds the capacity on entry
 *             and the count on return (upper triangle incl. diagonal, 0-based)
 *   emb_out   n x 3 embedding actually used (may be NULL) */
int rsoptsc_similarity(const double* data, int64_t m, int64_t n, const double* emb,
                       int32_t emb_dim, double lambda, int32_t K_override, double* W,
                       int64_t* nnz_inout, int32_t* rows, int32_t* cols, double* vals,
                       double* E, int32_t* iters, int32_t* K_out, double* emb_out);
/* Device-resident variant: d_data (m x n) and d_emb (n x emb_dim or NULL) are device pointers,
 * d_W (n x n, may be NULL) is a device pointer; scalars are host pointers. */
int rsoptsc_similarity_dev(const double* d_data, int64_t m, int64_t n, const double* d_emb,
                           int32_t emb_dim, double lambda, int32_t K_override, double* d_W,
                           double* E, int32_t* iters, int32_t* K_out);

/* SimilarityM with a caller-supplied neighbour table: the reference's `use_umap_indices = TRUE` branch
 * (R/SimilarityM.R:85-87, IDX <- data.umap$knn$indexes) and any other externally computed kNN graph.
 * idx: n x K row-major, 0-based; row j lists the columns of Z that row j may use (:90-92).  The PCA rule for K
 * (:41-72) does not apply (the table fixes the support); everything else as rsoptsc_similarity.  W: n x n. */
int rsoptsc_similarity_knn(const double* data, int64_t m, int64_t n, const int32_t* idx, int32_t K,
                           double lambda, double* W, double* E, int32_t* iters);

/* ---- ClusterCells stages ---------------------------------------------------------- */
/* a12-a13  NMF::nmf(V, rank=k, method='lee', seed='nndsvd') + row-argmax labels
 * (R/Clustering.R:29-37).  V: n x n non-negative (dense).  basis: n x k, coef: k x n
 * (either may be NULL), labels: n, 1-based like R; iters: iterations performed. */
int rsoptsc_nmf_lee(const double* V, int64_t n, int32_t k, int32_t max_iter, double* basis,
                    double* coef, int32_t* labels, int32_t* iters);
int rsoptsc_nmf_lee_dev(const double* d_V, int64_t n, int32_t k, int32_t max_iter,
                        double* basis, double* coef, int32_t* labels, int32_t* iters);

/* Rank sweep: nmf(V, k, 'lee', 'nndsvd') + labels for every k in k_lo .. k_hi on ONE similarity matrix (BASELINE.json
 * configs[3]: k = 5..30; the reference would call ClusterCells once per k, R/Clustering.R:29-37).  The independent
 * problems advance together, one batched launch set per iteration; every k gets exactly the labels and iteration
 * count of a separate rsoptsc_nmf_lee call.  labels: (k_hi - k_lo + 1) x n row-major, 1-based; iters: one per k. */
int rsoptsc_nmf_sweep(const double* V, int64_t n, int32_t k_lo, int32_t k_hi, int32_t max_iter,
                      int32_t* labels, int32_t* iters);
int rsoptsc_nmf_sweep_dev(const double* d_V, int64_t n, int32_t k_lo, int32_t k_hi, int32_t max_iter,
                          int32_t* labels, int32_t* iters);

/* ---- CountClusters stages --------------------------------------------------------- */
/* a14  GetComponents(data, tol)  R/Clustering.R:96-105: ascending |eigenvalues| of the
 * symmetric normalised Laplacian of S (n x n) and the count <= tol. */
int rsoptsc_get_components(const double* S, int64_t n, double tol, double* vals,
                           int32_t* n_eigs);
/* a17  GetConsensus + truncation of GetEnsemble (R/Clustering.R:143-153, :164-175):
 * labels: R x n row-major int32; consen: n x n; entries <= R*tau zeroed, then symmetrised. */
int rsoptsc_consensus(const int32_t* labels, int32_t R, int64_t n, double tau, double* consen);
/* a15-a17  GetEnsemble(data, tol, n_comp, tau, range = lo:hi)  R/Clustering.R:125-154: PCA scores of the
 * similarity S (n x n), cluster::pam for every k in the range, summed consensus, truncation at
 * (hi - lo + 1) * tau, symmetrisation.  consen: n x n.  (`tol` is unused by the reference's body.) */
int rsoptsc_get_ensemble(const double* S, int64_t n, int32_t n_comp, double tau, int32_t range_lo,
                         int32_t range_hi, double* consen);
/* a16  cluster::pam(P, k)$clustering (R/Clustering.R:137): P n x dim column-major;
 * labels n (1-based, numbered by first appearance), medoids k (0-based). */
int rsoptsc_pam(const double* P, int64_t n, int32_t dim, int32_t k, int32_t* labels,
                int32_t* medoids);
/* a14-a18  CountClusters(data, tol, range=lo:hi, n_comp)  R/Clustering.R:61-83.
 * vals: n ascending eigenvalues of the consensus Laplacian (may be NULL). */
int rsoptsc_count_clusters(const double* S, int64_t n, double tol, int32_t range_lo,
                           int32_t range_hi, int32_t n_comp, int32_t* upper_bound,
                           int32_t* lower_bound, double* vals, int32_t* n_eigs);

/* ---- RepresentationMap pieces (R/RepresentationMap.R) --------- */
/* stats::dist(flat_embedding, "euclidean") as a full matrix (:57); emb: n x dim, dist: n x n */
int rsoptsc_dist_flat(const double* emb, int64_t n, int32_t dim, double* dist);
/* adj[S > 0] = 1; diag(adj) = 0 (:59-62); S, adj: n x n */
int rsoptsc_adjacency(const double* S, int64_t n, double* adj);
/* igraph::components of the undirected graph of adj (:64-66): membership 1-based, numbered by
 * the lowest vertex of each component; edge {i,j} iff adj[i,j] != 0 or adj[j,i] != 0 */
int rsoptsc_graph_components(const double* adj, int64_t n, int32_t* membership,
                             int32_t* n_components);
/* igraph::distances(graph) (:89): all-pairs unweighted shortest path lengths, Inf if unreachable */
int rsoptsc_graph_distances(const double* adj, int64_t n, double* dist);

/* ---- preprocessing upstream of the path (R/DataSelection.R) ---- */
/* LogNormalize(M, scale, normalize) :117-133  (M, out: genes x cells) */
int rsoptsc_log_normalize(const double* M, int64_t m, int64_t n, double scale, int32_t normalize,
                          double* out);
/* ScaleCenterData(M) :63-76  per-gene z-score, sd == 0 -> 1 */
int rsoptsc_scale_center(const double* M, int64_t m, int64_t n, double* out);
/* ApplyCountThreshold(M, countL, countU, X) :91-103: keep[g] = 1 for genes to keep (the R function
 * returns which(keep)) */
int rsoptsc_count_threshold(const double* M, int64_t m, int64_t n, double countL, double countU,
                            double X, int32_t* keep);
/* SelectData(M, gene_expression_threshold, n_features) :17-51: gene_use (0-based rows of M, capacity
 * m), *n_selected their count, *max_component (may be NULL) the number of PCs used */
int rsoptsc_select_data(const double* M, int64_t m, int64_t n, double gene_expression_threshold,
                        int32_t n_features, int32_t* gene_use, int32_t* n_selected,
                        int32_t* max_component);

/* ---- GetSignalingPartners core (R/Signaling.R:75-163,291-293) ---- */
/* P[i,j] (n x n, i = ligand cell, j = receptor cell) for one ligand-receptor pair: lig, rec = the
 * two expression rows (n); beta / gamma = per-cell up / down target terms (getBetaj / getGammaj),
 * NULL when the receptor has no up / down targets; normalize != 0 divides every row by its sum
 * (NaN -> 0) as GetSignalingPartners does. */
int rsoptsc_signaling_pair(const double* lig, const double* rec, const double* beta,
                           const double* gamma, int64_t n, int32_t normalize, double* P);

/* ---- dense contraction backends (a4 Gram / reconstruction GEMMs) --------------------- */
/* backend: 0 = cuBLAS FP64 (library baseline), 1 = tcgen05 split-integer (Ozaki) GEMM for
 * contractions with every dimension >= 1536 (default), 2 = Ozaki wherever legal (>= 256). */
int rsoptsc_set_gemm_backend(int32_t backend);
int32_t rsoptsc_get_gemm_backend(void);
/* One contraction on host buffers (column-major), for parity tests and kernel benchmarks:
 *   op 0: C (M x M, lower triangle valid) = A^T A,  A: K x M
 *   op 1: C (M x N) = A B,    A: M x K, B: K x N
 *   op 2: C (M x N) = A B^T,  A: M x K, B: N x K
 *   op 3: C (M x N) = A^T B,  A: K x M, B: K x N
 *   op 4: C (M x M, in/out, lower triangle) -= A B^T + B A^T,  A, B: M x K -- the rank-2K trailing update of
 *         the tridiagonalisation, applied 1 + reps times (warm-up included)
 * reps >= 1 repetitions are timed with CUDA events; *ms_out (may be NULL) = mean ms per call. */
int rsoptsc_debug_gemm(int32_t op, const double* A, const double* B, double* C, int64_t M,
                       int64_t N, int64_t K, int32_t backend, int32_t reps, double* ms_out);

/* Symmetric eigen-decomposition of a host matrix (n x n column-major, lower triangle referenced) with
 * one of the two solvers behind the SVT of rsoptsc_computeM (R/SimilarityM.R:142-155):
 *   solver 0: cuSOLVER syevd;  solver 1: native (cooperative-kernel tridiagonalisation, tridiagonal
 *   divide and conquer with tcgen05 GEMMs, back-transformation);  solver 2: the native solver as a
 *   collective over the ranks of the active communicator (every rank passes the same matrix);
 *   solver 3: eigenvalues only, the path of the PCA spectrum (a2) and the Laplacian spectrum (a14): own
 *   tridiagonalisation + bisection on Sturm counts (V_out is not written).
 * lam_out (n) ascending eigenvalues, V_out (n x n, may be NULL) eigenvectors.  reps >= 1 repetitions
 * are timed with CUDA events; *ms_out (may be NULL) = mean ms per call. */
int rsoptsc_debug_sym_eig(const double* A, int64_t n, int32_t solver, double* lam_out, double* V_out,
                          int32_t reps, double* ms_out);

/* ---- host-only helpers (no GPU needed; exercised by the CPU test-suite) ---------------- */
/* Eigen-decomposition of a symmetric tridiagonal matrix (the small problem of every Lanczos
 * run): d (k) diagonal in / ascending eigenvalues out, e (k-1) off-diagonal, V (k x k
 * column-major eigenvectors, may be NULL). */
int rsoptsc_host_tridiag_eig(int32_t k, double* d, const double* e, double* V);
/* Eigenvalues of the same tridiagonal by bisection on Sturm counts, with the arithmetic of the device
 * kernel behind the values-only eigensolver (csrc/tridiag_bisect.h): lam (k) ascending. */
int rsoptsc_host_tridiag_bisect(int32_t k, const double* d, const double* e, double* lam);
/* The K rule of R/SimilarityM.R:61-72 applied to a descending PCA spectrum of length len. */
int32_t rsoptsc_host_choose_K(const double* eig, int64_t len);
/* Multi-GPU split of one computeM problem over nranks ranks (no GPU needed): for the neighbour table
 * idx (n x K row-major, as rsoptsc_computeM) owner[j] = rank that holds the column-local state of
 * cell j; component[j] = connected component of the support graph (numbered by lowest cell);
 * split[j] = 1 when that component has >= split_min cells and is cut across all ranks.  Any output
 * may be NULL. */
int rsoptsc_host_shard_owner(const int32_t* idx, int64_t n, int32_t K, int32_t nranks, int64_t split_min,
                             int32_t* owner, int32_t* component, int32_t* split);
/* Ownership tables of the fused multi-GPU tridiagonalisation (sytrd_mg.cuh; no GPU needed): the trailing matrix of an
 * n-row block is owned in 128-row groups dealt cyclically, 64-row block row I belongs to rank (I / 2) mod nranks.  For
 * `rank`: ownI (capacity ceil(n/64)) its block rows ascending, *nl their count, pref (capacity ceil(n/64) + 1) the strip
 * prefix sums, lfirst (capacity ceil(n/64) + 1) first own block row at or below every block row.  Used by the CPU
 * test of the kernel's unit map. */
int rsoptsc_host_mg_tables(int64_t n, int32_t nranks, int32_t rank, int32_t* ownI, int32_t* nl, int64_t* pref,
                           int32_t* lfirst);

#ifdef __cplusplus
}
#endif
#endif /* RSOPTSC_B200_H */
