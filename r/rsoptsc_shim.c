/* .Call shim between RSoptSC's R functions and librsoptsc_b200.so (include/rsoptsc_b200.h).
 *
 * Place in RSoptSC/src/ next to a Makevars with
 *     PKG_CPPFLAGS = -I<repo>/include
 *     PKG_LIBS     = -L<repo>/rsoptsc_b200 -lrsoptsc_b200 -Wl,-rpath,<repo>/rsoptsc_b200
 * and add `useDynLib(RSoptSC, .registration = TRUE)` to NAMESPACE.
 *
 * R's C API is single-threaded: every entry point runs on R's main thread.  Errors are raised
 * with Rf_error() only after the library call has returned (the library never longjmps and
 * frees its own device memory before returning).  R is not installed in the build image, so
 * this file is compile-checked against r/stub/Rinternals.h only; it has not been run under R.
 */
#ifdef RSOPTSC_STUB_R
#include "stub/Rinternals.h"
#else
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#endif
#include <stdint.h>
#include <string.h>

#include "rsoptsc_b200.h"

static void check(int status) {
  if (status == RSOPTSC_STATUS_INTERRUPTED) {
    /* the library has already released its device memory: now let R unwind (longjmp) if an interrupt is
     * pending; otherwise report the stop as an ordinary error */
    R_CheckUserInterrupt();
    Rf_error("rsoptsc_b200: interrupted");
  }
  if (status != 0) Rf_error("rsoptsc_b200: %s", rsoptsc_last_error());
}

/* Progress hook (rsoptsc_set_progress): prints the reference's per-iteration line (R/SimilarityM.R:191,
 * print(paste("Err = ", Err[iter,1]))) and turns a pending user interrupt into a clean stop.  R_CheckUserInterrupt
 * longjmps, which must not happen while the library holds device buffers, so the check runs under
 * R_ToplevelExec: it returns FALSE when the interrupt fired, and the callback then asks the library to stop. */
static void check_interrupt_fn(void* dummy) { (void)dummy; R_CheckUserInterrupt(); }
static int progress_cb(int32_t stage, int32_t iter, double value, void* user) {
  (void)user; (void)iter;
  if (stage == RSOPTSC_STAGE_ADMM) Rprintf("[1] \"Err =  %.15g\"\n", value);
  return R_ToplevelExec(check_interrupt_fn, NULL) == FALSE;
}

static void dims(SEXP x, int64_t* nr, int64_t* nc) {
  SEXP d = Rf_getAttrib(x, R_DimSymbol);
  if (Rf_isNull(d)) Rf_error("matrix expected");
  *nr = INTEGER(d)[0];
  *nc = INTEGER(d)[1];
}

static SEXP named_list(int n, const char** names) {
  SEXP out = PROTECT(Rf_allocVector(VECSXP, n));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, n));
  for (int i = 0; i < n; ++i) SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(2);
  return out;
}

/* .Call("rsoptsc_R_similarity", lambda, data, embedding_or_NULL, K_or_0, knn_indexes_or_NULL)
 * replaces the arithmetic of SimilarityM (R/SimilarityM.R:34-95); Rtsne/umap stay in R and
 * hand their layout in as `embedding`; knn_indexes (integer n x k, 1-based: data.umap$knn$indexes) is the
 * use_umap_indices = TRUE branch (:85-87).  Returns list(W, E, iters, K). */
SEXP rsoptsc_R_similarity(SEXP lambda, SEXP data, SEXP emb, SEXP K, SEXP knn) {
  int64_t m, n, en = 0, ed = 0;
  if (!Rf_isReal(data)) Rf_error("data must be a double matrix (use as.matrix)");
  dims(data, &m, &n);
  if ((double)n * (double)n > 2147483647.0 * 4096.0) Rf_error("n too large for a dense W; use the sparse interface");
  const double* e = NULL;
  if (!Rf_isNull(emb)) { dims(emb, &en, &ed); if (en != n) Rf_error("embedding must have one row per cell"); e = REAL(emb); }
  SEXP W = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)n));   /* R_xlen_t sized internally */
  double E = 0; int32_t iters = 0, Kout = 0;
  int st;
  rsoptsc_set_progress(progress_cb, NULL);
  if (!Rf_isNull(knn)) {
    int64_t kr, kc;
    if (!Rf_isInteger(knn)) Rf_error("knn indexes must be an integer matrix");
    dims(knn, &kr, &kc);
    if (kr != n) Rf_error("knn indexes must have one row per cell");
    /* R holds the table column-major and 1-based; the C ABI takes it row-major and 0-based */
    int32_t* idx = (int32_t*)R_alloc((size_t)(kr * kc), sizeof(int32_t));
    for (int64_t j = 0; j < kr; ++j)
      for (int64_t t = 0; t < kc; ++t) idx[j * kc + t] = (int32_t)(INTEGER(knn)[j + t * kr] - 1);
    st = rsoptsc_similarity_knn(REAL(data), m, n, idx, (int32_t)kc, Rf_asReal(lambda), REAL(W), &E, &iters);
    Kout = (int32_t)kc;
  } else {
    st = rsoptsc_similarity(REAL(data), m, n, e, (int32_t)ed, Rf_asReal(lambda), Rf_asInteger(K), REAL(W),
                            NULL, NULL, NULL, NULL, &E, &iters, &Kout, NULL);
  }
  rsoptsc_set_progress(NULL, NULL);
  if (st != 0) { UNPROTECT(1); check(st); }
  const char* nm[] = {"W", "E", "iters", "K"};
  SEXP out = PROTECT(named_list(4, nm));
  SET_VECTOR_ELT(out, 0, W);
  SET_VECTOR_ELT(out, 1, Rf_ScalarReal(E));
  SET_VECTOR_ELT(out, 2, Rf_ScalarInteger(iters));
  SET_VECTOR_ELT(out, 3, Rf_ScalarInteger(Kout));
  UNPROTECT(2);
  return out;
}

/* .Call("rsoptsc_R_computeM", D, X, lambda)  ->  list(Z, E)   (R/SimilarityM.R:115-197) */
SEXP rsoptsc_R_computeM(SEXP D, SEXP X, SEXP lambda) {
  int64_t m, n, dr, dc;
  dims(X, &m, &n);
  dims(D, &dr, &dc);
  if (dr != n || dc != n) Rf_error("D must be n x n");
  SEXP Z = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)n));
  double E = 0; int32_t iters = 0, stop = 0;
  rsoptsc_set_progress(progress_cb, NULL);
  int st = rsoptsc_computeM_dense_mask(REAL(D), REAL(X), m, n, Rf_asReal(lambda), REAL(Z), &E, &iters, &stop, NULL);
  rsoptsc_set_progress(NULL, NULL);
  if (st != 0) { UNPROTECT(1); check(st); }
  const char* nm[] = {"Z", "E"};
  SEXP out = PROTECT(named_list(2, nm));
  SET_VECTOR_ELT(out, 0, Z);
  SET_VECTOR_ELT(out, 1, Rf_ScalarReal(E));
  UNPROTECT(2);
  return out;
}

/* .Call("rsoptsc_R_nmf", V, k, maxIter)  ->  list(H = basis n x k, labels)   (R/Clustering.R:29-37; maxIter is the
 * one NMF option `...` can forward that changes the arithmetic, NMF's default is 2000) */
SEXP rsoptsc_R_nmf(SEXP V, SEXP k, SEXP max_iter) {
  int64_t n, nc;
  dims(V, &n, &nc);
  if (n != nc) Rf_error("similarity matrix must be square");
  const int kk = Rf_asInteger(k);
  SEXP H = PROTECT(Rf_allocMatrix(REALSXP, (int)n, kk));
  SEXP lab = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)n));
  int32_t iters = 0;
  rsoptsc_set_progress(progress_cb, NULL);
  int st = rsoptsc_nmf_lee(REAL(V), n, kk, Rf_asInteger(max_iter), REAL(H), NULL, (int32_t*)INTEGER(lab), &iters);
  rsoptsc_set_progress(NULL, NULL);
  if (st != 0) { UNPROTECT(2); check(st); }
  const char* nm[] = {"H", "labels"};
  SEXP out = PROTECT(named_list(2, nm));
  SET_VECTOR_ELT(out, 0, H);
  SET_VECTOR_ELT(out, 1, lab);
  UNPROTECT(3);
  return out;
}

/* .Call("rsoptsc_R_nmf_sweep", V, k_lo, k_hi, maxIter) -> list(labels = n x (k_hi - k_lo + 1) integer matrix, iters)
 * the rank sweep of BASELINE.json configs[3] (sapply(k_lo:k_hi, function(k) ClusterCells(V, k)$labels) in the
 * reference's terms), all ranks advanced together on the GPU */
SEXP rsoptsc_R_nmf_sweep(SEXP V, SEXP k_lo, SEXP k_hi, SEXP max_iter) {
  int64_t n, nc;
  dims(V, &n, &nc);
  if (n != nc) Rf_error("similarity matrix must be square");
  const int lo = Rf_asInteger(k_lo), hi = Rf_asInteger(k_hi);
  if (hi < lo) Rf_error("k_hi < k_lo");
  const int R = hi - lo + 1;
  /* the C ABI returns R x n row-major == n x R column-major: exactly R's layout of an n x R matrix */
  SEXP lab = PROTECT(Rf_allocMatrix(INTSXP, (int)n, R));
  SEXP its = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)R));
  rsoptsc_set_progress(progress_cb, NULL);
  int st = rsoptsc_nmf_sweep(REAL(V), n, lo, hi, Rf_asInteger(max_iter), (int32_t*)INTEGER(lab), (int32_t*)INTEGER(its));
  rsoptsc_set_progress(NULL, NULL);
  if (st != 0) { UNPROTECT(2); check(st); }
  const char* nm[] = {"labels", "iters"};
  SEXP out = PROTECT(named_list(2, nm));
  SET_VECTOR_ELT(out, 0, lab);
  SET_VECTOR_ELT(out, 1, its);
  UNPROTECT(3);
  return out;
}

/* .Call("rsoptsc_R_get_components", S, tol) -> list(val, n_eigs)   (R/Clustering.R:96-105) */
SEXP rsoptsc_R_get_components(SEXP S, SEXP tol) {
  int64_t n, nc;
  dims(S, &n, &nc);
  SEXP val = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)n));
  int32_t ne = 0;
  int st = rsoptsc_get_components(REAL(S), n, Rf_asReal(tol), REAL(val), &ne);
  if (st != 0) { UNPROTECT(1); check(st); }
  const char* nm[] = {"val", "n_eigs"};
  SEXP out = PROTECT(named_list(2, nm));
  SET_VECTOR_ELT(out, 0, val);
  SET_VECTOR_ELT(out, 1, Rf_ScalarInteger(ne));
  UNPROTECT(2);
  return out;
}

/* .Call("rsoptsc_R_count_clusters", S, tol, range_lo, range_hi, n_comp)
 *   -> list(upper_bound, lower_bound, eigs = list(val, n_eigs))   (R/Clustering.R:61-83) */
SEXP rsoptsc_R_count_clusters(SEXP S, SEXP tol, SEXP lo, SEXP hi, SEXP n_comp) {
  int64_t n, nc;
  dims(S, &n, &nc);
  SEXP val = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)n));
  int32_t up = 0, low = 0, ne = 0;
  int st = rsoptsc_count_clusters(REAL(S), n, Rf_asReal(tol), Rf_asInteger(lo), Rf_asInteger(hi), Rf_asInteger(n_comp),
                                  &up, &low, REAL(val), &ne);
  if (st != 0) { UNPROTECT(1); check(st); }
  const char* en[] = {"val", "n_eigs"};
  SEXP eigs = PROTECT(named_list(2, en));
  SET_VECTOR_ELT(eigs, 0, val);
  SET_VECTOR_ELT(eigs, 1, Rf_ScalarInteger(ne));
  const char* nm[] = {"upper_bound", "lower_bound", "eigs"};
  SEXP out = PROTECT(named_list(3, nm));
  SET_VECTOR_ELT(out, 0, Rf_ScalarInteger(up));
  SET_VECTOR_ELT(out, 1, Rf_ScalarInteger(low));
  SET_VECTOR_ELT(out, 2, eigs);
  UNPROTECT(3);
  return out;
}

/* .Call("rsoptsc_R_get_ensemble", S, n_comp, tau, range_lo, range_hi) -> truncated ensemble consensus matrix
 * (GetEnsemble, R/Clustering.R:125-154) */
SEXP rsoptsc_R_get_ensemble(SEXP S, SEXP n_comp, SEXP tau, SEXP lo, SEXP hi) {
  int64_t n, nc;
  dims(S, &n, &nc);
  if (n != nc) Rf_error("similarity matrix must be square");
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)n));
  int st = rsoptsc_get_ensemble(REAL(S), n, Rf_asInteger(n_comp), Rf_asReal(tau), Rf_asInteger(lo), Rf_asInteger(hi), REAL(out));
  if (st != 0) { UNPROTECT(1); check(st); }
  UNPROTECT(1);
  return out;
}

/* .Call("rsoptsc_R_get_consensus", clusters) -> n x n 0/1 matrix (GetConsensus, R/Clustering.R:164-175) */
SEXP rsoptsc_R_get_consensus(SEXP clusters) {
  const R_xlen_t n = Rf_xlength(clusters);
  if (!Rf_isInteger(clusters)) Rf_error("clusters must be an integer vector (use as.integer)");
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)n));
  int st = rsoptsc_consensus((const int32_t*)INTEGER(clusters), 1, (int64_t)n, -1.0, REAL(out));
  if (st != 0) { UNPROTECT(1); check(st); }
  UNPROTECT(1);
  return out;
}

/* R/DataSelection.R: .Call("rsoptsc_R_log_normalize", M, scale, normalize) (:117-133),
 * .Call("rsoptsc_R_scale_center", M) (:63-76), .Call("rsoptsc_R_count_threshold", M, countL, countU, X) -> which(keep)
 * (:91-103), .Call("rsoptsc_R_select_data", M, gene_expression_threshold, n_features) -> gene_use, 1-based (:17-51) */
SEXP rsoptsc_R_log_normalize(SEXP M, SEXP scale, SEXP normalize) {
  int64_t m, n;
  dims(M, &m, &n);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)m, (int)n));
  int st = rsoptsc_log_normalize(REAL(M), m, n, Rf_asReal(scale), Rf_asInteger(normalize), REAL(out));
  if (st != 0) { UNPROTECT(1); check(st); }
  UNPROTECT(1);
  return out;
}
SEXP rsoptsc_R_scale_center(SEXP M) {
  int64_t m, n;
  dims(M, &m, &n);
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, (int)m, (int)n));
  int st = rsoptsc_scale_center(REAL(M), m, n, REAL(out));
  if (st != 0) { UNPROTECT(1); check(st); }
  UNPROTECT(1);
  return out;
}
static SEXP which_1based(const int32_t* keep, int64_t m) {
  R_xlen_t cnt = 0;
  for (int64_t g = 0; g < m; ++g) cnt += keep[g] != 0;
  SEXP out = PROTECT(Rf_allocVector(INTSXP, cnt));
  R_xlen_t k = 0;
  for (int64_t g = 0; g < m; ++g) if (keep[g]) INTEGER(out)[k++] = (int)(g + 1);
  UNPROTECT(1);
  return out;
}
SEXP rsoptsc_R_count_threshold(SEXP M, SEXP countL, SEXP countU, SEXP X) {
  int64_t m, n;
  dims(M, &m, &n);
  int32_t* keep = (int32_t*)R_alloc((size_t)m, sizeof(int32_t));
  check(rsoptsc_count_threshold(REAL(M), m, n, Rf_asReal(countL), Rf_asReal(countU), Rf_asReal(X), keep));
  return which_1based(keep, m);
}
SEXP rsoptsc_R_select_data(SEXP M, SEXP thr, SEXP n_features) {
  int64_t m, n;
  dims(M, &m, &n);
  int32_t* use = (int32_t*)R_alloc((size_t)m, sizeof(int32_t));
  int32_t cnt = 0, maxc = 0;
  check(rsoptsc_select_data(REAL(M), m, n, Rf_asReal(thr), Rf_asInteger(n_features), use, &cnt, &maxc));
  SEXP out = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)cnt));
  for (int32_t k = 0; k < cnt; ++k) INTEGER(out)[k] = use[k] + 1;
  UNPROTECT(1);
  return out;
}

/* .Call("rsoptsc_R_graph_distances", adj) -> n x n unweighted shortest-path lengths (Inf if unreachable);
 * replaces igraph::distances(similarity_graph) in RepresentationMap (R/RepresentationMap.R:89) */
SEXP rsoptsc_R_graph_distances(SEXP adj) {
  int64_t n, nc;
  dims(adj, &n, &nc);
  if (n != nc) Rf_error("adjacency matrix must be square");
  SEXP D = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)n));
  int st = rsoptsc_graph_distances(REAL(adj), n, REAL(D));
  if (st != 0) { UNPROTECT(1); check(st); }
  UNPROTECT(1);
  return D;
}

/* .Call("rsoptsc_R_dist_flat", flat_embedding) -> as.matrix(dist(flat_embedding)) (R/RepresentationMap.R:57) */
SEXP rsoptsc_R_dist_flat(SEXP emb) {
  int64_t n, d;
  dims(emb, &n, &d);
  SEXP D = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)n));
  int st = rsoptsc_dist_flat(REAL(emb), n, (int32_t)d, REAL(D));
  if (st != 0) { UNPROTECT(1); check(st); }
  UNPROTECT(1);
  return D;
}

/* .Call("rsoptsc_R_signaling_pair", lig, rec, beta_or_NULL, gamma_or_NULL) -> row-normalised n x n P
 * (updateFlat* + normalisation of GetSignalingPartners, R/Signaling.R:75-163, :291-293) */
SEXP rsoptsc_R_signaling_pair(SEXP lig, SEXP rec, SEXP beta, SEXP gamma) {
  const R_xlen_t n = Rf_xlength(lig);
  if (Rf_xlength(rec) != n) Rf_error("lig and rec must have one value per cell");
  SEXP P = PROTECT(Rf_allocMatrix(REALSXP, (int)n, (int)n));
  int st = rsoptsc_signaling_pair(REAL(lig), REAL(rec), Rf_isNull(beta) ? NULL : REAL(beta),
                                  Rf_isNull(gamma) ? NULL : REAL(gamma), (int64_t)n, 1, REAL(P));
  if (st != 0) { UNPROTECT(1); check(st); }
  UNPROTECT(1);
  return P;
}

/* Multi-GPU: one R worker per GPU.  .Call("rsoptsc_R_init", device);
 * id <- .Call("rsoptsc_R_comm_unique_id") on rank 0 (raw vector of 128 bytes, handed to the other workers by the host's
 * own channel, e.g. parallel::clusterExport); .Call("rsoptsc_R_comm_init", nranks, rank, id) on every worker; from then
 * on every worker makes the SAME SimilarityM / computeM call and the library splits the one problem. */
SEXP rsoptsc_R_init(SEXP device) {
  check(rsoptsc_init(Rf_asInteger(device)));
  return R_NilValue;
}
SEXP rsoptsc_R_comm_unique_id(void) {
  SEXP id = PROTECT(Rf_allocVector(RAWSXP, 128));
  int st = rsoptsc_comm_unique_id((uint8_t*)RAW(id));
  if (st != 0) { UNPROTECT(1); check(st); }
  UNPROTECT(1);
  return id;
}
SEXP rsoptsc_R_comm_init(SEXP nranks, SEXP rank, SEXP id) {
  if (Rf_xlength(id) != 128) Rf_error("the NCCL unique id is a raw vector of 128 bytes");
  check(rsoptsc_comm_init(Rf_asInteger(nranks), Rf_asInteger(rank), (const uint8_t*)RAW(id)));
  return R_NilValue;
}
SEXP rsoptsc_R_comm_finalize(void) {
  check(rsoptsc_comm_finalize());
  return R_NilValue;
}

static const R_CallMethodDef call_methods[] = {
    {"rsoptsc_R_init", (DL_FUNC)&rsoptsc_R_init, 1},
    {"rsoptsc_R_comm_unique_id", (DL_FUNC)&rsoptsc_R_comm_unique_id, 0},
    {"rsoptsc_R_comm_init", (DL_FUNC)&rsoptsc_R_comm_init, 3},
    {"rsoptsc_R_comm_finalize", (DL_FUNC)&rsoptsc_R_comm_finalize, 0},
    {"rsoptsc_R_graph_distances", (DL_FUNC)&rsoptsc_R_graph_distances, 1},
    {"rsoptsc_R_dist_flat", (DL_FUNC)&rsoptsc_R_dist_flat, 1},
    {"rsoptsc_R_signaling_pair", (DL_FUNC)&rsoptsc_R_signaling_pair, 4},
    {"rsoptsc_R_similarity", (DL_FUNC)&rsoptsc_R_similarity, 5},
    {"rsoptsc_R_computeM", (DL_FUNC)&rsoptsc_R_computeM, 3},
    {"rsoptsc_R_nmf", (DL_FUNC)&rsoptsc_R_nmf, 3},
    {"rsoptsc_R_nmf_sweep", (DL_FUNC)&rsoptsc_R_nmf_sweep, 4},
    {"rsoptsc_R_get_ensemble", (DL_FUNC)&rsoptsc_R_get_ensemble, 5},
    {"rsoptsc_R_get_consensus", (DL_FUNC)&rsoptsc_R_get_consensus, 1},
    {"rsoptsc_R_log_normalize", (DL_FUNC)&rsoptsc_R_log_normalize, 3},
    {"rsoptsc_R_scale_center", (DL_FUNC)&rsoptsc_R_scale_center, 1},
    {"rsoptsc_R_count_threshold", (DL_FUNC)&rsoptsc_R_count_threshold, 4},
    {"rsoptsc_R_select_data", (DL_FUNC)&rsoptsc_R_select_data, 3},
    {"rsoptsc_R_get_components", (DL_FUNC)&rsoptsc_R_get_components, 2},
    {"rsoptsc_R_count_clusters", (DL_FUNC)&rsoptsc_R_count_clusters, 5},
    {NULL, NULL, 0}};

void R_init_RSoptSC(DllInfo* dll) {
  R_registerRoutines(dll, NULL, call_methods, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
  rsoptsc_init(-1); /* one process per GPU: current device; failure surfaces on first call */
}
