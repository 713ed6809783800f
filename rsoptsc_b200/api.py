"""Host-side mirror of the reference's R interface for the hot path, over the C ABI.

Same function names, argument meaning and return fields as the reference
(R/SimilarityM.R:27-28,96; R/Clustering.R:20-23,39-41,61,80-82,96-104,164), so the parity tests
read like the reference's own tests.  Matrices are genes x cells (column-major inside the
library, any numpy layout accepted here).  Indices/labels are 1-based like R where the reference
returns them (labels, upper_bound); neighbour index matrices are 0-based numpy arrays.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import RsoptscError, c_double_p, c_int32_p, c_int64_p  # noqa: F401

STOP_REASON = {1: "reached max iter", 2: "convergence by error min", 3: "reached epsilon"}


def _f64(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _ip(a):
    return a.ctypes.data_as(c_int32_p)


def normalize(data):
    """R/SimilarityM.R:34-39."""
    lib = _lib.init()
    d = _f64(data)
    m, n = d.shape
    X = np.empty((m, n), dtype=np.float64, order="F")
    _lib.check(lib.rsoptsc_normalize(_dp(d), m, n, _dp(X)))
    return X


def pca_spectrum(X, ncomp=0):
    """R/SimilarityM.R:41-72 -> dict(eig=sdev^2, K=..., emb=n x ncomp PCA scores or None)."""
    lib = _lib.init()
    x = _f64(X)
    m, n = x.shape
    eig = np.empty(min(m, n), dtype=np.float64)
    K = C.c_int32(0)
    emb = np.empty((n, ncomp), dtype=np.float64, order="F") if ncomp else None
    _lib.check(lib.rsoptsc_pca_spectrum(_dp(x), m, n, _dp(eig), C.byref(K), _dp(emb) if ncomp else None, ncomp))
    return dict(eig=eig, K=int(K.value), emb=emb)


def knn(emb, K, dim=3):
    """R/SimilarityM.R:76-77 (FNN::get.knnx on the first `dim` embedding columns); 0-based n x K."""
    lib = _lib.init()
    e = _f64(np.asarray(emb)[:, :dim])
    n, dim = e.shape
    idx = np.empty((n, K), dtype=np.int32)
    _lib.check(lib.rsoptsc_knn(_dp(e), n, dim, K, _ip(idx)))
    return idx


def _computeM_outputs(n, want_Z):
    Z = np.empty((n, n), dtype=np.float64, order="F") if want_Z else None
    return Z, C.c_double(0), C.c_int32(0), C.c_int32(0), np.zeros(200, dtype=np.float64)


def computeM(D, X, lam, idx=None, return_Z=True):
    """computeM(D, X, lambda) of R/SimilarityM.R:115 -> dict(Z, E, iters, stop, Err).
    Either the reference's dense mask D (n x n, entries > 0 forbidden) or a 0-based neighbour
    list `idx` (n x K) may be given."""
    lib = _lib.init()
    x = _f64(X)
    m, n = x.shape
    Z, E, iters, stop, trace = _computeM_outputs(n, return_Z)
    zp = _dp(Z) if return_Z else None
    if idx is not None:
        ix = np.ascontiguousarray(idx, dtype=np.int32)
        if ix.ndim != 2 or ix.shape[0] != n:
            raise ValueError("idx must be n x K")
        _lib.check(lib.rsoptsc_computeM(_dp(x), m, n, _ip(ix), ix.shape[1], float(lam), zp, C.byref(E),
                                        C.byref(iters), C.byref(stop), _dp(trace)))
    else:
        d = _f64(D)
        if d.shape != (n, n):
            raise ValueError("D must be n x n")
        _lib.check(lib.rsoptsc_computeM_dense_mask(_dp(d), _dp(x), m, n, float(lam), zp, C.byref(E),
                                                   C.byref(iters), C.byref(stop), _dp(trace)))
    it = int(iters.value)
    return dict(Z=Z, E=float(E.value), iters=it, stop=STOP_REASON.get(int(stop.value), "?"),
                Err=np.stack([trace[:100], trace[100:]], axis=1)[:max(it, 0)])


def threshold_symmetrise(Z):
    """R/SimilarityM.R:94-95."""
    lib = _lib.init()
    z = _f64(Z)
    n = z.shape[0]
    W = np.empty((n, n), dtype=np.float64, order="F")
    _lib.check(lib.rsoptsc_threshold_symmetrise(_dp(z), n, _dp(W)))
    return W


_progress_keepalive = None


def set_progress(fn):
    """Install (fn) or remove (None) the progress / interrupt hook of the C ABI (rsoptsc_set_progress).
    fn(stage, iter, value) is called once per computeM iteration (stage 0, value = Err[iter, 1], the number
    the reference prints at R/SimilarityM.R:191) and at every NMF stop check (stage 1); a true return value
    aborts the running call with RsoptscError (status 9, the R shim's R_CheckUserInterrupt)."""
    global _progress_keepalive
    lib = _lib.init()
    if fn is None:
        _lib.check(lib.rsoptsc_set_progress(None, None))
        _progress_keepalive = None
        return
    cb = _lib.PROGRESS_FN(lambda stage, it, value, user: 1 if fn(int(stage), int(it), float(value)) else 0)
    _lib.check(lib.rsoptsc_set_progress(C.cast(cb, C.c_void_p), None))
    _progress_keepalive = cb


def SimilarityM(lam=0.5, data=None, embedding=None, K=None, dense=True, sparse=False, W_out=None, knn_indices=None):
    """SimilarityM(lambda, data, ...) of R/SimilarityM.R:27 -> dict(W, E, nl_embedding, ...).

    `embedding` is what Rtsne/umap produce in the reference (n x >= 3); None selects the
    deterministic in-library PCA embedding.  `K` forces the neighbour count (config 5).
    `W_out` (optional, n x n float64 Fortran-ordered, e.g. pinned memory) receives W.
    `knn_indices` (n x K, 0-based) is the reference's `use_umap_indices = TRUE` branch (R/SimilarityM.R:85-87):
    the caller's neighbour table replaces the in-library kNN search."""
    lib = _lib.init()
    d = _f64(data)
    m, n = d.shape
    if knn_indices is not None:
        idx = np.ascontiguousarray(knn_indices, dtype=np.int32)
        if idx.ndim != 2 or idx.shape[0] != n:
            raise ValueError("knn_indices must be n x K")
        W = W_out if W_out is not None else np.empty((n, n), dtype=np.float64, order="F")
        E, iters = C.c_double(0), C.c_int32(0)
        _lib.check(lib.rsoptsc_similarity_knn(_dp(d), m, n, _ip(idx), idx.shape[1], float(lam), _dp(W), C.byref(E),
                                              C.byref(iters)))
        return dict(W=W, E=float(E.value), nl_embedding=embedding, iters=int(iters.value), K=int(idx.shape[1]))
    emb = None if embedding is None else _f64(embedding)
    if W_out is not None:
        if W_out.shape != (n, n) or W_out.dtype != np.float64 or not W_out.flags.f_contiguous:
            raise ValueError("W_out must be an n x n float64 Fortran-ordered array")
        dense = True
    W = W_out if W_out is not None else (np.empty((n, n), dtype=np.float64, order="F") if dense else None)
    E, iters, Kout = C.c_double(0), C.c_int32(0), C.c_int32(0)
    emb_out = np.zeros((n, 3), dtype=np.float64, order="F")
    cap = C.c_int64(2 * n * 32 if sparse else 0)
    rows = np.empty(cap.value, dtype=np.int32) if sparse else None
    cols = np.empty(cap.value, dtype=np.int32) if sparse else None
    vals = np.empty(cap.value, dtype=np.float64) if sparse else None
    _lib.check(lib.rsoptsc_similarity(
        _dp(d), m, n, _dp(emb) if emb is not None else None, emb.shape[1] if emb is not None else 0, float(lam),
        int(K or 0), _dp(W) if dense else None, C.byref(cap) if sparse else None,
        _ip(rows) if sparse else None, _ip(cols) if sparse else None, _dp(vals) if sparse else None,
        C.byref(E), C.byref(iters), C.byref(Kout), _dp(emb_out)))
    out = dict(W=W, E=float(E.value), nl_embedding=emb_out if emb is None else emb, iters=int(iters.value),
               K=int(Kout.value))
    if sparse:
        c = int(cap.value)
        out["W_coo"] = (rows[:c].copy(), cols[:c].copy(), vals[:c].copy())
    return out


def nmf_lee(V, rank, max_iter=2000):
    """NMF::nmf(V, rank, method='lee', seed='nndsvd') -> dict(W basis n x k, H coef k x n, labels, iters)."""
    lib = _lib.init()
    v = _f64(V)
    n = v.shape[0]
    if v.shape != (n, n):
        raise ValueError("V must be square (cells x cells similarity)")
    basis = np.empty((n, rank), dtype=np.float64, order="F")
    coef = np.empty((rank, n), dtype=np.float64, order="F")
    labels = np.empty(n, dtype=np.int32)
    iters = C.c_int32(0)
    _lib.check(lib.rsoptsc_nmf_lee(_dp(v), n, int(rank), int(max_iter), _dp(basis), _dp(coef), _ip(labels),
                                   C.byref(iters)))
    return dict(W=basis, H=coef, labels=labels, iters=int(iters.value))


def nmf_sweep(V, ranks=(5, 30), max_iter=2000):
    """nmf(V, k, 'lee', 'nndsvd') labels for every k in ranks[0] .. ranks[1] (BASELINE.json configs[3]), the
    independent problems batched on the GPU -> dict(ranks, labels (R x n, 1-based), iters (R))."""
    lib = _lib.init()
    v = _f64(V)
    n = v.shape[0]
    k_lo, k_hi = int(min(ranks)), int(max(ranks))
    R = k_hi - k_lo + 1
    labels = np.empty((R, n), dtype=np.int32)
    iters = np.empty(R, dtype=np.int32)
    _lib.check(lib.rsoptsc_nmf_sweep(_dp(v), n, k_lo, k_hi, int(max_iter), _ip(labels), _ip(iters)))
    return dict(ranks=list(range(k_lo, k_hi + 1)), labels=labels, iters=iters)


def GetComponents(data, tol=0.01):
    """R/Clustering.R:96-105 -> dict(val ascending, n_eigs)."""
    lib = _lib.init()
    s = _f64(data)
    n = s.shape[0]
    vals = np.empty(n, dtype=np.float64)
    ne = C.c_int32(0)
    _lib.check(lib.rsoptsc_get_components(_dp(s), n, float(tol), _dp(vals), C.byref(ne)))
    return dict(val=vals, n_eigs=int(ne.value))


def GetConsensusEnsemble(labels, tau):
    """Sum of GetConsensus over the label table + truncation + symmetrisation
    (R/Clustering.R:143-153, :164-175).  labels: R x n."""
    lib = _lib.init()
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    R, n = lab.shape
    out = np.empty((n, n), dtype=np.float64, order="F")
    _lib.check(lib.rsoptsc_consensus(_ip(lab), R, n, float(tau), _dp(out)))
    return out


def GetConsensus(clusters):
    """GetConsensus(clusters) of R/Clustering.R:164-175: consen[i, j] = 1 where the labels agree."""
    return GetConsensusEnsemble(np.asarray(clusters, dtype=np.int32)[None, :], tau=-1.0)


def GetEnsemble(data, tol=0.01, n_comp=3, tau=0.3, range=(2, 20)):
    """GetEnsemble(data, tol, n_comp, tau, range) of R/Clustering.R:125-154 -> truncated consensus matrix."""
    lib = _lib.init()
    s = _f64(data)
    n = s.shape[0]
    out = np.empty((n, n), dtype=np.float64, order="F")
    _lib.check(lib.rsoptsc_get_ensemble(_dp(s), n, int(n_comp), float(tau), int(min(range)), int(max(range)), _dp(out)))
    return out


def pam(P, k):
    """cluster::pam(P, k) -> (clustering 1-based, medoid indices 0-based)."""
    lib = _lib.init()
    p = _f64(P)
    n, dim = p.shape
    labels = np.empty(n, dtype=np.int32)
    med = np.empty(k, dtype=np.int32)
    _lib.check(lib.rsoptsc_pam(_dp(p), n, dim, int(k), _ip(labels), _ip(med)))
    return labels, med


def CountClusters(data, tol=0.01, range=(2, 20), eigengap=True, n_comp=3):
    """R/Clustering.R:61-83 -> dict(upper_bound, lower_bound, eigs=dict(val, n_eigs))."""
    lib = _lib.init()
    s = _f64(data)
    n = s.shape[0]
    vals = np.empty(n, dtype=np.float64)
    up, lo, ne = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    _lib.check(lib.rsoptsc_count_clusters(_dp(s), n, float(tol), int(range[0]), int(range[-1]), int(n_comp),
                                          C.byref(up), C.byref(lo), _dp(vals), C.byref(ne)))
    return dict(upper_bound=int(up.value), lower_bound=int(lo.value), eigs=dict(val=vals, n_eigs=int(ne.value)))


def ClusterCells(similarityMatrix=None, n_clusters=None, n_comp=3, max_iter=2000):
    """R/Clustering.R:20-42 -> dict(H basis n x k, labels 1-based, ensemble).  With n_clusters
    given the reference fails on an undefined variable (quirk Q11); here ensemble is None."""
    ensemble = None
    if n_clusters is None:
        ensemble = CountClusters(similarityMatrix, n_comp=n_comp)
        n_clusters = ensemble["upper_bound"]
    r = nmf_lee(similarityMatrix, n_clusters, max_iter=max_iter)
    return dict(H=r["W"], labels=r["labels"], ensemble=ensemble, iters=r["iters"], coef=r["H"])


def dist_flat(flat_embedding):
    """as.matrix(dist(flat_embedding, method='euclidean')) of R/RepresentationMap.R:57."""
    lib = _lib.init()
    e = _f64(flat_embedding)
    n, dim = e.shape
    out = np.empty((n, n), dtype=np.float64, order="F")
    _lib.check(lib.rsoptsc_dist_flat(_dp(e), n, dim, _dp(out)))
    return out


def adjacency(similarity_matrix):
    """adj[S > 0] = 1; diag(adj) = 0 of R/RepresentationMap.R:59-62."""
    lib = _lib.init()
    s = _f64(similarity_matrix)
    n = s.shape[0]
    out = np.empty((n, n), dtype=np.float64, order="F")
    _lib.check(lib.rsoptsc_adjacency(_dp(s), n, _dp(out)))
    return out


def graph_components(adj):
    """igraph::components(graph_from_adjacency_matrix(adj, 'undirected')) -> dict(membership 1-based, csize, no)."""
    lib = _lib.init()
    a = _f64(adj)
    n = a.shape[0]
    mem = np.empty(n, dtype=np.int32)
    no = C.c_int32(0)
    _lib.check(lib.rsoptsc_graph_components(_dp(a), n, _ip(mem), C.byref(no)))
    return dict(membership=mem, csize=np.bincount(mem)[1:], no=int(no.value))


def graph_distances(adj):
    """igraph::distances(graph) (unweighted all-pairs shortest paths, Inf when unreachable)."""
    lib = _lib.init()
    a = _f64(adj)
    n = a.shape[0]
    out = np.empty((n, n), dtype=np.float64, order="F")
    _lib.check(lib.rsoptsc_graph_distances(_dp(a), n, _dp(out)))
    return out


def RepresentationMap(flat_embedding, similarity_matrix, join_components=False, normalize_S=True):
    """Dense/graph part of RepresentationMap (R/RepresentationMap.R:30-100) for a given 2-D
    `flat_embedding` (umap/Rtsne stay in R, as does JoinGraphComponents): returns dist_flat,
    adj_matrix, components, n_components, dist_graph.  normalize_S only rescales S by its sum
    (:37-39) and does not change the S > 0 pattern."""
    if join_components:
        raise NotImplementedError("JoinGraphComponents stays in R host code (SURVEY.md section 2)")
    adj = adjacency(similarity_matrix)
    comps = graph_components(adj)
    return dict(dist_flat=dist_flat(flat_embedding), adj_matrix=adj, components=comps, n_components=comps["no"],
                dist_graph=graph_distances(adj), flat_embedding=flat_embedding)


def LogNormalize(M, scale=1e4, normalize=True):
    """R/DataSelection.R:117-133."""
    lib = _lib.init()
    a = _f64(M)
    m, n = a.shape
    out = np.empty((m, n), dtype=np.float64, order="F")
    _lib.check(lib.rsoptsc_log_normalize(_dp(a), m, n, float(scale), 1 if normalize else 0, _dp(out)))
    return out


def ScaleCenterData(M):
    """R/DataSelection.R:63-76."""
    lib = _lib.init()
    a = _f64(M)
    m, n = a.shape
    out = np.empty((m, n), dtype=np.float64, order="F")
    _lib.check(lib.rsoptsc_scale_center(_dp(a), m, n, _dp(out)))
    return out


def ApplyCountThreshold(M, countL=1, countU=3, X=3):
    """R/DataSelection.R:91-103 -> 1-based indices of the genes to keep (like which())."""
    lib = _lib.init()
    a = _f64(M)
    m, n = a.shape
    keep = np.zeros(m, dtype=np.int32)
    _lib.check(lib.rsoptsc_count_threshold(_dp(a), m, n, float(countL), float(countU), float(X), _ip(keep)))
    return np.flatnonzero(keep) + 1


def SelectData(M, gene_expression_threshold, n_features):
    """R/DataSelection.R:17-51 -> dict(gene_use 1-based, M_variable, max_component)."""
    lib = _lib.init()
    a = _f64(M)
    m, n = a.shape
    use = np.zeros(m, dtype=np.int32)
    cnt, mc = C.c_int32(0), C.c_int32(0)
    _lib.check(lib.rsoptsc_select_data(_dp(a), m, n, float(gene_expression_threshold), int(n_features), _ip(use),
                                       C.byref(cnt), C.byref(mc)))
    sel = use[:cnt.value].astype(np.int64)
    return dict(gene_use=sel + 1, M_variable=a[sel, :], max_component=int(mc.value))


def signaling_pair(lig, rec, beta=None, gamma=None, normalize=True):
    """P (n x n) of one ligand-receptor pair: updateFlat{Both,Up,Down,None} + row normalisation
    (R/Signaling.R:75-163, :291-293)."""
    lib = _lib.init()
    lg, rc = np.ascontiguousarray(lig, dtype=np.float64), np.ascontiguousarray(rec, dtype=np.float64)
    n = lg.shape[0]
    b = None if beta is None else np.ascontiguousarray(beta, dtype=np.float64)
    g = None if gamma is None else np.ascontiguousarray(gamma, dtype=np.float64)
    P = np.empty((n, n), dtype=np.float64, order="F")
    _lib.check(lib.rsoptsc_signaling_pair(_dp(lg), _dp(rc), _dp(b) if b is not None else None,
                                          _dp(g) if g is not None else None, n, 1 if normalize else 0, _dp(P)))
    return P


def GetSignalingPartners(data, gene_names, ligrec, rectarg):
    """GetSignalingPartners (R/Signaling.R:187-309) for an expression matrix `data` (genes x cells),
    its row names, a list of (ligand, receptor) pairs and a list of (receptor, target, direction)
    rows with direction in {'up', 'down'}.  Table handling (dplyr in the reference) and the O(t n)
    target terms getBetaj/getGammaj (:43-60) are host glue; every n x n matrix is computed on the GPU.
    Returns dict(P=[n x n per pair], P_tot=n x n, LR=[(lig, rec)])."""
    data = np.asarray(data, dtype=np.float64)
    row = {str(g).lower(): i for i, g in enumerate(gene_names)}
    P, LR = [], []
    for lig, rec in ligrec:
        lig, rec = str(lig).lower(), str(rec).lower()
        ups = sorted({str(t).lower() for r, t, d in rectarg if str(r).lower() == rec and d == "up"})
        downs = sorted({str(t).lower() for r, t, d in rectarg if str(r).lower() == rec and d == "down"})
        beta = gamma = None
        with np.errstate(divide="ignore", invalid="ignore"):
            if ups:
                s = data[[row[t] for t in ups]].sum(axis=0)
                beta = np.where(s == 0, 0.0, np.exp(-len(ups) / s))          # getBetaj
            if downs:
                s = data[[row[t] for t in downs]].sum(axis=0)
                gamma = np.where(s == 0, 0.0, np.exp(-s / len(downs)))       # getGammaj
        P.append(signaling_pair(data[row[lig]], data[row[rec]], beta, gamma, normalize=True))
        LR.append((lig, rec))
    tot = np.zeros_like(P[0])
    for p in P:
        tot = tot + p
    with np.errstate(divide="ignore", invalid="ignore"):
        tot = tot / tot.sum(axis=1, keepdims=True)
    tot[np.isnan(tot)] = 0.0
    return dict(P=P, P_tot=tot, LR=LR)


class DeviceBuffer:
    """Library-owned device allocation (for the *_dev entry points: inputs resident in HBM)."""

    def __init__(self, nbytes):
        lib = _lib.init()
        self.ptr = C.c_void_p(0)
        self.nbytes = int(nbytes)
        _lib.check(lib.rsoptsc_dev_alloc(C.byref(self.ptr), self.nbytes))

    @classmethod
    def from_host(cls, arr):
        a = _f64(arr)
        buf = cls(a.nbytes)
        _lib.check(_lib.load().rsoptsc_dev_upload(buf.ptr, a.ctypes.data_as(C.c_void_p), a.nbytes))
        return buf

    def to_host(self, shape, dtype=np.float64):
        out = np.empty(shape, dtype=dtype, order="F")
        _lib.check(_lib.load().rsoptsc_dev_download(out.ctypes.data_as(C.c_void_p), self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            _lib.load().rsoptsc_dev_free(self.ptr)
            self.ptr = C.c_void_p(0)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def SimilarityM_dev(lam, d_data, m, n, d_W=None, d_emb=None, emb_dim=0, K=None):
    """Device-resident SimilarityM: d_data (m x n) and d_W (n x n) are DeviceBuffers."""
    lib = _lib.init()
    E, iters, Kout = C.c_double(0), C.c_int32(0), C.c_int32(0)
    _lib.check(lib.rsoptsc_similarity_dev(d_data.ptr, m, n, d_emb.ptr if d_emb is not None else None, emb_dim,
                                          float(lam), int(K or 0), d_W.ptr if d_W is not None else None,
                                          C.byref(E), C.byref(iters), C.byref(Kout)))
    return dict(E=float(E.value), iters=int(iters.value), K=int(Kout.value))


def nmf_lee_dev(d_V, n, rank, max_iter=2000, want_factors=False):
    """Device-resident NMF: d_V is a DeviceBuffer holding the n x n similarity."""
    lib = _lib.init()
    labels = np.empty(n, dtype=np.int32)
    iters = C.c_int32(0)
    basis = np.empty((n, rank), dtype=np.float64, order="F") if want_factors else None
    _lib.check(lib.rsoptsc_nmf_lee_dev(d_V.ptr, n, int(rank), int(max_iter), _dp(basis) if want_factors else None,
                                       None, _ip(labels), C.byref(iters)))
    return dict(labels=labels, iters=int(iters.value), W=basis)


def timer_begin(name):
    _lib.check(_lib.load().rsoptsc_timer_begin(name.encode()))


def timer_end(name):
    _lib.check(_lib.load().rsoptsc_timer_end(name.encode()))


def phase_ms(name):
    return float(_lib.load().rsoptsc_phase_ms(name.encode()))


def kernel_launches():
    return int(_lib.load().rsoptsc_kernel_launches())


def timers(enable=True):
    _lib.load().rsoptsc_timers_enable(1 if enable else 0)
    _lib.load().rsoptsc_timers_reset()


def phase_report():
    """{phase: {ms, calls}} for every library phase timed since the last timers() reset."""
    lib = _lib.load()
    out = {}
    i = 0
    while True:
        name = lib.rsoptsc_phase_name(i)
        if name is None:
            break
        i += 1
        ms = lib.rsoptsc_phase_ms(name)
        calls = int(lib.rsoptsc_phase_calls(name))
        if calls > 0:
            out[name.decode()] = dict(ms=ms, calls=calls)
    return out
