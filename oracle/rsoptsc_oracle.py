"""CPU restatement (numpy/scipy, FP64) of RSoptSC's cell-by-cell hot path.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  The product never imports this.

Every function follows the reference line by line (file:line cited per function,
paths relative to /root/reference).  R is not installable in the build image, so the
restatement is pinned against the reference's own known answers instead
(tests/test_oracle_golden.py, fixture tests/golden/joost_fixture.npz extracted from
R/sysdata.rda by tests/golden/make_golden.py):

  PINNED
    * computeM            E = 16.7979 +- 1e-5 on joostTest$data/$D  (tests/testthat/testL2R2.R:4-15)
    * GetComponents       n_eigs == 6 on joostTest$S                (tests/testthat/testClusterNumber.R:4-7)
    * nmf lee/nndsvd      basis == joostTest$H (720x9) to ~1e-15, labels == joostTest$labels
                          (stored input/output pair in R/sysdata.rda; pins NNDSVD, update
                          order, eps, rescale and the connectivity stop rule)
    * CountClusters       upper_bound == 9 == ncol(joostTest$H) (consistency, not an assertion
                          of the reference)
    * GetSignalingPartners  P[[1]] == the hand-computed 3-cell answer of tests/testthat/testSignal.R:16-81
  PARITY UNPINNED (restated from the published algorithms of third-party packages that
  are not vendored in the reference: NMF (renozao/NMF@devel, >= 0.23), cluster >= 2.0.7-1,
  FNN >= 1.1.2.1):
    * cluster::pam label vectors themselves (BUILD/SWAP tie-breaking recalled from pam.c)
    * FNN::get.knnx tie order (only neighbour *sets* are compared)
    * SimilarityM end-to-end W (depends on Rtsne/umap RNG in the reference; here the
      embedding is an explicit input)
    * NMF maxIter cap (2000) is not exercised by the fixture
"""
import numpy as np
import scipy.linalg as sla

DBL_EPS = np.finfo(np.float64).eps


# --------------------------------------------------------------------------------------
# helpers mirroring base-R semantics
# --------------------------------------------------------------------------------------
def norm2(M, literal=True):
    """base::norm(M, "2"): largest singular value.  literal=True calls a values-only
    SVD like R does (La.svd -> dgesdd); literal=False uses the smaller Gram matrix
    (same value to ~1 ulp relative, much cheaper for m >> n)."""
    M = np.asarray(M, dtype=np.float64)
    if M.ndim == 1:
        return float(np.sqrt(np.dot(M, M)))
    if M.size == 0:
        return 0.0
    if literal:
        return float(sla.svdvals(M)[0])
    G = M.T @ M if M.shape[0] >= M.shape[1] else M @ M.T
    return float(np.sqrt(max(sla.eigvalsh(G, subset_by_index=[G.shape[0] - 1, G.shape[0] - 1])[0], 0.0)))


def prcomp_sdev2(T):
    """stats::prcomp(T)$sdev^2 with default center=TRUE, scale.=FALSE: T is obs x vars."""
    Tc = T - T.mean(axis=0, keepdims=True)
    d = sla.svdvals(Tc)
    return (d / np.sqrt(max(1, T.shape[0] - 1))) ** 2


def prcomp_scores(T, ncomp):
    """stats::prcomp(T, center=TRUE)$x[, 1:ncomp] (signs as returned by LAPACK)."""
    Tc = T - T.mean(axis=0, keepdims=True)
    U, d, _ = sla.svd(Tc, full_matrices=False, lapack_driver="gesdd")
    return U[:, :ncomp] * d[:ncomp]


# --------------------------------------------------------------------------------------
# SimilarityM  (R/SimilarityM.R:27-97)
# --------------------------------------------------------------------------------------
def normalize_columns(data):
    """R/SimilarityM.R:34-39: shift by global min, scale by global max, L2-normalise columns."""
    X = np.asarray(data, dtype=np.float64)
    X = X - X.min()
    X = X / X.max()
    nrm = np.sqrt((X * X).sum(axis=0))
    return X / nrm  # zero columns -> NaN exactly as R does (quirk Q10 upstream)


def choose_K(pca_eig):
    """R/SimilarityM.R:61-72 (K in [10, 30] from the cumulative PCA variance, first PC excluded)."""
    cc = np.cumsum(pca_eig[1:])
    dd = cc[1:] / np.sum(pca_eig[1:])
    K1 = int(np.sum(dd <= 0.3))
    if K1 <= 10:
        return 10
    if K1 >= 30:
        return 30
    return K1 + 1


NO_COMPS1 = 3  # R/SimilarityM.R:51-56 always yields 3 (quirk Q1)


def knn_indices(emb, K):
    """FNN::get.knnx(emb, emb, k=K)$nn.index restated as exact brute force
    (R/SimilarityM.R:76-77,83-84).  Self is among the K (quirk Q5).  Ties are broken
    by the lower index (FNN's own tie order is unpinned; only sets are compared).
    Squared distances are accumulated left to right without FMA so that a GPU kernel
    can reproduce them bit for bit.  Returns n x K int32, 0-based, ascending distance."""
    P = np.ascontiguousarray(np.asarray(emb, dtype=np.float64)[:, :NO_COMPS1])
    n = P.shape[0]
    out = np.empty((n, K), dtype=np.int32)
    for lo in range(0, n, 1024):
        hi = min(n, lo + 1024)
        d2 = None
        for c in range(P.shape[1]):
            diff = P[lo:hi, c][:, None] - P[:, c][None, :]
            sq = diff * diff
            d2 = sq if d2 is None else d2 + sq
        order = np.argsort(d2, axis=1, kind="stable")[:, :K]
        out[lo:hi] = order
    return out


def mask_from_idx(idx, n):
    """R/SimilarityM.R:75,90-92: D = ones; D[jj, IDX[jj,]] = 0.  Returns bool 'forbidden'."""
    forb = np.ones((n, n), dtype=bool)
    forb[np.arange(n)[:, None], idx] = False
    return forb


def computeM(forbidden, X, lam, literal=True, trace=None, maxiter=100):
    """R/SimilarityM.R:115-197, line by line.  `forbidden` is D > 0 (bool n x n).
    Returns dict(Z, E, iters, stop).  `trace` (list) receives per-iteration
    (iter, mu, rankJ, Err1, Err2, ||Z||_F)."""
    X = np.asarray(X, dtype=np.float64)
    m, n = X.shape
    Err = np.zeros((maxiter, 2))
    rho, mu, mumax, epsilon = 5.0, 1e-6, 1e6, 1e-5            # :118-121
    Z = np.zeros((n, n)); E = np.zeros((m, n))                  # :124-125
    Y1 = np.zeros((m, n)); Y2 = np.zeros((1, n)); Y3 = np.zeros((n, n))  # :126-128
    it = 0
    XXZ = X - X @ Z                                             # :130
    normX2 = None
    while True:
        it += 1
        if it >= maxiter:                                       # :134-137
            return dict(Z=Z, E=float(np.linalg.norm(XXZ, "fro")), iters=it, stop="max iter")
        mu = min(rho * mu, mumax)                               # :141
        A = Z - Y3 / mu
        U, s, Vt = sla.svd(A, full_matrices=True, lapack_driver="gesdd")  # :142
        shr = np.maximum(s - 1.0 / mu, 0.0)                     # :150
        J = (U * shr) @ Vt                                      # :155
        if it >= 3:                                             # :157-162
            if Err[it - 2, 0] >= Err[it - 3, 0] or norm2(XXZ, literal) <= epsilon:
                return dict(Z=Z, E=float(np.linalg.norm(XXZ, "fro")), iters=it, stop="error min")
        QQ = XXZ + Y1 / mu                                      # :164
        nq = np.sqrt((QQ * QQ).sum(axis=0))
        with np.errstate(invalid="ignore", divide="ignore"):
            E = ((nq - lam / mu) / nq) * QQ                     # :165-168
        E[:, nq <= lam / mu] = 0.0                              # :172-176
        if normX2 is None or literal:
            normX2 = norm2(X, literal)
        eta = normX2 ** 2 + np.sqrt(n) ** 2 + 1.0               # :178
        r = 1.0 - Z.sum(axis=0, keepdims=True) + Y2 / mu        # 1 x n
        H = -X.T @ (XXZ - E + Y1 / mu) - np.ones((n, 1)) @ r + (Z - J + Y3 / mu)  # :179
        Z = Z - H / eta                                         # :180
        Z[forbidden] = 0.0                                      # :181
        XXZ = X - X @ Z                                         # :182
        Y1 = Y1 + mu * (XXZ - E)                                # :185-187
        Y2 = Y2 + mu * (1.0 - Z.sum(axis=0, keepdims=True))
        Y3 = Y3 + mu * (Z - J)
        Err[it - 1, 0] = norm2(XXZ - E, literal)                # :189
        Err[it - 1, 1] = norm2(Z - J, literal)
        if trace is not None:
            trace.append((it, mu, int(np.sum(s > 1.0 / mu)), Err[it - 1, 0], Err[it - 1, 1],
                          float(np.linalg.norm(Z, "fro"))))
        if max(Err[it - 1]) <= epsilon:                         # :191-194
            return dict(Z=Z, E=float(np.linalg.norm(XXZ, "fro")), iters=it, stop="epsilon")


def threshold_symmetrise(Z):
    """R/SimilarityM.R:94-95."""
    Z = np.array(Z, dtype=np.float64, copy=True)
    Z[Z <= DBL_EPS] = 0.0
    return 0.5 * (np.abs(Z) + np.abs(Z.T))


def pca_embedding(X, ncomp=NO_COMPS1):
    """Deterministic stand-in for Rtsne/umap: first `ncomp` PCA score
    columns of t(X).  Sign of each column is LAPACK's; distances do not depend on it."""
    return prcomp_scores(np.asarray(X).T, ncomp)


def SimilarityM(lam, data, embedding=None, literal=True, K=None, trace=None):
    """R/SimilarityM.R:27-97 with the stochastic Rtsne/umap step replaced by an explicit
    `embedding` (n x >=3) argument (default: pca_embedding of the normalised X)."""
    X = normalize_columns(data)                                  # :34-39
    m, n = X.shape
    pca_eig = prcomp_sdev2(X.T)                                  # :41-49
    if K is None:
        K = choose_K(pca_eig)                                    # :61-72
    if embedding is None:
        embedding = pca_embedding(X)
    idx = knn_indices(embedding, K)                              # :76-77
    forb = mask_from_idx(idx, n)                                 # :90-92
    res = computeM(forb, X, lam, literal=literal, trace=trace)   # :93
    W = threshold_symmetrise(res["Z"])                           # :94-95
    return dict(W=W, E=res["E"], nl_embedding=embedding, K=K, idx=idx, Z=res["Z"],
                iters=res["iters"], stop=res["stop"], pca_eig=pca_eig)


# --------------------------------------------------------------------------------------
# ClusterCells -> NMF::nmf(method='lee', seed='nndsvd')   (R/Clustering.R:20-42)
# --------------------------------------------------------------------------------------
def nndsvd(V, k, svd=None):
    """NMF package seed 'nndsvd' (Boutsidis & Gallopoulos 2008, flag 0, densify none),
    call site R/Clustering.R:29-33.  Returns (W n x k, H k x n)."""
    V = np.asarray(V, dtype=np.float64)
    if svd is None:
        U, S, Vt = sla.svd(V, full_matrices=False, lapack_driver="gesdd")
    else:
        U, S, Vt = svd
    n, p = V.shape
    W = np.zeros((n, k)); H = np.zeros((k, p))
    W[:, 0] = np.sqrt(S[0]) * np.abs(U[:, 0])
    H[0, :] = np.sqrt(S[0]) * np.abs(Vt[0, :])
    for i in range(1, k):
        uu = U[:, i]; vv = Vt[i, :]
        uup = np.maximum(uu, 0); uun = np.maximum(-uu, 0)
        vvp = np.maximum(vv, 0); vvn = np.maximum(-vv, 0)
        n_uup = np.sqrt(uup @ uup); n_vvp = np.sqrt(vvp @ vvp)
        n_uun = np.sqrt(uun @ uun); n_vvn = np.sqrt(vvn @ vvn)
        termp = n_uup * n_vvp; termn = n_uun * n_vvn
        if termp >= termn:
            W[:, i] = np.sqrt(S[i] * termp) * uup / n_uup
            H[i, :] = np.sqrt(S[i] * termp) * vvp / n_vvp
        else:
            W[:, i] = np.sqrt(S[i] * termn) * uun / n_uun
            H[i, :] = np.sqrt(S[i] * termn) * vvn / n_vvn
    W[W < 1e-10] = 0.0
    H[H < 1e-10] = 0.0
    return W, H


def nmf_lee(V, k, max_iter=2000, eps=1e-9, stop_conv=40, check_interval=10, W0=None, H0=None):
    """NMF::nmf(V, rank=k, method='lee', seed='nndsvd') (call site R/Clustering.R:29-34):
    Lee-Seung Euclidean multiplicative updates, H then W, columns of W rescaled to sum 1
    every iteration, 'connectivity' stopping rule.  Returns dict(W, H, iters)."""
    V = np.asarray(V, dtype=np.float64)
    if W0 is None:
        W, H = nndsvd(V, k)
    else:
        W, H = np.array(W0, dtype=np.float64), np.array(H0, dtype=np.float64)
    prev = None
    inc = 0
    it = 0
    for it in range(1, max_iter + 1):
        H = np.maximum(H * (W.T @ V), eps) / ((W.T @ W) @ H + eps)
        W = np.maximum(W * (V @ H.T), eps) / (W @ (H @ H.T) + eps)
        W = W / W.sum(axis=0, keepdims=True)
        if it % check_interval == 0:
            idx = np.argmax(H, axis=0)          # which.max over each column of coef
            if prev is not None and np.array_equal(idx, prev):
                inc += 1
            else:
                prev = idx
                inc = 0
            if inc > stop_conv:
                break
    return dict(W=W, H=H, iters=it)


def labels_from_basis(Hb):
    """R/Clustering.R:36-37 (1-based; first maximum on ties, see quirk Q15)."""
    return (np.argmax(Hb, axis=1) + 1).astype(np.int32)


def ClusterCells(W, n_clusters=None, n_comp=3):
    """R/Clustering.R:20-42.  With n_clusters given the reference errors on an undefined
    variable (quirk Q11); the drop-in returns ensemble=None instead."""
    ensemble = None
    if n_clusters is None:
        ensemble = CountClusters(W, n_comp=n_comp)
        n_clusters = int(np.atleast_1d(ensemble["upper_bound"])[0])
    res = nmf_lee(W, n_clusters)
    return dict(H=res["W"], labels=labels_from_basis(res["W"]), ensemble=ensemble, iters=res["iters"])


# --------------------------------------------------------------------------------------
# CountClusters  (R/Clustering.R:61-175)
# --------------------------------------------------------------------------------------
def GetComponents(data, tol=0.01):
    """R/Clustering.R:96-105: eigenvalues of I - D^-1/2 S D^-1/2 (sqrtm + solve of a diagonal
    matrix == elementwise, quirk Q13), abs, ascending; count <= tol."""
    S = np.asarray(data, dtype=np.float64)
    n = S.shape[0]
    d = S @ np.ones(n)
    dd = 1.0 / np.sqrt(d)
    L = np.eye(n) - (dd[:, None] * S) * dd[None, :]
    ev = np.abs(sla.eigvalsh(0.5 * (L + L.T)))
    ev.sort()
    return dict(val=ev, n_eigs=int(np.sum(ev <= tol)))


def pam(P, k):
    """cluster::pam(P, k)$clustering (Kaufman & Rousseeuw BUILD + SWAP, variant 'original';
    call site R/Clustering.R:137).  Restated from the published algorithm / pam.c as
    recalled (parity unpinned): BUILD picks the later index on ties (`ammax <= beter`),
    SWAP takes the first strictly better (i, h) pair and stops when the best change is not
    below -16*eps*|sky|; objects go to the first nearest medoid; clusters are numbered by
    first appearance.  Returns (labels 1-based int32, medoid indices 0-based)."""
    P = np.asarray(P, dtype=np.float64)
    n = P.shape[0]
    dys = np.zeros((n, n))
    for c in range(P.shape[1]):
        diff = P[:, c][:, None] - P[:, c][None, :]
        dys += diff * diff
    dys = np.sqrt(dys)
    s = 1.1 * dys.max() + 1.0
    rep = np.zeros(n, dtype=bool)
    dysma = np.full(n, s)
    for _ in range(k):                                      # BUILD
        gain = np.maximum(dysma[None, :] - dys, 0.0).sum(axis=1)
        gain[rep] = -np.inf
        best = gain.max()
        nmax = int(np.flatnonzero(gain == best)[-1])        # `<=` -> last index wins
        rep[nmax] = True
        dysma = np.minimum(dysma, dys[nmax])
    sky = dysma.sum()
    while True:                                             # SWAP
        med = np.flatnonzero(rep)
        dm = dys[med]                                       # k x n
        order = np.sort(dm, axis=0)
        dysma = order[0]
        dysmb = order[1] if len(med) > 1 else np.full(n, s)
        non = np.flatnonzero(~rep)
        dh = dys[non]                                       # |non| x n
        T = np.empty((len(non), len(med)))                  # T[h, i] = sum_j C_jih  [KR p.104]
        for ii in range(len(med)):
            is_closest = dm[ii] == dysma                    # objects whose closest medoid is i
            small = np.minimum(dysmb[None, :], dh)
            c1 = np.where(is_closest[None, :], small - dysma[None, :], 0.0)
            c2 = np.where(~is_closest[None, :] & (dh < dysma[None, :]), dh - dysma[None, :], 0.0)
            T[:, ii] = (c1 + c2).sum(axis=1)
        # reference scans h outer, i inner and keeps the first strict minimum (`dzsky > dz`)
        flat = int(np.argmin(T))
        dzsky = float(T.flat[flat])
        hbest = int(non[flat // len(med)]); nbest = int(med[flat % len(med)])
        if dzsky < -16 * DBL_EPS * abs(sky):
            rep[hbest] = True
            rep[nbest] = False
            sky += dzsky
            continue
        break
    med = np.flatnonzero(rep)
    near = np.argmin(dys[med], axis=0)                      # first nearest medoid
    near[med] = np.arange(len(med))
    # number clusters by first appearance
    lab = np.zeros(n, dtype=np.int32)
    nxt = 0
    seen = {}
    for j in range(n):
        c = int(near[j])
        if c not in seen:
            nxt += 1
            seen[c] = nxt
        lab[j] = seen[c]
    return lab, med


def GetConsensus(clusters):
    """R/Clustering.R:164-175."""
    c = np.asarray(clusters)
    return (c[:, None] == c[None, :]).astype(np.float64)


def GetEnsemble(data, tol, tau, n_comp=3, rng=range(2, 21), return_labels=False):
    """R/Clustering.R:125-154."""
    S = np.asarray(data, dtype=np.float64)
    n = S.shape[1]
    scores = prcomp_scores(S.T, n_comp)                     # :133
    labs = [pam(scores, k)[0] for k in rng]                 # :136-138
    consen = np.zeros((n, n))
    for lab in labs:                                        # :143-147
        consen += GetConsensus(lab)
    consen[consen <= len(rng) * tau] = 0.0                  # :150
    consen = (consen + consen.T) / 2.0                      # :153
    if return_labels:
        return consen, np.stack(labs), scores
    return consen


def CountClusters(data, tol=0.01, rng=range(2, 21), n_comp=3):
    """R/Clustering.R:61-83."""
    rng = list(rng)
    solo = GetComponents(data, tol)["n_eigs"]               # :63
    tau = 0.3 if solo <= 5 else (0.4 if solo <= 10 else 0.5)  # :64-70
    cm = GetEnsemble(data, tol, tau, n_comp=n_comp, rng=rng)  # :71
    eigs = GetComponents(cm, tol)                           # :72
    gaps = np.diff(eigs["val"])                             # :74
    upper = np.flatnonzero(gaps == gaps.max()) + 1          # :75 (1-based, may be a vector)
    lower = int(np.sum(eigs["val"] <= tol))                 # :78
    return dict(upper_bound=upper if len(upper) > 1 else int(upper[0]), lower_bound=lower,
                eigs=eigs, tau=tau, solo_count=solo)


# --------------------------------------------------------------------------------------
# RepresentationMap dense / graph pieces  (R/RepresentationMap.R:37-39,57-66,89)
# parity unpinned: the reference's only test of this function is skipped (testPtime.R:4)
# --------------------------------------------------------------------------------------
def dist_flat(emb):
    """as.matrix(dist(flat_embedding, 'euclidean')) (:57): R accumulates dev*dev column by column."""
    P = np.asarray(emb, dtype=np.float64)
    acc = np.zeros((P.shape[0], P.shape[0]))
    for c in range(P.shape[1]):
        dev = P[:, c][:, None] - P[:, c][None, :]
        acc = acc + dev * dev
    return np.sqrt(acc)


def adjacency(S):
    """:59-62."""
    A = (np.asarray(S) > 0).astype(np.float64)
    np.fill_diagonal(A, 0.0)
    return A


def graph_components(adj):
    """igraph::components of the undirected graph (:64-66): membership numbered by lowest vertex, 1-based."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components
    A = sp.csr_matrix((np.asarray(adj) != 0) | (np.asarray(adj).T != 0))
    nc, lab = connected_components(A, directed=False)
    order = {}
    mem = np.empty(len(lab), dtype=np.int32)
    for v, c in enumerate(lab):
        if c not in order:
            order[c] = len(order) + 1
        mem[v] = order[c]
    return dict(membership=mem, csize=np.bincount(mem)[1:], no=int(nc))


# --------------------------------------------------------------------------------------
# Preprocessing upstream of the path  (R/DataSelection.R) -- parity unpinned (no reference test)
# --------------------------------------------------------------------------------------
def LogNormalize(M, scale=1e4, normalize=True):
    """R/DataSelection.R:117-133."""
    M = np.asarray(M, dtype=np.float64)
    MM = M.copy()
    if normalize:
        mx = M.max(axis=0)
        pos = mx > 0
        MM[:, pos] = M[:, pos] / mx[pos]
    return np.log10(scale * MM + 1.0)


def ScaleCenterData(M):
    """R/DataSelection.R:63-76 (sd = sample standard deviation; 0 -> 1)."""
    M = np.asarray(M, dtype=np.float64)
    means = M.mean(axis=1, keepdims=True)
    sds = M.std(axis=1, ddof=1, keepdims=True)
    sds[sds == 0] = 1.0
    return (M - means) / sds


def ApplyCountThreshold(M, countL=1, countU=3, X=3):
    """R/DataSelection.R:91-103 -> 1-based kept gene indices."""
    M = np.asarray(M)
    n = M.shape[1]
    upper = ~(((M >= countU).sum(axis=1) / n) >= (100 - X) * 0.01)
    lower = ((M >= countL).sum(axis=1) / n) >= X * 0.01
    return np.flatnonzero(upper & lower) + 1


def SelectData(M, gene_expression_threshold, n_features):
    """R/DataSelection.R:17-51 -> dict(gene_use 1-based, M_variable, max_component)."""
    M = np.asarray(M, dtype=np.float64)
    m, n = M.shape
    if gene_expression_threshold > 0:                                       # :22-30
        alpha = gene_expression_threshold / n
        nnz = (M != 0).sum(axis=1) / n
        gene_use = np.flatnonzero((nnz > alpha) & (nnz < 1 - alpha))
    else:
        gene_use = np.arange(m)
    Mv = M[gene_use]
    T = Mv.T                                                                # prcomp(t(M_variable)) :33
    Tc = T - T.mean(axis=0, keepdims=True)
    _, d, Vt = sla.svd(Tc, full_matrices=False, lapack_driver="gesdd")
    sdev = d / np.sqrt(max(1, n - 1))
    eigengaps = np.abs(sdev[1:-1] - sdev[2:])                               # :34
    max_component = int(np.flatnonzero(eigengaps == eigengaps.max())[0]) + 1 + 1   # :35 (1-based which + 1)
    rot = Vt.T                                                              # genes x components, sdev already descending
    max_coeffs = np.abs(rot[:, :max_component]).max(axis=1)                 # :39-41
    adj_n = min(Mv.shape[0], n_features)                                    # :45
    nth = np.sort(max_coeffs)[::-1][adj_n - 1]
    idx = np.flatnonzero(max_coeffs >= nth)                                 # :47
    return dict(gene_use=gene_use[idx] + 1, M_variable=Mv[idx], max_component=max_component)


# --------------------------------------------------------------------------------------
# GetSignalingPartners  (R/Signaling.R:1-60, 75-163, 187-309) -- pinned by the hand-computed
# known answer of tests/testthat/testSignal.R:16-81 (tests/test_oracle_golden.py)
# --------------------------------------------------------------------------------------
def signaling_pair(lig, rec, beta=None, gamma=None, normalize=True):
    """One ligand-receptor pair: updateFlat{Both,Up,Down,None} (:75-163) + row normalisation (:291-293).
    Scalar loops follow getAlpha/getK/getD literally (pure Python: small cases only)."""
    lig = np.asarray(lig, dtype=np.float64); rec = np.asarray(rec, dtype=np.float64)
    n = len(lig)
    P = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            a = 0.0 if (lig[i] == 0 or rec[j] == 0) else np.exp(-1.0 / (lig[i] * rec[j]))     # getAlpha
            v = a
            if beta is not None:
                b = beta[j]
                v = v * (0.0 if (a == 0 and b == 0) else a / (a + b))                         # getK
            if gamma is not None:
                g = gamma[j]
                v = v * (0.0 if (a == 0 and g == 0) else a / (a + g))                         # getD
            if beta is not None:
                v = v * beta[j]
            if gamma is not None:
                v = v * gamma[j]
            P[i, j] = v
    if normalize:
        with np.errstate(divide="ignore", invalid="ignore"):
            P = P / P.sum(axis=1, keepdims=True)
        P[np.isnan(P)] = 0.0
    return P


def GetSignalingPartners(data, gene_names, ligrec, rectarg):
    """R/Signaling.R:187-309 with the pathway given as (ligand, receptor) pairs and
    (receptor, target, direction) rows."""
    data = np.asarray(data, dtype=np.float64)
    row = {str(g).lower(): i for i, g in enumerate(gene_names)}
    P, LR = [], []
    for lig, rec in ligrec:
        lig, rec = str(lig).lower(), str(rec).lower()
        ups = sorted({str(t).lower() for r, t, d in rectarg if str(r).lower() == rec and d == "up"})
        downs = sorted({str(t).lower() for r, t, d in rectarg if str(r).lower() == rec and d == "down"})
        beta = gamma = None
        if ups:                                                                              # getBetaj :43-50
            s = data[[row[t] for t in ups]].sum(axis=0)
            beta = np.array([0.0 if v == 0 else np.exp(-len(ups) / v) for v in s])
        if downs:                                                                            # getGammaj :53-60
            s = data[[row[t] for t in downs]].sum(axis=0)
            gamma = np.array([0.0 if v == 0 else np.exp(-v / len(downs)) for v in s])
        P.append(signaling_pair(data[row[lig]], data[row[rec]], beta, gamma, normalize=True))
        LR.append((lig, rec))
    tot = sum(P)
    with np.errstate(divide="ignore", invalid="ignore"):
        tot = tot / tot.sum(axis=1, keepdims=True)
    tot[np.isnan(tot)] = 0.0
    return dict(P=P, P_tot=tot, LR=LR)


def graph_distances(adj):
    """igraph::distances (:89): unweighted all-pairs shortest paths, Inf when unreachable."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import shortest_path
    A = sp.csr_matrix(((np.asarray(adj) != 0) | (np.asarray(adj).T != 0)).astype(np.float64))
    return shortest_path(A, method="D", directed=False, unweighted=True)
