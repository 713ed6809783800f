"""Parity of the CUDA path (through the C ABI, rsoptsc_b200.api) against the CPU oracle and the
reference's golden fixture.  Run on the B200 box:  python -m pytest tests -m gpu -x -q

Tolerances (BASELINE.json north_star): Z within 1e-6 relative Frobenius error; neighbour index
sets, cluster labels and cluster counts identical; integer/index work bit-exact.
"""
import numpy as np
import pytest

from conftest import small_problem
from oracle import rsoptsc_oracle as O
from rsoptsc_b200 import api

pytestmark = pytest.mark.gpu


def relF(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


# ---- a1 / a2 / a3 ------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n", [(50, 33), (300, 257), (1000, 64)])
def test_normalize(m, n):
    p = small_problem(m, n, seed=2)
    X = api.normalize(p["data"])
    assert relF(X, p["X"]) < 1e-14
    assert np.allclose((X * X).sum(axis=0), 1.0, atol=1e-13)


@pytest.mark.parametrize("m,n", [(120, 300), (400, 150)])
def test_pca_spectrum_and_K(m, n):
    p = small_problem(m, n, seed=3)
    r = api.pca_spectrum(p["X"], ncomp=3)
    ref = p["eig"]
    assert np.allclose(r["eig"], ref, rtol=1e-8, atol=1e-12 * ref[0])
    assert r["K"] == O.choose_K(ref)
    # scores agree up to the sign of each component
    for c in range(3):
        a, b = r["emb"][:, c], p["emb"][:, c]
        assert min(relF(a, b), relF(-a, b)) < 1e-8


@pytest.mark.parametrize("n,K,dim", [(33, 5, 3), (1000, 10, 3), (2500, 30, 3), (700, 32, 2)])
def test_knn_index_sets_identical(n, K, dim):
    rng = np.random.default_rng(n)
    emb = rng.normal(size=(n, 3))
    emb[: n // 10] = emb[n // 10: 2 * (n // 10)][: n // 10]      # exact duplicates -> distance ties
    got = api.knn(emb, K, dim=dim)
    ref = O.knn_indices(emb[:, :dim] if dim < 3 else emb, K) if dim == 3 else None
    if ref is None:
        P = emb[:, :dim]
        d2 = ((P[:, None, :] - P[None, :, :]) ** 2)
        d2 = d2[:, :, 0] + d2[:, :, 1]
        ref = np.argsort(d2, axis=1, kind="stable")[:, :K].astype(np.int32)
    assert np.array_equal(got, ref)                               # bit-exact, order included


# ---- a4-a10 --------------------------------------------------------------------------------------
@pytest.mark.parametrize("m,n,lam", [(150, 200, 0.5), (400, 300, 0.05), (64, 500, 0.5)])
def test_computeM_matches_oracle(m, n, lam):
    p = small_problem(m, n, seed=m + n)
    ref = O.computeM(O.mask_from_idx(p["idx"], n), p["X"], lam, literal=False)
    got = api.computeM(None, p["X"], lam, idx=p["idx"])
    assert got["iters"] == ref["iters"] and got["stop"].endswith(ref["stop"].split()[-1])
    assert relF(got["Z"], ref["Z"]) < 1e-6
    assert abs(got["E"] - ref["E"]) < 1e-8 * max(1.0, ref["E"])
    # mask respected
    assert np.count_nonzero(got["Z"][O.mask_from_idx(p["idx"], n)]) == 0


def test_computeM_disconnected_support_matches_oracle():
    """kNN graph with several connected components (the library then runs one SVT per component on a
    block-diagonal dense state); the oracle works on the full n x n matrices."""
    from rsoptsc_b200.synthetic import synthetic_lognorm
    n, m, K = 420, 150, 10
    data, _ = synthetic_lognorm(m, n, seed=4)
    X = O.normalize_columns(data)
    rng = np.random.default_rng(4)
    sizes = [200, 150, 60, 9, 1]                      # includes a singleton component
    emb = np.concatenate([rng.normal(size=(s, 3)) + 1000.0 * g for g, s in enumerate(sizes)])
    perm = rng.permutation(n)                         # components interleaved in cell order
    emb = emb[perm]
    idx = np.empty((n, K), dtype=np.int32)
    d2 = ((emb[:, None, :] - emb[None, :, :]) ** 2).sum(-1)
    group = np.repeat(np.arange(len(sizes)), sizes)[perm]
    for i in range(n):
        same = np.flatnonzero(group == group[i])
        order = same[np.argsort(d2[i, same], kind="stable")]
        take = order[:K]
        idx[i] = np.resize(take, K)                   # small components: repeated (duplicate) neighbours
    ref = O.computeM(O.mask_from_idx(idx, n), X, 0.5, literal=False)
    got = api.computeM(None, X, 0.5, idx=idx)
    assert got["iters"] == ref["iters"]
    assert relF(got["Z"], ref["Z"]) < 1e-6
    assert abs(got["E"] - ref["E"]) < 1e-8 * ref["E"]


def test_computeM_default_tensor_core_path_vs_oracle():
    """One connected block of 1,800 cells: large enough that the DEFAULT backend sends the Gram and the
    reconstruction GEMMs of every SVT through the tcgen05 split-integer kernel (threshold 1,536).  The
    oracle takes about a minute here."""
    from rsoptsc_b200 import _lib
    from rsoptsc_b200.synthetic import synthetic_lognorm
    n, m, K = 1800, 250, 10
    data, _ = synthetic_lognorm(m, n, seed=9)
    X = O.normalize_columns(data)
    emb = np.random.default_rng(9).random((n, 3))            # uniform cloud -> a single kNN component
    idx = O.knn_indices(emb, K)
    assert api.graph_components(1.0 - O.mask_from_idx(idx, n))["no"] == 1
    assert _lib.load().rsoptsc_get_gemm_backend() == 1
    api.timers(True)
    got = api.computeM(None, X, 0.5, idx=idx)
    used = api.phase_report().get("ozaki_gemm", {"calls": 0})["calls"]
    api.timers(False)
    assert used > 0                                           # the tensor-core path really ran
    ref = O.computeM(O.mask_from_idx(idx, n), X, 0.5, literal=False)
    assert got["iters"] == ref["iters"]
    assert relF(got["Z"], ref["Z"]) < 1e-6
    assert abs(got["E"] - ref["E"]) < 1e-8 * ref["E"]


def test_computeM_dense_mask_signature():
    p = small_problem(100, 120, seed=5)
    n = 120
    D = O.mask_from_idx(p["idx"], n).astype(np.float64)
    a = api.computeM(D, p["X"], 0.5)
    b = api.computeM(None, p["X"], 0.5, idx=p["idx"])
    assert a["iters"] == b["iters"] and np.array_equal(a["Z"], b["Z"])


def test_computeM_joost_known_answer(joost):
    # tests/testthat/testL2R2.R:4-15
    X = O.normalize_columns(joost["data"].toarray())
    got = api.computeM(None, X, 0.5, idx=joost["idx"])
    assert abs(got["E"] - 16.7979) <= 1e-5
    assert got["iters"] == 27
    # Err1 trace
    assert abs(got["Err"][8, 0] - 0.38807218) < 1e-7
    assert abs(got["Err"][24, 0] - 0.017950117) < 1e-8
    assert np.count_nonzero(got["Z"]) == 7200


# ---- a11 -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 31, 32, 33, 500])
def test_threshold_symmetrise_bit_exact(n):
    rng = np.random.default_rng(n)
    Z = rng.normal(size=(n, n)) * (rng.random((n, n)) < 0.2)
    Z[rng.random((n, n)) < 0.05] = 1e-17                          # below the DBL_EPSILON threshold
    W = api.threshold_symmetrise(Z)
    assert np.array_equal(W, O.threshold_symmetrise(Z))
    assert np.array_equal(W, W.T) and W.min() >= 0


# ---- whole SimilarityM ------------------------------------------------------------------------------
def test_similarity_end_to_end():
    m, n = 300, 400
    p = small_problem(m, n, seed=11)
    ref = O.SimilarityM(0.5, p["data"], embedding=p["emb"], literal=False)
    got = api.SimilarityM(0.5, p["data"], embedding=p["emb"], sparse=True)
    assert got["K"] == ref["K"] and got["iters"] == ref["iters"]
    assert relF(got["W"], ref["W"]) < 1e-6
    assert abs(got["E"] - ref["E"]) < 1e-8 * ref["E"]
    assert np.array_equal(got["W"], got["W"].T)
    r, c, v = got["W_coo"]
    Ws = np.zeros((n, n))
    Ws[r, c] = v
    Ws[c, r] = v
    assert np.array_equal(Ws, got["W"])
    # in-library embedding: same neighbour sets as the oracle's PCA embedding -> same W
    got2 = api.SimilarityM(0.5, p["data"])
    assert relF(got2["W"], ref["W"]) < 1e-6


# ---- a12-a13 ----------------------------------------------------------------------------------------
def test_nmf_fixture(joost):
    r = api.nmf_lee(joost["S"], 9)
    assert r["iters"] == 510
    assert relF(r["W"], joost["H"]) < 1e-8
    assert np.array_equal(r["labels"], joost["labels"])
    assert np.allclose(r["W"].sum(axis=0), 1.0, atol=1e-12)


def test_nmf_matches_oracle_on_synthetic():
    p = small_problem(300, 400, seed=11)
    W = O.SimilarityM(0.5, p["data"], embedding=p["emb"], literal=False)["W"]
    ref = O.nmf_lee(W, 10)
    got = api.nmf_lee(W, 10)
    assert got["iters"] == ref["iters"]
    assert np.array_equal(got["labels"], O.labels_from_basis(ref["W"]))
    assert relF(got["W"], ref["W"]) < 1e-6


# ---- a14-a18 ----------------------------------------------------------------------------------------
def test_get_components_known_answer(joost):
    # tests/testthat/testClusterNumber.R:4-7
    r = api.GetComponents(joost["S"], tol=0.01)
    ref = O.GetComponents(joost["S"], tol=0.01)
    assert r["n_eigs"] == 6
    assert np.allclose(r["val"], ref["val"], atol=1e-12)


def test_consensus_bit_exact():
    rng = np.random.default_rng(0)
    n, R = 333, 19
    labs = np.stack([rng.integers(1, k + 1, size=n) for k in range(2, 2 + R)]).astype(np.int32)
    got = api.GetConsensusEnsemble(labs, 0.4)
    ref = np.zeros((n, n))
    for lab in labs:
        ref += O.GetConsensus(lab)
    ref[ref <= R * 0.4] = 0
    ref = (ref + ref.T) / 2
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("k", [2, 5, 12])
def test_pam_matches_oracle(k):
    rng = np.random.default_rng(k)
    centers = rng.normal(size=(6, 3)) * 4
    P = centers[rng.integers(0, 6, size=400)] + rng.normal(size=(400, 3))
    lab, med = api.pam(P, k)
    rlab, rmed = O.pam(P, k)
    assert np.array_equal(np.sort(med), np.sort(rmed))
    assert np.array_equal(lab, rlab)


def test_count_clusters_fixture(joost):
    r = api.CountClusters(joost["S"], n_comp=3)
    ref = O.CountClusters(joost["S"], n_comp=3)
    assert r["upper_bound"] == 9 == ref["upper_bound"]
    assert r["lower_bound"] == ref["lower_bound"]
    assert np.allclose(r["eigs"]["val"], ref["eigs"]["val"], atol=1e-10)


def test_cluster_cells_with_inferred_count(joost):
    r = api.ClusterCells(joost["S"], n_comp=3)
    assert r["ensemble"]["upper_bound"] == 9
    assert np.array_equal(r["labels"], joost["labels"])


# ---- full-size (BASELINE.json configs[1]) size-independent properties -------------------------------
def test_full_size_properties_config2():
    """10,000 cells x 2,000 genes: the oracle would need hours, so check properties the domain offers:
    mask respected, W symmetric non-negative with <= 2nK non-zeros, E == ||X - XZ||_F recomputed on
    the host from the returned sparse W support, NMF basis columns sum to 1, labels in range, and
    the planted populations are recovered (ARI)."""
    from rsoptsc_b200.synthetic import synthetic_lognorm
    m, n, k = 2000, 10000, 10
    data, planted = synthetic_lognorm(m, n, seed=0)
    r = api.SimilarityM(0.5, data, sparse=True)
    W = r["W"]
    K = r["K"]
    assert 10 <= K <= 30 and 3 <= r["iters"] < 100
    assert np.array_equal(W, W.T) and W.min() >= 0
    nnz = np.count_nonzero(W)
    assert n <= nnz <= 2 * n * K
    rows, cols, vals = r["W_coo"]
    assert len(vals) == (nnz + np.count_nonzero(np.diag(W))) // 2
    assert np.array_equal(W[rows, cols], vals)
    # neighbour structure: every non-zero of W is a kNN pair in one direction
    idx = api.knn(r["nl_embedding"], K)
    allowed = np.zeros((n, n), dtype=bool)
    allowed[np.arange(n)[:, None], idx] = True
    allowed |= allowed.T
    assert not np.any((W > 0) & ~allowed)
    # independent arithmetic for the dense contractions (cuBLAS FP64 instead of the tcgen05 split-integer
    # kernel): same iteration count and the same W to ~1e-12 -- a full-size check of the GEMM emulation
    from rsoptsc_b200 import _lib
    lib = _lib.load()
    prev = lib.rsoptsc_get_gemm_backend()
    lib.rsoptsc_set_gemm_backend(0)
    try:
        r0 = api.SimilarityM(0.5, data)
    finally:
        lib.rsoptsc_set_gemm_backend(prev)
    assert r0["iters"] == r["iters"] and r0["K"] == K
    assert relF(r0["W"], W) < 1e-9
    assert abs(r0["E"] - r["E"]) < 1e-9 * r["E"]
    c = api.ClusterCells(W, n_clusters=k)
    assert np.allclose(c["H"].sum(axis=0), 1.0, atol=1e-10)
    assert c["labels"].min() >= 1 and c["labels"].max() <= k
    assert np.array_equal(c["labels"], np.argmax(c["H"], axis=1) + 1)
    # agreement with the planted populations (pair-counting adjusted Rand index)
    from scipy.special import comb
    cont = np.zeros((k, planted.max() + 1))
    np.add.at(cont, (c["labels"] - 1, planted), 1)
    s_ij = comb(cont, 2).sum()
    s_a, s_b = comb(cont.sum(1), 2).sum(), comb(cont.sum(0), 2).sum()
    exp = s_a * s_b / comb(n, 2)
    ari = (s_ij - exp) / (0.5 * (s_a + s_b) - exp)
    assert ari > 0.5, ari


# ---- RepresentationMap pieces -----------------------------------------------------
def test_representation_map_pieces_fixture(joost):
    S = joost["S"]
    rng = np.random.default_rng(0)
    emb = rng.normal(size=(joost["n"], 2))
    got = api.RepresentationMap(emb, S)
    assert np.array_equal(got["dist_flat"], O.dist_flat(emb))               # bit-exact (no FMA, same order)
    assert np.array_equal(got["adj_matrix"], O.adjacency(S))
    ref_c = O.graph_components(O.adjacency(S))
    assert got["n_components"] == ref_c["no"] == 6                          # == n_eigs of GetComponents
    assert np.array_equal(got["components"]["membership"], ref_c["membership"])
    assert np.array_equal(got["components"]["csize"], ref_c["csize"])
    assert np.array_equal(got["dist_graph"], O.graph_distances(O.adjacency(S)))   # integers / Inf


@pytest.mark.parametrize("n", [1, 2, 257, 1500])
def test_graph_distances_random(n):
    rng = np.random.default_rng(n)
    A = (rng.random((n, n)) < min(1.0, 2.5 / max(n, 1))).astype(np.float64)   # sparse, directed input
    np.fill_diagonal(A, 0)
    got = api.graph_distances(A)
    ref = O.graph_distances(A)
    assert np.array_equal(got, ref)
    assert np.array_equal(got, got.T)
    c = api.graph_components(A)
    assert c["no"] == O.graph_components(A)["no"]
    assert np.array_equal(np.isfinite(got), c["membership"][:, None] == c["membership"][None, :])


# ---- preprocessing upstream of the path ---------------------------------------------
def test_preprocessing_matches_oracle():
    rng = np.random.default_rng(12)
    m, n = 700, 180
    counts = rng.poisson(rng.gamma(0.3, 2.0, size=(m, 1)), size=(m, n)).astype(float)
    counts[:5] = 0                      # never-expressed genes
    counts[5:8] = 7                     # constant genes (sd == 0)
    counts[:, 3] = 0                    # an empty cell (max == 0 branch of LogNormalize)
    assert np.allclose(api.LogNormalize(counts), O.LogNormalize(counts), rtol=1e-14, atol=0)
    assert np.allclose(api.LogNormalize(counts, normalize=False), O.LogNormalize(counts, normalize=False), rtol=1e-14)
    assert np.allclose(api.ScaleCenterData(counts), O.ScaleCenterData(counts), rtol=1e-12, atol=1e-13)
    assert np.array_equal(api.ApplyCountThreshold(counts, 1, 3, 3), O.ApplyCountThreshold(counts, 1, 3, 3))
    for thr, nf in [(3, 100), (0, 50), (10, 5000)]:
        got = api.SelectData(counts, thr, nf)
        ref = O.SelectData(counts, thr, nf)
        assert got["max_component"] == ref["max_component"]
        assert np.array_equal(got["gene_use"], ref["gene_use"])
        assert np.array_equal(got["M_variable"], ref["M_variable"])
    # cells > genes exercises the gene-side Gram branch of the loadings
    small = counts[8:60]
    got, ref = api.SelectData(small, 0, 10), O.SelectData(small, 0, 10)
    assert got["max_component"] == ref["max_component"] and np.array_equal(got["gene_use"], ref["gene_use"])


# ---- GetSignalingPartners core ------------------------------------------------------
def test_signaling_known_answer_and_oracle():
    from test_oracle_golden import (SIGNAL_EXPR, SIGNAL_GENES, SIGNAL_LIGREC, SIGNAL_RECTARG, signal_known_answer)
    got = api.GetSignalingPartners(SIGNAL_EXPR, SIGNAL_GENES, SIGNAL_LIGREC, SIGNAL_RECTARG)
    ref = O.GetSignalingPartners(SIGNAL_EXPR, SIGNAL_GENES, SIGNAL_LIGREC, SIGNAL_RECTARG)
    assert np.allclose(got["P"][0], signal_known_answer(), rtol=1e-13)      # tests/testthat/testSignal.R:16-81
    for a, b in zip(got["P"], ref["P"]):
        assert np.allclose(a, b, rtol=1e-13, atol=1e-300)
    assert np.allclose(got["P_tot"], ref["P_tot"], rtol=1e-13)
    # random expression with zeros (alpha = 0 branches, all-zero rows -> NaN -> 0), all four modes
    rng = np.random.default_rng(5)
    n = 150
    lig = rng.poisson(1.0, n).astype(float)
    rec = rng.poisson(1.5, n).astype(float)
    beta = np.where(rng.random(n) < 0.2, 0.0, rng.random(n))
    gamma = np.where(rng.random(n) < 0.2, 0.0, rng.random(n))
    for b, g in [(beta, gamma), (beta, None), (None, gamma), (None, None)]:
        for norm in (True, False):
            assert np.allclose(api.signaling_pair(lig, rec, b, g, normalize=norm),
                               O.signaling_pair(lig, rec, b, g, normalize=norm), rtol=1e-13, atol=1e-300)


def test_similarity_forced_K_and_lambda():
    """config 5 forces K = 30; the vignette uses lambda = 0.05."""
    p = small_problem(200, 260, seed=31)
    ref = O.SimilarityM(0.05, p["data"], embedding=p["emb"], literal=False, K=12)
    got = api.SimilarityM(0.05, p["data"], embedding=p["emb"], K=12)
    assert got["K"] == 12 and got["iters"] == ref["iters"]
    assert relF(got["W"], ref["W"]) < 1e-6
    assert (np.count_nonzero(got["W"], axis=1) >= 1).all()


def test_error_paths_return_status_not_crash():
    import ctypes as C
    from rsoptsc_b200 import _lib
    lib = _lib.init()
    x = np.ones((4, 4), order="F")
    out = np.empty_like(x)
    dp = lambda a: a.ctypes.data_as(_lib.c_double_p)  # noqa: E731
    assert lib.rsoptsc_normalize(dp(x), 0, 4, dp(out)) != 0                 # empty matrix
    assert b"empty" in lib.rsoptsc_last_error()
    idx = np.full((4, 2), 9, dtype=np.int32)                                  # neighbour index out of range
    E, it, st = C.c_double(0), C.c_int32(0), C.c_int32(0)
    rc = lib.rsoptsc_computeM(dp(x), 4, 4, idx.ctypes.data_as(_lib.c_int32_p), 2, 0.5, None, C.byref(E), C.byref(it),
                              C.byref(st), None)
    assert rc != 0 and b"out of range" in lib.rsoptsc_last_error()
    with pytest.raises(_lib.RsoptscError):
        api.nmf_lee(np.eye(5), 9)                                             # rank > n
    with pytest.raises(_lib.RsoptscError):
        api.knn(np.zeros((5, 3)), 40)                                         # K > 32
    # the library still works after errors (no poisoned state)
    assert np.allclose(api.normalize(np.arange(12.0).reshape(3, 4) + 1).sum(), api.normalize(np.arange(12.0).reshape(3, 4) + 1).sum())


def test_nmf_max_iter_cap_and_reinit(joost):
    """maxIter not a multiple of the stop-check interval (graph replay + tail iterations), then a
    finalize / re-init cycle of the library."""
    from rsoptsc_b200 import _lib
    ref = O.nmf_lee(joost["S"], 9, max_iter=25)
    got = api.nmf_lee(joost["S"], 9, max_iter=25)
    assert got["iters"] == ref["iters"] == 25
    assert relF(got["W"], ref["W"]) < 1e-10
    lib = _lib.load()
    assert lib.rsoptsc_finalize() == 0
    assert lib.rsoptsc_init(0) == 0
    again = api.nmf_lee(joost["S"], 9, max_iter=25)
    assert np.array_equal(again["W"], got["W"])


def test_large_block_eigensolver_path():
    """Blocks above cuSOLVER's 32,768-row limit go through cusolverMgSyevd; RSOPTSC_MG_MIN_N lowers the
    switch-over so that a small problem exercises that path (fresh process: the threshold is read once)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from conftest import small_problem\n"
        "from oracle import rsoptsc_oracle as O\n"
        "from rsoptsc_b200 import api\n"
        "p = small_problem(150, 260, seed=3)\n"
        "ref = O.computeM(O.mask_from_idx(p['idx'], 260), p['X'], 0.5, literal=False)\n"
        "got = api.computeM(None, p['X'], 0.5, idx=p['idx'])\n"
        "rel = np.linalg.norm(got['Z'] - ref['Z']) / np.linalg.norm(ref['Z'])\n"
        "assert got['iters'] == ref['iters'] and rel < 1e-6, (got['iters'], ref['iters'], rel)\n"
        "print('mg path ok', rel)\n" % (root, os.path.join(root, "tests")))
    env = dict(os.environ, RSOPTSC_MG_MIN_N="64")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mg path ok" in r.stdout, r.stdout + r.stderr


def test_repeated_calls_are_deterministic():
    """Caching allocator / persistent workspaces: same inputs -> bit-identical outputs, run to run."""
    p = small_problem(120, 300, seed=8)
    a = api.computeM(None, p["X"], 0.5, idx=p["idx"])
    api.nmf_lee(np.eye(40) + 0.01, 3)               # interleave a different workload
    b = api.computeM(None, p["X"], 0.5, idx=p["idx"])
    assert a["iters"] == b["iters"] and np.array_equal(a["Z"], b["Z"]) and a["E"] == b["E"]


def test_ozaki_gemm_matches_fp64():
    """tcgen05 split-integer GEMM vs numpy FP64, error measured against |A||B| (DGEMM-style bound)."""
    import ctypes as C
    from rsoptsc_b200 import _lib
    lib = _lib.init()
    rng = np.random.default_rng(0)
    dp = lambda a: a.ctypes.data_as(_lib.c_double_p)  # noqa: E731
    for op, (M, N, K) in [(0, (700, 700, 900)), (1, (640, 768, 1000)), (2, (513, 900, 515)), (3, (1300, 600, 2100))]:
        # wide dynamic range ACROSS output rows/columns (what the per-row power-of-two scaling of
        # the split-integer scheme absorbs); entries along the contraction index are O(1)
        sa, sb = np.exp(rng.normal(size=M) * 3), np.exp(rng.normal(size=N) * 2)
        if op in (0, 3):
            A = np.asfortranarray(rng.normal(size=(K, M)) * sa[None, :])
        else:
            A = np.asfortranarray(rng.normal(size=(M, K)) * sa[:, None])
        if op == 0:
            B = A
        elif op == 2:
            B = np.asfortranarray(rng.normal(size=(N, K)) * sb[:, None])
        else:
            B = np.asfortranarray(rng.normal(size=(K, N)) * sb[None, :])
        Nn = M if op == 0 else N
        out = np.zeros((M, Nn), order="F")
        _lib.check(lib.rsoptsc_debug_gemm(op, dp(A), dp(B), dp(out), M, Nn, K, 2, 1, None))
        ref = {0: lambda: A.T @ A, 1: lambda: A @ B, 2: lambda: A @ B.T, 3: lambda: A.T @ B}[op]()
        aa, bb = np.abs(A), np.abs(B)
        bound = {0: lambda: aa.T @ aa, 1: lambda: aa @ bb, 2: lambda: aa @ bb.T, 3: lambda: aa.T @ bb}[op]()
        err = np.abs(out - ref) / bound
        if op == 0:
            err = err[np.tril_indices(M)]
        assert err.max() < 1e-13, (op, err.max())


@pytest.mark.parametrize("backend", [0, 2])
def test_computeM_with_both_gemm_backends(backend):
    """The SVT contractions through the tcgen05 split-integer path (backend 2 forces it at this
    size) or cuBLAS FP64 (backend 0); both must reproduce the oracle."""
    from rsoptsc_b200 import _lib
    lib = _lib.init()
    p = small_problem(300, 700, seed=21)
    ref = O.computeM(O.mask_from_idx(p["idx"], 700), p["X"], 0.5, literal=False)
    prev = lib.rsoptsc_get_gemm_backend()
    lib.rsoptsc_set_gemm_backend(backend)
    try:
        got = api.computeM(None, p["X"], 0.5, idx=p["idx"])
    finally:
        lib.rsoptsc_set_gemm_backend(prev)
    assert got["iters"] == ref["iters"]
    assert relF(got["Z"], ref["Z"]) < 1e-6


# ---- native symmetric eigensolver (sytrd_kernel.cuh + stedc.cu) vs the library solver ---------------
def _sym_eig(G, solver):
    import ctypes as C
    from rsoptsc_b200 import _lib
    lib = _lib.load()
    dp = C.POINTER(C.c_double)
    n = G.shape[0]
    Gf = np.asfortranarray(G, dtype=np.float64)
    lam = np.zeros(n)
    V = np.zeros((n, n), order="F")
    _lib.check(lib.rsoptsc_debug_sym_eig(Gf.ctypes.data_as(dp), n, solver, lam.ctypes.data_as(dp),
                                         V.ctypes.data_as(dp), 1, None))
    return lam, V


def _gram_like(n, rng, K=30):
    Z = np.zeros((n, n))
    for i in range(n):
        Z[rng.choice(n, K, replace=False), i] = rng.random(K) / K
    return Z.T @ Z


@pytest.mark.parametrize("n,kind", [(65, "random"), (257, "random"), (700, "gram"), (1600, "gram"), (2300, "lowrank")])
def test_native_eigensolver(n, kind):
    """Own tridiagonalisation + tridiagonal divide and conquer + library back-transformation: eigenvalues
    agree with cuSOLVER to rounding, eigenvectors orthonormal, residual at rounding level."""
    rng = np.random.default_rng(n)
    if kind == "random":
        A = rng.standard_normal((n, n))
        G = (A + A.T) / 2
    elif kind == "gram":
        G = _gram_like(n, rng)
    else:  # numerically rank 40: one big cluster of tiny eigenvalues (heavy deflation)
        B = rng.standard_normal((n, 40))
        G = B @ B.T + 1e-10 * np.eye(n)
    l0, _ = _sym_eig(G, 0)
    l1, V1 = _sym_eig(G, 1)
    scale = np.abs(l0).max()
    assert np.all(np.diff(l1) >= 0)
    assert np.abs(l1 - l0).max() <= 1e-13 * n * scale
    assert np.abs(V1.T @ V1 - np.eye(n)).max() <= 1e-13 * n
    assert np.abs(G @ V1 - V1 * l1).max() <= 1e-13 * n * scale


def test_computeM_with_native_eigensolver():
    """The whole ADMM with every SVT through the native eigensolver (RSOPTSC_EIG=native lowers its size
    threshold to 256 rows; fresh process: the switch is read once)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from conftest import small_problem\n"
        "from oracle import rsoptsc_oracle as O\n"
        "from rsoptsc_b200 import api\n"
        "p = small_problem(200, 600, seed=5)\n"
        "ref = O.computeM(O.mask_from_idx(p['idx'], 600), p['X'], 0.5, literal=False)\n"
        "api.timers(True)\n"
        "got = api.computeM(None, p['X'], 0.5, idx=p['idx'])\n"
        "rep = api.phase_report()\n"
        "rel = np.linalg.norm(got['Z'] - ref['Z']) / np.linalg.norm(ref['Z'])\n"
        "assert 'eig_sytrd' in rep and rep['eig_sytrd']['calls'] > 0, rep\n"
        "assert got['iters'] == ref['iters'] and rel < 1e-6, (got['iters'], ref['iters'], rel)\n"
        "import ctypes as C\n"
        "from rsoptsc_b200 import _lib\n"
        "rng = np.random.default_rng(7); n = 1700\n"
        "B = rng.standard_normal((n, n)); G = np.asfortranarray(B @ B.T / n)\n"
        "lam = np.zeros(n); V = np.zeros((n, n), order='F'); dp = C.POINTER(C.c_double)\n"
        "_lib.check(_lib.load().rsoptsc_debug_sym_eig(G.ctypes.data_as(dp), n, 1, lam.ctypes.data_as(dp), "
        "V.ctypes.data_as(dp), 1, None))\n"
        "assert np.abs(G @ V - V * lam).max() < 1e-12 * lam.max() and np.abs(V.T @ V - np.eye(n)).max() < 1e-11\n"
        "print('native eig path ok', rel)\n" % (root, os.path.join(root, "tests")))
    # ... and the own blocked back-transformation instead of cusolverDnDormtr (the default below 32,768 rows)
    env = dict(os.environ, RSOPTSC_EIG="native", RSOPTSC_BACKTRANSFORM="native")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "native eig path ok" in r.stdout, r.stdout + r.stderr


# ---- values-only eigensolver (a2, a14): own tridiagonalisation + bisection on Sturm counts -----------
@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 3, 5, 31, 64, 65, 130, 700, 2300])
def test_values_only_eigensolver(n):
    rng = np.random.default_rng(100 + n)
    if n % 2:
        A = rng.standard_normal((n, n))
        G = (A + A.T) / 2
    else:  # Laplacian-like: eigenvalues in [0, 2] with a cluster at 1
        B = rng.random((n, max(n // 3, 1)))
        S = B @ B.T
        dd = 1.0 / np.sqrt(S.sum(axis=1))
        G = np.eye(n) - dd[:, None] * S * dd[None, :]
    lam, _ = _sym_eig(G, 3)
    ref = np.linalg.eigvalsh(G)
    assert np.all(np.diff(lam) >= 0)
    assert np.abs(lam - ref).max() <= 1e-13 * max(n, 8) * max(np.abs(ref).max(), 1e-300)


def _count_clusters_both_routes(S, n_comp=3):
    """CountClusters through the Lanczos count + signature-compressed consensus spectrum (default) and through the
    literal dense route (RSOPTSC_EIGENGAP_DENSE=1: every spectrum from a dense values-only eigensolve)."""
    import os
    new = api.CountClusters(S, n_comp=n_comp)
    os.environ["RSOPTSC_EIGENGAP_DENSE"] = "1"
    try:
        dense = api.CountClusters(S, n_comp=n_comp)
    finally:
        del os.environ["RSOPTSC_EIGENGAP_DENSE"]
    return new, dense


@pytest.mark.gpu
def test_count_clusters_lanczos_route_equals_dense_route_fixture(joost):
    new, dense = _count_clusters_both_routes(joost["S"])
    assert new["upper_bound"] == dense["upper_bound"] == 9
    assert new["lower_bound"] == dense["lower_bound"]
    assert np.allclose(new["eigs"]["val"], dense["eigs"]["val"], atol=1e-10)


@pytest.mark.gpu
def test_count_clusters_lanczos_route_large_component():
    """A similarity whose support has one component of > 1,024 cells: the n_eigs pre-estimate (R/Clustering.R:63)
    comes from the certified Lanczos count on the sparse scaled similarity, the consensus spectrum from the
    label signatures; both must reproduce the dense route."""
    g = _golden_scale("cut3000")
    n = int(g["cells"])
    W = _W_from_support(g["idx"], g["Zs"], n)
    api.timers(True)
    new, dense = _count_clusters_both_routes(W)
    rep = api.phase_report()
    api.timers(False)
    assert rep.get("lanczos_count", {}).get("calls", 0) > 0 and rep.get("consensus_spectrum", {}).get("calls", 0) > 0
    assert new["upper_bound"] == dense["upper_bound"]
    assert new["lower_bound"] == dense["lower_bound"]
    assert np.allclose(new["eigs"]["val"], dense["eigs"]["val"], atol=1e-10)
    # and the oracle's own GetComponents count on the same similarity
    assert api.GetComponents(W)["n_eigs"] == O.GetComponents(W)["n_eigs"]


# ---- parity at scale: golden vectors of the CPU oracle (tests/golden/make_golden_scale.py) -------------
def _golden_scale(name):
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scale_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("golden fixture %s not generated" % name)
    return np.load(path)


def _W_from_support(idx, Zs, n):
    Z = np.zeros((n, n))
    Z[np.arange(n)[:, None], idx] = Zs
    Z[Z <= 2.220446049250313e-16] = 0
    return (np.abs(Z) + np.abs(Z.T)) / 2


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["cut3000", "big6200"])
def test_similarity_matches_oracle_golden_at_scale(case):
    """Default path, no environment overrides (tcgen05 GEMMs; the 6,200-cell case has one connected block above
    4,000 rows, so its SVTs go through the native eigensolver): Z on its support within 1e-6 relative
    Frobenius of the oracle, same iteration count and stop reason, Err1 trace, K, neighbour table, labels."""
    from rsoptsc_b200.synthetic import synthetic_lognorm
    g = _golden_scale(case)
    n = int(g["cells"])
    data, _ = synthetic_lognorm(int(g["genes"]), int(g["gen_cells"]), seed=int(g["seed"]))
    data = np.asfortranarray(data[:, :n])
    api.timers(True)
    r = api.SimilarityM(float(g["lam"]), data)
    rep = api.phase_report()
    api.timers(False)
    assert r["K"] == int(g["K"])
    assert r["iters"] == int(g["iters"])
    assert abs(r["E"] - float(g["E"])) <= 1e-9 * float(g["E"])
    idx = g["idx"]
    Wref = _W_from_support(idx, g["Zs"], n)
    assert relF(r["W"], Wref) < 1e-6
    assert np.count_nonzero(r["W"]) == int(g["W_nnz"])
    assert rep.get("ozaki_gemm", {}).get("calls", 0) > 0
    if int(g["components"][0]) >= 4000:
        assert rep.get("sytrd_panel", {}).get("calls", 0) > 0      # native eigensolver on the default path
    c = api.ClusterCells(r["W"], n_clusters=int(g["rank"]))
    assert c["iters"] == int(g["nmf_iters"])
    assert np.array_equal(np.asarray(c["labels"]), g["labels"])


@pytest.mark.gpu
def test_computeM_err_trace_matches_oracle_golden():
    """computeM on the oracle's own neighbour table: the per-iteration Err1 = ||XXZ - E||_2 trace
    (R/SimilarityM.R:189) that drives the stop rule, and Z on the support."""
    from rsoptsc_b200.synthetic import synthetic_lognorm
    g = _golden_scale("cut3000")
    n = int(g["cells"])
    data, _ = synthetic_lognorm(int(g["genes"]), int(g["gen_cells"]), seed=int(g["seed"]))
    X = api.normalize(np.asfortranarray(data[:, :n]))
    r = api.computeM(None, X, float(g["lam"]), idx=g["idx"])
    assert r["iters"] == int(g["iters"])
    k = len(g["err1"])
    assert np.allclose(r["Err"][:k, 0], g["err1"], rtol=1e-8, atol=1e-12)
    Zs = r["Z"][np.arange(n)[:, None], g["idx"]]
    assert relF(Zs, g["Zs"]) < 1e-6


# ---- boundary completeness ----------------------------------------------------------
@pytest.mark.gpu
def test_similarity_with_caller_neighbour_table():
    """use_umap_indices = TRUE (R/SimilarityM.R:85-87): the caller's IDX replaces the kNN search.  The table here is
    NOT a kNN table of any embedding (random neighbours, K = 12), so only the supplied one can give the answer."""
    p = small_problem(120, 260, seed=17)
    n = p["data"].shape[1]
    rng = np.random.default_rng(4)
    idx = np.stack([np.concatenate(([j], rng.choice(np.delete(np.arange(n), j), 11, replace=False))) for j in range(n)])
    ref = O.computeM(O.mask_from_idx(idx, n), p["X"], 0.5, literal=False)
    Wref = O.threshold_symmetrise(ref["Z"])
    got = api.SimilarityM(0.5, p["data"], knn_indices=idx)
    assert got["iters"] == ref["iters"] and got["K"] == 12
    assert abs(got["E"] - ref["E"]) <= 1e-9 * ref["E"]
    assert relF(got["W"], Wref) < 1e-6
    with pytest.raises(Exception):
        api.SimilarityM(0.5, p["data"], knn_indices=idx + n)      # out-of-range indices are rejected


@pytest.mark.gpu
def test_progress_callback_reports_err_and_interrupts():
    """rsoptsc_set_progress: one call per ADMM iteration carrying Err[iter, 1] (the value R prints at
    R/SimilarityM.R:191), NMF stop checks every 10 iterations, and a non-zero return stops the run (status 9)."""
    from rsoptsc_b200 import _lib
    p = small_problem(100, 220, seed=23)
    seen = []
    api.set_progress(lambda stage, it, v: seen.append((stage, it, v)) and False)
    try:
        r = api.computeM(None, p["X"], 0.5, idx=p["idx"])
        admm = [s for s in seen if s[0] == 0]
        assert len(admm) in (r["iters"] - 1, r["iters"])                  # an "error min" stop returns before :190
        assert [s[1] for s in admm] == list(range(1, len(admm) + 1))
        assert np.allclose([s[2] for s in admm], r["Err"][:len(admm), 0], rtol=0, atol=0)
        seen.clear()
        W = api.threshold_symmetrise(r["Z"])
        c = api.nmf_lee(W, 5)
        nmf = [s for s in seen if s[0] == 1]
        assert [s[1] for s in nmf] == list(range(10, c["iters"] + 1, 10))
        api.set_progress(lambda stage, it, v: stage == 0 and it == 3)
        with pytest.raises(_lib.RsoptscError, match="error 9"):
            api.computeM(None, p["X"], 0.5, idx=p["idx"])
    finally:
        api.set_progress(None)
    again = api.computeM(None, p["X"], 0.5, idx=p["idx"])                   # the library is usable after an interrupt
    assert again["iters"] == r["iters"] and relF(again["Z"], r["Z"]) == 0.0


@pytest.mark.gpu
def test_get_ensemble_and_consensus_fixture(joost):
    """GetEnsemble / GetConsensus (R/Clustering.R:125-175) on the reference's own similarity matrix."""
    S = joost["S"]
    ref = O.GetEnsemble(S, tol=0.01, tau=0.4, n_comp=3)
    got = api.GetEnsemble(S, tol=0.01, n_comp=3, tau=0.4)
    assert np.array_equal(got, ref)
    lab = np.asarray(joost["labels"], dtype=np.int32)
    assert np.array_equal(api.GetConsensus(lab), O.GetConsensus(lab))


@pytest.mark.gpu
@pytest.mark.parametrize("backend", [0, 2])
def test_gemm_propagates_nonfinite_operands(backend):
    """a NaN / Inf operand entry must poison the outputs it takes part in on both GEMM backends (the
    split-integer kernel used to clamp it to a finite slice)."""
    import ctypes as C
    from rsoptsc_b200 import _lib
    lib = _lib.init()
    rng = np.random.default_rng(0)
    M = N = K = 300
    A = np.asfortranarray(rng.standard_normal((M, K)))
    B = np.asfortranarray(rng.standard_normal((K, N)))
    A[5, 7] = np.nan
    B[11, 3] = np.inf
    Cc = np.zeros((M, N), order="F")
    dp = _lib.c_double_p
    _lib.check(lib.rsoptsc_debug_gemm(1, A.ctypes.data_as(dp), B.ctypes.data_as(dp), Cc.ctypes.data_as(dp), M, N, K,
                                      backend, 1, None))
    assert np.isnan(Cc[5, :]).all()
    assert not np.isfinite(Cc[:, 3]).any()
    ok = np.ones((M, N), dtype=bool)
    ok[5, :] = False
    ok[:, 3] = False
    assert np.isfinite(Cc[ok]).all()
    ref = np.delete(np.delete(A, 5, axis=0) @ np.delete(B, 3, axis=1), [], axis=0)
    got = np.delete(np.delete(Cc, 5, axis=0), 3, axis=1)
    assert relF(got, ref) < 1e-12


@pytest.mark.gpu
def test_nmf_rank_sweep_equals_separate_calls(joost):
    """configs[3]: the batched rank sweep gives, for every k, exactly the labels and the iteration count of a separate
    nmf call (fixture similarity, k = 5..12; k = 9 is the reference's own stored answer: 510 iterations)."""
    S = joost["S"]
    sw = api.nmf_sweep(S, ranks=(5, 12))
    assert sw["ranks"] == list(range(5, 13))
    for p, k in enumerate(sw["ranks"]):
        one = api.nmf_lee(S, k)
        assert sw["iters"][p] == one["iters"], (k, sw["iters"][p], one["iters"])
        assert np.array_equal(sw["labels"][p], one["labels"]), k
    assert sw["iters"][sw["ranks"].index(9)] == 510
    assert np.array_equal(sw["labels"][sw["ranks"].index(9)], np.asarray(joost["labels"]))
    # a capped sweep stops every problem at the cap
    capped = api.nmf_sweep(S, ranks=(5, 6), max_iter=25)
    assert list(capped["iters"]) == [25, 25]
    assert np.array_equal(capped["labels"][0], api.nmf_lee(S, 5, max_iter=25)["labels"])
