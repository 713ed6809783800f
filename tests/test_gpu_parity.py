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
