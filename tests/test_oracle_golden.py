"""Pin the CPU oracle against the reference's own known answers.
CPU only; the slowest case (computeM on the 24,470 x 720 Joost matrix) takes ~45 s."""
import numpy as np

from oracle import rsoptsc_oracle as O


def test_get_components_known_answer(joost):
    # tests/testthat/testClusterNumber.R:4-7
    r = O.GetComponents(joost["S"], tol=0.01)
    assert r["n_eigs"] == 6
    assert np.all(np.diff(r["val"]) >= 0)


def test_nmf_reproduces_fixture_basis_and_labels(joost):
    # stored pair joostTest$S -> joostTest$H / $labels (R/sysdata.rda)
    r = O.nmf_lee(joost["S"], 9)
    assert r["iters"] == 510
    rel = np.linalg.norm(r["W"] - joost["H"]) / np.linalg.norm(joost["H"])
    assert rel < 1e-12
    assert np.array_equal(O.labels_from_basis(r["W"]), joost["labels"])
    assert np.allclose(r["W"].sum(axis=0), 1.0, atol=1e-12)


def test_count_clusters_matches_fixture_rank(joost):
    # rank of joostTest$H is 9; CountClusters(S) must give the same upper bound
    r = O.CountClusters(joost["S"], n_comp=3)
    assert r["upper_bound"] == 9 and r["solo_count"] == 6 and r["tau"] == 0.4


def test_computeM_known_answer(joost):
    # tests/testthat/testL2R2.R:4-15  E = 16.7979 +- 1e-5
    X = O.normalize_columns(joost["data"].toarray())
    forb = O.mask_from_idx(joost["idx"], joost["n"])
    trace = []
    r = O.computeM(forb, X, 0.5, literal=False, trace=trace)
    assert abs(r["E"] - 16.7979) <= 1e-5
    assert r["iters"] == 27 and r["stop"] == "error min"
    # SURVEY.md appendix B trace
    assert abs(trace[8][3] - 0.38807218) < 1e-7 and trace[9][2] == 59
    Z = r["Z"]
    assert np.count_nonzero(Z) == 7200 and Z.min() >= 0
    W = O.threshold_symmetrise(Z)
    assert np.count_nonzero(W) == 8622 and abs(W.sum() - 449.352) < 1e-3


# inst/extdata/test_gene_expression.txt, test_ligrec.txt, test_rectarg_updown.txt (3 cells)
SIGNAL_GENES = ["lig1", "lig2", "lig3", "rec1", "rec2", "targ_rec1_1", "targ_rec1_2", "targ_rec1_3", "targ_rec2_1"]
SIGNAL_EXPR = np.array([[1, 2, 3], [1, 2, 1], [1, 0, 1], [2, 1, 2], [1, 3, 1], [0, 1, 0], [3, 3, 3], [8, 9, 8],
                        [1, 4, 1]], dtype=np.float64)
SIGNAL_LIGREC = [("lig1", "rec1"), ("lig2", "rec1"), ("lig2", "rec2"), ("lig3", "rec2")]
SIGNAL_RECTARG = [("rec1", "targ_rec1_1", "up"), ("rec1", "targ_rec1_2", "up"), ("rec1", "targ_rec1_3", "down"),
                  ("rec2", "targ_rec2_1", "down")]


def signal_known_answer():
    """The hand-computed pLig1Rec1 of tests/testthat/testSignal.R:26-74."""
    e = np.exp
    L = [1.0, 2.0, 3.0]          # lig1 in cells 1..3
    R = [2.0, 1.0, 2.0]          # rec1
    up = [e(-2 / 3), e(-2 / 4), e(-2 / 3)]
    down = [e(-8 / 1), e(-9 / 1), e(-8 / 1)]
    P = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            a = e(-1 / (L[i] * R[j]))
            P[i, j] = a * (a / (a + up[j]) * up[j]) * (a / (a + down[j]) * down[j])
    return P / P.sum(axis=1, keepdims=True)


def test_signaling_known_answer():
    # tests/testthat/testSignal.R:16-81
    r = O.GetSignalingPartners(SIGNAL_EXPR, SIGNAL_GENES, SIGNAL_LIGREC, SIGNAL_RECTARG)
    assert np.allclose(r["P"][0], signal_known_answer(), rtol=1e-12, atol=0)
    for p in r["P"]:
        sums = p.sum(axis=1)
        assert np.all((np.abs(sums - 1) < 1e-12) | (sums == 0))
    assert np.allclose(r["P_tot"].sum(axis=1), 1.0)


def test_knn_self_included_and_mask():
    rng = np.random.default_rng(1)
    emb = rng.normal(size=(200, 3))
    idx = O.knn_indices(emb, 10)
    assert np.array_equal(idx[:, 0], np.arange(200))          # self is the nearest (distance 0)
    forb = O.mask_from_idx(idx, 200)
    assert np.all((~forb).sum(axis=1) == 10) and not forb.diagonal().any()


def test_choose_K_bounds():
    assert O.choose_K(np.array([5.0, 1.0] + [1e-3] * 50)) == 10
    flat = np.ones(400)
    assert O.choose_K(flat) == 30


def test_literal_and_restructured_oracle_agree():
    """literal=True mirrors R line by line (full-SVD spectral norms every iteration, SVT before the stop test,
    R/SimilarityM.R:142-163); literal=False is the restructured order every parity assertion uses (norms by
    eigvalsh of the small Gram, stop test hoisted).  Same Z, E, iteration count and stop reason."""
    from conftest import small_problem
    p = small_problem(200, 300, seed=4)
    forb = O.mask_from_idx(p["idx"], 300)
    a = O.computeM(forb, p["X"], 0.5, literal=True)
    b = O.computeM(forb, p["X"], 0.5, literal=False)
    assert a["iters"] == b["iters"] and a["stop"] == b["stop"]
    assert abs(a["E"] - b["E"]) <= 1e-10 * abs(a["E"])
    assert np.linalg.norm(a["Z"] - b["Z"]) <= 1e-9 * np.linalg.norm(a["Z"])
