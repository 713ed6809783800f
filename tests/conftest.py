import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def joost():
    """The reference's joostTest fixture (tests/golden/make_golden.py)."""
    import scipy.sparse as sp
    f = np.load(os.path.join(ROOT, "tests", "golden", "joost_fixture.npz"))
    n = int(f["S_n"])
    data = sp.csc_matrix((f["data_values"].astype(np.float64), f["data_indices"], f["data_indptr"]),
                         shape=tuple(f["data_shape"]))
    S = sp.coo_matrix((f["S_vals"], (f["S_rows"], f["S_cols"])), shape=(n, n)).toarray()
    return dict(data=data, idx=f["idx"], S=S, H=f["H"], labels=f["labels"], n=n)


@pytest.fixture(scope="session")
def lib():
    from rsoptsc_b200 import _lib
    return _lib.load()


def small_problem(m, n, seed=0, K=None):
    """Seeded synthetic SimilarityM inputs + oracle-side preprocessing shared by the parity tests."""
    from oracle import rsoptsc_oracle as O
    from rsoptsc_b200.synthetic import synthetic_lognorm
    data, labels = synthetic_lognorm(m, n, seed)
    X = O.normalize_columns(data)
    eig = O.prcomp_sdev2(X.T)
    K = K or O.choose_K(eig)
    emb = O.pca_embedding(X)
    idx = O.knn_indices(emb, K)
    return dict(data=data, X=X, eig=eig, K=K, emb=emb, idx=idx, labels=labels)
