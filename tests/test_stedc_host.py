"""CPU check of the tridiagonal divide-and-conquer numerics (rsoptsc_b200/csrc/stedc_core.h).

The CUDA driver (stedc.cu) and this harness (tests/stedc_host_check.cpp) share the merge planning,
the secular-equation solver and the Loewner weights; here they run with host loops and are compared
with numpy.linalg.eigh on matrices chosen to stress deflation and clustered spectra.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("stedc") / "libstedc_host_check.so"
    src = os.path.join(HERE, "stedc_host_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", src, "-o", str(out)], check=True)
    L = ctypes.CDLL(str(out))
    dp = ctypes.POINTER(ctypes.c_double)
    L.stedc_host.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int, dp, dp, ctypes.POINTER(ctypes.c_int)]
    L.stedc_host.restype = ctypes.c_int
    return L


def run(lib, d, e, leaf):
    n = len(d)
    d = np.ascontiguousarray(d, dtype=np.float64)
    e = np.ascontiguousarray(e if n > 1 else np.zeros(1), dtype=np.float64)
    lam = np.zeros(n)
    Q = np.zeros((n, n), order="F")
    stats = (ctypes.c_int * 3)()
    dp = ctypes.POINTER(ctypes.c_double)
    rc = lib.stedc_host(n, d.ctypes.data_as(dp), e.ctypes.data_as(dp), leaf, lam.ctypes.data_as(dp),
                        Q.ctypes.data_as(dp), stats)
    assert rc == 0
    return lam, Q, list(stats)


def check(lib, d, e, leaf, tol=2e-14):
    n = len(d)
    T = np.diag(d) + (np.diag(e, 1) + np.diag(e, -1) if n > 1 else 0)
    lam, Q, stats = run(lib, d, e, leaf)
    ref = np.linalg.eigvalsh(T)
    scale = max(np.abs(ref).max(), 1e-300)
    assert np.all(np.diff(lam) >= 0)
    assert np.abs(lam - ref).max() <= tol * n * scale
    assert np.abs(Q.T @ Q - np.eye(n)).max() <= tol * n
    assert np.abs(T @ Q - Q * lam).max() <= tol * n * scale
    return stats


def test_random_tridiagonal(lib):
    rng = np.random.default_rng(0)
    for n, leaf in [(1, 8), (2, 8), (3, 2), (17, 4), (100, 16), (257, 16), (300, 25)]:
        d = rng.standard_normal(n)
        e = rng.standard_normal(max(n - 1, 0))
        check(lib, d, e, leaf)


def test_wilkinson_and_toeplitz(lib):
    m = 50
    d = np.abs(np.arange(-m, m + 1)).astype(float)  # W(2m+1)+ : pairs of eigenvalues agreeing to 1e-15
    e = np.ones(2 * m)
    check(lib, d, e, 8)
    n = 200
    check(lib, 2 * np.ones(n), -np.ones(n - 1), 16)  # 1-D Laplacian


def test_decoupled_and_graded(lib):
    rng = np.random.default_rng(1)
    n = 150
    d = rng.standard_normal(n)
    e = rng.standard_normal(n - 1)
    e[::7] = 0.0           # exact splits, some at leaf boundaries
    e[15] = 0.0
    e[16] = 1e-300
    check(lib, d, e, 16)
    d = 10.0 ** (-np.arange(n) / 10.0)  # graded over 15 orders of magnitude
    e = 0.5 * np.sqrt(d[:-1] * d[1:])
    check(lib, d, e, 16)
    check(lib, np.ones(n), np.zeros(n - 1), 16)  # identity: everything deflates
    check(lib, np.ones(n), 1e-9 * np.ones(n - 1), 16)


def test_gram_spectrum_with_deflation(lib):
    """Tridiagonal form of A^T A for a numerically low-rank A: the shape the SVT feeds the solver
    (a cluster of tiny eigenvalues next to well separated large ones)."""
    from scipy.linalg import hessenberg
    rng = np.random.default_rng(2)
    n = 240
    A = rng.standard_normal((n, 40)) @ rng.standard_normal((40, n)) + 1e-7 * rng.standard_normal((n, n))
    G = A.T @ A
    H = hessenberg(G)
    d = np.diag(H).copy()
    e = np.diag(H, -1).copy()
    stats = check(lib, d, e, 16, tol=5e-14)
    assert stats[1] > 0  # deflation did happen


def test_glued_clusters(lib):
    # many copies of the same small block, weakly glued: massive clustering
    blk_d = np.array([1.0, 2.0, 3.0, 2.0, 1.0])
    blk_e = np.array([0.5, 0.25, 0.25, 0.5])
    d = np.tile(blk_d, 30)
    e = np.concatenate([np.concatenate([blk_e, [1e-8]]) for _ in range(30)])[:-1]
    check(lib, d, e, 8)


def _tridiag_with_spectrum(lam, rng):
    """Tridiagonal matrix (via a random orthogonal similarity + Hessenberg reduction) with the given spectrum."""
    from scipy.linalg import hessenberg
    n = len(lam)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    H = hessenberg((Q * lam) @ Q.T)
    return np.diag(H).copy(), np.diag(H, -1).copy()


@pytest.mark.parametrize("kind", ["repeated", "geometric", "two_clusters", "tiny_plus_huge"])
def test_prescribed_spectra(lib, kind):
    rng = np.random.default_rng(11)
    n = 180
    if kind == "repeated":          # exactly repeated eigenvalues: deflation by Givens rotations
        lam = np.repeat(np.arange(1.0, 10.0), n // 9)
    elif kind == "geometric":       # 16 orders of magnitude
        lam = 10.0 ** np.linspace(-14, 2, n)
    elif kind == "two_clusters":    # two tight clusters, relative width 1e-13
        lam = np.concatenate([1.0 + 1e-13 * rng.standard_normal(n // 2), 3.0 + 1e-13 * rng.standard_normal(n // 2)])
    else:                           # the SVT's shape: many values at rounding level next to O(1) ones
        lam = np.concatenate([1e-17 * rng.random(n - 20), 1.0 + rng.random(20)])
    lam = np.sort(lam)
    d, e = _tridiag_with_spectrum(lam, rng)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    got, Q, stats = run(lib, d, e, 16)
    scale = np.abs(lam).max()
    assert np.abs(got - np.linalg.eigvalsh(T)).max() <= 5e-14 * n * scale
    assert np.abs(Q.T @ Q - np.eye(n)).max() <= 5e-14 * n
    assert np.abs(T @ Q - Q * got).max() <= 5e-14 * n * scale


def test_larger_random(lib):
    rng = np.random.default_rng(12)
    n = 700
    check(lib, rng.standard_normal(n), rng.standard_normal(n - 1), 32)
