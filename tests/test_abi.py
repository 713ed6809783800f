"""The C-ABI library loads and exports every symbol include/rsoptsc_b200.h declares; without
a GPU the compute entry points fail loudly (no CPU fallback).  CPU only."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rsoptsc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "rsoptsc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rsoptsc_[a-zA-Z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib):
    names = _declared()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), "missing export " + name
    assert sorted(_lib.SIGNATURES) == names, "ctypes table out of sync with the header"
    assert lib.rsoptsc_abi_version() == 3


def test_no_reference_to_oracle_in_product():
    # the product path must never import the oracle
    pkg = os.path.join(ROOT, "rsoptsc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".c")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def _has_gpu(lib):
    return lib.rsoptsc_init(-1) == 0


def test_fails_loudly_without_gpu(lib):
    if _has_gpu(lib):
        pytest.skip("GPU present")
    x = np.ones((4, 4), order="F")
    out = np.empty_like(x)
    st = lib.rsoptsc_normalize(x.ctypes.data_as(_lib.c_double_p), 4, 4, out.ctypes.data_as(_lib.c_double_p))
    assert st != 0
    assert b"no CPU fallback" in lib.rsoptsc_last_error()
    from rsoptsc_b200 import api
    with pytest.raises(_lib.RsoptscError):
        api.normalize(x)


def test_r_shim_compiles_against_stub(tmp_path):
    # R headers are absent here: compile-check the .Call shim against the minimal stub only
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    out = tmp_path / "shim.o"
    r = subprocess.run(["gcc", "-std=c99", "-DRSOPTSC_STUB_R", "-I", os.path.join(ROOT, "r"), "-I",
                        os.path.join(ROOT, "include"), "-c", os.path.join(ROOT, "r", "rsoptsc_shim.c"), "-o", str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # every C-ABI function the shim calls is declared in the header
    src = open(os.path.join(ROOT, "r", "rsoptsc_shim.c")).read()
    used = set(re.findall(r"\b(rsoptsc_(?!R_)[a-z_A-Z0-9]+)\s*\(", src))
    assert used and used <= set(_declared())


def test_host_tridiag_eig(lib):
    rng = np.random.default_rng(0)
    for k in (1, 2, 7, 64, 300):
        d = rng.normal(size=k)
        e = rng.normal(size=max(k - 1, 1))
        T = np.diag(d) + (np.diag(e[:k - 1], 1) + np.diag(e[:k - 1], -1) if k > 1 else 0)
        dd = d.copy()
        V = np.zeros((k, k), order="F")
        st = lib.rsoptsc_host_tridiag_eig(k, dd.ctypes.data_as(_lib.c_double_p), e.ctypes.data_as(_lib.c_double_p),
                                          V.ctypes.data_as(_lib.c_double_p))
        assert st == 0
        assert np.allclose(dd, np.linalg.eigvalsh(T), atol=1e-12 * max(1, k))
        assert np.abs(T @ V - V * dd).max() < 1e-12 * max(1, k)


def test_host_choose_K_matches_oracle(lib):
    from oracle import rsoptsc_oracle as O
    rng = np.random.default_rng(3)
    for _ in range(20):
        L = int(rng.integers(5, 400))
        eig = np.sort(rng.gamma(0.5, size=L))[::-1].copy()
        assert lib.rsoptsc_host_choose_K(eig.ctypes.data_as(_lib.c_double_p), L) == O.choose_K(eig)


def test_host_tridiag_bisect(lib):
    """The Sturm-count bisection behind the values-only eigensolver (csrc/tridiag_bisect.h, shared with the device
    kernel): random, decoupled (zero off-diagonals), clustered and graded tridiagonals vs LAPACK."""
    from scipy.linalg import eigvalsh_tridiagonal
    rng = np.random.default_rng(0)
    dp = _lib.c_double_p
    for k in (1, 2, 3, 17, 200, 1000):
        for kind in range(4):
            d = rng.standard_normal(k)
            e = rng.standard_normal(max(k - 1, 1))
            if kind == 1:
                e[::3] = 0
            elif kind == 2:
                d[:] = 1.0
                e[:] = 1e-9
            elif kind == 3:
                d = np.linspace(-k, k, k) ** 3
                e = np.ones(max(k - 1, 1))
            lam = np.zeros(k)
            assert lib.rsoptsc_host_tridiag_bisect(k, d.ctypes.data_as(dp), e.ctypes.data_as(dp), lam.ctypes.data_as(dp)) == 0
            ref = eigvalsh_tridiagonal(d, e[:k - 1]) if k > 1 else d.copy()
            assert np.all(np.diff(lam) >= 0)
            assert np.abs(lam - ref).max() <= 5e-14 * max(np.abs(ref).max(), 1e-300), (k, kind)


def test_r_wrapper_calls_match_the_shim_registration():
    """Every .Call("name", ...) in r/RSoptSC_b200.R names a routine registered in r/rsoptsc_shim.c with the same number
    of arguments (R checks this at run time with .registration = TRUE; R is not available here)."""
    shim = open(os.path.join(ROOT, "r", "rsoptsc_shim.c")).read()
    reg = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(rsoptsc_R_\w+)",\s*\(DL_FUNC\)&\w+,\s*(\d+)\}', shim)}
    assert len(reg) >= 15
    for name in reg:                                       # registered routines exist with that arity
        sig = re.search(r"SEXP %s\(([^)]*)\)" % name, shim)
        assert sig, name
        args = [a for a in sig.group(1).split(",") if a.strip() and a.strip() != "void"]
        assert len(args) == reg[name], name
    rsrc = open(os.path.join(ROOT, "r", "RSoptSC_b200.R")).read()
    rsrc = re.sub(r"#.*", "", rsrc)
    calls = list(re.finditer(r'\.Call\("(rsoptsc_R_\w+)"', rsrc))
    assert len(calls) >= 8
    for m in calls:
        name = m.group(1)
        assert name in reg, name
        # count top-level commas of this .Call(...)
        i, depth, commas = m.end(), 1, 0
        while depth > 0:
            ch = rsrc[i]
            depth += ch in "([{"
            depth -= ch in ")]}"
            commas += (ch == "," and depth == 1)
            i += 1
        assert commas == reg[name], (name, commas, reg[name])
