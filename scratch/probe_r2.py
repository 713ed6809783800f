"""Round-2 probes on one B200:
    python scratch/probe_r2.py syr2k      rank-128 trailing update: tcgen05 split-integer kernel vs cublasDsyr2k
    python scratch/probe_r2.py panel      panel-shaped GEMMs of the back-transformation (K = 1024 / M = 1024)
    python scratch/probe_r2.py eig N...   native eigensolver phases at the given sizes
"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from rsoptsc_b200 import _lib, api  # noqa: E402

lib = _lib.init(0)
dp = lambda a: a.ctypes.data_as(_lib.c_double_p)  # noqa: E731


def gemm(op, A, B, Cc, M, N, K, backend, reps):
    ms = C.c_double(0)
    st = lib.rsoptsc_debug_gemm(op, dp(A), dp(B), dp(Cc), M, N, K, backend, reps, C.byref(ms))
    if st != 0:
        raise RuntimeError(lib.rsoptsc_last_error().decode())
    return ms.value


def syr2k():
    rng = np.random.default_rng(0)
    for n, reps in ((300, 1), (500, 1), (2000, 1), (7954, 1)):
        k = 64
        V = np.asfortranarray(rng.standard_normal((n, k)) * np.exp(rng.normal(size=(n, 1))))
        W = np.asfortranarray(rng.standard_normal((n, k)))
        C0 = np.asfortranarray(rng.standard_normal((n, n)))
        ref = C0 - 2 * (V @ W.T + W @ V.T)          # warm-up + 1 rep
        absref = 2 * (np.abs(V) @ np.abs(W).T + np.abs(W) @ np.abs(V).T) + np.abs(C0)
        il = np.tril_indices(n)
        for backend in (1, 0):
            Cc = C0.copy(order="F")
            gemm(4, V, W, Cc, n, n, k, backend, reps)
            err = (np.abs(Cc - ref)[il] / absref[il]).max()
            Cc = C0.copy(order="F")
            ms = gemm(4, V, W, Cc, n, n, k, backend, 20)
            print("syr2k n=%5d k=%d backend %d: max err/|.| %.2e, %.3f ms/call (%.1f TF/s-eq)"
                  % (n, k, backend, err, ms, 2.0 * n * n * k / ms / 1e9), flush=True)


def panel():
    rng = np.random.default_rng(1)
    for (op, M, N, K) in ((3, 1024, 7954, 6000), (1, 6000, 7954, 1024), (1, 7954, 7954, 7954), (0, 7954, 7954, 7954)):
        if op == 3:
            A = np.asfortranarray(rng.standard_normal((K, M)))
            B = np.asfortranarray(rng.standard_normal((K, N)))
        elif op == 1:
            A = np.asfortranarray(rng.standard_normal((M, K)))
            B = np.asfortranarray(rng.standard_normal((K, N)))
        else:
            A = np.asfortranarray(rng.standard_normal((K, M)))
            B = A
        Cc = np.zeros((M, N), order="F")
        for backend in (2, 0):
            ms = gemm(op, A, B, Cc, M, N, K, backend, 5)
            flops = 2.0 * M * N * K * (0.5 if op == 0 else 1.0)
            print("gemm op %d M=%d N=%d K=%d backend %d: %.3f ms (%.1f TF/s-eq)" % (op, M, N, K, backend, ms, flops / ms / 1e9),
                  flush=True)


def eig(sizes):
    for n in sizes:
        rng = np.random.default_rng(n)
        B = rng.standard_normal((n, n // 2 + 7))
        G = np.asfortranarray(B @ B.T)
        lam, V = np.zeros(n), np.zeros((n, n), order="F")
        for solver in (1, 0):
            ms = C.c_double(0)
            api.timers(True)
            st = lib.rsoptsc_debug_sym_eig(dp(G), n, solver, dp(lam), dp(V), 2, C.byref(ms))
            if st != 0:
                raise RuntimeError(lib.rsoptsc_last_error().decode())
            rep = api.phase_report()
            api.timers(False)
            res = np.abs(G @ V - V * lam).max() / np.abs(lam).max() if n <= 8000 else -1
            print("eig n=%d solver %d: %.1f ms  res %.1e  phases/3 calls: %s"
                  % (n, solver, ms.value, res, {k: round(v["ms"] / 3, 2) for k, v in rep.items()}), flush=True)


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "syr2k"
    if mode == "syr2k":
        syr2k()
    elif mode == "panel":
        panel()
    else:
        eig([int(a) for a in sys.argv[2:]] or [7954])
