"""Native symmetric eigensolver vs cuSOLVER on a B200: accuracy and time.
    python scratch/eig_native_check.py 300 1500 4000 7954
"""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from rsoptsc_b200 import _lib  # noqa: E402

from rsoptsc_b200 import api  # noqa: E402

lib = _lib.load()
_lib.init(0)
dp = C.POINTER(C.c_double)


def solve(G, solver, reps=1):
    n = G.shape[0]
    lam = np.zeros(n)
    V = np.zeros((n, n), order="F")
    ms = C.c_double(0)
    Gf = np.asfortranarray(G)
    rc = lib.rsoptsc_debug_sym_eig(Gf.ctypes.data_as(dp), n, solver, lam.ctypes.data_as(dp), V.ctypes.data_as(dp), reps,
                                   C.byref(ms))
    if rc != 0:
        raise RuntimeError(lib.rsoptsc_last_error().decode())
    return lam, V, ms.value


def quality(G, lam, V):
    g = torch.from_numpy(np.ascontiguousarray(G)).cuda()
    v = torch.from_numpy(np.ascontiguousarray(V)).cuda()
    l = torch.from_numpy(lam).cuda()
    res = (g @ v - v * l).abs().max().item() / max(abs(lam).max(), 1e-300)
    orth = (v.T @ v - torch.eye(G.shape[0], dtype=torch.float64, device="cuda")).abs().max().item()
    return res, orth


def gram_knn(n, rng, K=30):
    Z = np.zeros((n, n))
    for i in range(n):
        idx = rng.choice(n, K, replace=False)
        Z[idx, i] = rng.random(K) / K
    return Z.T @ Z


for arg in sys.argv[1:]:
    n = int(arg)
    rng = np.random.default_rng(n)
    for name in ("random", "gram_knn"):
        if name == "random":
            A = rng.standard_normal((n, n))
            G = (A + A.T) / 2
        else:
            t = torch.from_numpy(gram_knn(n, rng)).cuda()
            G = t.cpu().numpy()
        G = np.tril(G) + np.tril(G, -1).T
        l0, V0, ms0 = solve(G, 0, reps=2)
        api.timers(True)
        l1, V1, ms1 = solve(G, 1, reps=2)
        rep = api.phase_report()
        api.timers(False)
        r0, o0 = quality(G, l0, V0)
        r1, o1 = quality(G, l1, V1)
        print(f"n={n:6d} {name:9s} cusolver {ms0:8.1f} ms (res {r0:.1e} orth {o0:.1e}) | native {ms1:8.1f} ms "
              f"(res {r1:.1e} orth {o1:.1e}) | max|dlam|/|lam|max {np.abs(l0 - l1).max() / np.abs(l0).max():.1e} "
              f"speedup {ms0 / ms1:.2f}x", flush=True)
        print("        phases (ms over 3 calls):", {k: round(v["ms"], 1) for k, v in rep.items()}, flush=True)
