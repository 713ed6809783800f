"""Native symmetric eigensolver above cuSOLVER's 32,768-row limit: accuracy and time of one solve.
    python scratch/eig_native_big.py 33000
"""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from rsoptsc_b200 import _lib, api  # noqa: E402

lib = _lib.load()
_lib.init(0)
dp = C.POINTER(C.c_double)
n = int(sys.argv[1])
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.rand((n, 64), dtype=torch.float64, device="cuda", generator=g)
G = A @ A.T                                   # rank 64 + a full-rank part: wide spectrum
G += torch.diag(torch.rand(n, dtype=torch.float64, device="cuda", generator=g))
B = torch.rand((n, n), dtype=torch.float64, device="cuda", generator=g) * 1e-3
G += B + B.T
del A, B
Gh = np.asfortranarray(G.cpu().numpy())
lam = np.zeros(n)
V = np.zeros((n, n), order="F")
ms = C.c_double(0)
api.timers(True)
t = time.time()
_lib.check(lib.rsoptsc_debug_sym_eig(Gh.ctypes.data_as(dp), n, 1, lam.ctypes.data_as(dp), V.ctypes.data_as(dp), 1, C.byref(ms)))
print(f"n={n}: native solve {ms.value / 1e3:.2f} s (wall incl. warm-up and copies {time.time() - t:.1f} s)")
print({k: round(v["ms"] / 2e3, 2) for k, v in api.phase_report().items()}, "(s per solve)")
v = torch.from_numpy(np.ascontiguousarray(V)).cuda()
l = torch.from_numpy(lam).cuda()
res = (G @ v - v * l).abs().max().item() / lam.max()
orth = (v.T @ v - torch.eye(n, dtype=torch.float64, device="cuda")).abs().max().item()
ref = torch.linalg.eigvalsh(G) if n <= 20000 else None
print(f"residual {res:.2e}  orthogonality {orth:.2e}", "" if ref is None else f"eig diff {(ref - l).abs().max().item() / lam.max():.2e}")
