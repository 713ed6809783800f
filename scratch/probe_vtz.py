"""Panel GEMMs of the back-transformation at configs[2] scale: X = V^T Z (op 3) for several panel widths."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from rsoptsc_b200 import _lib  # noqa: E402

lib = _lib.init(0)
dp = lambda a: a.ctypes.data_as(_lib.c_double_p)  # noqa: E731
rng = np.random.default_rng(0)
K = int(sys.argv[1]) if len(sys.argv) > 1 else 17000
A = np.asfortranarray(rng.standard_normal((K, 1024)))
for N in (4312, 8624, 17250):
    B = np.asfortranarray(rng.standard_normal((K, N)))
    Cc = np.zeros((1024, N), order="F")
    for backend in (2, 0):
        ms = C.c_double(0)
        st = lib.rsoptsc_debug_gemm(3, dp(A), dp(B), dp(Cc), 1024, N, K, backend, 3, C.byref(ms))
        assert st == 0, lib.rsoptsc_last_error()
        print("V^T Z  M=1024 N=%d K=%d backend %d: %.2f ms (%.1f TF/s-eq)" % (N, K, backend, ms.value, 2.0 * 1024 * N * K / ms.value / 1e9),
              flush=True)
    if N == 4312:
        Bt = np.asfortranarray(rng.standard_normal((1024, N)))
        At = np.asfortranarray(rng.standard_normal((K, 1024)))
        Cu = np.zeros((K, N), order="F")
        for backend in (2, 0):
            ms = C.c_double(0)
            st = lib.rsoptsc_debug_gemm(1, dp(At), dp(Bt), dp(Cu), K, N, 1024, backend, 3, C.byref(ms))
            assert st == 0
            print("V X    M=%d N=%d K=1024 backend %d: %.2f ms (%.1f TF/s-eq)" % (K, N, backend, ms.value, 2.0 * 1024 * N * K / ms.value / 1e9),
                  flush=True)
