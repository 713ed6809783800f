import ctypes as C
import sys

import numpy as np

sys.path.insert(0, '.')
from rsoptsc_b200 import _lib  # noqa: E402

lib = _lib.init(0)
dp = lambda a: a.ctypes.data_as(_lib.c_double_p)  # noqa: E731


def run(op, M, N, K, backend, reps=1, seed=0, scale=True):
    rng = np.random.default_rng(seed)
    if op == 0:
        A = np.asfortranarray(rng.normal(size=(K, M)))
        B = A
        N = M
    elif op == 1:
        A = np.asfortranarray(rng.normal(size=(M, K)))
        B = np.asfortranarray(rng.normal(size=(K, N)))
    elif op == 2:
        A = np.asfortranarray(rng.normal(size=(M, K)))
        B = np.asfortranarray(rng.normal(size=(N, K)))
    else:
        A = np.asfortranarray(rng.normal(size=(K, M)))
        B = np.asfortranarray(rng.normal(size=(K, N)))
    if scale:  # wide dynamic range per row/column
        A *= np.exp(rng.normal(size=A.shape[1]) * 3)[None, :]
        if op != 0:
            B = np.asfortranarray(B * np.exp(rng.normal(size=B.shape[0]) * 2)[:, None])
        else:
            B = A
    Cc = np.zeros((M, N), order='F')
    ms = C.c_double(0)
    st = lib.rsoptsc_debug_gemm(op, dp(A), dp(B), dp(Cc), M, N, K, backend, reps, C.byref(ms))
    if st != 0:
        print('ERR', lib.rsoptsc_last_error())
        sys.exit(1)
    ref = {0: lambda: A.T @ A, 1: lambda: A @ B, 2: lambda: A @ B.T, 3: lambda: A.T @ B}[op]()
    aa, bb = np.abs(A), np.abs(B)
    absref = {0: lambda: aa.T @ aa, 1: lambda: aa @ bb, 2: lambda: aa @ bb.T, 3: lambda: aa.T @ bb}[op]()
    if op == 0:
        il = np.tril_indices(M)
        err = np.abs(Cc - ref)[il] / absref[il]
    else:
        err = np.abs(Cc - ref) / absref
    return err.max(), ms.value


if __name__ == '__main__':
    mode = sys.argv[1] if len(sys.argv) > 1 else 'check'
    if mode == 'check':
        for op in (0, 1, 2, 3):
            for (M, N, K) in [(512, 512, 512), (640, 768, 1000), (1300, 900, 2100)]:
                e1, _ = run(op, M, N, K, 2)
                e0, _ = run(op, M, N, K, 0)
                print('op', op, (M, N, K), 'ozaki max err/(|A||B|) %.2e   cublas %.2e' % (e1, e0), flush=True)
    elif mode == 'chunk':
        # K beyond one int32 accumulation unit (16608) exercises the K-chunk loop
        for op, (M, N, K) in [(3, (1024, 768, 20000)), (0, (1024, 1024, 40000)), (2, (600, 1100, 17000))]:
            e1, t1 = run(op, M, N, K, 2, scale=False)
            e0, t0 = run(op, M, N, K, 0, scale=False)
            print('op', op, (M, N, K), 'ozaki err %.2e (%.2f ms)  cublas err %.2e (%.2f ms)' % (e1, t1, e0, t0), flush=True)
    else:
        for n in (4096, 10000):
            for op in (0, 1):
                e1, t1 = run(op, n, n, n, 1, reps=3, scale=False)
                e0, t0 = run(op, n, n, n, 0, reps=3, scale=False)
                fl = (1 if op == 0 else 2) * n ** 3
                print('n', n, 'op', op, 'ozaki %.2f ms (%.1f TF/s-eq, err %.1e)  cublas %.2f ms (%.1f TF/s, err %.1e)'
                      % (t1, fl / t1 / 1e9, e1, t0, fl / t0 / 1e9, e0), flush=True)
