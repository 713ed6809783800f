"""Round-2 groundwork: two-stage tridiagonalisation prototype (numpy, CPU only).

Stage 1  dense -> band (bandwidth b): blocked Householder QR of the sub-diagonal block column, two-sided
         update of the trailing matrix  A22 <- A22 - V W^T - W V^T,  W = P - 1/2 V (T^T V^T P),  P = A22 V T
         (one symmetric x panel product + one rank-2b update per panel: all GEMM-shaped).
Stage 2  band -> tridiagonal by bulge chasing: sweep j annihilates column j below the sub-diagonal with a
         Householder reflector of length <= b, then chases the bulge it creates down the band in steps of b.
Both stages return their reflectors so that eigenvectors can be back-transformed: Z = Q1 Q2 Z_T.

The prototype checks (a) the tridiagonal matrix has the spectrum of A, (b) Q1 Q2 T Q2^T Q1^T = A, and
counts the reflectors of stage 2 (n^2 / (2b) of length b: the back-transformation applies them in
diamond-shaped groups on the GPU).

    python scratch/two_stage_proto.py 300 16
"""
import sys

import numpy as np


def house(x):
    """v (v[0] = 1), tau with (I - tau v v^T) x = beta e1."""
    alpha = x[0]
    xn = np.linalg.norm(x[1:])
    if xn == 0.0:
        return np.concatenate([[1.0], np.zeros(len(x) - 1)]), 0.0, alpha
    beta = -np.copysign(np.hypot(alpha, xn), alpha)
    v = x / (alpha - beta)
    v[0] = 1.0
    return v, (beta - alpha) / beta, beta


def dense_to_band(A, b):
    """Returns the band matrix B (full storage, bandwidth b) and the list of (row0, V, T) block reflectors."""
    A = A.copy()
    n = A.shape[0]
    blocks = []
    for k0 in range(0, n - b - 1, b):
        r0 = k0 + b                      # rows r0.. of block column k0..k0+b-1 get triangularised
        m = n - r0
        w = min(b, m)
        panel = A[r0:, k0:k0 + b].copy()
        V = np.zeros((m, w))
        taus = np.zeros(w)
        for c in range(w):               # Householder QR of the panel
            v, tau, beta = house(panel[c:, c].copy())
            V[c:, c] = v
            taus[c] = tau
            panel[c:, c:] -= tau * np.outer(v, v @ panel[c:, c:])
        T = np.zeros((w, w))             # compact WY: H0 H1 ... = I - V T V^T
        for c in range(w):
            T[c, c] = taus[c]
            if c:
                T[:c, c] = -taus[c] * (T[:c, :c] @ (V[:, :c].T @ V[:, c]))
        A[r0:, k0:k0 + b] = np.triu(panel)          # R (upper triangular block just below the band)
        A[k0:k0 + b, r0:] = A[r0:, k0:k0 + b].T
        A22 = A[r0:, r0:]
        P = A22 @ V @ T
        W = P - 0.5 * V @ (T.T @ (V.T @ P))
        A22 -= V @ W.T + W @ V.T
        blocks.append((r0, V, T))
    return A, blocks


def band_to_tridiag(B, b):
    """Bulge chasing on a full-storage band matrix; returns T (full storage) and reflectors (row0, v, tau)."""
    B = B.copy()
    n = B.shape[0]
    refl = []
    for j in range(n - 2):
        # annihilate B[j+2 : j+b+1, j]; then chase
        col = j
        r0 = j + 1
        while r0 < n - 1:
            r1 = min(r0 + b, n)
            x = B[r0:r1, col].copy()
            if len(x) < 2 or not np.any(x[1:]):
                break
            v, tau, beta = house(x)
            refl.append((r0, v, tau))
            # two-sided application on rows/cols r0..r1-1 (only the band neighbourhood is touched)
            lo, hi = max(0, r0 - b), min(n, r1 + b)
            blk = B[r0:r1, lo:hi]
            blk -= tau * np.outer(v, v @ blk)
            blk = B[lo:hi, r0:r1]
            blk -= tau * np.outer(blk @ v, v)
            # the bulge now sits in rows r1 .. r1+b-1 of column r0: next reflector starts there
            col = r0
            r0 = r1
    return B, refl


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, n))
    A = X @ X.T / n
    B, blocks = dense_to_band(A, b)
    off_band = np.abs(np.tril(B, -(b + 1))).max()
    T, refl = band_to_tridiag(B, b)
    off_tri = np.abs(np.tril(T, -2)).max()
    lam_A = np.linalg.eigvalsh(A)
    lam_T = np.linalg.eigvalsh(np.diag(np.diag(T)) + np.diag(np.diag(T, -1), -1) + np.diag(np.diag(T, -1), 1))
    # accumulate Q1, Q2 explicitly (prototype only)
    Q1 = np.eye(n)
    for r0, V, Tm in blocks:
        Q1[:, r0:] -= (Q1[:, r0:] @ V) @ Tm @ V.T
    Q2 = np.eye(n)
    for r0, v, tau in refl:
        r1 = r0 + len(v)
        Q2[:, r0:r1] -= tau * np.outer(Q2[:, r0:r1] @ v, v)
    Q = Q1 @ Q2
    Tt = np.diag(np.diag(T)) + np.diag(np.diag(T, -1), -1) + np.diag(np.diag(T, -1), 1)
    print(f"n={n} b={b}: off-band after stage 1 {off_band:.1e}, off-tridiagonal after stage 2 {off_tri:.1e}")
    print(f"  spectrum error {np.abs(lam_A - lam_T).max() / lam_A.max():.1e}, "
          f"|Q T Q^T - A| {np.abs(Q @ Tt @ Q.T - A).max():.1e}, |Q^T Q - I| {np.abs(Q.T @ Q - np.eye(n)).max():.1e}")
    print(f"  stage-2 reflectors: {len(refl)} (n^2/(2b) = {n * n / (2 * b):.0f}), "
          f"stage-2 flops ~ {sum(8.0 * len(v) * min(3 * b, n) for _, v, _ in refl):.2e} (6 n^2 b = {6.0 * n * n * b:.2e})")


if __name__ == "__main__":
    main()
