"""Round-2 groundwork (CPU, numpy): can the SVTs of the saturated ADMM phase (mu = 1e6, 15 of 23) reuse the
eigenvectors of the previous iteration?  One step of the Ogita-Aishima refinement (4 GEMMs) is applied to
the previous eigenvector matrix instead of a fresh eigen-decomposition, with the exact solver as fallback when
the refined basis is not accurate enough (a posteriori: ||X^T G X - diag||, ||X^T X - I||).

Reports, per saturated iteration, the a-posteriori residuals and the error of J against the exact SVT, and
the final relative Frobenius error of Z against the exact run.

    python scratch/refine_svt.py 500 600 [steps]
"""
import sys

import numpy as np

sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/scratch")
import proto_admm as P  # noqa: E402


def refine(G, X, normG):
    n = G.shape[0]
    R = np.eye(n) - X.T @ X
    S = X.T @ (G @ X)
    lam = np.diag(S) / (1.0 - np.diag(R))
    off = S - np.diag(lam)
    delta = 2.0 * (np.linalg.norm(off, 2) + normG * np.linalg.norm(R, 2))
    diff = lam[None, :] - lam[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        E = np.where(np.abs(diff) > delta, (S + R * lam[None, :]) / diff, R / 2.0)
    np.fill_diagonal(E, np.diag(R) / 2.0)
    return X + X @ E, lam, np.linalg.norm(off, "fro") / normG, np.linalg.norm(R, "fro"), delta


class WarmSVT:
    def __init__(self, steps=1, tol_off=1e-9):
        self.X = None
        self.steps = steps
        self.tol_off = tol_off
        self.log = []
        self.exact_calls = 0
        self.refined_calls = 0

    def __call__(self, A, tau):
        G = A.T @ A
        normG = np.linalg.norm(G, 2)
        saturated = tau <= 1.0000001e-6
        Jex = None
        if self.X is not None and saturated:
            X = self.X
            for _ in range(self.steps):
                X, lam, off, orth, delta = refine(G, X, normG)
            # a-posteriori check of the refined basis
            S = X.T @ (G @ X)
            lam = np.diag(S).copy()
            off2 = np.linalg.norm(S - np.diag(lam), "fro") / normG
            orth2 = np.linalg.norm(np.eye(len(lam)) - X.T @ X, "fro")
            sig = np.sqrt(np.maximum(lam, 0))
            with np.errstate(divide="ignore"):
                h = np.where(sig > tau, 1 - tau / sig, 0.0)
            J = (A @ (X * h)) @ X.T
            Jex = P.svt_gram(A, tau)
            C, Cex = A - J, A - Jex
            self.log.append((off, orth, off2, orth2, np.linalg.norm(J - Jex) / np.linalg.norm(Jex),
                             np.linalg.norm(C - Cex) / np.linalg.norm(Cex), delta / normG))
            if off2 < self.tol_off and orth2 < 1e-10:
                self.X = X
                self.refined_calls += 1
                return J
        lam, V = np.linalg.eigh(G)
        self.X = V
        self.exact_calls += 1
        sig = np.sqrt(np.maximum(lam, 0))
        with np.errstate(divide="ignore"):
            h = np.where(sig > tau, 1 - tau / sig, 0.0)
        keep = h > 0
        return (A @ (V[:, keep] * h[keep])) @ V[:, keep].T


def main():
    m, n = int(sys.argv[1]), int(sys.argv[2])
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    X, idx, K = P.make_problem(m, n)
    ref = P.admm(X, idx, 0.5, svt=P.svt_gram)
    for tol in (1e-6, 1e-9):
        w = WarmSVT(steps=steps, tol_off=tol)
        r = P.admm(X, idx, 0.5, svt=w)
        print(f"steps={steps} accept tol {tol:g}: iters {r['iters']} (ref {ref['iters']}), exact solves {w.exact_calls}, refined {w.refined_calls}, "
              f"relF(Z) {np.linalg.norm(r['Z'] - ref['Z']) / np.linalg.norm(ref['Z']):.2e}")
        for k, row in enumerate(w.log):
            print("   sat it %2d: before off %.1e orth %.1e | after off %.1e orth %.1e | relerr J %.1e  relerr C %.1e  delta/|G| %.1e" % ((k,) + row))


if __name__ == "__main__":
    main()
