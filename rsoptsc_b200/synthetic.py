"""Seeded synthetic single-cell count matrices.

k_true populations; gene base means lognormal(0,1); per (gene, population) log fold
change N(0,1) on a random 10 % of the genes; uniform cell labels; library-size factor
lognormal(0,0.3); Poisson counts; log-normalise (per cell / total * 1e4, log1p).
Returns genes x cells float64 (column-major friendly: Fortran order) and the planted labels.
"""
import numpy as np


def synthetic_lognorm(m_genes, n_cells, seed=0, k_true=10, frac_de=0.10, chunk=4096):
    rng = np.random.default_rng(seed)
    base = rng.lognormal(0.0, 1.0, size=m_genes)
    lfc = np.zeros((m_genes, k_true))
    for p in range(k_true):
        de = rng.random(m_genes) < frac_de
        lfc[de, p] = rng.normal(0.0, 1.0, size=int(de.sum()))
    labels = rng.integers(0, k_true, size=n_cells)
    lib = rng.lognormal(0.0, 0.3, size=n_cells)
    out = np.empty((m_genes, n_cells), dtype=np.float64, order="F")
    for lo in range(0, n_cells, chunk):
        hi = min(n_cells, lo + chunk)
        mean = base[:, None] * np.exp(lfc[:, labels[lo:hi]]) * lib[None, lo:hi]
        cnt = rng.poisson(mean).astype(np.float64)
        tot = cnt.sum(axis=0, keepdims=True)
        tot[tot == 0] = 1.0
        out[:, lo:hi] = np.log1p(cnt / tot * 1e4)
    # avoid all-zero cells (reference quirk Q10: NaN columns)
    dead = np.flatnonzero(out.sum(axis=0) == 0)
    for j in dead:
        out[rng.integers(0, m_genes), j] = 1.0
    return out, labels.astype(np.int32)
