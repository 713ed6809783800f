"""Golden vectors of the CPU oracle at sizes where the default CUDA path runs its native
eigensolver and the tcgen05 GEMMs.  Run in the build container only
(CPU, ~1 h):

    OPENBLAS_NUM_THREADS=6 python tests/golden/make_golden_scale.py [case ...]

Cases
  cut3000   first 3,000 cells of the bench matrix (synthetic_lognorm(2000, 10000, seed=0))
  big6200   synthetic_lognorm(2000, 6200, seed=21): kNN support with one connected component of
            > 4,200 cells (native eigensolver + tcgen05 by default) and three small ones

Stored per case (tests/golden/scale_<case>.npz, ~1 MB): the generator arguments, K, the kNN index
table, Z on its support (Z[j, idx[j, t]]), E, the iteration count, the stop reason, the Err1/Err2
traces (R/SimilarityM.R:189), the component sizes, and ClusterCells(k=10) labels + NMF iterations.
The input matrix itself is regenerated from the seed by the test.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import rsoptsc_oracle as O  # noqa: E402
from rsoptsc_b200.synthetic import synthetic_lognorm  # noqa: E402

CASES = {
    "cut3000": dict(genes=2000, gen_cells=10000, cells=3000, seed=0),
    "big6200": dict(genes=2000, gen_cells=6200, cells=6200, seed=21),
}
LAMBDA = 0.5
RANK = 10


def make_input(cfg):
    data, _ = synthetic_lognorm(cfg["genes"], cfg["gen_cells"], seed=cfg["seed"])
    return np.asfortranarray(data[:, :cfg["cells"]])


def run_case(name):
    cfg = CASES[name]
    data = make_input(cfg)
    n = data.shape[1]
    t0 = time.time()
    trace = []
    r = O.SimilarityM(LAMBDA, data, literal=False, trace=trace)
    t1 = time.time()
    idx = r["idx"]
    Z = r["Z"]
    Zs = Z[np.arange(n)[:, None], idx]
    assert np.count_nonzero(Z) <= idx.size
    c = O.ClusterCells(r["W"], n_clusters=RANK)
    t2 = time.time()
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components
    G = sp.coo_matrix((np.ones(idx.size), (np.repeat(np.arange(n), idx.shape[1]), idx.ravel())), shape=(n, n))
    _, lab = connected_components(G, directed=False)
    comps = np.sort(np.bincount(lab))[::-1]
    out = os.path.join(HERE, "scale_%s.npz" % name)
    np.savez_compressed(
        out, genes=np.int64(cfg["genes"]), gen_cells=np.int64(cfg["gen_cells"]), cells=np.int64(cfg["cells"]),
        seed=np.int64(cfg["seed"]), lam=np.float64(LAMBDA), K=np.int32(r["K"]), idx=idx.astype(np.int32), Zs=Zs,
        E=np.float64(r["E"]), iters=np.int32(r["iters"]), stop=np.asarray(r["stop"]),
        err1=np.asarray([t[3] for t in trace]), err2=np.asarray([t[4] for t in trace]),
        rankJ=np.asarray([t[2] for t in trace], dtype=np.int32), components=comps.astype(np.int64),
        W_sum=np.float64(r["W"].sum()), W_nnz=np.int64(np.count_nonzero(r["W"])),
        labels=np.asarray(c["labels"], dtype=np.int32), nmf_iters=np.int32(c["iters"]), rank=np.int32(RANK))
    print("%s: n=%d K=%d iters=%d stop=%s E=%.10f comps=%s nmf_iters=%d | SimilarityM %.0f s, ClusterCells %.0f s -> %s (%d B)"
          % (name, n, r["K"], r["iters"], r["stop"], r["E"], comps[:5].tolist(), c["iters"], t1 - t0, t2 - t1, out,
             os.path.getsize(out)), flush=True)


if __name__ == "__main__":
    for nm in (sys.argv[1:] or list(CASES)):
        run_case(nm)
