"""BASELINE.json configs[3]: CountClusters eigengap + NMF rank sweep k = 5..30 on 20,000 cells.
Scratch driver: prints timings and the inferred cluster count (planted: 10 populations)."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from rsoptsc_b200 import _lib, api  # noqa: E402
from rsoptsc_b200.synthetic import synthetic_lognorm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
m = 2000
_lib.init(0)
data, planted = synthetic_lognorm(m, n, seed=0)
t = time.time()
r = api.SimilarityM(0.5, data)
t_sim = time.time() - t
W = r["W"]
print("SimilarityM n=%d: %.1f s, K=%d, ADMM iters=%d, nnz(W)=%d" % (n, t_sim, r["K"], r["iters"], np.count_nonzero(W)), flush=True)
api.timers(True)
t = time.time()
cc = api.CountClusters(W, n_comp=3)
t_cc = time.time() - t
print("CountClusters: %.1f s upper=%d lower=%d  phases=%s" % (t_cc, cc["upper_bound"], cc["lower_bound"],
      {k: round(v["ms"]) for k, v in api.phase_report().items()}), flush=True)
api.timers(False)
t = time.time()
sw = api.nmf_sweep(W, ranks=(5, 30))
t_sweep = time.time() - t
print("NMF rank sweep k=5..30 (batched, rsoptsc_nmf_sweep): %.1f s, iters=%s" % (t_sweep, list(sw["iters"])), flush=True)
t = time.time()
one = api.nmf_lee(W, 10)
print("single nmf_lee k=10: %.2f s, iters=%d, labels equal to the sweep's: %s"
      % (time.time() - t, one["iters"], bool(np.array_equal(one["labels"], sw["labels"][5]))), flush=True)
lab = sw["labels"][5]
from scipy.special import comb  # noqa: E402
cont = np.zeros((10, planted.max() + 1))
np.add.at(cont, (lab - 1, planted), 1)
s_ij = comb(cont, 2).sum(); s_a = comb(cont.sum(1), 2).sum(); s_b = comb(cont.sum(0), 2).sum()
exp = s_a * s_b / comb(n, 2)
print("ARI(k=10 vs planted) = %.3f" % ((s_ij - exp) / (0.5 * (s_a + s_b) - exp)))
