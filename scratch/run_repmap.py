"""Timing of the RepresentationMap pieces at the configs[1] size."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from rsoptsc_b200 import _lib, api  # noqa: E402
from rsoptsc_b200.synthetic import synthetic_lognorm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
_lib.init(0)
data, _ = synthetic_lognorm(2000, n, seed=0)
W = api.SimilarityM(0.5, data)["W"]
emb = np.random.default_rng(0).normal(size=(n, 2))
api.timers(True)
for rep in range(2):
    t = time.time()
    r = api.RepresentationMap(emb, W)
    dt = time.time() - t
ph = api.phase_report()
print("n=%d RepresentationMap pieces: wall %.2f s (host<->device copies of n x n matrices included)" % (n, dt))
print({k: (round(v["ms"] / 2, 2), v["calls"] // 2) for k, v in ph.items()})
print("components:", r["n_components"], "sizes", np.sort(r["components"]["csize"])[::-1][:6],
      "finite distances: max", np.nanmax(np.where(np.isfinite(r["dist_graph"]), r["dist_graph"], np.nan)))
print("algorithmic bytes: dist_flat 8 n^2 = %.2f GB, adjacency 16 n^2 = %.2f GB, distances 8 n^2 = %.2f GB"
      % (8e-9 * n * n, 16e-9 * n * n, 8e-9 * n * n))
