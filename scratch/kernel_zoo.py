"""Runs each standalone HBM/ALU-bound kernel of the library once at a realistic size (for ncu captures and
CUDA-event timings): kNN, dense threshold+symmetrise, consensus, dist_flat, adjacency, BFS, signalling, normalise."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from rsoptsc_b200 import _lib, api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
_lib.init(0)
rng = np.random.default_rng(0)
api.timers(True)
emb = rng.normal(size=(n, 3))
idx = api.knn(emb, 30)
Z = np.zeros((n, n))
Z[np.arange(n)[:, None], idx] = rng.random((n, 30))
W = api.threshold_symmetrise(Z)
labs = np.stack([rng.integers(1, k + 1, size=n) for k in range(2, 21)]).astype(np.int32)
C = api.GetConsensusEnsemble(labs, 0.4)
D = api.dist_flat(emb[:, :2])
A = api.adjacency(W)
G = api.graph_distances(A)
P = api.signaling_pair(rng.poisson(1.0, n).astype(float), rng.poisson(1.0, n).astype(float), rng.random(n), rng.random(n))
X = api.normalize(rng.random((2000, n)))
rep = api.phase_report()
print("n =", n)
for k, v in rep.items():
    print("%-12s %8.3f ms x %d" % (k, v["ms"] / max(v["calls"], 1), v["calls"]))
print("algorithmic GB: symm %.2f  consensus(write) %.2f  dist_flat(write) %.2f  adjacency %.2f  bfs(write) %.2f  signaling(write) %.2f  normalize %.2f"
      % (24e-9 * n * n, 8e-9 * n * n, 8e-9 * n * n, 16e-9 * n * n, 8e-9 * n * n, 8e-9 * n * n, 24e-9 * 2000 * n))
