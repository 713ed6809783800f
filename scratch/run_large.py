"""Large-n feasibility run (towards BASELINE.json configs[2]): SimilarityM with the sparse return only,
then ClusterCells on the device-resident W.  Prints timings, phase split and the planted-label ARI."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from rsoptsc_b200 import _lib, api  # noqa: E402
from rsoptsc_b200.synthetic import synthetic_lognorm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
m = 2000
_lib.init(0)
data, planted = synthetic_lognorm(m, n, seed=0)
d_data = api.DeviceBuffer.from_host(data)
d_W = api.DeviceBuffer(n * n * 8)
api.timers(True)
t = time.time()
r = api.SimilarityM_dev(0.5, d_data, m, n, d_W=d_W)
t_sim = time.time() - t
print("SimilarityM n=%d: %.1f s, K=%d, ADMM iters=%d, E=%.6f" % (n, t_sim, r["K"], r["iters"], r["E"]), flush=True)
t = time.time()
c = api.nmf_lee_dev(d_W, n, 10)
t_nmf = time.time() - t
print("ClusterCells k=10: %.1f s, iters=%d" % (t_nmf, c["iters"]), flush=True)
print({k: round(v["ms"]) for k, v in api.phase_report().items()})
from scipy.special import comb  # noqa: E402
cont = np.zeros((10, planted.max() + 1))
np.add.at(cont, (c["labels"] - 1, planted), 1)
s_ij = comb(cont, 2).sum(); s_a = comb(cont.sum(1), 2).sum(); s_b = comb(cont.sum(0), 2).sum()
exp = s_a * s_b / comb(n, 2)
print("ARI(k=10 vs planted) = %.3f ; cells/s = %.1f" % ((s_ij - exp) / (0.5 * (s_a + s_b) - exp), n / (t_sim + t_nmf)))
