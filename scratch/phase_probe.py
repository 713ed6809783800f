import sys, time
import numpy as np
sys.path.insert(0, '.')
from rsoptsc_b200 import _lib, api
from rsoptsc_b200.synthetic import synthetic_lognorm
n = int(sys.argv[1]); m = 2000
_lib.init(0)
data, _ = synthetic_lognorm(m, n, seed=0)
for rep in range(2):
    api.timers(True)
    t = time.time(); r = api.SimilarityM(0.5, data, dense=False); dt = time.time() - t
    ph = api.phase_report()
    print('rep', rep, 'n', n, '%.2f s' % dt, 'iters', r['iters'], {k: (round(v['ms']), v['calls']) for k, v in ph.items()}, flush=True)
