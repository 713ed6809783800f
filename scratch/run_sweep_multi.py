"""NMF rank sweep k = 5..30 (BASELINE.json configs[3]) dealt round-robin to the ranks of a torchrun job
(one process per GPU, no data-path collective; only label vectors are gathered).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 scratch/run_sweep_multi.py [cells]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from rsoptsc_b200 import _lib, api, parallel  # noqa: E402
from rsoptsc_b200.synthetic import synthetic_lognorm  # noqa: E402

rank, local_rank, world = parallel.world()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
_lib.init(local_rank)
data, _ = synthetic_lognorm(2000, n, seed=0)          # every rank builds the same similarity matrix
W = api.SimilarityM(0.5, data)["W"]
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
res = parallel.nmf_rank_sweep(W, range(5, 31))
torch.cuda.synchronize()
t = parallel.max_over_ranks(time.perf_counter() - t0, device="cuda")
if rank == 0:
    assert sorted(res) == list(range(5, 31))
    print("world=%d n=%d rank sweep k=5..30: %.2f s (max over ranks); iters=%s"
          % (world, n, t, [res[k]["iters"] for k in sorted(res)]))
    np.save(os.path.join("gpurun_out", "sweep_labels_world%d.npy" % world),
            np.stack([res[k]["labels"] for k in sorted(res)]))
if world > 1:
    dist.destroy_process_group()
