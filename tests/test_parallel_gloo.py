"""World-size-2 test of the N > 1 host logic on CPU (gloo): unit partitioning, result gather and
max-over-ranks timing.  No GPU involved (the per-unit function is a stand-in)."""
import os
import socket

import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, out_q):
    import torch.distributed as dist
    from rsoptsc_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    units = list(range(5, 31))
    mine = parallel.partition(units, rank, world_size)
    res = parallel.sweep(units, lambda k: dict(k2=k * k, by=rank))
    tmax = parallel.max_over_ranks(1.0 + rank)
    out_q.put((rank, mine, res, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_sweep_partition_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    units = list(range(5, 31))
    assert sorted(got[0][1] + got[1][1]) == units and not set(got[0][1]) & set(got[1][1])
    for rank, mine, res, tmax in got:
        assert sorted(res) == units                       # every rank sees every unit
        assert all(res[k]["k2"] == k * k for k in units)
        assert all(res[k]["by"] == (units.index(k) % 2) for k in units)
        assert tmax == 2.0                                 # slowest rank defines the time


def test_partition_covers_all_units():
    from rsoptsc_b200 import parallel
    for ws in (1, 2, 3, 8):
        units = list(range(26))
        parts = [parallel.partition(units, r, ws) for r in range(ws)]
        assert sorted(sum(parts, [])) == units
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
