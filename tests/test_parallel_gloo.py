"""World-size-2 tests of the N > 1 host logic on CPU (no GPU involved):
  * the multi-GPU split of one computeM problem (rsoptsc_host_shard_owner): both ranks derive the same
    plan, every cell has exactly one owner, big components are cut, small ones stay whole;
  * the NCCL-id rendezvous of parallel.init_comm (standard-library TCP group), with a stand-in payload;
  * the replica-style sweep over independent problems, gathered through a gloo process group.
"""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ring_idx(sizes, K=4):
    """Neighbour table whose support graph has one ring-shaped component per entry of `sizes`."""
    rows = []
    base = 0
    for s in sizes:
        for i in range(s):
            rows.append([base + (i + t) % s for t in range(K)] if s >= K else [base + (i + t) % s for t in range(K)])
        base += s
    return np.asarray(rows, dtype=np.int32)


class GlooGroup:
    """all_gather of python objects over torch.distributed (gloo), the interface parallel.sweep expects."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world_size = dist.get_rank(), dist.get_world_size()

    def all_gather(self, obj):
        parts = [None] * self.world_size
        self.dist.all_gather_object(parts, obj)
        return parts


def _worker(rank, world_size, port, out_q):
    import torch.distributed as dist
    from rsoptsc_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    group = GlooGroup()
    units = list(range(5, 31))
    mine = parallel.partition(units, rank, world_size)
    res = parallel.sweep(units, lambda k: dict(k2=k * k, by=rank), group=group)
    # the id rendezvous init_comm uses by default (TCP on MASTER_PORT + 1), with a stand-in 128-byte payload
    tcp = parallel.TcpGroup(rank, world_size)
    token = tcp.broadcast(bytes(range(128)) if rank == 0 else None)
    gathered = tcp.all_gather(("rank", rank))
    tcp.close()
    # the split of one problem: 300 + 40 + 25 + 7 cells, components of >= 100 cells are cut across the ranks
    idx = _ring_idx([300, 40, 25, 7])
    plan = parallel.shard_owner(idx, world_size, split_min=100)
    plans = group.all_gather({k: v.tolist() for k, v in plan.items()})
    out_q.put((rank, mine, res, token, gathered, plans))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_split_rendezvous_and_sweep():
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
    for rank, mine, res, token, gathered, plans in got:
        assert sorted(res) == units                       # every rank sees every unit
        assert all(res[k]["k2"] == k * k for k in units)
        assert all(res[k]["by"] == (units.index(k) % 2) for k in units)
        assert token == bytes(range(128))                 # rank 0's payload reached every rank
        assert gathered == [("rank", 0), ("rank", 1)]
        assert plans[0] == plans[1]                       # both ranks derived the same split
        owner = np.asarray(plans[0]["owner"])
        comp = np.asarray(plans[0]["component"])
        split = np.asarray(plans[0]["split"])
        assert set(owner.tolist()) == {0, 1}              # every cell has exactly one owner in range
        assert split[:300].all() and not split[300:].any()
        # the cut component: contiguous halves; small components: whole, largest first onto the lighter rank
        assert (owner[:150] == 0).all() and (owner[150:300] == 1).all()
        for c in np.unique(comp[300:]):
            assert len(set(owner[comp == c].tolist())) == 1
        assert owner[300] != owner[340]                   # 40-cell and 25-cell components on different ranks


def test_partition_covers_all_units():
    from rsoptsc_b200 import parallel
    for ws in (1, 2, 3, 8):
        units = list(range(26))
        parts = [parallel.partition(units, r, ws) for r in range(ws)]
        assert sorted(sum(parts, [])) == units
        assert max(map(len, parts)) - min(map(len, parts)) <= 1


def test_shard_owner_balances_components():
    from rsoptsc_b200 import parallel
    idx = _ring_idx([4100, 900, 800, 50, 50, 50])
    for P in (1, 2, 4, 8):
        plan = parallel.shard_owner(idx, P, split_min=2048)
        owner = plan["owner"]
        assert owner.min() >= 0 and owner.max() < P
        if P == 1:
            assert not plan["split"].any()
            continue
        assert plan["split"][:4100].all() and not plan["split"][4100:].any()
        counts = np.bincount(owner[:4100], minlength=P)
        assert counts.max() - counts.min() <= 1           # equal column ranges of the cut component
        if P >= 2:
            assert owner[4100] != owner[5000]             # the 900- and 800-cell blocks go to different ranks
