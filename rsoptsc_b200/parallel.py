"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed for rendezvous only.

The hot path shards where it has independent units: independent matrices
(replicas), the NMF rank sweep k = 5..30 (config 4) and the PAM sweep are embarrassingly
parallel, so units are dealt round-robin to ranks and only the small per-unit results (label
vectors, scalars) are exchanged with all_gather_object -- no data-path collective.
"""
import os


def world():
    """(rank, local_rank, world_size) from the torchrun environment."""
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)),
            int(os.environ.get("WORLD_SIZE", 1)))


def partition(units, rank, world_size):
    """Round-robin share of `units` for `rank` (deterministic, covers every unit exactly once)."""
    units = list(units)
    return [u for i, u in enumerate(units) if i % world_size == rank]


def gather_results(local, group=None):
    """Merge the per-rank {unit: result} dicts on every rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(local)
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, dict(local), group=group)
    out = {}
    for p in parts:
        out.update(p)
    return out


def max_over_ranks(value, group=None, device=None):
    """Max of a python float over ranks (device-time reporting: slowest rank defines the step)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sweep(units, fn, group=None):
    """Run fn(unit) for this rank's share of `units`, return {unit: result} for ALL units."""
    rank, _, ws = world()
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank, ws = dist.get_rank(group), dist.get_world_size(group)
    except ImportError:
        pass
    local = {u: fn(u) for u in partition(units, rank, ws)}
    return gather_results(local, group)


def nmf_rank_sweep(W, ranks=range(5, 31), group=None):
    """Config 4: ClusterCells for every rank k, the 26 independent NMF problems dealt to the GPUs."""
    from . import api

    def one(k):
        r = api.nmf_lee(W, k)
        return dict(labels=r["labels"], iters=r["iters"])
    return sweep(list(ranks), one, group)
