"""Host-side logic of the multi-GPU mode: independent KKT instances (multistart / parameter sweeps) are
assigned to ranks statically, one process per GPU, no collective on the data path.
`torch.distributed` is used only for the barrier and for reducing timings / gathering results."""
from __future__ import annotations


def assign_instances(num_instances: int, world: int, rank: int) -> list[int]:
    """Round-robin: instance i runs on rank i mod world (gpu = i mod G)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, num_instances, world))


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Timing rule of the benchmark contract: the job takes as long as its slowest rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_objects(obj, dist=None):
    """All ranks' python objects on every rank (results of the instances each rank solved)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out
