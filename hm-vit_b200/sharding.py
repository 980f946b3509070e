"""Scene sharding for multi-GPU runs: scenes are independent, so a batch is split into
contiguous per-rank blocks and every rank runs the whole fusion forward on its block.  There is NO
collective on the data path; torch.distributed is only used to agree on the timing."""
from typing import Tuple


def scene_shard(n_scenes: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[start, stop) of the scenes rank `rank` owns; blocks differ in size by at most one scene."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, rem = divmod(n_scenes, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """max of a per-rank scalar (e.g. the device time of the timed region); identity without a group."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
