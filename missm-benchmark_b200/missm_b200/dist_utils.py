"""Small torch.distributed helpers shared by bench.py and the multi-process tests."""
import torch
import torch.distributed as dist


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def max_over_ranks(value, device="cpu"):
    """Device-timed durations are reported as the MAX over ranks (never a wall clock)."""
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def rank_seed(base, rank):
    """Every rank draws its own synthetic shard (weak scaling: B samples per GPU)."""
    return int(base) + int(rank)
