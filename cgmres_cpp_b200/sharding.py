"""Instance sharding across the GPUs of one box: contiguous, disjoint index ranges, no collective
on the data path.  The only cross-rank traffic of a run is the barrier around the timed region and a MAX
reduction of two scalars (elapsed times); both go through whatever torch.distributed backend is initialised
(NCCL on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations


def shard_range(total: int, world: int, rank: int) -> tuple[int, int]:
    """Half-open range of global instance indices owned by `rank` (sizes differ by at most one)."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def weak_scaling_range(per_gpu: int, world: int, rank: int) -> tuple[int, int]:
    """Weak scaling: every rank owns `per_gpu` instances of a global batch of per_gpu*world."""
    return shard_range(per_gpu * world, world, rank)


def max_over_ranks(values, dist=None, device=None):
    """Element-wise MAX over ranks of a short list of floats (timings); identity without a process group."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch

    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def aggregate_updates_per_second(per_gpu: int, world: int, steps: int, max_elapsed_ms: float) -> float:
    """Whole-job throughput: all instances of all ranks advanced `steps` times in the slowest rank's time."""
    return per_gpu * world * steps / (max_elapsed_ms * 1e-3)
