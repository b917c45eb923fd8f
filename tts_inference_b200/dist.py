"""Multi-GPU plumbing: streams are independent, so they shard across GPUs with NO data-path
collective ("replicas only" per stream; the reference is single-GPU,
tensor_parallel_size=1, vllm_inference/modal_audio_stream.py:226).  One process per GPU;
``torch.distributed`` is used only for the barrier and the max-over-ranks of a timing."""
from __future__ import annotations

from typing import Tuple


def shard_range(n_streams: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced shard [start, start+count) of n_streams for this rank."""
    assert 0 <= rank < world
    base, rem = divmod(n_streams, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def stream_owner(stream_id: int, world: int) -> int:
    """Sticky placement of a long-lived stream: a stream's windows always go to the same GPU."""
    return stream_id % world


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (e.g. device milliseconds) over the job; identity without a group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_over_ranks(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])
