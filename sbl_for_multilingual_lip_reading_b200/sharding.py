"""Batch sharding of the visual-encoder path across the GPUs of one box.

Clips are independent in eval mode, so the multi-GPU path is contiguous batch shards with replicated weights and no
data-path collective; the only exchange the reference has is `nn.DataParallel`'s scatter of the clip batch and gather
of the outputs (SBL/train.py:114-115, test.py:163-174).  One process per GPU (`torchrun`); these helpers are
backend-agnostic (NCCL on B200s, gloo in the CPU tests) and carry no arithmetic.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    """Clip range [begin, end) of `rank`, chunked like `DataParallel.scatter` (= `Tensor.chunk(world)`): chunks of
    ceil(n / world) clips, the last ranks may get fewer or none."""
    if n < 0 or world < 1 or not 0 <= rank < world:
        raise ValueError(f"shard_bounds: bad arguments n={n} world={world} rank={rank}")
    per = -(-n // world) if n else 0
    begin = min(n, rank * per)
    return begin, min(n, begin + per)


def gather_outputs(local_out: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All ranks contribute their shard's encoder output [n_local, T, D] (n_local may differ per rank, `shard_bounds`
    order) and receive the whole batch [n_total, T, D] — the `DataParallel.gather` of the reference on every rank.
    One `all_gather_into_tensor` of equal-sized (zero-padded) blocks."""
    world = dist.get_world_size(group)
    if world == 1:
        return local_out
    per = -(-n_total // world)
    t, d = local_out.shape[1], local_out.shape[2]
    block = local_out
    if local_out.shape[0] != per:
        block = local_out.new_zeros((per, t, d))
        block[:local_out.shape[0]] = local_out
    full = local_out.new_empty((world * per, t, d))
    dist.all_gather_into_tensor(full, block.contiguous(), group=group)
    return full[:n_total]


def max_over_ranks(value: float, device, group=None) -> float:
    """Device-side timing is reported as the MAX over ranks (bench.py)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
