"""Multi-GPU planning: words are independent, so the word axis is split into
contiguous shards, one process per GPU, weights replicated, and NO data-path collective runs during
planning.  The only traffic is one final gather of the planned cps and the loss logs.

Backend: ``torch.distributed`` -- NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_words: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of the word axis; the first (n_words % world_size) ranks get one extra word."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, rem = divmod(n_words, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t: torch.Tensor, world_size: int, rank: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], world_size, rank)
    return t[lo:hi]


def gather_words(local: torch.Tensor, n_words: int, dst: Optional[int] = None,
                 group: Optional[dist.ProcessGroup] = None) -> Optional[torch.Tensor]:
    """Concatenate per-rank word shards (axis 0) in rank order.

    dst=None: all_gather (every rank gets the full tensor); dst=r: only rank r gets it (others None).
    Shards may be ragged (n_words not divisible by world size): they are padded to the largest shard
    for the collective and trimmed afterwards."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(n_words, ws, r) for r in range(ws)]
    mx = max(hi - lo for lo, hi in sizes)
    lo, hi = sizes[rank]
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank} holds {local.shape[0]} words, expected {hi - lo}")
    pad = local
    if hi - lo < mx:
        pad = torch.cat((local, local.new_zeros((mx - (hi - lo),) + tuple(local.shape[1:]))), dim=0)
    pad = pad.contiguous()
    if dst is None:
        bufs: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(ws)]
        dist.all_gather(bufs, pad, group=group)
    else:
        bufs = [torch.empty_like(pad) for _ in range(ws)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst, group=group)
        if rank != dst:
            return None
    return torch.cat([b[: h - l] for b, (l, h) in zip(bufs, sizes)], dim=0)


def plan_sharded(make_planner, initial_cp: torch.Tensor, target_mel: torch.Tensor,
                 target_semvec: Optional[torch.Tensor], n_steps: int, gather_dst: Optional[int] = None,
                 group: Optional[dist.ProcessGroup] = None):
    """Shard the words of a job over the ranks of ``group``, plan locally, gather at the end.

    ``make_planner(cp_shard, mel_shard, sv_shard)`` builds this rank's BatchPlanner (any object with
    ``step(n)``, ``planned_cp()`` and ``losses()``).  Returns (planned_cp [B,T,C], total_loss [steps,B]) on
    the gathering rank(s)."""
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_words = initial_cp.shape[0]
    lo, hi = shard_bounds(n_words, ws, rank)
    planner = make_planner(initial_cp[lo:hi], target_mel[lo:hi], None if target_semvec is None else target_semvec[lo:hi])
    planner.step(n_steps)
    cps = gather_words(planner.planned_cp(), n_words, gather_dst, group)
    loss = gather_words(planner.losses()["total"].transpose(0, 1).contiguous(), n_words, gather_dst, group)
    return cps, (None if loss is None else loss.transpose(0, 1))
