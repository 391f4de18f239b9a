"""Multi-GPU planning: words are independent, so the word axis is split into
contiguous shards, one process per GPU, weights replicated, and NO data-path collective runs during
planning.  The only traffic is one final gather of the planned cps and the loss logs.

Backend: ``torch.distributed`` -- NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from collections import namedtuple
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

# what plan_resynth_sharded returns: this rank's PlanningResults plus the gathered job
ShardedPlanningResults = namedtuple("ShardedPlanningResults", "local, planned_cp, planned_loss_steps, word_range")


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


def _host(t: torch.Tensor) -> np.ndarray:
    if not t.is_cuda:
        return t.numpy()
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)      # pinned staging: the gathered job is ~100 MB
    h.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return h.numpy()


def plan_resynth_sharded(paule, *, target_acoustic, initial_cp=None, target_semvec=None, gather_dst: Optional[int] = None,
                         group: Optional[dist.ProcessGroup] = None, **kwargs) -> ShardedPlanningResults:
    """``Paule.plan_resynth`` for a batch of words over all ranks of ``group`` (one process per GPU): every rank passes the
    WHOLE job (host arrays ``target_acoustic`` [B,Tm,60], ``initial_cp`` [B,T,30], ``target_semvec`` [B,300]), plans its
    contiguous shard of the word axis through ``paule.plan_resynth(**kwargs)`` and one final NCCL ``all_gather`` (``gather``
    with ``gather_dst``) returns the planned cps [B,T,30] and the per-step total loss [steps,B] of the whole job as host
    arrays.  No collective runs inside the planning loop.  ``local`` is this rank's full ``PlanningResults``."""
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_words = len(target_acoustic)
    lo, hi = shard_bounds(n_words, ws, rank)
    if hi <= lo:
        raise ValueError(f"rank {rank} of {ws} has no words: plan at least one word per GPU")
    res = paule.plan_resynth(target_acoustic=target_acoustic[lo:hi],
                             initial_cp=None if initial_cp is None else initial_cp[lo:hi],
                             target_semvec=None if target_semvec is None else target_semvec[lo:hi], **kwargs)
    pl = paule.last_planner
    cps = gather_words(pl.planned_cp(), n_words, gather_dst, group)
    loss = gather_words(pl.losses()["total"].transpose(0, 1).contiguous(), n_words, gather_dst, group)
    if cps is None:
        return ShardedPlanningResults(res, None, None, (lo, hi))
    return ShardedPlanningResults(res, _host(cps), np.ascontiguousarray(_host(loss).T), (lo, hi))


# ---- ragged jobs: length-bucketed sharding ----------------------------------------------
def length_buckets(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Assign words of different lengths to ranks.  A rank plans its words in lock-step, padded to ITS longest word, so
    its cost is (longest word) x (number of words): words are sorted by length (longest first, ties by index) and cut
    into ``world_size`` contiguous runs whose largest cost is minimal -- ranks with long words get fewer of them, and no
    word is padded beyond the longest word of its own rank.  Returns the word indices of every rank (possibly empty)."""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    L = [int(lengths[i]) for i in order]

    def cuts(limit):   # greedy: longest run from `start` whose cost L[start] * n stays within limit
        out, start = [], 0
        while start < len(L):
            n = max(1, limit // L[start])
            out.append((start, min(len(L), start + n)))
            start += n
        return out

    if not L:
        return [[] for _ in range(world_size)]
    lo, hi = L[0], L[0] * len(L)
    while lo < hi:   # smallest cost limit that needs at most world_size runs
        mid = (lo + hi) // 2
        if len(cuts(mid)) <= world_size:
            hi = mid
        else:
            lo = mid + 1
    runs = cuts(lo)
    return [[order[i] for i in range(a, b)] for a, b in runs] + [[] for _ in range(world_size - len(runs))]


def plan_sharded_ragged(make_planner, cps: Sequence[torch.Tensor], mels: Sequence[torch.Tensor], n_steps: int,
                        group: Optional[dist.ProcessGroup] = None):
    """Ragged job over the ranks of ``group``: ``cps[b]`` is [T_b, C], ``mels[b]`` [T_b // 2, Cm].  Every rank plans its
    length bucket padded to the bucket's longest word (``make_planner(cp_pad, mel_pad, lengths)``) and one all_gather
    returns, on every rank, the planned cps as a list in the ORIGINAL word order plus the loss log [steps, B]."""
    ws = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lengths = [int(c.shape[0]) for c in cps]
    buckets = length_buckets(lengths, ws)
    mine = buckets[rank]
    T_glob, C = max(lengths), cps[0].shape[1]
    dev = cps[0].device
    if mine:
        T_loc = max(lengths[i] for i in mine)
        cp_pad = cps[0].new_zeros((len(mine), T_loc, C))
        mel_pad = mels[0].new_zeros((len(mine), T_loc // 2, mels[0].shape[1]))
        for k, i in enumerate(mine):
            cp_pad[k, :lengths[i]] = cps[i]
            mel_pad[k, :lengths[i] // 2] = mels[i]
        planner = make_planner(cp_pad, mel_pad, [lengths[i] for i in mine])
        planner.step(n_steps)
        out_cp, out_loss = planner.planned_cp(), planner.losses()["total"].transpose(0, 1).contiguous()
    else:
        T_loc = 0
        out_cp, out_loss = cps[0].new_zeros((0, 0, C)), cps[0].new_zeros((0, n_steps))
    # one collective: every rank's block padded to [largest bucket, T_glob * C + n_steps]
    mx = max(len(b) for b in buckets)
    block = torch.zeros((mx, T_glob * C + n_steps), dtype=out_cp.dtype, device=dev)
    if mine:
        block[:len(mine), :T_loc * C] = out_cp.reshape(len(mine), -1).to(dev)
        block[:len(mine), T_glob * C:] = out_loss.to(dev)
    if ws > 1:
        bufs = [torch.empty_like(block) for _ in range(ws)]
        dist.all_gather(bufs, block, group=group)
    else:
        bufs = [block]
    planned: List[Optional[torch.Tensor]] = [None] * len(lengths)
    loss = torch.zeros((n_steps, len(lengths)), dtype=out_cp.dtype, device=dev)
    for r, idx in enumerate(buckets):
        for k, i in enumerate(idx):
            # rank r laid word i out with row stride T_r * C (its own padding), then zero-filled up to T_glob * C
            T_r = max(lengths[j] for j in idx)
            planned[i] = bufs[r][k, :T_r * C].reshape(T_r, C)[:lengths[i]].clone()
            loss[:, i] = bufs[r][k, T_glob * C:]
    return planned, loss
