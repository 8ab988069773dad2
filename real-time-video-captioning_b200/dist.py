"""Multi-GPU plumbing: clips are independent units, so they are sharded contiguously across ranks (one
process per GPU) and the only collective is the gather of the caption tokens at the end.
No activation ever crosses NVLink on this path."""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of ceil(n/world)-sized shards: deterministic, so the gathered token matrix is
    identical at 1/2/4/8 GPUs."""
    per = (n_items + world - 1) // world
    b = min(n_items, rank * per)
    return b, min(n_items, b + per)


def caption_sharded(caption_fn: Callable[[int, int], Tuple[torch.Tensor, torch.Tensor]], n_clips: int, group=None):
    """Run ``caption_fn(begin, end) -> (tokens int32 [m, keep, L], logprobs fp32 [m, keep])`` on this rank's
    shard and all-gather the results in clip order.  Works with NCCL (device tensors) and gloo (CPU tensors)."""
    if not dist.is_available() or not dist.is_initialized():
        return caption_fn(0, n_clips)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (n_clips + world - 1) // world
    b, e = shard_range(n_clips, rank, world)
    tokens, logprobs = caption_fn(b, e)
    pad_t = tokens.new_zeros((per,) + tuple(tokens.shape[1:]))
    pad_l = logprobs.new_zeros((per,) + tuple(logprobs.shape[1:]))
    pad_t[: e - b] = tokens
    pad_l[: e - b] = logprobs
    all_t = [torch.empty_like(pad_t) for _ in range(world)]
    all_l = [torch.empty_like(pad_l) for _ in range(world)]
    dist.all_gather(all_t, pad_t, group=group)
    dist.all_gather(all_l, pad_l, group=group)
    return torch.cat(all_t)[:n_clips], torch.cat(all_l)[:n_clips]
