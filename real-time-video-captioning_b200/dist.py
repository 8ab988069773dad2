"""Multi-GPU plumbing: clips are independent units, so they are sharded contiguously across ranks (one
process per GPU) and the only collective is the gather of the caption tokens at the end.
No activation ever crosses NVLink on this path.

The gather is ONE collective per call: tokens (int32) and log-probabilities (fp32, bit-cast) of a shard travel in one
packed int32 buffer through ``all_gather_into_tensor``.  ``async_op=True`` returns a handle instead of blocking, so a
caller that captions batch after batch (bench.py) enqueues the gather of batch i behind its decode and carries on with
batch i+1; ranks then meet once, when the handles are waited for, instead of in lock-step after every batch."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of ceil(n/world)-sized shards: deterministic, so the gathered token matrix is
    identical at 1/2/4/8 GPUs."""
    per = (n_items + world - 1) // world
    b = min(n_items, rank * per)
    return b, min(n_items, b + per)


class GatherHandle:
    """Result of ``caption_sharded(..., async_op=True)``: ``wait()`` -> (tokens int32 [n, keep, L], logprobs fp32 [n, keep])."""

    def __init__(self, work, packed: torch.Tensor, n_clips: int, keep: int, length: int):
        self._work, self._packed, self._n, self._keep, self._len = work, packed, n_clips, keep, length
        self._out = None

    def wait(self):
        if self._out is None:
            if self._work is not None:
                self._work.wait()
            flat = self._packed.view(-1, self._keep * (self._len + 1))[: self._n]
            tokens = flat[:, : self._keep * self._len].reshape(self._n, self._keep, self._len)
            logprobs = flat[:, self._keep * self._len:].contiguous().view(torch.float32).reshape(self._n, self._keep)
            self._out = (tokens, logprobs)
        return self._out


def _tail_shape(tokens: Optional[torch.Tensor], group) -> Tuple[int, int, torch.device]:
    """(keep, L) of the token matrix.  A rank whose shard is empty has no tensor to read them from: rank 0's shard is
    never empty (n_clips >= 1), so it announces the shape."""
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    meta = torch.zeros(2, dtype=torch.int64, device=dev)
    if dist.get_rank(group) == 0:
        meta[0], meta[1] = tokens.shape[1], tokens.shape[2]
    dist.broadcast(meta, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    keep, length = (int(v) for v in meta.tolist())
    return keep, length, dev


def caption_sharded(caption_fn: Callable[[int, int], Tuple[torch.Tensor, torch.Tensor]], n_clips: int, group=None,
                    async_op: bool = False, tail: Optional[Tuple[int, int]] = None):
    """Run ``caption_fn(begin, end) -> (tokens int32 [m, keep, L], logprobs fp32 [m, keep])`` on this rank's
    shard and all-gather the results in clip order.  Works with NCCL (device tensors) and gloo (CPU tensors).

    A rank whose shard is empty (n_clips < world, or the last ranks of a ragged split) does not call ``caption_fn`` and
    contributes padding.  ``tail=(keep, L)`` spares the shape broadcast that case otherwise needs.
    ``async_op=True``: returns a GatherHandle; the collective is enqueued behind the caption and not waited for."""
    if n_clips < 1:
        raise ValueError("caption_sharded: n_clips must be >= 1")
    if not dist.is_available() or not dist.is_initialized():
        tokens, logprobs = caption_fn(0, n_clips)
        if not async_op:
            return tokens, logprobs
        h = GatherHandle(None, torch.empty(0, dtype=torch.int32), n_clips, tokens.shape[1], tokens.shape[2])
        h._out = (tokens, logprobs)
        return h
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = (n_clips + world - 1) // world
    b, e = shard_range(n_clips, rank, world)
    tokens = logprobs = None
    if e > b:
        tokens, logprobs = caption_fn(b, e)
    if tail is not None:
        keep, length = tail
        dev = tokens.device if tokens is not None else (
            torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    elif (world - 1) * per >= n_clips:
        keep, length, dev = _tail_shape(tokens, group)   # some rank has an empty shard and cannot know the shape
    else:
        keep, length, dev = tokens.shape[1], tokens.shape[2], tokens.device
    width = keep * (length + 1)
    mine = torch.zeros(per, width, dtype=torch.int32, device=dev)
    if e > b:
        mine[: e - b, : keep * length] = tokens.reshape(e - b, keep * length).to(torch.int32)
        mine[: e - b, keep * length:] = logprobs.reshape(e - b, keep).to(torch.float32).contiguous().view(torch.int32)
    packed = torch.empty(world * per, width, dtype=torch.int32, device=dev)
    work = dist.all_gather_into_tensor(packed, mine, group=group, async_op=async_op)
    handle = GatherHandle(work if async_op else None, packed, n_clips, keep, length)
    return handle if async_op else handle.wait()


def all_reduce_bucket(flat: torch.Tensor, begin: int, end: int, group=None):
    """Start the SUM all-reduce of ``flat[begin:end]`` (a view: reduced in place) and return the work handle, or None when
    no process group is initialised.  The distillation step (reference: Lightning DDP, train.py:217-221)
    calls it once for the vocabulary-head bucket as soon as that part of the backward pass is enqueued and once for the
    rest, so the first reduction overlaps the remaining backward kernels; ``finish_all_reduce`` waits for both."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1 or end <= begin:
        return None
    return dist.all_reduce(flat[begin:end], op=dist.ReduceOp.SUM, group=group, async_op=True)


def finish_all_reduce(handles, group=None) -> float:
    """Wait for the bucket reductions; returns the factor that turns the summed gradients into the DDP average."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    for h in handles:
        if h is not None:
            h.wait()
    return 1.0 / world
