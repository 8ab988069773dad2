"""N>1 host logic on CPU: two gloo ranks shard the clips, each 'captions' its shard, and the all-gather must
return the clips in order and identical on every rank (same tokens at 1/2/4/8 GPUs)."""
import importlib
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_clips, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = importlib.import_module("real-time-video-captioning_b200")

    def fake_caption(b, e):  # token row i encodes the clip id: any mis-ordering is visible
        assert e > b, "an empty shard must not reach the caption function (Engine.caption rejects 0 clips)"
        ids = torch.arange(b, e)
        return (ids.view(-1, 1, 1) * 10 + torch.arange(4).view(1, 1, 4)).int(), -ids.view(-1, 1).float() - 0.25

    tok, lp = g.caption_sharded(fake_caption, n_clips)
    # the asynchronous form (bench.py: gathers of consecutive batches in flight, waited for once) must agree
    handles = [g.caption_sharded(fake_caption, n_clips, async_op=True, tail=(1, 4)) for _ in range(3)]
    for h in handles:
        t2, l2 = h.wait()
        assert torch.equal(t2, tok) and torch.equal(l2, lp)
    q.put((rank, tok.flatten().tolist(), lp.flatten().tolist()))
    dist.destroy_process_group()


def _run(world, n_clips):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, n_clips, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_two_ranks_gather_in_clip_order():
    for n in (5, 2, 1):  # n = 1: rank 1's shard is empty
        out = _run(2, n)
        want_tok = [i * 10 + k for i in range(n) for k in range(4)]
        for rank, tok, lp in out:
            assert tok == want_tok, (rank, tok)
            assert lp == [-float(i) - 0.25 for i in range(n)]


def test_three_ranks_ragged_split_with_an_empty_shard():
    """4 clips on 3 ranks: shards [0,2), [2,4), [4,4) -- the last rank contributes padding only (it used to
    call the caption function with 0 clips, raise, and leave the others blocked in the collective)."""
    out = _run(3, 4)
    want_tok = [i * 10 + k for i in range(4) for k in range(4)]
    assert len(out) == 3
    for rank, tok, lp in out:
        assert tok == want_tok, (rank, tok)
        assert lp == [-float(i) - 0.25 for i in range(4)]
