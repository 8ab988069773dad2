"""Persistent single-clip decode kernel (csrc/decode_mega.cu): the whole caption search of ONE clip as one cooperative
launch.  Its contract is bit-identity with the launch-per-kernel sequence (which the other GPU tests compare with the oracle):
same tokens, same log-probabilities to the last bit, for greedy and beam search, both cache-reorder modes, every key-split
geometry (1 / 2 / 6 frames), the early-exit branch (model.py:640) and the streaming captioner; plus the oracle directly."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import git_oracle as go  # noqa: E402
from oracle import search_oracle as so  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g():
    import gitb200
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return gitb200


def make_engine(g, n_frames, tied, seed, eos_boost=0.0):
    cfg = go.GitConfig(num_image_with_embedding=n_frames, tie_output=tied)
    sd = go.init_state_dict(cfg, seed=seed, temporal_std=0.02, perturb=True)
    if eos_boost:
        sd["textual.output.bias"] = sd["textual.output.bias"].clone()
        sd["textual.output.bias"][cfg.eos_index] += eos_boost
    eng = g.Engine(g.make_config({"num_image_with_embedding": n_frames}, cfg.sos_index, cfg.eos_index), 0)
    eng.load_state_dict(sd)
    return cfg, sd, eng


def both_ways(eng, frames, sp):
    """(tokens, logprobs, launches) with the persistent kernel and with the launch sequence."""
    out = []
    for persistent in (True, False):
        eng.set_persistent_decode(persistent)
        eng.launch_count(reset=True)
        tok, lp, _ = eng.caption(frames, sp)
        torch.cuda.synchronize()
        out.append((tok.cpu(), lp.cpu(), eng.launch_count(reset=True)))
    eng.set_persistent_decode(True)
    return out


@pytest.mark.parametrize("n_frames", [1, 2, 6])
def test_persistent_decode_is_bit_identical_to_the_launch_sequence(g, n_frames):
    # untied head: top-2 gaps of ~0.2 sigma, so the beams really branch; 197 / 394 / 1182 visual keys = 3 / 6 / 16 key splits
    cfg, sd, eng = make_engine(g, n_frames, tied=False, seed=70 + n_frames)
    gen = torch.Generator().manual_seed(5 + n_frames)
    for clip in range(2):
        frames = torch.randn(1, n_frames, 3, 224, 224, generator=gen).cuda()
        for nb in (1, 2, 3, 4):
            for reorder in (False, True):
                for max_steps in (6, 15) if nb in (1, 4) else (9,):
                    sp = g.SearchConfig(beam_size=nb, max_steps=max_steps, reorder_cache=reorder)
                    (t_p, l_p, n_p), (t_l, l_l, n_l) = both_ways(eng, frames, sp)
                    assert torch.equal(t_p, t_l), (n_frames, nb, reorder, max_steps, t_p, t_l)
                    assert torch.equal(l_p, l_l), (n_frames, nb, reorder, max_steps, l_p, l_l)
                    assert n_p < n_l and n_l - n_p >= 30 * (max_steps - 1), (n_p, n_l)  # the step loop is one launch


def test_persistent_decode_matches_the_oracle(g):
    cfg, sd, eng = make_engine(g, 2, tied=True, seed=81)
    frames = torch.randn(1, 2, 3, 224, 224, generator=torch.Generator().manual_seed(9))
    with torch.no_grad():
        vf = go.encode_clip(sd, cfg, frames[0])
    for nb in (1, 4):
        sp = g.SearchConfig(beam_size=nb, max_steps=8)
        eng.set_persistent_decode(True)
        tok, lp, _ = eng.caption(frames.cuda(), sp)
        with torch.no_grad():
            ref = so.infer(sd, cfg, vf, beam_size=nb, max_steps=8, save_logits=False)
        if nb == 1:  # tied head: the copy margin makes the greedy sequence exact
            assert torch.equal(tok[:, 0].cpu().long(), ref["predictions"])
            assert torch.allclose(lp.cpu(), ref["logprobs"], atol=0.02, rtol=0.01)
        else:
            assert torch.allclose(lp[:, :1].cpu(), ref["logprobs"][:, :1], atol=0.05, rtol=0.02)


def test_persistent_decode_leaves_its_loop_when_the_clip_is_done(g):
    """EOS-boosted head: the search finishes after a few steps; the kernel breaks out of its step loop on the device
    (model.py:640) and returns exactly what the launch sequence (polled early exit, and no early exit) returns."""
    cfg, sd, eng = make_engine(g, 2, tied=False, seed=62, eos_boost=4.0)
    gen = torch.Generator().manual_seed(8)
    for clip in range(3):
        frames = torch.randn(1, 2, 3, 224, 224, generator=gen).cuda()
        for nb in (1, 4):
            sp = g.SearchConfig(beam_size=nb, max_steps=20)
            (t_p, l_p, _), (t_l, l_l, _) = both_ways(eng, frames, sp)
            eng.set_persistent_decode(False)
            eng.set_early_exit(0)
            t_f, l_f, _ = eng.caption(frames, sp)
            eng.set_early_exit(4)
            eng.set_persistent_decode(True)
            assert torch.equal(t_p, t_l) and torch.equal(l_p, l_l)
            assert torch.equal(t_p, t_f.cpu()) and torch.equal(l_p, l_f.cpu())


def test_persistent_decode_under_cuda_graph_replay_and_streaming(g):
    """Latency mode end to end: the small-batch call on a side stream is captured into a CUDA graph (memset of the barrier
    counter + the cooperative launch are graph nodes) and replayed; the streaming captioner decodes through the same kernel."""
    cfg, sd, eng = make_engine(g, 2, tied=False, seed=91)
    frames = torch.randn(1, 2, 3, 224, 224, generator=torch.Generator().manual_seed(3)).cuda()
    sp = g.SearchConfig(beam_size=4, max_steps=10)
    eng.set_persistent_decode(False)
    t_ref, l_ref, _ = eng.caption(frames, sp)
    t_ref, l_ref = t_ref.cpu(), l_ref.cpu()
    eng.set_persistent_decode(True)
    import ctypes
    tok = torch.empty_like(t_ref, device="cuda")
    lp = torch.empty_like(l_ref, device="cuda")
    side = torch.cuda.Stream()
    c = sp.to_c()
    before = eng.lib.gitb200_graph_launches(eng.h)
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        for i in range(5):  # eager (sizing), capture + launch, replays
            rc = eng.lib.gitb200_caption(eng.h, ctypes.c_void_p(frames.data_ptr()), 1, frames.shape[1], ctypes.byref(c),
                                         ctypes.c_void_p(tok.data_ptr()), ctypes.c_void_p(lp.data_ptr()), None,
                                         ctypes.c_void_p(side.cuda_stream))
            assert rc == 0, eng.lib.gitb200_last_error(eng.h)
            side.synchronize()
            assert torch.equal(tok.cpu(), t_ref) and torch.equal(lp.cpu(), l_ref), i
            tok.zero_()
    assert eng.lib.gitb200_graph_launches(eng.h) - before >= 3  # the cooperative launch did not break graph capture
    # streaming window: push the two frames, caption; equals the batch call on the same frames
    eng.stream_reset()
    for f in range(2):
        eng.stream_push(frames[0, f])
    res = []
    for persistent in (True, False):
        eng.set_persistent_decode(persistent)
        tok, lp = eng.stream_caption(sp)
        res.append((tok.cpu(), lp.cpu()))
    eng.set_persistent_decode(True)
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    # the window's frames were encoded one by one (other GEMM tiles than the 2-frame batch): same tokens, log-probabilities within
    # the batch-invariance tolerance of test_full_size_properties_bench_geometry
    assert torch.equal(res[0][0], t_ref) and torch.allclose(res[0][1], l_ref, rtol=2e-2, atol=5e-3)
