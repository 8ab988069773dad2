"""Development script: one small caption pass that touches every kernel family (CTA-pair GEMM with >= 1024 rows, 1-CTA GEMM,
skinny GEMM, both tcgen05 attention shapes, decode-step attention with split keys, LayerNorm, search, preprocessing,
sub-batch sweeps, host path).  Written for `compute-sanitizer --tool memcheck`, which is closed on the GPU pool (runs under it
left GPUs needing a reset), so it serves as a plain smoke pass: device and host paths must agree token for token."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import git_oracle as go  # noqa: E402  (weight initialiser only)

g = importlib.import_module("real-time-video-captioning_b200")


def main():
    F = 2
    param = {"num_image_with_embedding": F}
    ocfg = go.GitConfig.from_param(param)
    sd = go.init_state_dict(ocfg, seed=0, temporal_std=0.02, perturb=True)
    eng = g.Engine(g.make_config(param, ocfg.sos_index, ocfg.eos_index), 0)
    eng.load_state_dict(sd)
    gen = torch.Generator().manual_seed(3)
    frames = torch.randn(7, F, 3, 224, 224, generator=gen)          # 7 x 394 = 2758 rows: CTA-pair GEMMs, ragged tiles
    eng.set_sweep_rows(3 * F * 197)                                  # sub-batches of 3 + 3 + 1 clips (tail: 1-CTA GEMM)
    for beam in (1, 4):
        sp = g.SearchConfig(beam_size=beam, max_steps=6, reorder_cache=(beam > 1))
        tok, lp, _ = eng.caption(frames.cuda(), sp)
        th, lh = eng.caption_host(frames.pin_memory(), sp, chunk_clips=4)
        torch.cuda.synchronize()
        assert torch.equal(tok.cpu(), th)
    one, _, _ = eng.caption(frames[:1].cuda(), g.SearchConfig(beam_size=1, max_steps=6))   # skinny-GEMM decode rows
    logits, vf, hidden = eng.forward_logits(frames[:2].cuda(), torch.full((2, 4), 1012, dtype=torch.int32))
    torch.cuda.synchronize()
    print("memcheck case done:", tok[:, 0].tolist()[:2], one[0, 0].tolist(), tuple(logits.shape), eng.launch_count(), "launches")


if __name__ == "__main__":
    main()
