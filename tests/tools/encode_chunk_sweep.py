"""Development script: ViT encode + decoder visual pass time per clip as a function of the clips per call (does keeping a
chunk's activations L2-resident beat one 512-clip sweep per kernel?)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import git_oracle as go  # noqa: E402  (weight initialiser only)

g = importlib.import_module("real-time-video-captioning_b200")


def main():
    param = {"num_image_with_embedding": 6}
    ocfg = go.GitConfig.from_param(param)
    sd = go.init_state_dict(ocfg, seed=0, temporal_std=0.02, perturb=True)
    eng = g.Engine(g.make_config(param, ocfg.sos_index, ocfg.eos_index), 0)
    eng.load_state_dict(sd)
    dev = torch.device("cuda", 0)
    total = 512
    frames = torch.randn(total, 6, 3, 224, 224, device=dev)
    sp = g.SearchConfig(beam_size=1, max_steps=2)   # encode + visual pass + ONE decode step
    for chunk in (8, 16, 32, 64, 128, 256, 512):
        def run():
            for i in range(0, total, chunk):
                eng.caption(frames[i:i + chunk], sp)
        run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            run()
        b.record()
        b.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f"chunk {chunk:3d}: {ms:.1f} ms per {total} clips  ({total / ms * 1e3:.0f} clips/s encode+prefill)")


if __name__ == "__main__":
    main()
