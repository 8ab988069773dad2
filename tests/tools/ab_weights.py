"""Development script: does the random-init distribution change the step time?  Same box, same process: one engine with the
oracle's seeded initialiser (what bench.py used until the third session), one with the package's get_git_model init
(what it uses now), 512-clip greedy captions timed alternately; plus the split encode / decode time of each."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import git_oracle as go  # noqa: E402

g = importlib.import_module("real-time-video-captioning_b200")
gm = importlib.import_module("real-time-video-captioning_b200.model")


def package_sd(param):
    tok = gm.SyntheticTokenizer()
    torch.manual_seed(0)
    model = gm.get_git_model(tok, param)
    gen_w = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for name, p_ in model.named_parameters():
            if name.endswith("bias") or "img_temperal_embedding" in name:
                p_.add_(torch.randn(p_.shape, generator=gen_w) * 0.02)
            elif p_.dim() == 1 and name.endswith("weight"):
                p_.add_(torch.randn(p_.shape, generator=gen_w) * 0.1)
    return model.state_dict()


def main():
    param = {"num_image_with_embedding": 6}
    ocfg = go.GitConfig.from_param(param)
    sds = {"oracle-init": go.init_state_dict(ocfg, seed=0, temporal_std=0.02, perturb=True), "package-init": package_sd(param)}
    for k in ("textual.embedding.words.weight", "textual.transformer.encoder.layer.0.attention.self.query.weight",
              "image_encoder.transformer.resblocks.0.attn.in_proj_weight", "image_encoder.conv1.weight",
              "textual.visual_projection.0.weight", "image_encoder.transformer.resblocks.0.mlp.c_fc.weight"):
        print(k, {n: round(float(sd[k].std()), 4) for n, sd in sds.items()})
    B = 512
    frames = torch.randn(B, 6, 3, 224, 224, device="cuda", generator=torch.Generator(device="cuda").manual_seed(100))
    sp = g.SearchConfig(beam_size=1, max_steps=15)
    engs = {}
    for n, sd in sds.items():
        e = g.Engine(g.make_config(param, 101, 102), 0)
        e.load_state_dict(sd)
        e.reserve(B, 6, 1, 15)
        engs[n] = e
    for rnd in range(3):
        for n, e in engs.items():
            for _ in range(2 if rnd == 0 else 0):
                e.caption(frames, sp)
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            e.encode(frames, want_features=False)
            b.record()
            tok, lp, _ = e.caption(frames, sp)
            c.record()
            c.synchronize()
            print(f"round {rnd} {n}: encode {a.elapsed_time(b):.1f} ms, full caption {b.elapsed_time(c):.1f} ms, "
                  f"tokens/clip ended by EOS: {(tok[:, 0] == 102).any(dim=-1).float().mean().item():.2f}")


if __name__ == "__main__":
    main()
