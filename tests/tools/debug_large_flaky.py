"""Repro probe for the GIT-large device-vs-host-u8 caption log-prob mismatch (tests/test_gpu_parity.py:563)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gitb200 as g
from oracle import git_oracle as go

param = {"image_encoder_type": "CLIPViT_L_14", "visual_feature_size": 1024, "num_image_with_embedding": 2}
cfg = go.GitConfig.from_param(param)
sd = go.init_state_dict(cfg, seed=31, temporal_std=0.02, perturb=True)
poison = sys.argv[2] if len(sys.argv) > 2 else "none"
if poison != "none":
    torch.cuda.init()
    bufs = []
    for _ in range(24):  # 24 x 1 GiB of garbage, returned to the driver before the engine allocates
        t = torch.empty(512 * 1024 * 1024, dtype=torch.bfloat16, device="cuda")
        if poison == "randn":
            t.normal_(0, 3.0)
        elif poison == "nan":
            t.fill_(float("nan"))
        else:
            t.fill_(1.0)
        bufs.append(t)
    torch.cuda.synchronize()
    del bufs, t
    torch.cuda.empty_cache()
eng = g.Engine(g.make_config(param, cfg.sos_index, cfg.eos_index), 0)
eng.load_state_dict(sd)
raw = torch.randint(0, 256, (2, 3, 180, 240, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(29))
sp1 = g.SearchConfig(beam_size=1, max_steps=5)
rawp = raw.pin_memory()
frames = torch.randn(2, 3, 3, 224, 224, generator=torch.Generator().manual_seed(3))
tokens = torch.tensor([[101, 2023, 2003, 1037], [101, 7, 8, 9]])
mode = sys.argv[1] if len(sys.argv) > 1 else "ABC"

def final(tag):
    dev = g.preprocess_frames(raw.view(-1, 180, 240, 3).cuda()).view(2, 3, 3, 224, 224).contiguous()
    td, ld, _ = eng.caption(dev, sp1)
    ld = ld.cpu().flatten().tolist()
    th, lh = eng.caption_host_u8(rawp, sp1, chunk_clips=2)
    print(tag, "dev", td.cpu().flatten().tolist(), [round(x, 4) for x in ld], "host_u8", [round(x, 4) for x in lh.flatten().tolist()], flush=True)

final("fresh")
if "A" in mode:
    eng.forward_logits(frames.cuda(), tokens.cuda())
    final("after A (forward_logits 2 clips)")
if "B" in mode:
    for nb, reorder in ((1, False), (4, True), (4, False)):
        eng.caption(frames.cuda(), g.SearchConfig(beam_size=nb, max_steps=6, reorder_cache=reorder))
        final(f"after B caption nb={nb} reorder={reorder}")
if "C" in mode:
    for b in range(2):
        hyp = torch.tensor([101, 101, 24013, 24013, 24013])
        eng.forward_logits(frames[b:b + 1].cuda(), hyp[None].cuda(), want_hidden=False, want_features=False)
        final(f"after C forward_logits 1 clip b={b}")
for i in range(3):
    final(f"repeat {i}")
