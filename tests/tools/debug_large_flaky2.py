"""Probe 2: pollute the process like tests/test_gpu_baseline_geometry.py does, then look for run-to-run differences in the
GIT-large F=2 engine: hidden states of the teacher-forced forward (per layer, visual / text rows, per clip) and captions."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import gitb200 as g
from oracle import git_oracle as go
from importlib import import_module
engm = import_module("real-time-video-captioning_b200.engine")
pre = sys.argv[1] if len(sys.argv) > 1 else "attn,large6,large24"

if "attn" in pre:
    for n_groups, group_len, heads in [(2, 1542, 12), (1, 6168, 12), (1, 4097, 16)]:
        W = heads * 64
        qkv = torch.randn(n_groups * group_len, 3 * W, device="cuda").bfloat16()
        engm.op_attention_groups(qkv, n_groups, group_len, heads, 0.125)
for nf in (6, 24):
    if f"large{nf}" in pre:
        param = {"image_encoder_type": "CLIPViT_L_14", "visual_feature_size": 1024, "num_image_with_embedding": nf}
        cfg = go.GitConfig.from_param(param)
        sd = go.init_state_dict(cfg, seed=50 + nf, temporal_std=0.02, perturb=True)
        e = g.Engine(g.make_config(param, cfg.sos_index, cfg.eos_index), 0)
        e.load_state_dict(sd)
        fr = torch.randn(1, nf, 3, 224, 224, generator=torch.Generator().manual_seed(nf))
        e.forward_logits(fr.cuda(), torch.tensor([[101, 2023, 2003, 1037, 3231]]).cuda())
        e.caption(fr.cuda(), g.SearchConfig(beam_size=1, max_steps=6))
        del e
torch.cuda.synchronize()

param = {"image_encoder_type": "CLIPViT_L_14", "visual_feature_size": 1024, "num_image_with_embedding": 2}
cfg = go.GitConfig.from_param(param)
sd = go.init_state_dict(cfg, seed=31, temporal_std=0.02, perturb=True)
eng = g.Engine(g.make_config(param, cfg.sos_index, cfg.eos_index), 0)
eng.load_state_dict(sd)
raw = torch.randint(0, 256, (2, 3, 180, 240, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(29))
sp1 = g.SearchConfig(beam_size=1, max_steps=5)
rawp = raw.pin_memory()
dev = g.preprocess_frames(raw.view(-1, 180, 240, 3).cuda()).view(2, 3, 3, 224, 224).contiguous()
tokens = torch.tensor([[101, 101, 101, 101], [101, 101, 101, 101]]).cuda()
ref = None
for it in range(6):
    logits, vf, hidden = eng.forward_logits(dev, tokens)
    torch.cuda.synchronize()
    nv = vf.shape[1]
    cur = {"vf": vf.clone(), "logits": logits.clone(), "hidden": hidden.clone()}
    # step-wise decode of two fixed steps: logits of the decode-only path (rows = 2: the fused single-clip launch sequence)
    eng.encode(dev)
    eng.decode_begin(1)
    st0 = eng.decode_step(torch.tensor([101, 101]), 0).clone()
    st1 = eng.decode_step(torch.tensor([101, 101]), 1).clone()
    cur["st0"], cur["st1"] = st0, st1
    td, ld, _ = eng.caption(dev, sp1)
    th, lh = eng.caption_host_u8(rawp, sp1, chunk_clips=2)
    msg = [f"it {it} cap dev {[round(x, 4) for x in ld.cpu().flatten().tolist()]} host {[round(x, 4) for x in lh.flatten().tolist()]}"]
    if ref is None:
        ref = cur
    else:
        msg.append(f"vf_eq {torch.equal(cur['vf'], ref['vf'])} logits_maxdiff {[(cur['logits'][b] - ref['logits'][b]).abs().max().item() for b in range(2)]}")
        msg.append(f"  step0 logits maxdiff per row {[(cur['st0'][r] - ref['st0'][r]).abs().max().item() for r in range(2)]} step1 {[(cur['st1'][r] - ref['st1'][r]).abs().max().item() for r in range(2)]}")
        for l in range(0):
            dv = [(cur["hidden"][b, l, :nv] - ref["hidden"][b, l, :nv]).abs().max().item() for b in range(2)]
            dt = [(cur["hidden"][b, l, nv:] - ref["hidden"][b, l, nv:]).abs().max().item() for b in range(2)]
            msg.append(f"  hidden[{l}] visual maxdiff {dv} text maxdiff {dt}")
    print("\n".join(msg), flush=True)
