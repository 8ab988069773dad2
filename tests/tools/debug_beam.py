import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
g = importlib.import_module("real-time-video-captioning_b200")
from oracle import git_oracle as go, search_oracle as so
param = {"image_encoder_type": "CLIPViT_L_14", "visual_feature_size": 1024, "num_image_with_embedding": 2}
cfg = go.GitConfig.from_param(param)
sd = go.init_state_dict(cfg, seed=31, temporal_std=0.02, perturb=True)
eng = g.Engine(g.make_config(param, cfg.sos_index, cfg.eos_index), 0)
eng.load_state_dict(sd)
frames = torch.randn(2, 3, 3, 224, 224, generator=torch.Generator().manual_seed(3))[:1]
nb, ms = 4, 6
for reorder in (True, False):
    sp = g.SearchConfig(beam_size=nb, max_steps=ms, reorder_cache=reorder)
    tok, lp, logits = eng.caption(frames.cuda(), sp, save_logits=True)
    logits = logits.cpu()[:, :, :cfg.vocab_size]
    t = {"i": 0}
    def step(ids):
        s = logits[t["i"]]; t["i"] += 1; return s.clone()
    dec, olp, _ = so.search(torch.full((1, 1), 101, dtype=torch.long), step, eos_index=102, max_steps=ms, beam_size=nb, length_penalty=0.6, save_logits=False)
    print("reorder", reorder, "engine", tok[0, 0].tolist(), lp.tolist(), "| oracle search on ENGINE logits", dec.tolist(), olp.tolist())
    with torch.no_grad():
        rvf = go.encode_clip(sd, cfg, frames[0])
        ref = so.infer(sd, cfg, rvf, beam_size=nb, max_steps=ms, reorder_cache=reorder, save_logits=True)
    rl = torch.from_numpy(np.array(ref["logits_dict"]))
    print("   oracle", ref["predictions"].tolist(), ref["logprobs"].tolist())
    for st in range(ms - 1):
        d = (logits[st] - rl[st]).abs().max(dim=1).values
        e_top = torch.log_softmax(logits[st], -1).topk(3, dim=-1)
        o_top = torch.log_softmax(rl[st], -1).topk(3, dim=-1)
        print("   step", st, "max|dlogit| per row", [round(x, 3) for x in d.tolist()])
        print("      eng top3", [[(int(i), round(float(v), 2)) for i, v in zip(e_top.indices[r], e_top.values[r])] for r in range(nb)])
        print("      ora top3", [[(int(i), round(float(v), 2)) for i, v in zip(o_top.indices[r], o_top.values[r])] for r in range(nb)])
