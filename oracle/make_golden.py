"""Generate tests/golden/*.npz from the CPU oracle (TEST INFRASTRUCTURE; see oracle/git_oracle.py header).

The reference ships no golden vectors for this path and cannot be imported here, so these
fixtures pin the ORACLE (after it has been cross-checked against transformers.GitForCausalLM by
tests/test_oracle_vs_hf.py): a later change to the oracle, to torch's kernels, or to the CUDA path shows up as
a diff against frozen numbers.  Everything is seeded; re-running this script must reproduce the files.

    python -m oracle.make_golden            # writes tests/golden/git_base_f2.npz (~1.5 MB)
"""
import os

import numpy as np
import torch

from . import git_oracle as go
from . import search_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SPEC = dict(weights_seed=21, frames_seed=22, tokens_seed=23, n_frames=2, n_clips=2, caption_len=6, max_steps=7)


def inputs(tied=True):
    cfg = go.GitConfig(num_image_with_embedding=SPEC["n_frames"], tie_output=tied)
    sd = go.init_state_dict(cfg, seed=SPEC["weights_seed"], temporal_std=0.02, perturb=True)
    frames = torch.randn(SPEC["n_clips"], SPEC["n_frames"], 3, 224, 224, generator=torch.Generator().manual_seed(SPEC["frames_seed"]))
    tokens = torch.randint(1000, 30000, (SPEC["n_clips"], SPEC["caption_len"]), generator=torch.Generator().manual_seed(SPEC["tokens_seed"]))
    tokens[:, 0] = cfg.sos_index
    return cfg, sd, frames, tokens


def compute():
    torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    out = {}
    cfg, sd, frames, tokens = inputs(True)
    cols = torch.arange(0, cfg.vocab_size, 97)  # every 97th vocabulary column keeps the file small
    with torch.no_grad():
        for b in range(SPEC["n_clips"]):
            logits, vf, hidden = go.forward_one_custom(sd, cfg, frames[b], tokens[b:b + 1])
            out[f"vf_rows_{b}"] = vf[0, ::37].numpy()                       # every 37th visual token, all channels
            out[f"vf_sum_{b}"] = np.array([vf.double().sum().item(), vf.double().abs().sum().item()])
            out[f"logits_cols_{b}"] = logits[0][:, cols].numpy()
            out[f"logits_argmax_{b}"] = logits[0].argmax(-1).numpy()
            out[f"logits_stats_{b}"] = np.array([logits.double().mean().item(), logits.double().std().item()])
            out[f"hidden_text_{b}"] = hidden[:, -SPEC["caption_len"]:, ::16].numpy()  # text rows of the 7 states
            out[f"hidden_norms_{b}"] = hidden.double().flatten(1).norm(dim=1).numpy()
        vf_all = torch.cat([go.encode_clip(sd, cfg, f) for f in frames])
        for nb in (1, 4):
            r = so.infer(sd, cfg, vf_all, beam_size=nb, max_steps=SPEC["max_steps"], save_logits=False)
            out[f"tokens_beam{nb}"] = r["predictions"].numpy()
            out[f"logprobs_beam{nb}"] = r["logprobs"].numpy()
    out["vocab_cols"] = cols.numpy()
    return out


def main():
    out = compute()
    path = os.path.join(ROOT, "tests", "golden", "git_base_f2.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
