"""Decode-step attention alone (development script, GPU box): microseconds per launch and fraction of the copy bandwidth
for greedy (1 row per clip) and beam-4 geometries.  Inputs exceed the L2 (0.9-1.9 GB of visual K/V per launch)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
eng = importlib.import_module("real-time-video-captioning_b200.engine")


def main():
    dev = torch.device("cuda", 0)
    cases = [(512, 1, 12, 1182, 8, 1), (256, 4, 12, 1182, 10, 1), (256, 4, 12, 1182, 10, 2), (256, 4, 12, 1182, 10, 3), (256, 2, 12, 1182, 10, 1),
             (128, 4, 16, 1542, 10, 1)]
    if "--one" in sys.argv:
        cases = cases[1:2]
    for n_clips, rpc, heads, n_vis, n_text, splits in cases:
        W, rows = heads * 64, n_clips * rpc
        q = torch.randn(rows, W, device=dev).bfloat16()
        vis = torch.randn(n_clips, n_vis, 2 * W, device=dev).bfloat16()
        txt = torch.randn(n_text, rows, 2 * W, device=dev).bfloat16()
        base = torch.arange(rows, device=dev).div(rpc, rounding_mode="floor") * rpc
        anc = (base[:, None] + torch.randint(0, rpc, (rows, n_text), device=dev)).int()
        for _ in range(5):
            eng.op_text_attention(q, vis, txt, anc, n_clips, rpc, heads, 0.125, splits)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            eng.op_text_attention(q, vis, txt, anc, n_clips, rpc, heads, 0.125, splits)
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / 20 * 1e3
        gb = (vis.numel() + txt.numel()) * 2 / 1e9
        print(f"clips {n_clips} rows/clip {rpc} heads {heads} n_vis {n_vis} splits {splits}: {us:8.1f} us per launch, {gb / us * 1e6:7.0f} GB/s algorithmic", flush=True)


if __name__ == "__main__":
    main()
