"""Where GenerativeImageTextTeacher.forward spends its time (development tool, GPU box)."""
import importlib
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
g = importlib.import_module("real-time-video-captioning_b200")
gm = importlib.import_module("real-time-video-captioning_b200.model")

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
teacher = gm.GenerativeImageTextTeacher.from_random_init({"num_image_with_embedding": 6})
m = teacher.model
eng = m.engine()
sp = g.SearchConfig(beam_size=4, max_steps=15)
eng.reserve(B, 6, 4, 15)
x = torch.randn(B, 6, 3, 224, 224).pin_memory()


def t(fn, n=3, label=""):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        r = fn()
        del r
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n * 1e3
    print(f"{label:60s} {dt:8.2f} ms  ({B / dt * 1e3:7.1f} clips/s)", flush=True)


t(lambda: eng.caption_host(x, sp, chunk_clips=64), label="Engine.caption_host (host tokens)")
t(lambda: eng.caption_from_host(x, sp, chunk_clips=64), label="Engine.caption_from_host (device tokens)")
t(lambda: eng.caption_from_host(x, sp, chunk_clips=64, save_logits=True), label="  + save_logits")
t(lambda: eng.caption_from_host(x, sp, chunk_clips=64, save_logits=True, want_features=True), label="  + visual features fp32")
t(lambda: m.forward_host_frames(x), label="model.forward_host_frames")
t(lambda: teacher(x), label="teacher.forward")
xd = x.cuda()
t(lambda: teacher(xd), label="teacher.forward (device frames)")
t(lambda: eng.caption(xd, sp), label="Engine.caption (device frames)")
