"""Experiment: two engines (half batches) on two streams from two host threads -- does the decode phase of one half
overlap the encode phase of the other?"""
import importlib, os, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
g = importlib.import_module("real-time-video-captioning_b200")
from oracle import git_oracle as go

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = 6
cfg = go.GitConfig(num_image_with_embedding=6)
sd = go.init_state_dict(cfg, seed=0)
engs = []
for i in range(2):
    e = g.Engine(g.make_config({"num_image_with_embedding": 6}, 101, 102), 0)
    e.load_state_dict(sd)
    engs.append(e)
sp = g.SearchConfig(beam_size=1, max_steps=15)
frames = [torch.randn(B, 6, 3, 224, 224, device="cuda") for _ in range(2)]
streams = [torch.cuda.Stream() for _ in range(2)]

def run(i, n, offset_first=False):
    with torch.cuda.stream(streams[i]):
        for k in range(n):
            engs[i].caption(frames[i], sp)
    streams[i].synchronize()

for i in range(2):
    run(i, 2)
torch.cuda.synchronize()
# single engine, sequential
t0 = time.perf_counter(); run(0, steps); t1 = time.perf_counter()
print(f"one engine  B={B}: {B * steps / (t1 - t0):.1f} clips/s")
# two engines concurrently
t0 = time.perf_counter()
th = [threading.Thread(target=run, args=(i, steps)) for i in range(2)]
for t in th: t.start()
for t in th: t.join()
torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"two engines B={B} each, concurrent streams: {2 * B * steps / (t1 - t0):.1f} clips/s")
