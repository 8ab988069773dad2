"""Single-clip latency breakdown (development script): graph-replayed caption at several max_steps -> per-step slope and
the encode + prefill intercept."""
import importlib
import os
import statistics
import sys
import ctypes as ct

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
    one = torch.randn(1, 6, 3, 224, 224, device=dev)
    side = torch.cuda.Stream(dev)
    for ms in (2, 3, 8, 15, 25):
        sp = g.SearchConfig(beam_size=1, max_steps=ms)
        csp = sp.to_c()
        tok = torch.empty(1, 1, ms, dtype=torch.int32, device=dev)
        lp = torch.empty(1, 1, dtype=torch.float32, device=dev)
        lat = []
        torch.cuda.synchronize()
        with torch.cuda.stream(side):
            for i in range(30):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(side)
                rc = eng.lib.gitb200_caption(eng.h, ct.c_void_p(one.data_ptr()), 1, 6, ct.byref(csp), ct.c_void_p(tok.data_ptr()),
                                             ct.c_void_p(lp.data_ptr()), None, ct.c_void_p(side.cuda_stream))
                assert rc == 0
                b.record(side)
                b.synchronize()
                if i >= 8:
                    lat.append(a.elapsed_time(b))
        print(f"max_steps={ms:2d}: p50 {statistics.median(lat):.3f} ms  ({eng.launch_count()} launches so far)")


if __name__ == "__main__":
    main()
