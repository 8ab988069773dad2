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
    import itertools
    for persistent, nb, ms in itertools.product((True, False), (1, 4), (2, 3, 8, 15, 25)):
        eng.set_persistent_decode(persistent)
        sp = g.SearchConfig(beam_size=nb, max_steps=ms)
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
        print(f"persistent={int(persistent)} beam={nb} max_steps={ms:2d}: p50 {statistics.median(lat):.3f} ms  min {min(lat):.3f}", flush=True)


def trace():
    """Per-phase SM cycles of CTA 0 inside the persistent decode kernel (work / barrier wait), greedy and beam 4."""
    import numpy as np
    param = {"num_image_with_embedding": 6}
    ocfg = go.GitConfig.from_param(param)
    sd = go.init_state_dict(ocfg, seed=0, temporal_std=0.02, perturb=True)
    eng = g.Engine(g.make_config(param, ocfg.sos_index, ocfg.eos_index), 0)
    eng.load_state_dict(sd)
    one = torch.randn(1, 6, 3, 224, 224, device="cuda")
    names = ["QKV", "attention", "out-proj", "fc1", "fc2", "vocab", "search rows", "search walk"]
    for nb in (1, 4):
        sp = g.SearchConfig(beam_size=nb, max_steps=15)
        eng.caption(one, sp)
        eng.lib.gitb200_debug_persistent_decode_trace(eng.h, None, 1)
        n = 5
        for _ in range(n):
            eng.caption(one, sp)
        out = np.zeros(32, dtype=np.uint64)
        eng.lib.gitb200_debug_persistent_decode_trace(eng.h, ct.c_void_p(out.ctypes.data), 0)
        steps = 14 * n
        tot = 0
        for i, nm in enumerate(names):
            per = 6 if i < 5 else 1
            w, bw = out[i] / steps, out[16 + i] / steps
            tot += w + bw
            print(f"beam={nb} {nm:10s}: work {w / per:8.0f} cyc  barrier wait {bw / per:8.0f} cyc per phase instance  ({per} per step; {w + bw:8.0f} cyc per step)")
        print(f"beam={nb} total {tot:.0f} cycles per step ({int(out[15])} launches); search-row sections (stats, thread best, rounds, list, rank) "
              f"{[int(out[8 + k] / steps) for k in range(5)]} cycles per step", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "trace":
        trace()
        sys.exit(0)
    main()
