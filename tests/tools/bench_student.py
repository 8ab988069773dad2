"""Student decoder micro-benchmark: cached greedy decode on the GPU library vs the oracle (the reference's
stock torch.nn decoder with full re-decode per step) on the host cores.  Usage: python tests/tools/bench_student.py [B] [max_len]"""
import importlib
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import student_oracle as st  # noqa: E402

g = importlib.import_module("real-time-video-captioning_b200")


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    max_len = int(sys.argv[2]) if len(sys.argv) > 2 else 25   # real_time_inference.py:58
    cfg = st.StudentConfig()
    m = st.init_student(cfg, seed=8)
    s = g.StudentCandidateV1(None, cfg.d_model, cfg.n_head, cfg.d_ffn, cfg.dropout, cfg.num_decoder_layers, cfg.vocab_length,
                             cfg.cls_token_id, cfg.sep_token_id)
    s.load_state_dict(st.state_dict_of(m))
    s = s.to("cuda")
    mem = torch.randn(B, 6, cfg.d_model).cuda()
    for _ in range(3):
        s.greedy_decode_from_memory(mem, max_len)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    iters = 10
    for _ in range(iters):
        s.greedy_decode_from_memory(mem, max_len)
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / iters
    print(f"student greedy decode (GPU, K/V cache): B={B} max_len={max_len}: {ms:.3f} ms  {B / ms * 1e3:.0f} captions/s  "
          f"{ms / max_len * 1e3:.1f} us/step")
    one = mem[:1]
    for _ in range(3):
        s.greedy_decode_from_memory(one, max_len)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        s.greedy_decode_from_memory(one, max_len)
    torch.cuda.synchronize()
    print(f"single clip latency: {(time.perf_counter() - t0) * 100:.2f} ms")
    torch.set_num_threads(os.cpu_count())
    nb = min(B, 8)
    t0 = time.perf_counter()
    m.greedy_decode_from_memory(mem[:nb].cpu(), max_len)
    dt = time.perf_counter() - t0
    print(f"oracle (stock torch.nn decoder, re-decode per step, {torch.get_num_threads()} threads): {nb} clips in {dt * 1e3:.1f} ms  "
          f"{nb / dt:.1f} captions/s")


if __name__ == "__main__":
    main()
