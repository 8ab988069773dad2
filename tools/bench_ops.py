"""Micro-benchmarks of single operators through the C ABI (CUDA events, warm, L2-exceeding inputs)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
eng = importlib.import_module("real-time-video-captioning_b200.engine")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


def attention(B=64):
    for name, n_groups, glen, heads in (("vit", B * 6, 197, 12), ("dec", B, 1182, 12)):
        qkv = torch.randn(n_groups * glen, 3 * heads * 64, device="cuda").bfloat16()
        flops = 4.0 * glen * glen * 64 * heads * n_groups
        for legacy in (True, False):
            ms = timeit(lambda: eng.op_attention_groups(qkv, n_groups, glen, heads, 0.125, legacy_mma=legacy))
            print(f"attention {name} B={B} {'mma.sync' if legacy else 'tcgen05 '}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  ({ms * 1e3 / B:.1f} us/clip/layer)")




def layernorm(rows=605184):
    import torch.nn.functional as F  # noqa: F401
    for cols in (768, 1024):
        x = torch.randn(rows, cols, device="cuda").bfloat16()
        g = torch.ones(cols, device="cuda")
        b = torch.zeros(cols, device="cuda")
        ms = timeit(lambda: eng.op_layernorm(x, g, b, 1e-5))
        print(f"layernorm rows={rows} cols={cols}: {ms:.3f} ms  {2 * rows * cols * 2 / ms / 1e9:.2f} TB/s")


def decode_gemms():
    """Decode-step GEMM shapes (rows = clips x beams) per tile choice: kernel durations from CUPTI records (the launches are a
    few microseconds: host launch cost would dominate CUDA-event timing of eager calls)."""
    from torch.profiler import ProfilerActivity, profile
    for M in (512, 1024, 2048):
        for N, K in ((2304, 768), (768, 768), (3072, 768), (768, 3072)):
            a = torch.randn(M, K, device="cuda").bfloat16()
            w = torch.randn(N, K, device="cuda").bfloat16()
            bias = torch.zeros(N, device="cuda")
            res = torch.randn(M, N, device="cuda").bfloat16()
            line = f"M={M:5d} N={N:5d} K={K:5d}:"
            for tile in (0, 128, 256):
                for _ in range(3):
                    eng.op_gemm(a, w, bias, res, 0, False, tile)
                torch.cuda.synchronize()
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    for _ in range(10):
                        eng.op_gemm(a, w, bias, res, 0, False, tile)
                    torch.cuda.synchronize()
                ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "gemm" in e.name]
                us = sum(e.time_range.end - e.time_range.start for e in ev) / max(1, len(ev))
                name = ev[0].name.split("(")[0][-28:] if ev else "?"
                line += f"  tile {tile:3d}: {us:6.1f} us [{name}]"
            print(line, flush=True)


def vocab_gemm():
    """Vocabulary head (fp32 logits, N = 30720 padded columns, K = 768) per tile width: CUPTI kernel durations."""
    from torch.profiler import ProfilerActivity, profile
    N, K = 30720, 768
    w = torch.randn(N, K, device="cuda").bfloat16()
    bias = torch.zeros(N, device="cuda")
    for M in (512, 1024, 2048):
        a = torch.randn(M, K, device="cuda").bfloat16()
        line = f"M={M:5d} N={N} K={K} fp32 out:"
        for tile in (0, 128, 256):
            for _ in range(3):
                eng.op_gemm(a, w, bias, None, 0, True, tile)
            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for _ in range(10):
                    eng.op_gemm(a, w, bias, None, 0, True, tile)
                torch.cuda.synchronize()
            ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "gemm" in e.name]
            us = sum(e.time_range.end - e.time_range.start for e in ev) / max(1, len(ev))
            line += f"  tile {tile:3d}: {us:6.1f} us = {2.0 * M * N * K / us / 1e6:6.0f} TFLOP/s"
        print(line, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "vocab_gemm":
        vocab_gemm()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "decode_gemms":
        decode_gemms()
        sys.exit(0)
    if len(sys.argv) > 2 and sys.argv[2] == "ln":
        layernorm(int(sys.argv[1]) * 6 * 197)
    else:
        attention(int(sys.argv[1]) if len(sys.argv) > 1 else 64)
