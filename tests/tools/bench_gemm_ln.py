"""gemm2 LNOUT vs plain GEMM + LayerNorm kernel at the bench's sub-batch shape (development tool, GPU box)."""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
eng = importlib.import_module("real-time-video-captioning_b200.engine")


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


M = int(sys.argv[1]) if len(sys.argv) > 1 else 151296
for N, K in ((768, 768), (768, 3072), (1024, 1024), (1024, 4096)):
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda").bfloat16()
    gamma, beta = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
    x = torch.randn(M, N, device="cuda").bfloat16()
    t_g = t(lambda: eng.op_gemm(a, w, bias, r, 0))
    t_l = t(lambda: eng.op_layernorm(x, gamma, beta, 1e-5))
    t_f = t(lambda: eng.op_gemm_ln(a, w, bias, r, gamma, beta, 1e-5))
    fl = 2.0 * M * N * K
    print(f"M={M} N={N} K={K}: gemm {t_g:7.1f} us ({fl / t_g / 1e6:6.0f} TF/s)  layernorm {t_l:6.1f} us  gemm+ln fused {t_f:7.1f} us  "
          f"(separate sum {t_g + t_l:7.1f}; fused saves {t_g + t_l - t_f:6.1f} us)", flush=True)
