// Standalone micro-benchmark / phase tracer of the tcgen05 attention kernel (development tool, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DATTN_TRACE=<group>] -o build/attn_bench tools/attn_bench.cu \
//        real-time-video-captioning_b200/csrc/gemm_tcgen05.cu tests/cuda/note_launch_stub.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#ifdef ATTN_SRC  // A/B against an older revision of the kernel: -DATTN_SRC='"/tmp/old/attention_tc_v5.cu"' -I real-time-video-captioning_b200/csrc
#include ATTN_SRC
#else
#include "../real-time-video-captioning_b200/csrc/attention_tc.cu"
#endif

static void run(const char* name, int n_groups, int glen, int heads, int iters) {
  const size_t rows = (size_t)n_groups * glen, W = (size_t)heads * 64;
  std::vector<bf16> h(rows * 3 * W);
  unsigned s = 12345u;
  for (auto& x : h) {
    s = s * 1664525u + 1013904223u;
    float u = 0.f;
    for (int k = 0; k < 4; ++k) { s = s * 1664525u + 1013904223u; u += (float)(s >> 8) / 16777216.f; }
    x = __float2bfloat16((u - 2.f) * 1.7320508f);  // ~N(0,1)
  }
  bf16 *qkv, *out;
  cudaMalloc(&qkv, h.size() * sizeof(bf16));
  cudaMalloc(&out, rows * W * sizeof(bf16));
  cudaMemcpy(qkv, h.data(), h.size() * sizeof(bf16), cudaMemcpyHostToDevice);
  for (int i = 0; i < 3; ++i) attention_groups_tc(qkv, 3 * W, out, W, n_groups, glen, heads, 0.125f, 0);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < iters; ++i) attention_groups_tc(qkv, 3 * W, out, W, n_groups, glen, heads, 0.125f, 0);
  cudaEventRecord(b);
  cudaError_t e = cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  ms /= iters;
  const double flops = 4.0 * glen * (double)glen * 64 * heads * n_groups;
  printf("%s groups=%d len=%d: %.3f ms  %.1f TFLOP/s  (%s)\n", name, n_groups, glen, ms, flops / ms / 1e9, cudaGetErrorString(e));
#ifdef ATTN_TRACE
  static long long t[12 * 32 * 12];
  cudaMemcpyFromSymbol(t, g_attn_trace, sizeof(t));
  const int nb = (glen + 63) / 64;
  long long t0 = 0;
  for (int w = 4; w < 12; ++w) {
    const long long v = t[(w * 32 + 0) * 12];
    if (v && (!t0 || v < t0)) t0 = v;
  }
  printf("clk since the pair's first block started | softmax warp: s_ready ld_done max_done exp_or_wait_done p_done\n");
  for (int j = 0; j < nb && j < 30; ++j) {
    for (int g = 0; g < 2; ++g) {
      const long long* r = &t[((1 + g) * 32 + j) * 12];
      printf("blk %2d g%d S-issue: start %lld operands_ready %lld mma_issued %lld committed %lld | PV-issue: wait_start %lld p_full %lld committed %lld\n",
             j, g, r[0] - t0, r[1] - t0, r[2] - t0, r[3] - t0, r[5] - t0, r[6] - t0, r[7] - t0);
      for (int w = 4 + 4 * g; w < 8 + 4 * g; ++w) {
        r = &t[(w * 32 + j) * 12];
        printf("blk %2d g%d warp %2d:", j, g, w);
        for (int p = 0; p < 5; ++p) printf(" %7lld", r[p] ? r[p] - t0 : -1);
        printf("\n");
      }
    }
  }
  for (int w = 4; w < 12; ++w) {
    const long long* r = &t[(w * 32 + 31) * 12];
    printf("epilogue warp %2d: enter %lld last_product_done %lld stored %lld\n", w, r[5] - t0, r[6] - t0, r[10] - t0);
  }
#endif
  cudaFree(qkv); cudaFree(out);
}

int main(int argc, char** argv) {
  const int B = argc > 1 ? atoi(argv[1]) : 64;
  const int which = argc > 2 ? atoi(argv[2]) : 3;
  const int iters = argc > 3 ? atoi(argv[3]) : 10;
  if (which & 1) run("vit", B * 6, 197, 12, iters);
  if (which & 2) run("dec", B, 1182, 12, iters);
  if (which & 4) run("vit-L", B * 6, 257, 16, iters);
  if (which & 8) run("dec-L24", B / 8 > 0 ? B / 8 : 1, 6168, 12, iters);
  return 0;
}
