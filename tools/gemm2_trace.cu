// Phase tracer of the CTA-pair GEMM (development tool): per-tile timestamps of the MMA warp and one epilogue warp of CTA 0.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DGEMM2_TRACE -o build/gemm2_trace tools/gemm2_trace.cu \
//        real-time-video-captioning_b200/csrc/gemm_tcgen05.cu real-time-video-captioning_b200/csrc/gemv_skinny.cu tests/cuda/note_launch_stub.cu -lcuda
#include <cstdio>
#include <cstdlib>

#include "../real-time-video-captioning_b200/csrc/gemm2_tcgen05.cu"

static void run(int M, int N, int K, int act, int use_res) {
  bf16 *A, *W, *out;
  float* bias;
  cudaMalloc(&A, (size_t)M * K * 2);
  cudaMalloc(&W, (size_t)N * K * 2);
  cudaMalloc(&out, (size_t)M * N * 2);
  cudaMalloc(&bias, (size_t)N * 4);
  cudaMemset(A, 0, (size_t)M * K * 2);
  cudaMemset(W, 0, (size_t)N * K * 2);
  cudaMemset(out, 0, (size_t)M * N * 2);
  cudaMemset(bias, 0, (size_t)N * 4);
  GemmArgs g;
  g.A = A; g.lda = K; g.W = W; g.ldw = K; g.M = M; g.N = N; g.K = K; g.bias = bias; g.act = act;
  g.residual = use_res ? out : nullptr; g.ldr = N; g.out = out; g.ldo = N;
  for (int i = 0; i < 3; ++i) gemm2_bf16(g, 0);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int i = 0; i < 10; ++i) gemm2_bf16(g, 0);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  ms /= 10;
  printf("M=%d N=%d K=%d act=%d res=%d: %.3f ms %.1f TFLOP/s (%s)\n", M, N, K, act, use_res, ms, 2.0 * M * N * K / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
#ifdef GEMM2_TRACE
  static long long t[10 * 16 * 8];
  cudaMemcpyFromSymbol(t, g_gemm2_trace, sizeof(t));
  const long long t0 = t[(1 * 16 + 0) * 8 + 0];
  printf("tile | MMA: start tempty first_kb issued_all | EPI warp2: start tfull | blk0: buf_free tmem_ld stored | blk1: buf_free tmem_ld stored\n");
  for (int it = 0; it < 10; ++it) {
    const long long* m = &t[(1 * 16 + it) * 8];
    const long long* e = &t[(2 * 16 + it) * 8];
    printf("%2d | %7lld %7lld %7lld %7lld | %7lld %7lld | %7lld %7lld %7lld | %7lld %7lld %7lld\n", it, m[0] - t0, m[1] - t0, m[2] - t0, m[3] - t0,
           e[0] - t0, e[1] - t0, e[2] - t0, e[3] - t0, e[4] - t0, e[5] - t0, e[6] - t0, e[7] - t0);
  }
#endif
  cudaFree(A); cudaFree(W); cudaFree(out); cudaFree(bias);
}

int main() {
  const int M = 1182 * 64;
  run(M, 3072, 768, ACT_QUICK_GELU, 0);
  run(M, 2304, 768, ACT_NONE, 0);
  run(M, 768, 768, ACT_NONE, 1);
  run(M, 768, 3072, ACT_NONE, 1);
  return 0;
}
