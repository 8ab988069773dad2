// Standalone bring-up / regression binary for the tcgen05 GEMM (no torch).  Compares against a naive
// CUDA-core GEMM on the same bf16 inputs and prints achieved TFLOP/s for the path's real shapes.
//   build: see real-time-video-captioning_b200/build.py (target test_gemm)
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../real-time-video-captioning_b200/csrc/kernels.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d (%s)\n", cudaGetErrorString(e_), __FILE__, __LINE__, gemm_last_error()); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

__global__ void fill_kernel(bf16* p, size_t n, uint32_t seed, float scale) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = (uint32_t)i * 2654435761u ^ seed;
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  float u = (float)(x & 0xffffff) / 16777216.0f - 0.5f;
  p[i] = __float2bfloat16(u * scale);
}
__global__ void fill_f32(float* p, size_t n, uint32_t seed, float scale) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t x = (uint32_t)i * 2654435761u ^ seed;
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  p[i] = ((float)(x & 0xffffff) / 16777216.0f - 0.5f) * scale;
}

__global__ void ref_gemm(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const float* bias,
                         const bf16* res, int ldr, int res_periodic, int act, int gin, int gout, int goff,
                         float* out, int ldo) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __bfloat162float(A[(size_t)m * lda + k]) * __bfloat162float(W[(size_t)n * ldw + k]);
  if (bias) acc += bias[n];
  if (act == ACT_QUICK_GELU) acc = acc / (1.f + expf(-1.702f * acc));
  if (act == ACT_GELU_ERF) acc = 0.5f * acc * (1.f + erff(acc * 0.70710678f));
  int orow = m, rrow = m;
  if (gin > 0) {
    orow = (m / gin) * gout + (m % gin) + goff;
    rrow = res_periodic ? (m % gin) + goff : orow;
  }
  if (res) acc += __bfloat162float(res[(size_t)rrow * ldr + n]);
  out[(size_t)orow * ldo + n] = acc;
}

struct Case {
  int M, N, K, act, use_bias, use_res, gin, gout, goff, periodic, f32out, bn;
};

static int run_case(const Case& c) {
  const int out_rows = c.gin > 0 ? (c.M / c.gin + 1) * c.gout + c.goff : c.M;
  const int res_rows = c.periodic ? c.gin + c.goff : out_rows;
  bf16 *A, *W, *res, *out;
  float *bias, *ref, *out32;
  CK(cudaMalloc(&A, (size_t)c.M * c.K * 2));
  CK(cudaMalloc(&W, (size_t)c.N * c.K * 2));
  CK(cudaMalloc(&res, (size_t)res_rows * c.N * 2));
  CK(cudaMalloc(&out, (size_t)out_rows * c.N * 2));
  CK(cudaMalloc(&out32, (size_t)out_rows * c.N * 4));
  CK(cudaMalloc(&ref, (size_t)out_rows * c.N * 4));
  CK(cudaMalloc(&bias, (size_t)c.N * 4));
  CK(cudaMemset(out, 0, (size_t)out_rows * c.N * 2));
  CK(cudaMemset(out32, 0, (size_t)out_rows * c.N * 4));
  CK(cudaMemset(ref, 0, (size_t)out_rows * c.N * 4));
  auto fill = [&](bf16* p, size_t n, uint32_t s, float sc) { fill_kernel<<<(unsigned)((n + 255) / 256), 256>>>(p, n, s, sc); };
  fill(A, (size_t)c.M * c.K, 1u, 2.0f);
  fill(W, (size_t)c.N * c.K, 2u, 0.25f);
  fill(res, (size_t)res_rows * c.N, 3u, 2.0f);
  fill_f32<<<(c.N + 255) / 256, 256>>>(bias, c.N, 4u, 1.0f);
  CK(cudaDeviceSynchronize());

  GemmArgs g;
  g.A = A; g.lda = c.K; g.W = W; g.ldw = c.K; g.M = c.M; g.N = c.N; g.K = c.K;
  g.bias = c.use_bias ? bias : nullptr;
  g.residual = c.use_res ? res : nullptr; g.ldr = c.N; g.res_periodic = c.periodic;
  g.act = c.act; g.out = out; g.ldo = c.N;
  g.out_f32 = c.f32out ? out32 : nullptr; g.ldo32 = c.N;
  g.gin = c.gin; g.gout = c.gout; g.goff = c.goff;
  CK(gemm_bf16(g, 0, c.bn));
  CK(cudaDeviceSynchronize());
  dim3 grid((c.N + 127) / 128, c.M);
  ref_gemm<<<grid, 128>>>(A, c.K, W, c.K, c.M, c.N, c.K, g.bias, g.residual, c.N, c.periodic, c.act, c.gin, c.gout,
                          c.goff, ref, c.N);
  CK(cudaDeviceSynchronize());

  std::vector<bf16> h_out((size_t)out_rows * c.N);
  std::vector<float> h_ref((size_t)out_rows * c.N), h_out32((size_t)out_rows * c.N);
  CK(cudaMemcpy(h_out.data(), out, h_out.size() * 2, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_ref.data(), ref, h_ref.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h_out32.data(), out32, h_out32.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, max_err32 = 0, max_ref = 0;
  size_t bad = 0, first_bad = (size_t)-1;
  for (size_t i = 0; i < h_ref.size(); ++i) {
    const double r = h_ref[i];
    const double o = __bfloat162float(h_out[i]);
    const double e = fabs(o - r);
    const double tol = 0.02 + 0.01 * fabs(r);
    if (e > tol) {
      ++bad;
      if (first_bad == (size_t)-1) first_bad = i;
    }
    if (e > max_err) max_err = e;
    if (fabs(r) > max_ref) max_ref = fabs(r);
    if (c.f32out) {
      const double e32 = fabs((double)h_out32[i] - r);
      if (e32 > max_err32) max_err32 = e32;
      if (e32 > 1e-2 + 1e-3 * fabs(r)) ++bad;
    }
  }
  printf("case M=%d N=%d K=%d act=%d bias=%d res=%d remap=%d/%d/%d per=%d f32=%d bn=%d : max|ref|=%.3f max_err_bf16=%.4f max_err_f32=%.5f bad=%zu %s\n",
         c.M, c.N, c.K, c.act, c.use_bias, c.use_res, c.gin, c.gout, c.goff, c.periodic, c.f32out, c.bn, max_ref,
         max_err, max_err32, bad, bad ? "FAIL" : "ok");
  if (bad) {
    const size_t i = first_bad;
    printf("   first bad at row %zu col %zu: got %.4f want %.4f\n", i / c.N, i % c.N, __bfloat162float(h_out[i]), h_ref[i]);
    // error map by 8-row x 16-col blocks of the first tile to expose layout / descriptor mistakes
    for (int r = 0; r < 16 && r < out_rows; ++r) {
      printf("   row %2d:", r);
      for (int cc = 0; cc < 8; ++cc) printf(" %8.3f/%8.3f", __bfloat162float(h_out[(size_t)r * c.N + cc * 8]), h_ref[(size_t)r * c.N + cc * 8]);
      printf("\n");
    }
  }
  cudaFree(A); cudaFree(W); cudaFree(res); cudaFree(out); cudaFree(out32); cudaFree(ref); cudaFree(bias);
  return bad ? 1 : 0;
}

static void bench_case(int M, int N, int K, int bn, int act, int use_res) {
  bf16 *A, *W, *out, *res;
  float* bias;
  CK(cudaMalloc(&A, (size_t)M * K * 2));
  CK(cudaMalloc(&W, (size_t)N * K * 2));
  CK(cudaMalloc(&out, (size_t)M * N * 2));
  CK(cudaMalloc(&res, (size_t)M * N * 2));
  CK(cudaMalloc(&bias, (size_t)N * 4));
  fill_kernel<<<(unsigned)(((size_t)M * K + 255) / 256), 256>>>(A, (size_t)M * K, 1u, 2.0f);
  fill_kernel<<<(unsigned)(((size_t)N * K + 255) / 256), 256>>>(W, (size_t)N * K, 2u, 0.25f);
  fill_kernel<<<(unsigned)(((size_t)M * N + 255) / 256), 256>>>(res, (size_t)M * N, 3u, 1.0f);
  fill_f32<<<(N + 255) / 256, 256>>>(bias, N, 4u, 1.0f);
  GemmArgs g;
  g.A = A; g.lda = K; g.W = W; g.ldw = K; g.M = M; g.N = N; g.K = K; g.bias = bias; g.act = act;
  g.residual = use_res ? res : nullptr; g.ldr = N; g.out = out; g.ldo = N;
  for (int i = 0; i < 3; ++i) CK(gemm_bf16(g, 0, bn));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20;
  CK(cudaEventRecord(e0));
  for (int i = 0; i < iters; ++i) CK(gemm_bf16(g, 0, bn));
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= iters;
  printf("bench M=%d N=%d K=%d bn=%d act=%d res=%d : %.3f ms  %.1f TFLOP/s\n", M, N, K, bn, act, use_res, ms,
         2.0 * M * N * K / (ms * 1e-3) / 1e12);
  cudaFree(A); cudaFree(W); cudaFree(out); cudaFree(res); cudaFree(bias);
}

int main(int argc, char** argv) {
  int fails = 0;
  const Case cases[] = {
      {128, 128, 64, 0, 0, 0, 0, 0, 0, 0, 1, 128},
      {128, 256, 64, 0, 0, 0, 0, 0, 0, 0, 1, 256},
      {128, 128, 256, 0, 0, 0, 0, 0, 0, 0, 1, 128},
      {256, 256, 768, 0, 1, 0, 0, 0, 0, 0, 1, 256},
      {1182, 768, 768, 0, 1, 1, 0, 0, 0, 0, 0, 128},
      {1182, 2304, 768, 0, 1, 0, 0, 0, 0, 0, 0, 256},
      {1182 * 3, 3072, 768, ACT_QUICK_GELU, 1, 0, 0, 0, 0, 0, 0, 256},
      {1182 * 3, 768, 3072, 0, 1, 1, 0, 0, 0, 0, 0, 128},
      {1182 * 2, 3072, 768, ACT_GELU_ERF, 1, 0, 0, 0, 0, 0, 0, 0},
      {1176, 768, 768, 0, 0, 1, 196, 197, 1, 1, 0, 0},  // patch-embed: CLS gap + periodic positional add
      {4, 30720, 768, 0, 1, 0, 0, 0, 0, 0, 1, 0},       // vocabulary head, skinny M
      {100, 768, 3072, 0, 1, 1, 0, 0, 0, 0, 0, 0},
      {1542, 1024, 640, 0, 1, 0, 0, 0, 0, 0, 0, 0},     // ViT-L/14 patch K=588 padded to 640 (K % 64 != 0 path via 8-multiple)
      {777, 1024, 584, 0, 1, 0, 0, 0, 0, 0, 1, 128},    // K tail (584 = 9*64 + 8)
      // CTA-pair kernel (bn == 2): 256x256 pair tiles, TMA-store epilogue, TMA residual load
      {256, 256, 64, 0, 0, 0, 0, 0, 0, 0, 0, 2},
      {256, 256, 768, 0, 1, 0, 0, 0, 0, 0, 0, 2},
      {1182, 768, 768, 0, 1, 1, 0, 0, 0, 0, 0, 2},
      {1182 * 4, 2304, 768, 0, 1, 0, 0, 0, 0, 0, 0, 2},
      {1182 * 3, 3072, 768, ACT_QUICK_GELU, 1, 0, 0, 0, 0, 0, 0, 2},
      {1182 * 3 + 77, 768, 3072, 0, 1, 1, 0, 0, 0, 0, 0, 2},
      {100, 1536, 768, ACT_GELU_ERF, 1, 1, 0, 0, 0, 0, 0, 2},
      {1542 * 2, 1024, 640, 0, 1, 0, 0, 0, 0, 0, 0, 2},
      // weight-streaming skinny kernel (bn == 1): M <= 8 decode rows
      {1, 768, 768, 0, 1, 1, 0, 0, 0, 0, 1, 1},
      {1, 30720, 768, 0, 1, 0, 0, 0, 0, 0, 1, 1},
      {4, 3072, 768, ACT_GELU_ERF, 1, 0, 0, 0, 0, 0, 0, 1},
      {4, 768, 3072, 0, 1, 1, 0, 0, 0, 0, 0, 1},
      {3, 2304, 768, 0, 1, 0, 0, 0, 0, 0, 1, 1},
      {8, 768, 3072, ACT_QUICK_GELU, 0, 1, 0, 0, 0, 0, 1, 1},
  };
  for (const Case& c : cases) fails += run_case(c);
  if (argc > 1 && fails == 0) {
    const int B = atoi(argv[1]);
    const int M = 1182 * B;
    for (int bn : {256, 2}) {
      bench_case(M, 2304, 768, bn, 0, 0);
      bench_case(M, 768, 768, bn, 0, 1);
      bench_case(M, 3072, 768, bn, ACT_QUICK_GELU, 0);
      bench_case(M, 768, 3072, bn, 0, 1);
    }
    bench_case(8192, 8192, 8192, 256, 0, 0);
    bench_case(8192, 8192, 8192, 2, 0, 0);
    bench_case(64, 30720, 768, 128, 0, 0);
    for (int m : {1, 4}) {
      bench_case(m, 2304, 768, 1, 0, 0);
      bench_case(m, 768, 768, 1, 0, 1);
      bench_case(m, 3072, 768, 1, ACT_GELU_ERF, 0);
      bench_case(m, 768, 3072, 1, 0, 1);
      bench_case(m, 30720, 768, 1, 0, 0);
      bench_case(m, 30720, 768, 128, 0, 0);
    }
  }
  printf("test_gemm: %s (%d failing cases)\n", fails ? "FAILED" : "PASSED", fails);
  return fails ? 1 : 0;
}
