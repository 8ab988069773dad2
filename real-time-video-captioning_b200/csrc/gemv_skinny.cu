// Weight-streaming skinny GEMM for decode steps with a handful of rows (single-clip latency mode: 1 row for greedy,
// 4 for beam search) -- SURVEY K12.  out[m, n] = act(sum_k x[m, k] * W[n, k] + bias[n]) + residual[m, n], M <= 8.
//
// HBM bound: every weight byte is read exactly once per step (131.8 MB per decode step for GIT-base) while the M
// activation rows stay in registers / L1.  One CTA owns NC = 4 consecutive output columns (4 weight rows); its warps
// split K in 256-element slices (so N = 768 still spreads over 192 CTAs), all 16-byte loads of a slice are issued
// before any FMA (4 x 512 B in flight per warp), fp32 accumulation, warp-shuffle + shared-memory reduction, fused
// bias / GELU / residual epilogue.  The 128-row tcgen05 tile is the wrong tool here: with M = 1 it
// fills 6-24 of the 148 SMs and spends ~10-20 us per matrix on what is 0.2-7 us of weight streaming.
#include "common.cuh"
#include "kernels.h"

namespace {

constexpr int NC = 4;        // output columns per warp
constexpr int MAX_WARPS = 8; // warps per CTA = min(8, K / 256): each takes every MAX-th 256-wide slice of K

// SPLITK = true : the warps of a CTA share NC columns and split K (few columns, long K: fc2 768 x 3072)
// SPLITK = false: every warp owns its own NC columns and walks the whole K (everything else)
template <int MT, bool SPLITK>
__global__ void __launch_bounds__(MAX_WARPS * 32)
gemv_skinny_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ w, int ldw, int M, int N, int K,
                   const float* __restrict__ bias, const bf16* __restrict__ residual, int ldr, int act,
                   bf16* __restrict__ out, int ldo, float* __restrict__ out_f32, int ldo32) {
  __shared__ float red[MAX_WARPS][MT * NC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  const int n0 = SPLITK ? blockIdx.x * NC : (blockIdx.x * n_warps + warp) * NC;
  if (!SPLITK && n0 >= N) return;
  float acc[MT][NC];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[m][c] = 0.f;

  for (int k0 = (SPLITK ? warp * 256 : 0) + lane * 8; k0 < K; k0 += (SPLITK ? n_warps : 1) * 256) {  // K % 8 == 0; a warp covers 256 k per iteration
    uint4 wv[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) wv[c] = __ldg(reinterpret_cast<const uint4*>(w + (size_t)(n0 + c) * ldw + k0));
    float xf[MT][8];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (m < M) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (size_t)m * ldx + k0));
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c2 = unpack_bf16(u.z), d = unpack_bf16(u.w);
        xf[m][0] = a.x; xf[m][1] = a.y; xf[m][2] = b.x; xf[m][3] = b.y;
        xf[m][4] = c2.x; xf[m][5] = c2.y; xf[m][6] = d.x; xf[m][7] = d.y;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) xf[m][i] = 0.f;
      }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float2 a = unpack_bf16(wv[c].x), b = unpack_bf16(wv[c].y), c2 = unpack_bf16(wv[c].z), d = unpack_bf16(wv[c].w);
      const float wf[8] = {a.x, a.y, b.x, b.y, c2.x, c2.y, d.x, d.y};
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[m][c] = fmaf(xf[m][i], wf[i], acc[m][c]);
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float v = warp_sum(acc[m][c]);
      if (lane == 0) red[warp][m * NC + c] = v;
    }
  if (SPLITK) __syncthreads(); else __syncwarp();
  // thread t < MT * NC finalises (m, c) = (t / NC, t % NC)
#pragma unroll
  for (int m = 0; m < MT; ++m) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if ((SPLITK ? (int)threadIdx.x : lane) == m * NC + c && m < M) {
        const int n = n0 + c;
        float v = 0.f;
        if (SPLITK) {
          for (int w2 = 0; w2 < n_warps; ++w2) v += red[w2][m * NC + c];
        } else {
          v = red[warp][m * NC + c];
        }
        if (bias) v += bias[n];
        if (act == ACT_QUICK_GELU) v = quick_gelu(v);
        else if (act == ACT_GELU_ERF) v = gelu_erf(v);
        else if (act == ACT_RELU) v = fmaxf(v, 0.f);
        if (residual) v += __bfloat162float(residual[(size_t)m * ldr + n]);
        if (out) out[(size_t)m * ldo + n] = __float2bfloat16(v);
        if (out_f32) out_f32[(size_t)m * ldo32 + n] = v;
      }
    }
  }
}

}  // namespace

// Returns cudaErrorNotSupported when the shape is not a skinny one (caller uses the tcgen05 kernels).
cudaError_t gemv_skinny(const GemmArgs& a, cudaStream_t stream) {
  if (a.M > 8 || a.N % NC != 0 || a.K % 8 != 0 || a.gin > 0 || a.res_periodic) return cudaErrorNotSupported;
  const bool splitk = a.N / (MAX_WARPS * NC) < 64 && a.K >= 2048;
  const int warps = MAX_WARPS;
  const int grid = splitk ? a.N / NC : (a.N + warps * NC - 1) / (warps * NC);
#define LAUNCH(MT)                                                                                                       \
  do {                                                                                                                   \
    if (splitk)                                                                                                          \
      gemv_skinny_kernel<MT, true><<<grid, warps * 32, 0, stream>>>(a.A, a.lda, a.W, a.ldw, a.M, a.N, a.K, a.bias,       \
                                                                    a.residual, a.ldr, a.act, a.out, a.ldo, a.out_f32, a.ldo32); \
    else                                                                                                                 \
      gemv_skinny_kernel<MT, false><<<grid, warps * 32, 0, stream>>>(a.A, a.lda, a.W, a.ldw, a.M, a.N, a.K, a.bias,      \
                                                                     a.residual, a.ldr, a.act, a.out, a.ldo, a.out_f32, a.ldo32); \
  } while (0)
  if (a.M <= 1) LAUNCH(1);
  else if (a.M <= 2) LAUNCH(2);
  else if (a.M <= 4) LAUNCH(4);
  else LAUNCH(8);
#undef LAUNCH
  note_launch();
  return cudaGetLastError();
}
