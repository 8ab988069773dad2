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
struct SkinnyFused {  // optional fusions of the single-clip decode step (kernels.h: GemmArgs::lnl_* / sc_*)
  const float* gamma;
  const float* beta;
  float eps;
  bf16* ln_out;
  int ln_ldo;
  bf16* kv;
  int q_width, kv_width;
  const int* pos;
  int pos_const, slot_div, n_slots;
};

// LNL: LayerNorm-on-load of the A rows (K == 768 = 3 slices of 256 per warp, rows kept normalised in registers)
template <int MT, bool SPLITK, bool LNL>
__global__ void __launch_bounds__(MAX_WARPS * 32)
gemv_skinny_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ w, int ldw, int M, int N, int K,
                   const float* __restrict__ bias, const bf16* __restrict__ residual, int ldr, int act,
                   bf16* __restrict__ out, int ldo, float* __restrict__ out_f32, int ldo32, SkinnyFused fz) {
  __shared__ float red[MAX_WARPS][MT * NC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  const int n0 = SPLITK ? blockIdx.x * NC : (blockIdx.x * n_warps + warp) * NC;
  if (!SPLITK && n0 >= N) return;
  float acc[MT][NC];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[m][c] = 0.f;

  constexpr int LN_SLICES = 3;  // K == 768
  float xn[LNL ? MT : 1][LNL ? LN_SLICES * 8 : 1];
  if (LNL) {
    // same arithmetic and the same lane -> column assignment as layernorm_kernel: bit-identical normalised rows
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      float v[LN_SLICES * 8];
#pragma unroll
      for (int i = 0; i < LN_SLICES; ++i) {
        uint4 u = make_uint4(0, 0, 0, 0);
        if (m < M) u = __ldg(reinterpret_cast<const uint4*>(x + (size_t)m * ldx + (i * 32 + lane) * 8));
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c2 = unpack_bf16(u.z), d = unpack_bf16(u.w);
        v[i * 8 + 0] = a.x; v[i * 8 + 1] = a.y; v[i * 8 + 2] = b.x; v[i * 8 + 3] = b.y;
        v[i * 8 + 4] = c2.x; v[i * 8 + 5] = c2.y; v[i * 8 + 6] = d.x; v[i * 8 + 7] = d.y;
      }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < LN_SLICES * 8; ++i) sum += v[i];
      const float mean = warp_sum(sum) / (float)K;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < LN_SLICES * 8; ++i) {
        const float dlt = v[i] - mean;
        q += dlt * dlt;
      }
      const float rstd = rsqrtf(warp_sum(q) / (float)K + fz.eps);
#pragma unroll
      for (int i = 0; i < LN_SLICES; ++i) {
        const int c = (i * 32 + lane) * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(fz.gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(fz.gamma + c + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(fz.beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(fz.beta + c + 4));
        float o[8];
        o[0] = (v[i * 8 + 0] - mean) * rstd * g0.x + b0.x;
        o[1] = (v[i * 8 + 1] - mean) * rstd * g0.y + b0.y;
        o[2] = (v[i * 8 + 2] - mean) * rstd * g0.z + b0.z;
        o[3] = (v[i * 8 + 3] - mean) * rstd * g0.w + b0.w;
        o[4] = (v[i * 8 + 4] - mean) * rstd * g1.x + b1.x;
        o[5] = (v[i * 8 + 5] - mean) * rstd * g1.y + b1.y;
        o[6] = (v[i * 8 + 6] - mean) * rstd * g1.z + b1.z;
        o[7] = (v[i * 8 + 7] - mean) * rstd * g1.w + b1.w;
        uint4 u;
        u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]); u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
        if (fz.ln_out != nullptr && blockIdx.x == 0 && warp == 0 && m < M)
          *reinterpret_cast<uint4*>(fz.ln_out + (size_t)m * fz.ln_ldo + c) = u;
        const float2 r0 = unpack_bf16(u.x), r1 = unpack_bf16(u.y), r2 = unpack_bf16(u.z), r3 = unpack_bf16(u.w);
        xn[m][i * 8 + 0] = r0.x; xn[m][i * 8 + 1] = r0.y; xn[m][i * 8 + 2] = r1.x; xn[m][i * 8 + 3] = r1.y;
        xn[m][i * 8 + 4] = r2.x; xn[m][i * 8 + 5] = r2.y; xn[m][i * 8 + 6] = r3.x; xn[m][i * 8 + 7] = r3.y;
      }
    }
  }

  // one 256-wide K slice of this warp: 4 weight rows x 16 bytes per lane against the M activation rows
  auto slice_step = [&](int k0, int slice) {
    uint4 wv[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) wv[c] = __ldg(reinterpret_cast<const uint4*>(w + (size_t)(n0 + c) * ldw + k0));
    float xf[MT][8];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      if (LNL) {
#pragma unroll
        for (int i = 0; i < 8; ++i) xf[m][i] = xn[m][(LNL ? slice : 0) * 8 + i];
      } else if (m < M) {
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
  };
  if (LNL) {
#pragma unroll
    for (int it = 0; it < LN_SLICES; ++it) slice_step(lane * 8 + it * 256, it);  // fully unrolled: xn stays in registers
  } else {
    for (int k0 = (SPLITK ? warp * 256 : 0) + lane * 8; k0 < K; k0 += (SPLITK ? n_warps : 1) * 256)  // K % 8 == 0; a warp covers 256 k per iteration
      slice_step(k0, 0);
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
        if (fz.kv != nullptr && n >= fz.q_width) {  // this step's K | V also go straight into the text K/V plane
          const int p = fz.pos ? fz.pos[m] : fz.pos_const;
          fz.kv[((size_t)p * fz.n_slots + m / fz.slot_div) * fz.kv_width + (n - fz.q_width)] = __float2bfloat16(v);
        }
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
  const bool lnl = a.lnl_gamma != nullptr;
  if (lnl && (a.M > 4 || a.K != 768 || splitk || a.lnl_beta == nullptr)) return cudaErrorInvalidValue;
  SkinnyFused fz{a.lnl_gamma, a.lnl_beta, a.lnl_eps, a.lnl_out, a.lnl_ldo, a.sc_kv, a.sc_q_width, a.sc_kv_width, a.sc_pos,
                 a.sc_pos_const, a.sc_slot_div, a.sc_n_slots};
#define LAUNCH(MT)                                                                                                       \
  do {                                                                                                                   \
    if (splitk)                                                                                                          \
      gemv_skinny_kernel<MT, true, false><<<grid, warps * 32, 0, stream>>>(a.A, a.lda, a.W, a.ldw, a.M, a.N, a.K, a.bias, \
                                                                    a.residual, a.ldr, a.act, a.out, a.ldo, a.out_f32, a.ldo32, fz); \
    else if (lnl)                                                                                                        \
      gemv_skinny_kernel<(MT > 4 ? 4 : MT), false, true><<<grid, warps * 32, 0, stream>>>(a.A, a.lda, a.W, a.ldw, a.M, a.N, a.K, a.bias, \
                                                                     a.residual, a.ldr, a.act, a.out, a.ldo, a.out_f32, a.ldo32, fz); \
    else                                                                                                                 \
      gemv_skinny_kernel<MT, false, false><<<grid, warps * 32, 0, stream>>>(a.A, a.lda, a.W, a.ldw, a.M, a.N, a.K, a.bias, \
                                                                     a.residual, a.ldr, a.act, a.out, a.ldo, a.out_f32, a.ldo32, fz); \
  } while (0)
  if (a.M <= 1) LAUNCH(1);
  else if (a.M <= 2) LAUNCH(2);
  else if (a.M <= 4) LAUNCH(4);
  else LAUNCH(8);
#undef LAUNCH
  note_launch();
  return cudaGetLastError();
}
