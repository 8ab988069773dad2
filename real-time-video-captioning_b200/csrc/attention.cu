// Attention kernels of the captioning path.
//
//  attention_groups : bidirectional softmax(QK^T/sqrt(d))V inside fixed-length row groups (one
//                     ViT frame = 197/257 rows; K10 visual block: one clip's Nv visual rows, which under
//                     the prefix-LM mask only ever see each other).  Flash-style single pass: K/V blocks
//                     of 64 keys are staged with cp.async (double buffered, XOR-swizzled rows), QK^T and
//                     PV run on the warp-level tensor path (mma.sync m16n8k16 bf16, fp32 accumulate),
//                     softmax statistics stay in registers and are reduced with warp shuffles.
//  text_attention   : the decode-step / text-row kernel: a handful of query rows per clip
//                     against the clip's cached visual K/V (read once for all beams of the clip) plus the
//                     causal text prefix found through the per-position ancestor table.  HBM bound:
//                     16-byte coalesced loads, 8 lanes per key, warp-shuffle dot products and softmax.
#include <math.h>

#include "common.cuh"
#include "kernels.h"
#include <stdlib.h>

#include "text_attention_dev.cuh"

namespace {

using text_attn_dev::HD;  // head dim of both towers (width / heads = 64)

// ================================================================== attention_groups
constexpr int AG_WARPS = 4;
constexpr int AG_BQ = 16 * AG_WARPS;  // 64 query rows per CTA
constexpr int AG_BKV = 64;

// tile row = 128 bytes (64 bf16); 16-byte chunk index XOR (row & 7) avoids ldmatrix bank conflicts
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

__global__ void __launch_bounds__(AG_WARPS * 32)
attention_groups_kernel(const bf16* __restrict__ qkv, int ld_qkv, bf16* __restrict__ out, int ldo, int group_len,
                        int heads, float scale_log2) {
  __shared__ __align__(128) uint8_t smem[AG_BQ * 128 + 2 * 2 * AG_BKV * 128];  // Q | K0 V0 | K1 V1  (40 KB)
  const uint32_t sQ = ptx::smem_u32(smem);
  auto sK = [&](int b) { return sQ + AG_BQ * 128 + b * (2 * AG_BKV * 128); };
  auto sV = [&](int b) { return sK(b) + AG_BKV * 128; };

  const int qt = blockIdx.x, h = blockIdx.y, g = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int width = heads * HD;
  const size_t row0 = (size_t)g * group_len;
  const bf16* qbase = qkv + row0 * ld_qkv + h * HD;
  const bf16* kbase = qbase + width;
  const bf16* vbase = qbase + 2 * width;
  const int q0 = qt * AG_BQ;
  const int n_kv = (group_len + AG_BKV - 1) / AG_BKV;

  // ---- stage Q tile and first K/V block
  for (int i = tid; i < AG_BQ * 8; i += AG_WARPS * 32) {
    const int r = i >> 3, c = i & 7;
    const bool ok = q0 + r < group_len;
    ptx::cp_async_16(sQ + tile_off(r, c), qbase + (size_t)(ok ? q0 + r : 0) * ld_qkv + c * 8, ok);
  }
  auto load_kv = [&](int blk, int buf) {
    const int k0 = blk * AG_BKV;
    for (int i = tid; i < AG_BKV * 8; i += AG_WARPS * 32) {
      const int r = i >> 3, c = i & 7;
      const bool ok = k0 + r < group_len;
      const size_t off = (size_t)(ok ? k0 + r : 0) * ld_qkv + c * 8;
      ptx::cp_async_16(sK(buf) + tile_off(r, c), kbase + off, ok);
      ptx::cp_async_16(sV(buf) + tile_off(r, c), vbase + off, ok);
    }
  };
  load_kv(0, 0);
  ptx::cp_async_commit();

  uint32_t qf[4][4];  // Q fragments: 4 k-steps of 16 dims
  float o[8][4];      // O accumulator: 8 dim tiles of 8
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  for (int blk = 0; blk < n_kv; ++blk) {
    const int buf = blk & 1;
    if (blk + 1 < n_kv) load_kv(blk + 1, buf ^ 1);
    ptx::cp_async_commit();
    ptx::cp_async_wait<1>();
    __syncthreads();
    if (blk == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int m = lane >> 3, r = lane & 7;
        const int row = warp * 16 + (m & 1) * 8 + r;
        ptx::ldmatrix_x4(sQ + tile_off(row, ks * 2 + (m >> 1)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    // ---- S = Q K^T  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key tiles
        const int m = lane >> 3, r = lane & 7;
        const int key = np * 16 + (m >> 1) * 8 + r;
        uint32_t b0, b1, b2, b3;
        ptx::ldmatrix_x4(sK(buf) + tile_off(key, ks * 2 + (m & 1)), b0, b1, b2, b3);
        ptx::mma_bf16_16816(s[np * 2], qf[ks], b0, b1);
        ptx::mma_bf16_16816(s[np * 2 + 1], qf[ks], b2, b3);
      }
    }
    // ---- mask keys beyond the group, online softmax (rows lane/4 and lane/4 + 8)
    const int kbase_idx = blk * AG_BKV + (lane & 3) * 2;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kbase_idx + nt * 8 + (e & 1);
        const float v = key < group_len ? s[nt][e] * scale_log2 : -INFINITY;
        s[nt][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float corr[2];
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      mx[rr] = fmaxf(mx[rr], __shfl_xor_sync(0xffffffffu, mx[rr], 1));
      mx[rr] = fmaxf(mx[rr], __shfl_xor_sync(0xffffffffu, mx[rr], 2));
      const float m_new = fmaxf(m_run[rr], mx[rr]);  // finite: every block holds at least one valid key
      corr[rr] = exp2f(m_run[rr] - m_new);
      m_run[rr] = m_new;
      l_run[rr] *= corr[rr];
    }
    float ls[2] = {0.f, 0.f};
    uint32_t pf[4][4];  // P as A fragments for the PV product: 4 k-steps of 16 keys
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - m_run[0]);
      const float p1 = exp2f(s[nt][1] - m_run[0]);
      const float p2 = exp2f(s[nt][2] - m_run[1]);
      const float p3 = exp2f(s[nt][3] - m_run[1]);
      ls[0] += p0 + p1;
      ls[1] += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    l_run[0] += ls[0];
    l_run[1] += ls[1];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      o[dt][0] *= corr[0];
      o[dt][1] *= corr[0];
      o[dt][2] *= corr[1];
      o[dt][3] *= corr[1];
    }
    // ---- O += P V
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {    // 16 keys per step
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {  // pairs of 8-dim tiles
        const int m = lane >> 3, r = lane & 7;
        const int key = ks * 16 + (m & 1) * 8 + r;
        uint32_t b0, b1, b2, b3;
        ptx::ldmatrix_x4_trans(sV(buf) + tile_off(key, dp * 2 + (m >> 1)), b0, b1, b2, b3);
        ptx::mma_bf16_16816(o[dp * 2], pf[ks], b0, b1);
        ptx::mma_bf16_16816(o[dp * 2 + 1], pf[ks], b2, b3);
      }
    }
    __syncthreads();  // all warps done with this buffer before it is refilled
  }

  // ---- finalise: reduce row sums across the quad, normalise, store
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    l_run[rr] += __shfl_xor_sync(0xffffffffu, l_run[rr], 1);
    l_run[rr] += __shfl_xor_sync(0xffffffffu, l_run[rr], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  bf16* obase = out + row0 * ldo + h * HD + (lane & 3) * 2;
#pragma unroll
  for (int dt = 0; dt < 8; ++dt) {
    if (r0 < group_len) *reinterpret_cast<uint32_t*>(obase + (size_t)r0 * ldo + dt * 8) = pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
    if (r1 < group_len) *reinterpret_cast<uint32_t*>(obase + (size_t)r1 * ldo + dt * 8) = pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
  }
}

// ================================================================== text_attention
// (body: text_attention_dev.cuh)
using text_attn_dev::TA_THREADS;
using text_attn_dev::TA_GROUPS;

template <int NB, int UT = 2>
__global__ void __launch_bounds__(TA_THREADS, NB == 1 ? 8 : 6) text_attention_kernel(TextAttnArgs a, float scale_log2, int kcap) {
  extern __shared__ float sm[];
  const int chunks = (a.rows_per_clip + NB - 1) / NB;
  text_attn_dev::text_attention_body<NB, false, UT, 5>(a, scale_log2, kcap, sm, blockIdx.y / chunks, blockIdx.y % chunks, blockIdx.x,
                                                       blockIdx.z, threadIdx.x, [] { __syncthreads(); });
}

// combine split partials: one warp per (row, head)
__global__ void text_attention_combine_kernel(const float* __restrict__ partial, int n_rows, int heads, int splits,
                                              bf16* __restrict__ out, int ldo) {
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (idx >= n_rows * heads) return;
  const int row = idx / heads, h = idx % heads;
  const float* p = partial + (size_t)idx * splits * (HD + 2);
  *reinterpret_cast<uint32_t*>(out + (size_t)row * ldo + h * HD + lane * 2) = text_attn_dev::combine_partials<false>(p, splits, lane);
}

}  // namespace

cudaError_t attention_groups(const bf16* qkv, int ld_qkv, bf16* out, int ldo, int n_groups, int group_len, int heads,
                             float scale, cudaStream_t stream) {
  if (n_groups <= 0 || group_len <= 0) return cudaSuccess;
  if (ld_qkv % 8 != 0 || ldo % 2 != 0) return cudaErrorInvalidValue;
  dim3 grid((group_len + AG_BQ - 1) / AG_BQ, heads, n_groups);
  attention_groups_kernel<<<grid, AG_WARPS * 32, 0, stream>>>(qkv, ld_qkv, out, ldo, group_len, heads,
                                                              scale * 1.4426950408889634f);
  note_launch();
  return cudaGetLastError();
}

size_t text_attention_workspace_floats(int n_rows, int heads, int splits) {
  return (size_t)n_rows * heads * splits * (HD + 2);
}

cudaError_t text_attention(const TextAttnArgs& a, cudaStream_t stream) {
  const int n_rows = a.n_clips * a.rows_per_clip;
  if (n_rows <= 0) return cudaSuccess;
  if (a.splits < 1 || (a.splits > 1 && a.partial == nullptr) || a.ld_vis % 8 != 0 || a.ldq % 8 != 0)
    return cudaErrorInvalidValue;
  const int nb = a.rows_per_clip >= 4 ? 4 : (a.rows_per_clip >= 2 ? 2 : 1);
  const int chunks = (a.rows_per_clip + nb - 1) / nb;
  const int per = (a.Nv + a.splits - 1) / a.splits;
  const int kcap = ((per + a.max_text + 3) / 4) * 4;
  const size_t smem = (size_t)text_attn_dev::ta_smem_floats(nb, kcap) * sizeof(float);
  if (smem > 200 * 1024 || (long)a.n_clips * chunks > 65535) return cudaErrorInvalidValue;  // grid.y limit
  // heads are the FASTEST grid dimension: the CTAs of one clip's heads are resident together and jointly read whole
  // 1.5 KB K (then V) runs of every cache row instead of one 128-byte piece of rows 4.6 KB apart
  dim3 grid(a.heads, a.n_clips * chunks, a.splits);
  const float sl2 = a.scale * 1.4426950408889634f;
  cudaError_t e = cudaSuccess;
  auto launch = [&](auto kernel) {
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) kernel<<<grid, TA_THREADS, smem, stream>>>(a, sl2, kcap);
  };
  // throughput batches: 2 tiles per warp and step; a few clips (key splits, the persistent decode kernel's twin): TA_UT_LATENCY
  const int ut = a.n_clips * chunks * a.heads >= 2 * 148 ? 2 : text_attn_dev::TA_UT_LATENCY;
  if (nb == 4 && ut == 2) launch(text_attention_kernel<4, 2>);
  else if (nb == 4) launch(text_attention_kernel<4, text_attn_dev::TA_UT_LATENCY>);
  else if (nb == 2 && ut == 2) launch(text_attention_kernel<2, 2>);
  else if (nb == 2) launch(text_attention_kernel<2, text_attn_dev::TA_UT_LATENCY>);
  else launch(text_attention_kernel<1>);
  if (e != cudaSuccess) return e;
  note_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess || a.splits == 1) return e;
  const int warps = n_rows * a.heads;
  text_attention_combine_kernel<<<(warps * 32 + 127) / 128, 128, 0, stream>>>(a.partial, n_rows, a.heads, a.splits, a.out,
                                                                              a.ldo);
  note_launch();
  return cudaGetLastError();
}
