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

namespace {

constexpr int HD = 64;  // head dim of both towers (width / heads = 64)

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
constexpr int TA_NB = 4;       // query rows (beams / text positions of one clip) handled together
constexpr int TA_THREADS = 128;
constexpr int TA_GROUPS = TA_THREADS / 8;  // 16 key lanes-groups per CTA, 8 lanes (8 dims each) per key

struct Acc {
  float m, l, o[8];
};
__device__ __forceinline__ void acc_add(Acc& a, float s, const float (&v)[8]) {
  const float m_new = fmaxf(a.m, s);
  const float c = exp2f(a.m - m_new), p = exp2f(s - m_new);
  a.l = a.l * c + p;
#pragma unroll
  for (int i = 0; i < 8; ++i) a.o[i] = a.o[i] * c + p * v[i];
  a.m = m_new;
}
__device__ __forceinline__ void acc_merge(Acc& a, float m2, float l2, const float (&o2)[8]) {
  const float m_new = fmaxf(a.m, m2);
  const float c1 = (a.m == -INFINITY) ? 0.f : exp2f(a.m - m_new);
  const float c2 = (m2 == -INFINITY) ? 0.f : exp2f(m2 - m_new);
  a.l = a.l * c1 + l2 * c2;
#pragma unroll
  for (int i = 0; i < 8; ++i) a.o[i] = a.o[i] * c1 + o2[i] * c2;
  a.m = m_new;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
// Dot product over the 8 lanes of one key group.  `gmask` names exactly those 8 lanes: the groups of a warp run
// different trip counts at the tails of the key loops, so a full-warp shuffle mask would be undefined behaviour.
__device__ __forceinline__ float dot8_group(const float (&q)[8], const float (&k)[8], uint32_t gmask) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s = fmaf(q[i], k[i], s);
  s += __shfl_xor_sync(gmask, s, 1);
  s += __shfl_xor_sync(gmask, s, 2);
  s += __shfl_xor_sync(gmask, s, 4);
  return s;
}

__global__ void __launch_bounds__(TA_THREADS) text_attention_kernel(TextAttnArgs a, float scale_log2) {
  __shared__ float sm_m[TA_NB][4], sm_l[TA_NB][4], sm_o[TA_NB][4][HD];
  const int chunks = (a.rows_per_clip + TA_NB - 1) / TA_NB;
  const int clip = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const int h = blockIdx.y, split = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = tid >> 3;       // 0..15: which key of a 16-key stride this lane group handles
  const int d0 = (lane & 7) * 8;  // this lane's 8 dims of the 64-dim head
  const uint32_t gmask = 0xFFu << (lane & 24);
  const int rl0 = chunk * TA_NB;
  const int n_loc = min(TA_NB, a.rows_per_clip - rl0);
  const int width = a.heads * HD;

  float q[TA_NB][8];
  Acc acc[TA_NB];
#pragma unroll
  for (int r = 0; r < TA_NB; ++r) {
    acc[r].m = -INFINITY;
    acc[r].l = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[r].o[i] = 0.f;
    if (r < n_loc) {
      load8(a.q + (size_t)(clip * a.rows_per_clip + rl0 + r) * a.ldq + h * HD + d0, q[r]);
#pragma unroll
      for (int i = 0; i < 8; ++i) q[r][i] *= scale_log2;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) q[r][i] = 0.f;
    }
  }

  // ---- visual keys of this clip, shared by all rows of the chunk; this split's slice
  const int per = (a.Nv + a.splits - 1) / a.splits;
  const int k_begin = split * per, k_end = min(a.Nv, k_begin + per);
  const bf16* kp = a.vis_kv + (size_t)clip * a.Nv * a.ld_vis + a.k_off + h * HD + d0;
  const bf16* vp = a.vis_kv + (size_t)clip * a.Nv * a.ld_vis + a.v_off + h * HD + d0;
  int key = k_begin + grp;
  for (; key + TA_GROUPS < k_end; key += 2 * TA_GROUPS) {  // two keys in flight per lane group
    float k0[8], v0[8], k1[8], v1[8];
    load8(kp + (size_t)key * a.ld_vis, k0);
    load8(kp + (size_t)(key + TA_GROUPS) * a.ld_vis, k1);
    load8(vp + (size_t)key * a.ld_vis, v0);
    load8(vp + (size_t)(key + TA_GROUPS) * a.ld_vis, v1);
#pragma unroll
    for (int r = 0; r < TA_NB; ++r) {
      if (r < n_loc) {  // block-uniform
        acc_add(acc[r], dot8_group(q[r], k0, gmask), v0);
        acc_add(acc[r], dot8_group(q[r], k1, gmask), v1);
      }
    }
  }
  for (; key < k_end; key += TA_GROUPS) {
    float k0[8], v0[8];
    load8(kp + (size_t)key * a.ld_vis, k0);
    load8(vp + (size_t)key * a.ld_vis, v0);
#pragma unroll
    for (int r = 0; r < TA_NB; ++r)
      if (r < n_loc) acc_add(acc[r], dot8_group(q[r], k0, gmask), v0);
  }

  // ---- causal text prefix (handled by the last split); per-row keys via the ancestor table
  if (split == a.splits - 1) {
#pragma unroll
    for (int r = 0; r < TA_NB; ++r) {
      if (r >= n_loc) break;
      const int row = clip * a.rows_per_clip + rl0 + r;
      const int nt = a.n_text ? a.n_text[row] : a.n_text_const;
      for (int s = grp; s < nt; s += TA_GROUPS) {
        int slot = row;
        if (a.anc != nullptr && s < nt - 1) slot = a.anc[(size_t)row * a.anc_ld + s];
        if (a.text_slot_is_clip) slot = clip;
        const bf16* t = a.txt_kv + ((size_t)s * a.txt_slots + slot) * (2 * width) + h * HD + d0;
        float k0[8], v0[8];
        load8(t, k0);
        load8(t + width, v0);
        acc_add(acc[r], dot8_group(q[r], k0, gmask), v0);
      }
    }
  }

  __syncwarp();
  // ---- merge the 4 lane groups of each warp (lanes differing in bits 3,4), then the 4 warps through smem
#pragma unroll
  for (int r = 0; r < TA_NB; ++r) {
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {
      const float m2 = __shfl_xor_sync(0xffffffffu, acc[r].m, off);
      const float l2 = __shfl_xor_sync(0xffffffffu, acc[r].l, off);
      float o2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o2[i] = __shfl_xor_sync(0xffffffffu, acc[r].o[i], off);
      acc_merge(acc[r], m2, l2, o2);
    }
    if (lane < 8) {
      sm_m[r][warp] = acc[r].m;
      sm_l[r][warp] = acc[r].l;
#pragma unroll
      for (int i = 0; i < 8; ++i) sm_o[r][warp][d0 + i] = acc[r].o[i];
    }
  }
  __syncthreads();
  // thread t: row r = t / 32 (TA_NB == 4 rows x 32 lanes), dims 2*lane, 2*lane+1
  {
    const int r = warp, d = lane * 2;
    if (r < n_loc) {
      float m = -INFINITY;
#pragma unroll
      for (int w = 0; w < 4; ++w) m = fmaxf(m, sm_m[r][w]);
      float l = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const float c = (sm_m[r][w] == -INFINITY) ? 0.f : exp2f(sm_m[r][w] - m);
        l += sm_l[r][w] * c;
        o0 += sm_o[r][w][d] * c;
        o1 += sm_o[r][w][d + 1] * c;
      }
      const int row = clip * a.rows_per_clip + rl0 + r;
      if (a.splits == 1) {
        const float inv = 1.f / l;
        *reinterpret_cast<uint32_t*>(a.out + (size_t)row * a.ldo + h * HD + d) = pack_bf16(o0 * inv, o1 * inv);
      } else {
        float* p = a.partial + (((size_t)row * a.heads + h) * a.splits + split) * (HD + 2);
        if (lane == 0) {
          p[0] = m;
          p[1] = l;
        }
        p[2 + d] = o0;
        p[2 + d + 1] = o1;
      }
    }
  }
}

// combine split partials: one warp per (row, head)
__global__ void text_attention_combine_kernel(const float* __restrict__ partial, int n_rows, int heads, int splits,
                                              bf16* __restrict__ out, int ldo) {
  const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (idx >= n_rows * heads) return;
  const int row = idx / heads, h = idx % heads;
  const float* p = partial + (size_t)idx * splits * (HD + 2);
  float m = -INFINITY;
  for (int s = 0; s < splits; ++s) m = fmaxf(m, p[s * (HD + 2)]);
  float l = 0.f, o0 = 0.f, o1 = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float* ps = p + s * (HD + 2);
    const float c = (ps[0] == -INFINITY) ? 0.f : exp2f(ps[0] - m);
    l += ps[1] * c;
    o0 += ps[2 + lane * 2] * c;
    o1 += ps[2 + lane * 2 + 1] * c;
  }
  const float inv = 1.f / l;
  *reinterpret_cast<uint32_t*>(out + (size_t)row * ldo + h * HD + lane * 2) = pack_bf16(o0 * inv, o1 * inv);
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
  const int chunks = (a.rows_per_clip + TA_NB - 1) / TA_NB;
  dim3 grid(a.n_clips * chunks, a.heads, a.splits);
  text_attention_kernel<<<grid, TA_THREADS, 0, stream>>>(a, a.scale * 1.4426950408889634f);
  note_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || a.splits == 1) return e;
  const int warps = n_rows * a.heads;
  text_attention_combine_kernel<<<(warps * 32 + 127) / 128, 128, 0, stream>>>(a.partial, n_rows, a.heads, a.splits, a.out,
                                                                              a.ldo);
  note_launch();
  return cudaGetLastError();
}
