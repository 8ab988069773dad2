// Device-side body of the decode-step / text-row attention, shared by text_attention_kernel (attention.cu: one
// CTA per (clip, row chunk, head, key split)) and the persistent single-clip decode kernel (decode_mega.cu: several
// 128-thread "virtual CTAs" per CTA, same arithmetic in the same order -> bit-identical results).
#pragma once
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace text_attn_dev {

constexpr int HD = 64;  // head dim (width / heads = 64)

// Decode-step attention, HBM bound.  One CTA = (clip, chunk of <= NB query rows, head, key split).  Two streaming
// phases so that the load loops carry no softmax dependency chain:
//   1. scores: 8 lanes per key (16-byte loads = one full 128-byte line per key and group), 8 keys in flight per
//      lane, partial dot products combined with a transposing butterfly (7 shuffles per 8 keys instead of 24);
//      scores go to shared memory, then one block-wide max / exp2 / sum pass;
//   2. values: the same 8-lanes-per-key streaming over V, probabilities broadcast from shared memory.
// All NB rows (beams of one clip) share every K/V byte that is loaded.
constexpr int TA_THREADS = 128;
constexpr int TA_GROUPS = TA_THREADS / 8;   // 16 lane groups
constexpr int TA_KPB = TA_GROUPS * 8;       // 128 keys per block-wide step

template <bool COHERENT = false>
__device__ __forceinline__ uint4 ld16(const bf16* p) {
  return COHERENT ? __ldcg(reinterpret_cast<const uint4*>(p)) : __ldg(reinterpret_cast<const uint4*>(p));
}
template <bool COHERENT = false>
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  const uint4 u = ld16<COHERENT>(p);
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ float dot8u(const float (&q)[8], const uint4& u) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  float s = q[0] * a.x;
  s = fmaf(q[1], a.y, s); s = fmaf(q[2], b.x, s); s = fmaf(q[3], b.y, s);
  s = fmaf(q[4], c.x, s); s = fmaf(q[5], c.y, s); s = fmaf(q[6], d.x, s); s = fmaf(q[7], d.y, s);
  return s;
}
// 8 partial sums per lane (one per key) -> lane `sub` of the 8-lane group ends with the full sum of key `sub`
__device__ __forceinline__ float transpose_reduce8(const float (&p)[8], int lane, uint32_t gmask) {
  const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4;
  float v1[4], v2[2];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float keep = b0 ? p[2 * t + 1] : p[2 * t], send = b0 ? p[2 * t] : p[2 * t + 1];
    v1[t] = keep + __shfl_xor_sync(gmask, send, 1);
  }
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const float keep = b1 ? v1[2 * t + 1] : v1[2 * t], send = b1 ? v1[2 * t] : v1[2 * t + 1];
    v2[t] = keep + __shfl_xor_sync(gmask, send, 2);
  }
  const float keep = b2 ? v2[1] : v2[0], send = b2 ? v2[0] : v2[1];
  return keep + __shfl_xor_sync(gmask, send, 4);
}

// COHERENT: q, the text K/V plane and the ancestor table are written by OTHER CTAs of the SAME kernel launch (persistent
// decode kernel): they are read with ld.global.cg (L2) instead of the non-coherent / L1-cached path.
// sync(): barrier over the TA_THREADS threads that execute this body together (tid = 0 .. TA_THREADS-1).
template <int NB, bool COHERENT, class Sync>
__device__ __forceinline__ void text_attention_body(const TextAttnArgs& a, float scale_log2, int kcap, float* sm, int clip, int chunk,
                                                    int h, int split, int tid, Sync sync) {
  float* sc = sm;                              // [NB][kcap] scores -> probabilities
  float* red = sm + NB * kcap;                 // [TA_GROUPS][NB][HD] value partials (also max/sum scratch)
  const int warp = tid >> 5, lane = tid & 31;
  const int grp = tid >> 3, sub = lane & 7;
  const int d0 = sub * 8;
  const uint32_t gmask = 0xFFu << (lane & 24);
  const int rl0 = chunk * NB;
  const int n_loc = min(NB, a.rows_per_clip - rl0);
  const int width = a.heads * HD;

  float q[NB][8];
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    if (r < n_loc) {
      load8<COHERENT>(a.q + (size_t)(clip * a.rows_per_clip + rl0 + r) * a.ldq + h * HD + d0, q[r]);
#pragma unroll
      for (int i = 0; i < 8; ++i) q[r][i] *= scale_log2;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) q[r][i] = 0.f;
    }
  }
  const int per = (a.Nv + a.splits - 1) / a.splits;
  const int k_begin = split * per, k_end = min(a.Nv, k_begin + per);
  const int n_vis = max(0, k_end - k_begin);
  const bool do_text = split == a.splits - 1;
  int nt_max = 0;
  int nt[NB];
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    nt[r] = 0;
    if (do_text && r < n_loc) {
      const int row = clip * a.rows_per_clip + rl0 + r;
      nt[r] = a.n_text ? a.n_text[row] : a.n_text_const;
    }
    nt_max = max(nt_max, nt[r]);
  }

  // ---- phase 1a: visual-key scores
  const bf16* kp = a.vis_kv + (size_t)clip * a.Nv * a.ld_vis + a.k_off + h * HD + d0;
  for (int base = k_begin; base < k_end; base += TA_KPB) {
    const int k0 = base + grp * 8;
    uint4 kv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = min(k0 + i, k_end - 1);  // clamp: tail lanes re-read the last key, their scores are not stored
      kv[i] = ld16<false>(kp + (size_t)k * a.ld_vis);
    }
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      if (r < n_loc) {
        float part[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) part[i] = dot8u(q[r], kv[i]);
        const float sv = transpose_reduce8(part, lane, gmask);
        if (k0 + sub < k_end) sc[r * kcap + (k0 + sub - k_begin)] = sv;
      }
    }
  }
  // ---- phase 1b: text-key scores (few keys, per-row slots through the ancestor table)
  if (do_text) {
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      if (r >= n_loc) break;
      const int row = clip * a.rows_per_clip + rl0 + r;
      for (int s = grp; s < nt[r]; s += TA_GROUPS) {
        int slot = row;
        if (a.anc != nullptr && s < nt[r] - 1) slot = COHERENT ? __ldcg(a.anc + (size_t)row * a.anc_ld + s) : a.anc[(size_t)row * a.anc_ld + s];
        if (a.text_slot_is_clip) slot = clip;
        const uint4 u = ld16<COHERENT>(a.txt_kv + ((size_t)s * a.txt_slots + slot) * (2 * width) + h * HD + d0);
        float sv = dot8u(q[r], u);
        sv += __shfl_xor_sync(gmask, sv, 1);
        sv += __shfl_xor_sync(gmask, sv, 2);
        sv += __shfl_xor_sync(gmask, sv, 4);
        if (sub == 0) sc[r * kcap + n_vis + s] = sv;
      }
    }
  }
  sync();

  // ---- softmax statistics per row (block-wide), probabilities written back in place
  float m_row[NB], l_row[NB];
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    const int n = n_vis + nt[r];
    float m = -INFINITY;
    if (r < n_loc)
      for (int i = tid; i < n; i += TA_THREADS) m = fmaxf(m, sc[r * kcap + i]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    sync();
    m = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    sync();
    float l = 0.f;
    if (r < n_loc && m > -INFINITY)
      for (int i = tid; i < n; i += TA_THREADS) {
        const float p = exp2f(sc[r * kcap + i] - m);
        sc[r * kcap + i] = p;
        l += p;
      }
    l = warp_sum(l);
    if (lane == 0) red[warp] = l;
    sync();
    l = red[0] + red[1] + red[2] + red[3];
    sync();
    m_row[r] = m;
    l_row[r] = l;
  }

  // ---- phase 2: values
  float acc[NB][8];
#pragma unroll
  for (int r = 0; r < NB; ++r)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[r][i] = 0.f;
  const bf16* vp = a.vis_kv + (size_t)clip * a.Nv * a.ld_vis + a.v_off + h * HD + d0;
  for (int base = k_begin; base < k_end; base += TA_KPB) {
    const int k0 = base + grp * 8;
    uint4 vv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = min(k0 + i, k_end - 1);
      vv[i] = ld16<false>(vp + (size_t)k * a.ld_vis);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (k0 + i < k_end) {
        const float2 x0 = unpack_bf16(vv[i].x), x1 = unpack_bf16(vv[i].y), x2 = unpack_bf16(vv[i].z), x3 = unpack_bf16(vv[i].w);
#pragma unroll
        for (int r = 0; r < NB; ++r) {
          if (r < n_loc) {
            const float p = sc[r * kcap + (k0 + i - k_begin)];
            acc[r][0] = fmaf(p, x0.x, acc[r][0]); acc[r][1] = fmaf(p, x0.y, acc[r][1]);
            acc[r][2] = fmaf(p, x1.x, acc[r][2]); acc[r][3] = fmaf(p, x1.y, acc[r][3]);
            acc[r][4] = fmaf(p, x2.x, acc[r][4]); acc[r][5] = fmaf(p, x2.y, acc[r][5]);
            acc[r][6] = fmaf(p, x3.x, acc[r][6]); acc[r][7] = fmaf(p, x3.y, acc[r][7]);
          }
        }
      }
    }
  }
  if (do_text) {
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      if (r >= n_loc) break;
      const int row = clip * a.rows_per_clip + rl0 + r;
      for (int s = grp; s < nt[r]; s += TA_GROUPS) {
        int slot = row;
        if (a.anc != nullptr && s < nt[r] - 1) slot = COHERENT ? __ldcg(a.anc + (size_t)row * a.anc_ld + s) : a.anc[(size_t)row * a.anc_ld + s];
        if (a.text_slot_is_clip) slot = clip;
        float v[8];
        load8<COHERENT>(a.txt_kv + ((size_t)s * a.txt_slots + slot) * (2 * width) + width + h * HD + d0, v);
        const float p = sc[r * kcap + n_vis + s];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[r][i] = fmaf(p, v[i], acc[r][i]);
      }
    }
  }
  // ---- reduce the 16 lane groups through shared memory
#pragma unroll
  for (int r = 0; r < NB; ++r)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[(grp * NB + r) * HD + d0 + i] = acc[r][i];
  sync();
  for (int idx = tid; idx < n_loc * HD; idx += TA_THREADS) {
    const int r = idx / HD, d = idx % HD;
    float o = 0.f;
#pragma unroll
    for (int g2 = 0; g2 < TA_GROUPS; ++g2) o += red[(g2 * NB + r) * HD + d];
    const int row = clip * a.rows_per_clip + rl0 + r;
    if (a.splits == 1) {
      a.out[(size_t)row * a.ldo + h * HD + d] = __float2bfloat16(o / l_row[r]);
    } else {
      float* p = a.partial + (((size_t)row * a.heads + h) * a.splits + split) * (HD + 2);
      if (d == 0) {
        p[0] = m_row[r];
        p[1] = l_row[r];
      }
      p[2 + d] = o;
    }
  }
}

// Combine of the key-split partials (m, l, o[HD]) of one (row, head) by one warp: lane -> dims 2 * lane, 2 * lane + 1.
// splits <= MAX_SPLITS.  Every load is issued before the first use (one L2 round trip): lane s fetches (m_s, l_s) and the warp
// reads them by shuffle, the o pairs of all splits go to registers.  Same operations in the same order as the plain loop.
constexpr int MAX_SPLITS = 16;
template <bool COHERENT>
__device__ __forceinline__ uint32_t combine_partials(const float* p, int splits, int lane) {
  auto ld = [](const float* q) { return COHERENT ? __ldcg(q) : *q; };
  const float pm_l = lane < splits ? ld(p + lane * (HD + 2)) : -INFINITY;
  const float pl_l = lane < splits ? ld(p + lane * (HD + 2) + 1) : 0.f;
  float2 po[MAX_SPLITS];
#pragma unroll
  for (int j = 0; j < MAX_SPLITS; ++j) {
    const float2* q = reinterpret_cast<const float2*>(p + j * (HD + 2) + 2) + lane;
    po[j] = j < splits ? (COHERENT ? __ldcg(q) : *q) : make_float2(0.f, 0.f);
  }
  float m = -INFINITY;
  for (int s = 0; s < splits; ++s) m = fmaxf(m, __shfl_sync(0xffffffffu, pm_l, s));
  float l = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
  for (int j = 0; j < MAX_SPLITS; ++j) {
    if (j < splits) {  // warp-uniform
      const float pm = __shfl_sync(0xffffffffu, pm_l, j);
      const float c = (pm == -INFINITY) ? 0.f : exp2f(pm - m);
      l += __shfl_sync(0xffffffffu, pl_l, j) * c;
      o0 += po[j].x * c;
      o1 += po[j].y * c;
    }
  }
  const float inv = 1.f / l;
  return pack_bf16(o0 * inv, o1 * inv);
}

}  // namespace text_attn_dev
