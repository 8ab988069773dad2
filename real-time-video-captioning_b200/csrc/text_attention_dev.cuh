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
// shared-memory floats of one (virtual) CTA: scores + value partials (scalar body: one partial per lane group; tensor body: per warp)
__host__ __device__ constexpr int ta_smem_floats(int nb, int kcap) { return nb * kcap + (nb >= 2 ? 4 : TA_GROUPS) * nb * HD; }

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
__device__ __forceinline__ void text_attention_body_scalar(const TextAttnArgs& a, float scale_log2, int kcap, float* sm, int clip, int chunk,
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

// 16-byte streaming load of read-only K/V data; the L2 fetches the whole 128-byte line the piece lies in
__device__ __forceinline__ uint4 ld16_line(const bf16* p) {
  uint4 u;
  asm volatile("ld.global.nc.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "l"(p));
  return u;
}

// Several query rows per clip (NB >= 2: the beams of a clip, or the text rows of a teacher-forced pass).  The scalar body
// above spends NB x 16 FMAs + the transposing butterflies per loaded 16 bytes: at NB = 4 it ran at a third of the HBM rate
// (399 us per 0.93 GB launch at 256 clips x 4 beams, profiles/r02_timeline_beam4_before.md).  Here the visual keys go through the
// warp-level tensor path (mma.sync m16n8k16 bf16, fp32 accumulate; the work is HBM bound, tcgen05's 64-row minimum would
// compute 16x padding) with fragments assembled in registers straight from 16-byte global loads -- no shared-memory staging:
//   scores  S^T[16 keys x 8 beams] = K[16 keys x 64 dims] . Q^T : thread (g, t) of a warp loads 2 x 16 bytes of key rows g and
//           g + 8; the contraction order over the 64 dims is a permutation that is the same for K and Q, so the loaded words
//           ARE the A fragments (no shuffles);
//   values  O^T[64 dims x 8 beams] += V^T[64 dims x 16 keys] . P^T : thread (g, t) loads the 16-byte chunk g of keys 2t, 2t+1,
//           2t+8, 2t+9; the two dims of one loaded word are the rows g / g + 8 of one 16-dim tile (a permutation of the output
//           rows), so an A fragment is one byte-permute of two loaded words.  P is split in two bf16 (p = hi + lo, 2^-17
//           relative) so the probabilities carry fp32-like precision as in the scalar body; the MMA count does not matter here.
// The few text keys (own slots per row through the ancestor table) keep the scalar 8-lanes-per-key code.
// Shared memory: sc [NB][kcap], red [4 warps][NB][HD] (ta_smem_floats).
template <int NB, bool COHERENT, int UT, int TP, class Sync>
__device__ __forceinline__ void text_attention_body_mma(const TextAttnArgs& a, float scale_log2, int kcap, float* sm, int clip, int chunk,
                                                        int h, int split, int tid, Sync sync) {
  static_assert(NB == 2 || NB == 4, "beam rows are columns of the N = 8 MMA; lane groups map to rows by grp % NB");
  float* sc = sm;               // [NB][kcap] scores -> probabilities
  float* red = sm + NB * kcap;  // [4 warps][NB][HD] value partials (also max/sum scratch)
  const int warp = tid >> 5, lane = tid & 31;
  const int grp = tid >> 3, sub = lane & 7;
  const int g = lane >> 2, t = lane & 3;
  const int d0 = sub * 8;
  const uint32_t gmask = 0xFFu << (lane & 24);
  const int rl0 = chunk * NB;
  const int n_loc = min(NB, a.rows_per_clip - rl0);
  const int width = a.heads * HD;
  const int row0 = clip * a.rows_per_clip + rl0;

  const int per = (a.Nv + a.splits - 1) / a.splits;
  const int k_begin = split * per, k_end = min(a.Nv, k_begin + per);
  const int n_vis = max(0, k_end - k_begin);
  const bool do_text = split == a.splits - 1;
  int nt[NB];
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    nt[r] = 0;
    if (do_text && r < n_loc) nt[r] = a.n_text ? a.n_text[row0 + r] : a.n_text_const;
  }
  // Text keys (few; own slots per row through the ancestor table): lane group grp serves row r_g = grp % NB at the text positions
  // s = grp / NB + (16 / NB) j.  A CTA that walked them in place paid two dependent global loads (ancestor slot -> K or V row) per
  // row and phase with nothing else in flight: ~18 us of a 73 us CTA at 4 rows.  So the slots of the first TP positions are fetched
  // at the top, their K rows right after the visual-key loop, their V rows before the softmax pass, each a full phase ahead of use.
  // TP: positions per lane group held in registers (arithmetic does not depend on it)
  constexpr int SSTEP = TA_GROUPS / NB;
  const int r_g = grp % NB, s_g = grp / NB, row_g = row0 + r_g;
  int nt_g = 0;
  if (do_text && r_g < n_loc) nt_g = a.n_text ? a.n_text[row_g] : a.n_text_const;
  auto text_slot = [&](int s) -> int {
    int slot = row_g;
    if (a.anc != nullptr && s < nt_g - 1) slot = COHERENT ? __ldcg(a.anc + (size_t)row_g * a.anc_ld + s) : a.anc[(size_t)row_g * a.anc_ld + s];
    if (a.text_slot_is_clip) slot = clip;
    return slot;
  };
  auto text_row = [&](int s, int slot) { return a.txt_kv + ((size_t)s * a.txt_slots + slot) * (2 * width) + h * HD + d0; };
  int tslot[TP];
#pragma unroll
  for (int j = 0; j < TP; ++j) tslot[j] = s_g + SSTEP * j < nt_g ? text_slot(s_g + SSTEP * j) : 0;
  // UT: 16-key tiles per warp and block-wide step (4 UT 16-byte loads in flight per thread)
  constexpr int KPB = 4 * UT * 16;       // keys per block-wide step (4 warps)

  // Not kept: the two streaming loops on two register buffers (loop unrolled by two, the loads of step i + 1 issued before step i
  // is computed, the first V step before the softmax pass): 128 registers / 4 CTAs per SM 212 us, one tile per step at 80 registers
  // 216 us, against 178-182 us for this form at 80 registers / 6 CTAs per SM (256 clips x 4 beams); this form at 64 registers /
  // 8 CTAs per SM: 187 us (profiles/r02_text_attention_mma.md).
  const bf16* vp = a.vis_kv + (size_t)clip * a.Nv * a.ld_vis + a.v_off + h * HD + g * 8;
  auto load_v = [&](uint4 (&vv)[UT][4], int base) {
#pragma unroll
    for (int u = 0; u < UT; ++u) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const int k = min(base + u * 16 + 2 * t + (kk & 1) + (kk >> 1) * 8, k_end - 1);  // clamped: its probability is 0
        vv[u][kk] = ld16_line(vp + (size_t)k * a.ld_vis);
      }
    }
  };
  const int base0 = k_begin + warp * (UT * 16);

  // ---- phase 1a: visual-key scores.  B fragments: beam g's query, words (t, t + 4 of the row's eight 16-byte chunks)
  {
    const bf16* kp = a.vis_kv + (size_t)clip * a.Nv * a.ld_vis + a.k_off + h * HD + t * 8;
    auto load_k = [&](uint4 (&kv)[UT][4], int base) {
#pragma unroll
      for (int u = 0; u < UT; ++u) {
        const int ka = min(base + u * 16 + g, k_end - 1), kb = min(base + u * 16 + g + 8, k_end - 1);  // clamped: not stored
        kv[u][0] = ld16_line(kp + (size_t)ka * a.ld_vis);
        kv[u][1] = ld16_line(kp + (size_t)ka * a.ld_vis + 32);
        kv[u][2] = ld16_line(kp + (size_t)kb * a.ld_vis);
        kv[u][3] = ld16_line(kp + (size_t)kb * a.ld_vis + 32);
      }
    };
    uint32_t qb[8];
    if (g < n_loc) {
      const bf16* qp = a.q + (size_t)(row0 + g) * a.ldq + h * HD;
      const uint4 u0 = ld16<COHERENT>(qp + t * 8), u1 = ld16<COHERENT>(qp + 32 + t * 8);
      qb[0] = u0.x; qb[1] = u0.y; qb[2] = u0.z; qb[3] = u0.w; qb[4] = u1.x; qb[5] = u1.y; qb[6] = u1.z; qb[7] = u1.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) qb[i] = 0u;
    }
    auto scores = [&](const uint4 (&kv)[UT][4], int base) {
#pragma unroll
      for (int u = 0; u < UT; ++u) {
        const uint32_t wa[8] = {kv[u][0].x, kv[u][0].y, kv[u][0].z, kv[u][0].w, kv[u][1].x, kv[u][1].y, kv[u][1].z, kv[u][1].w};
        const uint32_t wb[8] = {kv[u][2].x, kv[u][2].y, kv[u][2].z, kv[u][2].w, kv[u][3].x, kv[u][3].y, kv[u][3].z, kv[u][3].w};
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t af[4] = {wa[2 * ks], wb[2 * ks], wa[2 * ks + 1], wb[2 * ks + 1]};
          ptx::mma_bf16_16816(s4, af, qb[2 * ks], qb[2 * ks + 1]);
        }
        // s4: (key g, beams 2t, 2t+1), (key g + 8, beams 2t, 2t+1)
        const int ka = base + u * 16 + g, kb = ka + 8;
        if (2 * t < n_loc) {
          if (ka < k_end) sc[(2 * t) * kcap + (ka - k_begin)] = s4[0] * scale_log2;
          if (kb < k_end) sc[(2 * t) * kcap + (kb - k_begin)] = s4[2] * scale_log2;
        }
        if (2 * t + 1 < n_loc) {
          if (ka < k_end) sc[(2 * t + 1) * kcap + (ka - k_begin)] = s4[1] * scale_log2;
          if (kb < k_end) sc[(2 * t + 1) * kcap + (kb - k_begin)] = s4[3] * scale_log2;
        }
      }
    };
    for (int base = base0; base < k_end; base += KPB) {
      uint4 kv[UT][4];
      load_k(kv, base);
      scores(kv, base);
    }
  }
  // ---- phase 1b: text-key scores
  uint4 tkv[TP];
  if (do_text) {
#pragma unroll
    for (int j = 0; j < TP; ++j)
      if (s_g + SSTEP * j < nt_g) tkv[j] = ld16<COHERENT>(text_row(s_g + SSTEP * j, tslot[j]));
    float q[8];
    if (nt_g > 0) {
      load8<COHERENT>(a.q + (size_t)row_g * a.ldq + h * HD + d0, q);
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] *= scale_log2;
    }
    auto score = [&](int s, const uint4& u) {
      float sv = dot8u(q, u);
      sv += __shfl_xor_sync(gmask, sv, 1);
      sv += __shfl_xor_sync(gmask, sv, 2);
      sv += __shfl_xor_sync(gmask, sv, 4);
      if (sub == 0) sc[r_g * kcap + n_vis + s] = sv;
    };
#pragma unroll
    for (int j = 0; j < TP; ++j)
      if (s_g + SSTEP * j < nt_g) score(s_g + SSTEP * j, tkv[j]);
    for (int s2 = s_g + SSTEP * TP; s2 < nt_g; s2 += SSTEP) score(s2, ld16<COHERENT>(text_row(s2, text_slot(s2))));
    // the V rows of the same positions: in flight during the softmax pass
#pragma unroll
    for (int j = 0; j < TP; ++j)
      if (s_g + SSTEP * j < nt_g) tkv[j] = ld16<COHERENT>(text_row(s_g + SSTEP * j, tslot[j]) + width);
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
    if (lane == 0) red[r * 8 + warp] = m;
    m_row[r] = m;
  }
  sync();
#pragma unroll
  for (int r = 0; r < NB; ++r) {
    const int n = n_vis + nt[r];
    const float m = fmaxf(fmaxf(red[r * 8], red[r * 8 + 1]), fmaxf(red[r * 8 + 2], red[r * 8 + 3]));
    float l = 0.f;
    if (r < n_loc && m > -INFINITY)
      for (int i = tid; i < n; i += TA_THREADS) {
        const float p = exp2f(sc[r * kcap + i] - m);
        sc[r * kcap + i] = p;
        l += p;
      }
    l = warp_sum(l);
    if (lane == 0) red[r * 8 + 4 + warp] = l;
    m_row[r] = m;
  }
  sync();
#pragma unroll
  for (int r = 0; r < NB; ++r) l_row[r] = red[r * 8 + 4] + red[r * 8 + 5] + red[r * 8 + 6] + red[r * 8 + 7];
  sync();  // red is reused for the value partials

  // ---- text values: group grp's share of row r_g -> red[warp][r_g][dims of the lane] (the warp's visual partial is added below)
  if (do_text) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    auto value = [&](int s, const uint4& u) {
      const float p = sc[r_g * kcap + n_vis + s];
      const float2 x0 = unpack_bf16(u.x), x1 = unpack_bf16(u.y), x2 = unpack_bf16(u.z), x3 = unpack_bf16(u.w);
      acc[0] = fmaf(p, x0.x, acc[0]); acc[1] = fmaf(p, x0.y, acc[1]); acc[2] = fmaf(p, x1.x, acc[2]); acc[3] = fmaf(p, x1.y, acc[3]);
      acc[4] = fmaf(p, x2.x, acc[4]); acc[5] = fmaf(p, x2.y, acc[5]); acc[6] = fmaf(p, x3.x, acc[6]); acc[7] = fmaf(p, x3.y, acc[7]);
    };
#pragma unroll
    for (int j = 0; j < TP; ++j)
      if (s_g + SSTEP * j < nt_g) value(s_g + SSTEP * j, tkv[j]);
    for (int s2 = s_g + SSTEP * TP; s2 < nt_g; s2 += SSTEP) value(s2, ld16<COHERENT>(text_row(s2, text_slot(s2)) + width));
    if (NB == 2) {  // lane groups q and q + 2 of a warp serve the same row
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    }
    if (NB == 4 || lane < 16) {
#pragma unroll
      for (int i = 0; i < 8; ++i) red[(warp * NB + r_g) * HD + d0 + i] = acc[i];
    }
    __syncwarp();
  }

  // ---- phase 2: values.  o[j]: (dim 8g + 2j, beams 2t, 2t+1), (dim 8g + 2j + 1, beams 2t, 2t+1)
  float o[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[j][i] = 0.f;
  {
    const float* pr = sc + min(g, NB - 1) * kcap - k_begin;
    const bool live = g < n_loc;
    auto values = [&](const uint4 (&vv)[UT][4], int base) {
#pragma unroll
      for (int u = 0; u < UT; ++u) {
        const int k0 = base + u * 16 + 2 * t;
        float p[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const int k = k0 + (kk & 1) + (kk >> 1) * 8;
          p[kk] = (live && k < k_end) ? pr[k] : 0.f;
        }
        const uint32_t ph0 = pack_bf16(p[0], p[1]), ph1 = pack_bf16(p[2], p[3]);
        const float2 h0 = unpack_bf16(ph0), h1 = unpack_bf16(ph1);
        const uint32_t pl0 = pack_bf16(p[0] - h0.x, p[1] - h0.y), pl1 = pack_bf16(p[2] - h1.x, p[3] - h1.y);
        const uint32_t x0[4] = {vv[u][0].x, vv[u][0].y, vv[u][0].z, vv[u][0].w};
        const uint32_t x1[4] = {vv[u][1].x, vv[u][1].y, vv[u][1].z, vv[u][1].w};
        const uint32_t x2[4] = {vv[u][2].x, vv[u][2].y, vv[u][2].z, vv[u][2].w};
        const uint32_t x3[4] = {vv[u][3].x, vv[u][3].y, vv[u][3].z, vv[u][3].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t af[4] = {__byte_perm(x0[j], x1[j], 0x5410), __byte_perm(x0[j], x1[j], 0x7632),
                                  __byte_perm(x2[j], x3[j], 0x5410), __byte_perm(x2[j], x3[j], 0x7632)};
          ptx::mma_bf16_16816(o[j], af, ph0, ph1);
          ptx::mma_bf16_16816(o[j], af, pl0, pl1);
        }
      }
    };
    for (int base = base0; base < k_end; base += KPB) {
      uint4 vv[UT][4];
      load_v(vv, base);
      values(vv, base);
    }
  }
  // this warp's visual partial -> red[warp][beam][dim], on top of the text share
  if (2 * t < NB) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float* rp = red + (warp * NB + 2 * t) * HD + 8 * g + 2 * j;
      if (do_text) {
        rp[0] += o[j][0];
        rp[HD] += o[j][1];
        rp[1] += o[j][2];
        rp[HD + 1] += o[j][3];
      } else {
        rp[0] = o[j][0];
        rp[HD] = o[j][1];
        rp[1] = o[j][2];
        rp[HD + 1] = o[j][3];
      }
    }
  }
  sync();
  for (int idx = tid; idx < n_loc * HD; idx += TA_THREADS) {
    const int r = idx / HD, d = idx % HD;
    const float ov = ((red[(0 * NB + r) * HD + d] + red[(1 * NB + r) * HD + d]) + red[(2 * NB + r) * HD + d]) + red[(3 * NB + r) * HD + d];
    const int row = row0 + r;
    float l = l_row[0];
#pragma unroll
    for (int rr = 1; rr < NB; ++rr) l = (r == rr) ? l_row[rr] : l;
    if (a.splits == 1) {
      a.out[(size_t)row * a.ldo + h * HD + d] = __float2bfloat16(ov / l);
    } else {
      float m = m_row[0];
#pragma unroll
      for (int rr = 1; rr < NB; ++rr) m = (r == rr) ? m_row[rr] : m;
      float* p = a.partial + (((size_t)row * a.heads + h) * a.splits + split) * (HD + 2);
      if (d == 0) {
        p[0] = m;
        p[1] = l;
      }
      p[2 + d] = ov;
    }
  }
}

// One query row per clip (greedy decode): the scalar body streams at 92 % of the copy bandwidth.  Several rows: tensor path.
// UT (16-key tiles per warp and step) fixes which warp sums which keys: callers that must agree bit for bit use the same UT
// (TA_UT_LATENCY: the launch path on a few clips == the persistent decode kernel).
constexpr int TA_UT_LATENCY = 1;
template <int NB, bool COHERENT, int UT = TA_UT_LATENCY, int TP = 1, class Sync>
__device__ __forceinline__ void text_attention_body(const TextAttnArgs& a, float scale_log2, int kcap, float* sm, int clip, int chunk,
                                                    int h, int split, int tid, Sync sync) {
  if constexpr (NB >= 2) text_attention_body_mma<NB, COHERENT, UT, TP>(a, scale_log2, kcap, sm, clip, chunk, h, split, tid, sync);
  else text_attention_body_scalar<NB, COHERENT>(a, scale_log2, kcap, sm, clip, chunk, h, split, tid, sync);
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
