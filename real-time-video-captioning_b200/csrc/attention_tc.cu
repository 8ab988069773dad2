// tcgen05 flash attention over fixed-length row groups (ViT frames, K10: the decoder's visual block).
//
// One CTA = one (group, head, 128-query tile); two CTAs share an SM so that one CTA's softmax overlaps the other's
// MMAs and barrier latencies.  Per 64-key block:
//     S  = Q K^T          tcgen05.mma  M=128, N<=64, K=64   (Q, K: K-major 128B-swizzled TMA tiles)      -> TMEM
//     P  = exp2(S*c - m)  4 softmax warps, one query row per thread (tcgen05.ld 32x32b), running max / sum in registers,
//                         P written as the bf16 K-major A operand into 128B-swizzled shared memory
//     Oj = P V            tcgen05.mma  M=128, N=64, K<=64   (V: the TMA tile used as an MN-major B operand)   -> TMEM
//     O  = O*alpha + Oj   folded in registers by the softmax threads (no TMEM read-modify-write, the MMA warp never
//                         waits for a rescale)
// The last key block of a group issues only the 16-key steps that hold valid keys (197 = 3*64 + 5 -> N = 16).
// The mma.sync predecessor of this kernel reached 123-222 TFLOP/s (the legacy tensor path peaks near 510 TFLOP/s on
// B200, ncu: hmma pipe 40-45 % active); this version is bound by the exp2 rate of the softmax instead.
#include <cuda.h>
#include <math.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"

bool gemm_get_tensor_map(const bf16* ptr, int rows, int cols, int ld, int box_cols, int box_rows, CUtensorMap* out);

namespace {

constexpr int HD = 64;
constexpr int BQ = 128;   // query rows per CTA = TMEM lanes
constexpr int BKV = 64;   // keys per block
constexpr int KV_STAGES = 3;
constexpr int Q_BYTES = BQ * 128;         // 16 KB
constexpr int KV_TILE_BYTES = BKV * 128;  // 8 KB each for K and V
constexpr int P_BYTES = BQ * 128;         // 16 KB: P[128 x 64 keys] bf16
constexpr int SMEM_BYTES = Q_BYTES + KV_STAGES * 2 * KV_TILE_BYTES + P_BYTES + 1024 + 128 + 2 * BQ * 4;
constexpr int NUM_SM_WARPS = 8;   // softmax warps: (lane quarter q, column half ch) = (warp & 3, (warp - 2) >> 2)
constexpr int NUM_THREADS = 64 + 32 * NUM_SM_WARPS;  // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2-9: softmax
constexpr int TMEM_COLS = 256;    // S double buffer: columns [0,64) and [64,128); Oj: columns [128,192)

// V tile [keys][64 dims] (rows of 128 B, 128B swizzle) read as an MN-major B operand (N = dims, K = keys):
// 8 key rows form one 1024-byte swizzle atom, atoms follow each other along K every 1024 B (SBO).
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr) {
  const uint32_t lo = ((smem_addr >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
__device__ __forceinline__ uint32_t idesc_qk(int n) {  // A, B K-major
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
}
__device__ __forceinline__ uint32_t idesc_pv() {  // A K-major, B MN-major, N = 64
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(NUM_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    bf16* __restrict__ out, int ldo, int group_len, int heads, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  auto sK = [&](int s) { return smem_base + Q_BYTES + s * 2 * KV_TILE_BYTES; };
  auto sV = [&](int s) { return sK(s) + KV_TILE_BYTES; };
  const uint32_t sP = smem_base + Q_BYTES + KV_STAGES * 2 * KV_TILE_BYTES;
  const uint32_t bar_base = sP + P_BYTES;
  const uint32_t q_full = bar_base;
  auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar_base + 8u * (1 + KV_STAGES + s); };
  auto s_full = [&](int b) { return bar_base + 8u * (1 + 2 * KV_STAGES + b); };
  const uint32_t p_full = bar_base + 8u * (3 + 2 * KV_STAGES), o_full = bar_base + 8u * (4 + 2 * KV_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (5 + 2 * KV_STAGES);
  const uint32_t sMax = bar_base + 128;  // float [2 halves][128 rows]: per-block row maxima exchanged between column halves

  const int warp = __shfl_sync(0xffffffffu, (int)threadIdx.x / 32, 0);
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, g = blockIdx.z;
  const int width = heads * HD;
  const int row0 = g * group_len;  // first row of the group in the qkv matrix
  const int q0 = qt * BQ;
  const int n_blocks = (group_len + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_kv);
  }
  if (warp == 1) {
    if (lane == 0) {
      ptx::mbar_init(q_full, 1);
      for (int s = 0; s < KV_STAGES; ++s) {
        ptx::mbar_init(kv_full(s), 1);
        ptx::mbar_init(kv_empty(s), 1);
      }
      ptx::mbar_init(s_full(0), 1);
      ptx::mbar_init(s_full(1), 1);
      ptx::mbar_init(p_full, NUM_SM_WARPS);  // one arrival per softmax warp
      ptx::mbar_init(o_full, 1);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const uint32_t tS = tmem_base, tO = tmem_base + 128;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, Q_BYTES);
      ptx::tma_load_2d(sQ, &tmap_q, q_full, h * HD, row0 + q0);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j % KV_STAGES;
        ptx::mbar_wait(kv_empty(st), (uint32_t)(((j / KV_STAGES) & 1) ^ 1));
        ptx::mbar_arrive_expect_tx(kv_full(st), 2 * KV_TILE_BYTES);
        ptx::tma_load_2d(sK(st), &tmap_kv, kv_full(st), width + h * HD, row0 + j * BKV);
        ptx::tma_load_2d(sV(st), &tmap_kv, kv_full(st), 2 * width + h * HD, row0 + j * BKV);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      ptx::mbar_wait(q_full, 0);
      ptx::tc_fence_after();
      const uint64_t dq = ptx::umma_desc_sw128_kmajor(sQ);
      const uint64_t dp = ptx::umma_desc_sw128_kmajor(sP);
      // S_j = Q K_j^T over the head dim (4 steps of 16) into S buffer (j & 1).  The buffer is free: every softmax
      // warp finished reading S_{j-2} before it arrived on p_full(j-2), which this thread has already waited for.
      auto issue_s = [&](int j) {
        const int st = j % KV_STAGES;
        const int nk = min(BKV, group_len - j * BKV);
        const int nk16 = (nk + 15) & ~15;
        ptx::mbar_wait(kv_full(st), (uint32_t)((j / KV_STAGES) & 1));
        ptx::tc_fence_after();
        const uint64_t dk = ptx::umma_desc_sw128_kmajor(sK(st));
        const uint32_t id_s = idesc_qk(nk16);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          ptx::umma_bf16(tS + (uint32_t)((j & 1) * 64), dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), id_s, k != 0);
        ptx::umma_commit(s_full(j & 1));
      };
      issue_s(0);
      for (int j = 0; j < n_blocks; ++j) {
        const int st = j % KV_STAGES;
        const int nk = min(BKV, group_len - j * BKV);
        const int nk16 = (nk + 15) & ~15;
        if (j + 1 < n_blocks) issue_s(j + 1);  // runs while the softmax warps work on block j
        // Oj = P V once the softmax warps have written P (and folded the previous Oj out of TMEM)
        ptx::mbar_wait(p_full, (uint32_t)(j & 1));
        ptx::tc_fence_after();
        const uint64_t dv = umma_desc_sw128_mnmajor(sV(st));
        const uint32_t id_o = idesc_pv();
        for (int k = 0; k < nk16 / 16; ++k)  // 16 keys per step: P advances 32 B along its rows, V advances 2 atoms
          ptx::umma_bf16(tO, dp + (uint64_t)(2 * k), dv + (uint64_t)(128 * k), id_o, k != 0);
        ptx::umma_commit(o_full);
        ptx::umma_commit(kv_empty(st));
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warps: two threads per query row
    // thread (row r, half ch): keys [32 ch, 32 ch + 32) of every block and output dims [32 ch, 32 ch + 32)
    const int quarter = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int r = quarter * 32 + lane;  // row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const bool warp_has_rows = q0 + quarter * 32 < group_len;  // warp-uniform, same for both halves of a row
    const uint32_t my_max = sMax + (uint32_t)(ch * BQ + r) * 4u, peer_max = sMax + (uint32_t)((ch ^ 1) * BQ + r) * 4u;
    const int pair_bar = 1 + quarter;  // named barrier shared by the two warps that hold the same rows
    float o_acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;

    auto fold = [&](float alpha) {  // o_acc = o_acc * alpha + Oj (this thread's 32 dims of TMEM columns [64,128))
#pragma unroll
      for (int c = 0; c < 32; c += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tO + lane_off + (uint32_t)(ch * 32 + c), v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o_acc[c + i] = fmaf(o_acc[c + i], alpha, __uint_as_float(v[i]));
      }
    };

    // One 64-key block for this thread's 32 columns.  TAIL = the last block of a group (masked keys, possibly fewer
    // 16-key steps); every other block takes the straight-line path.
    auto block = [&](int j, auto tail_tag) {
      constexpr bool TAIL = decltype(tail_tag)::value;
      const int nk = TAIL ? group_len - j * BKV : BKV;
      const int nk16 = (nk + 15) & ~15;
      const int c0 = ch * 32;
      const uint32_t tSj = tS + (uint32_t)((j & 1) * 64) + lane_off + (uint32_t)c0;
      const bool have_cols = !TAIL || c0 < nk16;  // warp-uniform
      float s[32];
      float mx = -INFINITY;
      if (have_cols) {
        uint32_t v0[16], v1[16];
        tmem_ld_32x32b_x16(tSj, v0);
        tmem_ld_32x32b_x16(tSj + 16u, v1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          s[i] = __uint_as_float(v0[i]);
          s[16 + i] = __uint_as_float(v1[i]);
        }
        if (TAIL) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c0 + i >= nk) s[i] = -INFINITY;
        }
        mx = s[0];
#pragma unroll
        for (int i = 1; i < 32; ++i) mx = fmaxf(mx, s[i]);
      }
      // exchange the raw block maximum with the thread that owns the other half of this row
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(my_max), "f"(mx) : "memory");
      named_bar_sync(pair_bar, 64);
      float mx_peer;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(mx_peer) : "r"(peer_max) : "memory");
      const float m_new = fmaxf(m_run, fmaxf(mx, mx_peer) * scale_log2);  // scale > 0: max commutes with it
      const float alpha = ex2_approx(m_run - m_new);                      // 0 on the first block
      m_run = m_new;
      if (j > 0) {  // the previous Oj is complete; P and the previous K/V stage are no longer being read
        ptx::mbar_wait(o_full, (uint32_t)((j - 1) & 1));
        ptx::tc_fence_after();
        fold(alpha_prev);
      }
      alpha_prev = alpha;
      float sum = 0.f;
      if (have_cols) {
        const float neg_m = -m_new;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          float p[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            p[i] = ex2_approx(fmaf(s[c8 * 8 + i], scale_log2, neg_m));
            sum += p[i];
          }
          const uint32_t p0 = pack_bf16(p[0], p[1]), p1 = pack_bf16(p[2], p[3]), p2 = pack_bf16(p[4], p[5]), p3 = pack_bf16(p[6], p[7]);
          const uint32_t addr = sP + (uint32_t)r * 128u + (uint32_t)(((ch * 4 + c8) ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
        }
      }
      l_run = fmaf(l_run, alpha, sum);
      named_bar_sync(pair_bar, 64);  // both halves have read the exchanged maxima before the next block overwrites them
    };

    for (int j = 0; j < n_blocks; ++j) {
      ptx::mbar_wait(s_full(j & 1), (uint32_t)((j >> 1) & 1));
      ptx::tc_fence_after();
      if (warp_has_rows) {
        if (j + 1 < n_blocks || group_len % BKV == 0)
          block(j, std::false_type{});
        else
          block(j, std::true_type{});
      } else if (j > 0) {
        ptx::mbar_wait(o_full, (uint32_t)((j - 1) & 1));  // keep this warp's view of the barrier phases in step
      }
      ptx::fence_proxy_async();  // P (generic-proxy stores) -> visible to the tensor core's async proxy
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
    }
    ptx::mbar_wait(o_full, (uint32_t)((n_blocks - 1) & 1));
    ptx::tc_fence_after();
    if (warp_has_rows) {
      fold(alpha_prev);
      // total row sum = this half's partial + the other half's (same max sequence, so the partials simply add)
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(my_max), "f"(l_run) : "memory");
      named_bar_sync(pair_bar, 64);
      float l_peer;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(l_peer) : "r"(peer_max) : "memory");
      const int q = q0 + r;
      if (q < group_len) {
        const float inv = 1.f / (l_run + l_peer);
        uint4* dst = reinterpret_cast<uint4*>(out + (size_t)(row0 + q) * ldo + h * HD + ch * 32);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 u;
          u.x = pack_bf16(o_acc[c * 8 + 0] * inv, o_acc[c * 8 + 1] * inv);
          u.y = pack_bf16(o_acc[c * 8 + 2] * inv, o_acc[c * 8 + 3] * inv);
          u.z = pack_bf16(o_acc[c * 8 + 4] * inv, o_acc[c * 8 + 5] * inv);
          u.w = pack_bf16(o_acc[c * 8 + 6] * inv, o_acc[c * 8 + 7] * inv);
          dst[c] = u;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

cudaError_t attention_groups_tc(const bf16* qkv, int ld_qkv, bf16* out, int ldo, int n_groups, int group_len, int heads,
                                float scale, cudaStream_t stream) {
  if (n_groups <= 0 || group_len <= 0) return cudaSuccess;
  if (ld_qkv % 8 != 0 || ldo % 8 != 0) return cudaErrorInvalidValue;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int rows = n_groups * group_len;
  CUtensorMap tq, tkv;
  if (!gemm_get_tensor_map(qkv, rows, 3 * heads * HD, ld_qkv, HD, BQ, &tq)) return cudaErrorInvalidValue;
  if (!gemm_get_tensor_map(qkv, rows, 3 * heads * HD, ld_qkv, HD, BKV, &tkv)) return cudaErrorInvalidValue;
  dim3 grid((group_len + BQ - 1) / BQ, heads, n_groups);
  attention_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tkv, out, ldo, group_len, heads, scale * 1.4426950408889634f);
  note_launch();
  return cudaGetLastError();
}
