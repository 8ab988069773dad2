// 2-CTA (cta_group::2) persistent bf16 GEMM for the large-M contractions of the path (ViT and decoder
// QKV / out-proj / MLP over tens of thousands of rows):  C = act(A * W^T + bias) + residual.
//
// A CTA pair (cluster of 2 on one TPC) computes a 256 x 256 tile with tcgen05.mma.cta_group::2 (M = 256):
//   * each CTA TMA-loads its own 128 rows of A and HALF of the 256 W rows (128) per K block, so the pair moves
//     64 KB per 256x256x64 MACs -- 2/3 of the L2->SM traffic of two independent 128x256 tiles, and each SM reads
//     half as much operand data from shared memory per MMA (ncu on the 1-CTA kernel: tensor pipe 48-71 %, L2-feed
//     and epilogue bound),
//   * both CTAs' TMA transactions complete on the LEADER's mbarrier; the leader's single MMA thread issues the
//     instruction for the pair and releases smem stages / publishes accumulators with multicast commits,
//   * the epilogue (8 warps per CTA) converts 32 x 64 blocks, stages them in 128B-swizzled shared memory and writes
//     them with TMA stores (full 128-byte lines, rows beyond M clipped by the tensor map); the residual block is
//     brought in by a TMA load into the same staging buffer while the accumulator is read from TMEM.
//     (The 1-CTA kernel's per-thread 16-byte row stores cost ~13.5k cycles per tile -- more than the K=768 mainloop.)
#include <cuda.h>

#include <string>

#include "common.cuh"
#include "kernels.h"

// shared with gemm_tcgen05.cu
bool gemm_get_tensor_map(const bf16* ptr, int rows, int cols, int ld, int box_cols, int box_rows, CUtensorMap* out);
int gemm_sm_count();

namespace {

constexpr int BM = 128;          // rows per CTA (256 per pair)
constexpr int BN = 256;          // columns per pair tile
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int STAGES = 6;  // (the LNOUT instantiation uses LN_STAGES)
constexpr int PLAIN_STAGES = STAGES;
constexpr int A_BYTES = BM * BK * 2;        // 16 KB
constexpr int B_BYTES = (BN / 2) * BK * 2;  // 16 KB: this CTA's half of W
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int EPI_COLS = 64;                       // columns per staged block (128-byte rows)
constexpr int STAGING_BYTES = 32 * EPI_COLS * 2;   // 4 KB per epilogue warp
constexpr int TMEM_COLS = 2 * BN;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NUM_EPI_WARPS * STAGING_BYTES + 1024 + 512;
constexpr int LN_STAGES = 5;  // LNOUT instantiation: one operand stage fewer pays for a second staging buffer per epilogue warp
constexpr int SMEM_BYTES_LN = LN_STAGES * STAGE_BYTES + 2 * NUM_EPI_WARPS * STAGING_BYTES + 1024 + 512;
static_assert(SMEM_BYTES_LN <= 227 * 1024, "shared memory");

struct Epi2 {
  const float* bias;
  int act;
  int has_res;
  const float* ln_stats;   // folded LayerNorm on the A rows: [M][ln_slots] partial (sum, sum of squares) per row, or nullptr
  const float* ln_colsum;  // [N]
  float ln_inv_k, ln_eps;
  float* stats_out;        // [M][2 * N / BN] partial (sum, sum of squares) of the output rows, one slot per 128-column half tile, or nullptr
  int ln_slots;            // partial-sum slots per row of ln_stats (= 2 * K / BN: the producer's column halves)
  int out_slots;           // = 2 * N / BN
  // LNOUT: LayerNorm of the OUTPUT rows written as a second tensor (tmap_ln) by the CTA pair that owns the whole row block
  const float* lnout_gamma;
  const float* lnout_beta;
  float lnout_eps;
};

__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// Warp-uniform issue (see common.cuh): all lanes call with identical operands, one elected lane executes.
// TMA load whose completion bytes are signalled on an mbarrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair_elect(uint32_t smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
      :
      : "r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair_elect(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior tcgen05 ops of the elected thread have completed) on the same barrier offset in both CTAs
__device__ __forceinline__ void umma_commit_pair_elect(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar), "h"(mask)
      : "memory");
}

#ifdef GEMM2_TRACE
// Debug build only (tools/gemm2_trace.cu): phase timestamps of CTA 0, [warp][tile iteration][phase]
__device__ long long g_gemm2_trace[10 * 16 * 8];
#define G2TRACE(it, ph)                                                                                   \
  do {                                                                                                    \
    if (lane == 0 && blockIdx.x == 0 && (it) < 16) g_gemm2_trace[(warp * 16 + (it)) * 8 + (ph)] = clock64(); \
  } while (0)
#else
#define G2TRACE(it, ph) do { } while (0)
#endif

// FOLD: the optional folded-LayerNorm epilogue (ln_stats / stats_out) is compiled into a separate instantiation.  As run-time
// flags the compiler if-converted it into predicated FFMA / FMUL / PRMT that every element of every GEMM still issued
// (ncu: 10.6 issued instructions per output element for a plain bias epilogue), and the 8 epilogue warps -- not the
// tensor pipe -- set the tile period of the K = 768 GEMMs.
// LNOUT: the pair walks whole ROW BLOCKS (the N / 256 tiles of 256 rows one after the other), so that its epilogue threads
// see complete output rows: they accumulate the rows' (sum, sum of squares) over the tiles, then re-read the just-written
// output blocks (L2 hits, TMA loads double-buffered per warp), normalise them and TMA-store them to a second tensor -- the
// LayerNorm that follows a residual GEMM costs one extra write from the GEMM instead of a read + write by its own kernel.
template <bool FOLD, bool LNOUT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
             const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
             const __grid_constant__ CUtensorMap tmap_ln, int M, int N, int K, Epi2 ep) {
  constexpr int STAGES = LNOUT ? LN_STAGES : PLAIN_STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = ptx::warp_uniform((ptx::smem_u32(smem_raw) + 1023u) & ~1023u);
  const uint32_t staging_base = smem_base + STAGES * STAGE_BYTES;  // 1024-aligned (32 KB multiples)
  const uint32_t staging2_base = staging_base + NUM_EPI_WARPS * STAGING_BYTES;  // LNOUT only
  const uint32_t bar_base = staging_base + (LNOUT ? 2 : 1) * NUM_EPI_WARPS * STAGING_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  auto res_bar = [&](int w) { return bar_base + 8u * (2 * STAGES + 4 + w); };
  auto ln_bar = [&](int w) { return bar_base + 8u * (2 * STAGES + 4 + NUM_EPI_WARPS + w); };  // LNOUT: second staging buffer's loads
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4 + 2 * NUM_EPI_WARPS);
  auto smem_a = [&](int s) { return smem_base + s * STAGE_BYTES; };
  auto smem_b = [&](int s) { return smem_base + s * STAGE_BYTES + A_BYTES; };

  const int warp = __shfl_sync(0xffffffffu, (int)threadIdx.x / 32, 0);
  const int lane = threadIdx.x & 31;
  // == %cluster_ctarank for __cluster_dims__(2, 1, 1); written this way so that ptxas sees a uniform value (a branch on
  // the special register read, even broadcast with shfl, made it wrap every UTCHMMA in a waterfall loop)
  const uint32_t rank = blockIdx.x & 1u;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  const int m_pairs = (M + 2 * BM - 1) / (2 * BM);
  const int n_tiles = N / BN;
  const int num_tiles = m_pairs * n_tiles;
  const int num_kb = (K + BK - 1) / BK;
  // i-th tile of this pair: round-robin over all tiles, or (LNOUT) the n_tiles tiles of row block cluster_id + k * num_clusters in turn
  auto tile_at = [&](int i, int& tile) -> bool {
    if (LNOUT) {
      const int rb = cluster_id + (i / n_tiles) * num_clusters;
      tile = rb * n_tiles + i % n_tiles;
      return rb < m_pairs;
    }
    tile = cluster_id + i * num_clusters;
    return tile < num_tiles;
  };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_a);
    ptx::prefetch_tensormap(&tmap_b);
    ptx::prefetch_tensormap(&tmap_out);
    if (ep.has_res) ptx::prefetch_tensormap(&tmap_res);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        ptx::mbar_init(full_bar(s), 1);
        ptx::mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        ptx::mbar_init(tfull_bar(a), 1);
        ptx::mbar_init(tempty_bar(a), 2 * NUM_EPI_WARPS);  // every epilogue warp of BOTH CTAs arrives on the leader's
      }
      for (int w = 0; w < NUM_EPI_WARPS; ++w) {
        ptx::mbar_init(res_bar(w), 1);
        if (LNOUT) ptx::mbar_init(ln_bar(w), 1);
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_pair(tmem_slot, TMEM_COLS);
    tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  cluster_sync_all();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = ptx::warp_uniform(tmem_base);

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp of each CTA, elected lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      int tile;
      for (int ti = 0; tile_at(ti, tile); ++ti) {
        const int m0 = (tile / n_tiles) * 2 * BM + (int)rank * BM;
        const int n0 = (tile % n_tiles) * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t leader_full = mapa_rank(full_bar(stage), 0);
          if (rank == 0) ptx::mbar_arrive_expect_tx_elect(full_bar(stage), 2 * STAGE_BYTES);  // both CTAs' bytes
          tma_load_2d_pair_elect(smem_a(stage), &tmap_a, leader_full, kb * BK, m0);
          tma_load_2d_pair_elect(smem_b(stage), &tmap_b, leader_full, kb * BK, n0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: warp 1 of the leader CTA.  The whole warp
    // runs the loop (warp-uniform control flow keeps descriptors in uniform registers: no per-instruction
    // ELECT / R2UR / BRA.U.ANY waterfall around every UTCHMMA); one elected lane executes the tcgen05 instructions.
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(2 * BM, BN);
      uint32_t it = 0;       // k blocks issued so far: stage = it % STAGES, phase = (it / STAGES) & 1
      uint32_t tile_it = 0;  // tiles issued so far: accumulator = tile_it & 1, phase = (tile_it >> 1) & 1
      int tile;
      for (int ti = 0; tile_at(ti, tile); ++ti, ++tile_it) {
        const uint32_t acc = tile_it & 1u;
        G2TRACE(tile_it, 0);
        ptx::mbar_wait(tempty_bar(acc), ((tile_it >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        G2TRACE(tile_it, 1);
        const uint32_t d_tmem = tmem_base + acc * (uint32_t)BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t stage = it % STAGES;
          ptx::mbar_wait(full_bar(stage), (it / STAGES) & 1u);
          ptx::tc_fence_after();
          const uint64_t da = ptx::umma_desc_sw128_kmajor(smem_a(stage));
          const uint64_t db = ptx::umma_desc_sw128_kmajor(smem_b(stage));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16_pair_elect(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          umma_commit_pair_elect(empty_bar(stage));
          if (kb == 0) G2TRACE(tile_it, 2);
          if (kb == num_kb - 1) umma_commit_pair_elect(tfull_bar(acc));
        }
        G2TRACE(tile_it, 3);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int c_begin = (ew >> 2) * (BN / 2), c_end = c_begin + BN / 2;
    const uint32_t stg = staging_base + ew * STAGING_BYTES;
    const uint32_t rbar = res_bar(ew);
    const uint32_t my_row_off = (uint32_t)lane * 128u;
    uint32_t res_phase = 0, ln_phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int e_it = -1;
    float ln_s1 = 0.f, ln_s2 = 0.f;  // LNOUT: this thread's row, its 128-column halves of the row block's tiles so far
    int tile;
    for (int ti = 0; tile_at(ti, tile); ++ti) {
      ++e_it;
      const int m0 = (tile / n_tiles) * 2 * BM + (int)rank * BM;
      const int n0 = (tile % n_tiles) * BN;
      const int row0 = m0 + quarter * 32;
      G2TRACE(e_it, 0);
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      G2TRACE(e_it, 1);
      const uint32_t t_row = tmem_base + (uint32_t)(acc * BN) + ((uint32_t)(quarter * 32) << 16);
      const int my_row = row0 + lane;
      float ln_mean = 0.f, ln_rstd = 1.f;
      if (FOLD && ep.ln_stats != nullptr && my_row < M) {
        // the row's statistics: partial sums of the producer's column halves, added in slot order (deterministic)
        float2 st = make_float2(0.f, 0.f);
        for (int sl = 0; sl < ep.ln_slots; ++sl) {
          const float2 pp = __ldg(reinterpret_cast<const float2*>(ep.ln_stats) + (size_t)my_row * ep.ln_slots + sl);
          st.x += pp.x;
          st.y += pp.y;
        }
        ln_mean = st.x * ep.ln_inv_k;
        ln_rstd = rsqrtf(fmaxf(st.y * ep.ln_inv_k - ln_mean * ln_mean, 0.f) + ep.ln_eps);
      }
      float so1 = 0.f, so2 = 0.f;
      if (row0 < M) {  // warp-uniform: this warp's 32 rows are not entirely beyond M
#pragma unroll 1
        for (int c = c_begin; c < c_end; c += EPI_COLS) {
          const int col = n0 + c;
          // this block's 64 bias values go to registers BEFORE the waits below: inside the chunk loop every load sat
          // behind the previous chunk's st.shared (asm memory clobber) and its latency was paid 8 times per block
          float4 bia[16];
          if (ep.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) bia[i] = __ldg(reinterpret_cast<const float4*>(ep.bias + col) + i);
          }
          if (lane == 0) ptx::tma_store_wait_read<0>();  // the previous block has left the staging buffer
          __syncwarp();
          G2TRACE(e_it, c == c_begin ? 2 : 5);
          if (ep.has_res) {
            ptx::mbar_arrive_expect_tx_elect(rbar, STAGING_BYTES);
            ptx::tma_load_2d_elect(stg, &tmap_res, rbar, col, row0);
          }
          uint32_t v0[32], v1[32];
          ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)c, v0);
          ptx::tmem_ld_32x32b_x32(t_row + (uint32_t)(c + 32), v1);
          ptx::tmem_ld_wait();
          G2TRACE(e_it, c == c_begin ? 3 : 6);
          if (ep.has_res) {
            ptx::mbar_wait(rbar, res_phase);
            res_phase ^= 1u;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {  // 8 chunks of 8 columns = one 128-byte staged row
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(j < 4 ? v0[j * 8 + i] : v1[(j - 4) * 8 + i]);
            if (FOLD && ep.ln_stats != nullptr) {  // rstd * (acc - mean * colsum[n])
              const float4 c0 = __ldg(reinterpret_cast<const float4*>(ep.ln_colsum + col + j * 8));
              const float4 c1 = __ldg(reinterpret_cast<const float4*>(ep.ln_colsum + col + j * 8 + 4));
              const float nm = -ln_mean;
              f[0] = fmaf(nm, c0.x, f[0]) * ln_rstd; f[1] = fmaf(nm, c0.y, f[1]) * ln_rstd;
              f[2] = fmaf(nm, c0.z, f[2]) * ln_rstd; f[3] = fmaf(nm, c0.w, f[3]) * ln_rstd;
              f[4] = fmaf(nm, c1.x, f[4]) * ln_rstd; f[5] = fmaf(nm, c1.y, f[5]) * ln_rstd;
              f[6] = fmaf(nm, c1.z, f[6]) * ln_rstd; f[7] = fmaf(nm, c1.w, f[7]) * ln_rstd;
            }
            if (ep.bias != nullptr) {
              const float4 b0 = bia[2 * j], b1 = bia[2 * j + 1];
              f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
              f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
            }
            if (ep.act == ACT_QUICK_GELU) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = quick_gelu(f[i]);
            } else if (ep.act == ACT_GELU_ERF) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = gelu_erf(f[i]);
            } else if (ep.act == ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = fmaxf(f[i], 0.f);
            }
            const uint32_t saddr = stg + my_row_off + (uint32_t)((j ^ (lane & 7)) << 4);
            if (ep.has_res) {
              uint4 u;
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(saddr));
              const float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y), a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
              f[0] += a0.x; f[1] += a0.y; f[2] += a1.x; f[3] += a1.y;
              f[4] += a2.x; f[5] += a2.y; f[6] += a3.x; f[7] += a3.y;
            }
            const uint32_t p0 = pack_bf16(f[0], f[1]), p1 = pack_bf16(f[2], f[3]), p2 = pack_bf16(f[4], f[5]), p3 = pack_bf16(f[6], f[7]);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
            if (LNOUT) {  // statistics of the row as stored (bf16-rounded), accumulated over the row block's tiles
              const float2 r0 = unpack_bf16(p0), r1 = unpack_bf16(p1), r2 = unpack_bf16(p2), r3 = unpack_bf16(p3);
              ln_s1 += (r0.x + r0.y) + (r1.x + r1.y) + (r2.x + r2.y) + (r3.x + r3.y);
              ln_s2 = fmaf(r0.x, r0.x, ln_s2); ln_s2 = fmaf(r0.y, r0.y, ln_s2); ln_s2 = fmaf(r1.x, r1.x, ln_s2); ln_s2 = fmaf(r1.y, r1.y, ln_s2);
              ln_s2 = fmaf(r2.x, r2.x, ln_s2); ln_s2 = fmaf(r2.y, r2.y, ln_s2); ln_s2 = fmaf(r3.x, r3.x, ln_s2); ln_s2 = fmaf(r3.y, r3.y, ln_s2);
            }
            if (FOLD && ep.stats_out != nullptr) {  // statistics of the values as stored (bf16-rounded)
              const float2 r0 = unpack_bf16(p0), r1 = unpack_bf16(p1), r2 = unpack_bf16(p2), r3 = unpack_bf16(p3);
              so1 += (r0.x + r0.y) + (r1.x + r1.y) + (r2.x + r2.y) + (r3.x + r3.y);
              so2 = fmaf(r0.x, r0.x, so2); so2 = fmaf(r0.y, r0.y, so2); so2 = fmaf(r1.x, r1.x, so2); so2 = fmaf(r1.y, r1.y, so2);
              so2 = fmaf(r2.x, r2.x, so2); so2 = fmaf(r2.y, r2.y, so2); so2 = fmaf(r3.x, r3.x, so2); so2 = fmaf(r3.y, r3.y, so2);
            }
          }
          ptx::fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmap_out, stg, col, row0);
            ptx::tma_store_commit();
          }
          G2TRACE(e_it, c == c_begin ? 4 : 7);
        }
      }
      if (FOLD && ep.stats_out != nullptr && my_row < M)  // this warp's 128 columns of the row: its own slot, a plain store
        reinterpret_cast<float2*>(ep.stats_out)[(size_t)my_row * ep.out_slots + (n0 / BN) * 2 + (ew >> 2)] = make_float2(so1, so2);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_rank(tempty_bar(acc), 0));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
      if (LNOUT && (tile % n_tiles) == n_tiles - 1) {
        // ---- the row block is complete: LayerNorm of its rows into the second output tensor
        const uint32_t stg2 = staging2_base + ew * STAGING_BYTES;
        const uint32_t lbar = ln_bar(ew);
        // (1) the other column half of every row lives in warp ew ^ 4: exchange (sum, sum of squares) through the second staging
        //     buffer (free here), halves added in a fixed order
        if (lane == 0) ptx::tma_store_wait<0>();  // this warp's output blocks are in L2 / HBM (and nothing reads the staging buffers)
        __syncwarp();
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(stg2 + (uint32_t)lane * 8u), "f"(ln_s1), "f"(ln_s2) : "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");
        float o1, o2;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(o1), "=f"(o2) : "r"(staging2_base + (ew ^ 4) * STAGING_BYTES + (uint32_t)lane * 8u) : "memory");
        const float t1 = (ew < 4) ? ln_s1 + o1 : o1 + ln_s1, t2 = (ew < 4) ? ln_s2 + o2 : o2 + ln_s2;
        const float inv_n = 1.0f / (float)N;
        const float mean = t1 * inv_n;
        const float rstd = rsqrtf(fmaxf(t2 * inv_n - mean * mean, 0.f) + ep.lnout_eps);
        asm volatile("bar.sync 1, %0;" ::"n"(32 * NUM_EPI_WARPS) : "memory");  // everybody has read its partner's pair: the buffer is free
        ln_s1 = 0.f;
        ln_s2 = 0.f;
        // (2) this warp's 2 * n_tiles blocks (32 rows x 64 columns) come back by TMA, two in flight
        const int n_blk = 2 * n_tiles;
        auto blk_col = [&](int q) { return (q >> 1) * BN + c_begin + (q & 1) * EPI_COLS; };
        uint32_t ph_a = res_phase, ph_b = 0;  // parity of the next completion of res_bar / ln_bar (ln_bar: tracked across row blocks below)
        ph_b = ln_phase;
        if (row0 < M) {
          ptx::mbar_arrive_expect_tx_elect(rbar, STAGING_BYTES);
          ptx::tma_load_2d_elect(stg, &tmap_out, rbar, blk_col(0), row0);
#pragma unroll 1
          for (int q = 0; q < n_blk; ++q) {
            const uint32_t buf = (q & 1) ? stg2 : stg;
            if (q + 1 < n_blk) {  // next block into the other buffer, once the store that last read it has let go of it
              if (lane == 0) ptx::tma_store_wait_read<0>();
              __syncwarp();
              const uint32_t nb = (q & 1) ? stg : stg2, nbar = (q & 1) ? rbar : lbar;
              ptx::mbar_arrive_expect_tx_elect(nbar, STAGING_BYTES);
              ptx::tma_load_2d_elect(nb, &tmap_out, nbar, blk_col(q + 1), row0);
            }
            if (q & 1) {
              ptx::mbar_wait(lbar, ph_b);
              ph_b ^= 1u;
            } else {
              ptx::mbar_wait(rbar, ph_a);
              ph_a ^= 1u;
            }
            const int col = blk_col(q);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t saddr = buf + my_row_off + (uint32_t)((j ^ (lane & 7)) << 4);
              uint4 u;
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(saddr));
              const float4 g0 = __ldg(reinterpret_cast<const float4*>(ep.lnout_gamma + col + j * 8));
              const float4 g1 = __ldg(reinterpret_cast<const float4*>(ep.lnout_gamma + col + j * 8 + 4));
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(ep.lnout_beta + col + j * 8));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(ep.lnout_beta + col + j * 8 + 4));
              const float2 a0 = unpack_bf16(u.x), a1 = unpack_bf16(u.y), a2 = unpack_bf16(u.z), a3 = unpack_bf16(u.w);
              const uint32_t p0 = pack_bf16(fmaf((a0.x - mean) * rstd, g0.x, b0.x), fmaf((a0.y - mean) * rstd, g0.y, b0.y));
              const uint32_t p1 = pack_bf16(fmaf((a1.x - mean) * rstd, g0.z, b0.z), fmaf((a1.y - mean) * rstd, g0.w, b0.w));
              const uint32_t p2 = pack_bf16(fmaf((a2.x - mean) * rstd, g1.x, b1.x), fmaf((a2.y - mean) * rstd, g1.y, b1.y));
              const uint32_t p3 = pack_bf16(fmaf((a3.x - mean) * rstd, g1.z, b1.z), fmaf((a3.y - mean) * rstd, g1.w, b1.w));
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
            }
            ptx::fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_2d(&tmap_ln, buf, col, row0);
              ptx::tma_store_commit();
            }
          }
        }
        res_phase = ph_a;
        ln_phase = ph_b;
      }
    }
    if (lane == 0) ptx::tma_store_wait<0>();  // all output bytes written before the CTA may exit
  }

  ptx::tc_fence_before();
  cluster_sync_all();  // the peer may still read this CTA's smem / signal its barriers until here
  if (warp == 1) {
    ptx::tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

}  // namespace

// Returns cudaErrorNotSupported when the shape does not suit the pair kernel (caller falls back to the 1-CTA kernel).
cudaError_t gemm2_bf16(const GemmArgs& a, cudaStream_t stream) {
  if (a.N % BN != 0 || a.out == nullptr || a.out_f32 != nullptr || a.gin > 0 || a.res_periodic) return cudaErrorNotSupported;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES_LN);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  CUtensorMap ta, tb, to, tr;
  if (!gemm_get_tensor_map(a.A, a.M, a.K, a.lda, BK, BM, &ta)) return cudaErrorInvalidValue;
  if (!gemm_get_tensor_map(a.W, a.N, a.K, a.ldw, BK, BN / 2, &tb)) return cudaErrorInvalidValue;
  if (!gemm_get_tensor_map(a.out, a.M, a.N, a.ldo, EPI_COLS, 32, &to)) return cudaErrorInvalidValue;
  tr = to;
  if (a.residual != nullptr && !gemm_get_tensor_map(a.residual, a.M, a.N, a.ldr, EPI_COLS, 32, &tr)) return cudaErrorInvalidValue;
  const int tiles = ((a.M + 2 * BM - 1) / (2 * BM)) * (a.N / BN);
  int clusters = gemm_sm_count() / 2;
  if (tiles < clusters) clusters = tiles;
  Epi2 ep{a.bias, a.act, a.residual != nullptr ? 1 : 0, a.ln_stats, a.ln_colsum, 1.0f / (float)a.K, a.ln_eps, a.stats_out, 2 * a.K / BN, 2 * a.N / BN,
          a.lnout_gamma, a.lnout_beta, a.lnout_eps};
  CUtensorMap tl = to;
  if (a.lnout != nullptr) {
    // LayerNorm of the output rows as a second output: the pair must own whole rows (row-block tile order), N <= 1024
    if (a.ln_stats != nullptr || a.stats_out != nullptr || a.N > 1024 || a.lnout_gamma == nullptr || a.lnout_beta == nullptr || a.lnout_ld % 8 != 0)
      return cudaErrorInvalidValue;
    if (!gemm_get_tensor_map(a.lnout, a.M, a.N, a.lnout_ld, EPI_COLS, 32, &tl)) return cudaErrorInvalidValue;
    const int row_blocks = (a.M + 2 * BM - 1) / (2 * BM);
    if (row_blocks < clusters) clusters = row_blocks;
    gemm2_kernel<false, true><<<2 * clusters, NUM_THREADS, SMEM_BYTES_LN, stream>>>(ta, tb, to, tr, tl, a.M, a.N, a.K, ep);
  } else if (a.ln_stats != nullptr || a.stats_out != nullptr)
    gemm2_kernel<true, false><<<2 * clusters, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, to, tr, tl, a.M, a.N, a.K, ep);
  else
    gemm2_kernel<false, false><<<2 * clusters, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, to, tr, tl, a.M, a.N, a.K, ep);
  note_launch();
  return cudaGetLastError();
}
