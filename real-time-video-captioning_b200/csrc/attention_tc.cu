// tcgen05 flash attention over fixed-length row groups (ViT frames, K10: the decoder's visual block).
//
// Work item = one (group, head, 128-query tile).  A CTA (two per SM) is persistent and keeps TWO items in flight: softmax
// warpgroup g (4 warps, one query row per thread) owns item 2 * pair + g from its first key block to its output rows -- its
// own Q tile, running max m_g / row sum l_g, output accumulator O[g] and probability buffer P[g], both in tensor memory.
// The two groups never merge, exchange or wait for each other.  Per 64-key block j of a group's item:
//     S_j = Q K_j^T           tcgen05.mma  M=128, N<=64, K=64 (Q, K: K-major 128B-swizzled TMA tiles)  -> TMEM S
//     P_j = exp2(S_j*c - m_g) one query row per thread (tcgen05.ld 32x32b.x32), packed f32x2 FMA / add, P written as bf16
//                             pairs straight back into TENSOR MEMORY (tcgen05.st, 32 columns per group)
//     O_g += P_j V_j          tcgen05.mma with the A operand FROM TMEM (the "TS" form: lane = query row, one 32-bit column =
//                             two consecutive keys), M=128, N=64, K<=64 (V: the TMA tile used as an MN-major B operand);
//                             the accumulation stays in TMEM
// TMEM map of a CTA (256 columns): S [0,64) -- ONE score buffer handed to the two groups alternately; a group keeps it only
// for the ~100 cycles of its tcgen05.ld --, O[0] [64,128), O[1] [128,192), P[0] [192,224), P[1] [224,256).
// The running max is LAZY: m_g only moves when a block's maximum exceeds it by more than 2^8; then (rarely, in practice
// during the first blocks only) the warp rescales O_g in TMEM (tcgen05.ld / tcgen05.st) before it releases P_j.  No
// per-block read-back of O, no cross-warp max exchange, no named barrier anywhere.
// 12 warps: two TMA producers (Q + K, and V: tiles of the two items alternate in the rings), one MMA issuer per group (S_j: operands first,
// then the hand-over of the score buffer, then 4 UTCHMMA; after it the product of the group's previous block), 2 x 4
// softmax warps (setmaxnreg 32 / 104).  Output rows leave through a small swizzled staging slice per warp so that every
// global store instruction writes whole 64-byte runs.  The last key block of a group issues only the 16-key steps that hold valid keys
// (197 = 3*64 + 5 -> N = 16); for >= 512 keys the exponentials run before the wait on the group's previous product and one
// pair of every eight goes through a polynomial on the FMA pipe.
//
// History (profiles/): mma.sync kernel 123-222 TFLOP/s; first tcgen05 version (two threads per row, per-block O read-back,
// shared-memory max exchange) 176 / 405 (ViT / decoder shape); round 1 final (two groups on alternate blocks of ONE item,
// P staged in shared memory, partial softmaxes merged per item) 268 / 657; round 2: P in TMEM + TS-form MMA 290 / 670, then
// one item per group (this file): see profiles/r02_attention_summary.md.
#include <cuda.h>
#include <math.h>

#include <type_traits>

#include "common.cuh"
#include "kernels.h"

bool gemm_get_tensor_map(const bf16* ptr, int rows, int cols, int ld, int box_cols, int box_rows, CUtensorMap* out);
int gemm_sm_count();

namespace {

constexpr int HD = 64;
constexpr int BQ = 128;   // query rows per CTA = TMEM lanes
constexpr int BKV = 64;   // keys per block
constexpr int K_STAGES = 3, V_STAGES = 4;  // tiles of the two items in flight alternate in the rings
constexpr int STAGE_BYTES = 8 * 32 * 64;    // output staging: per softmax warp 32 rows x 32 dims bf16 (coalesced epilogue stores)
constexpr int Q_BYTES = BQ * 128;         // 16 KB
constexpr int KV_TILE_BYTES = BKV * 128;  // 8 KB
constexpr int N_BARRIERS = 4 + 2 * K_STAGES + 2 * V_STAGES + 8;
constexpr int SMEM_BYTES = 1024 + 2 * Q_BYTES + (K_STAGES + V_STAGES) * KV_TILE_BYTES + STAGE_BYTES + 256 /*barriers + tmem slot*/;
static_assert(8 * N_BARRIERS + 4 <= 256, "barrier block");
static_assert(2 * (SMEM_BYTES + 1024) <= 228 * 1024, "two CTAs per SM");
constexpr int NUM_SM_WARPS = 8;   // warps 4-7: softmax group 0, warps 8-11: group 1; TMEM lane quarter = warp & 3
constexpr int NUM_THREADS = 128 + 32 * NUM_SM_WARPS;  // warp 0: TMA, warps 1 / 2: MMA issuers of group 0 / 1 (warp 1 owns TMEM), warp 3: TMA producer of V
constexpr int SERVICE_REGS = 32, SOFTMAX_REGS = 104;  // setmaxnreg split: the increase is served from the CTA's OWN pool (what its service
                                                      // warps released; more than that deadlocks), so 128 x 32 + 256 x 104 = 30720 = the launch allocation 384 x 80
constexpr int TMEM_COLS = 256;    // S: [0,64), O[0]: [64,128), O[1]: [128,192), P[0]: [192,224), P[1]: [224,256)
constexpr uint32_t TM_S = 0, TM_O = 64, TM_P = 192;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units: P never exceeds 2^8 before the running max moves

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
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[TMEM] * B[smem desc]: the A operand (bf16, K-major: lane = row, 32-bit column = two consecutive K elements)
// is read from tensor memory.  Warp-uniform issue like ptx::umma_bf16_elect.
__device__ __forceinline__ void umma_bf16_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for a pair of scores on the FMA pipe (no MUFU): x = n + f with n = round(x), f in [-0.5, 0.5];
// 2^f by a degree-3 minimax polynomial (max relative error 7.5e-5, far below the bf16 rounding of P), 2^n by adding n
// to the exponent field.  x <= RESCALE_THRESHOLD by construction; x is clamped at -126 (result ~ 0).
// The softmax is bound by the MUFU (XU) rate -- 16 exp2 per clock per SM against 128 FMA lanes -- so moving a
// fraction of the exponentials here shortens the XU critical path (the FlashAttention-4 trick).
// Measured on B200 (tools/attn_bench.cu, 256 clips): 1 of every 4 pairs on the polynomial = +6.7 % on the 1182-key
// decoder shape (578 -> 617 TFLOP/s), -2 % on the 197-key ViT shape (4 key blocks per CTA: start-up bound, the extra
// FMA work only adds latency); 2 pairs = no gain, 3 pairs = -7 %.  Hence: 1 pair for long groups, 0 for short ones.
constexpr int LONG_GROUP = 512;
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  const float magic = 12582912.f;  // 1.5 * 2^23: x + magic has round(x) in its low mantissa bits
  x.x = fmaxf(x.x, -126.f);
  x.y = fmaxf(x.y, -126.f);
  const float2 t = __fadd2_rn(x, make_float2(magic, magic));
  const float2 n = __fadd2_rn(t, make_float2(-magic, -magic));
  const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
  float2 p = __ffma2_rn(make_float2(0.0551716685f, 0.0551716685f), f, make_float2(0.2426111251f, 0.2426111251f));
  p = __ffma2_rn(p, f, make_float2(0.6932609677f, 0.6932609677f));
  p = __ffma2_rn(p, f, make_float2(0.9999280572f, 0.9999280572f));
  float2 r;
  r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return r;
}
#ifndef ATTN_TRACE_ITEM
#define ATTN_TRACE_ITEM 0
#endif
#ifdef ATTN_TRACE
// Debug build only (tools/attn_bench.cu): per-warp phase timestamps of one CTA, [warp][block][phase]
__device__ long long g_attn_trace[12 * 32 * 12];
#define TRACE(blk, ph)                                                                                              \
  do {                                                                                                              \
    if (lane == 0 && blockIdx.x == ATTN_TRACE && (blk) < 32)                  \
      g_attn_trace[(warp * 32 + (blk)) * 12 + (ph)] = clock64();                                                     \
  } while (0)
#else
#define TRACE(blk, ph) do { } while (0)
#endif

// SHARE_KV: the two items of a pair that are query tiles of the same (group, head) read ONE copy of every K / V tile (half the
// TMA traffic, twice the ring depth in key blocks).  Measured (tools/attn_bench.cu, same box): +4 % at 1182 keys, +5 % at 6168
// keys, but -8 % / -12 % on the 197- / 257-key ViT shapes (a stage is then released by the slower of the two groups) ->
// on for long groups only, together with EXP_FIRST.
template <int ATTN_POLY_PAIRS, bool EXP_FIRST>
__global__ void __launch_bounds__(NUM_THREADS, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    bf16* __restrict__ out, int ldo, int group_len, int heads, int n_groups, float scale_log2) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = ptx::warp_uniform((ptx::smem_u32(smem_raw) + 1023u) & ~1023u);
  auto sQ = [&](int g) { return smem_base + g * Q_BYTES; };
  auto sK = [&](int s) { return smem_base + 2 * Q_BYTES + s * KV_TILE_BYTES; };
  auto sV = [&](int s) { return smem_base + 2 * Q_BYTES + (K_STAGES + s) * KV_TILE_BYTES; };
  const uint32_t sStage = smem_base + 2 * Q_BYTES + (K_STAGES + V_STAGES) * KV_TILE_BYTES;
  const uint32_t bar_base = sStage + STAGE_BYTES;
  auto q_full = [&](int g) { return bar_base + 8u * g; };
  auto q_empty = [&](int g) { return bar_base + 8u * (2 + g); };
  auto k_full = [&](int s) { return bar_base + 8u * (4 + s); };
  auto k_empty = [&](int s) { return bar_base + 8u * (4 + K_STAGES + s); };
  auto v_full = [&](int s) { return bar_base + 8u * (4 + 2 * K_STAGES + s); };
  auto v_empty = [&](int s) { return bar_base + 8u * (4 + 2 * K_STAGES + V_STAGES + s); };
  const uint32_t bar_g = bar_base + 8u * (4 + 2 * K_STAGES + 2 * V_STAGES);
  auto s_full = [&](int g) { return bar_g + 8u * g; };        // MMA -> softmax group g: S holds a new block of this group's item
  auto s_free = [&](int g) { return bar_g + 8u * (2 + g); };  // group g -> MMA: S has been read into registers
  auto p_full = [&](int g) { return bar_g + 8u * (4 + g); };  // group g -> its P V issuer: P[g] written to TMEM (and O[g] rescaled)
  auto o_full = [&](int g) { return bar_g + 8u * (6 + g); };  // MMA -> group g: O[g] += P V finished, P[g] is free
  const uint32_t tmem_slot = bar_base + 8u * N_BARRIERS;

  const int warp = __shfl_sync(0xffffffffu, (int)threadIdx.x / 32, 0);
  const int lane = threadIdx.x & 31;
  const int width = heads * HD;
  const int n_blocks = (group_len + BKV - 1) / BKV;
  // PERSISTENT, TWO ITEMS IN FLIGHT: the CTA walks PAIRS of work items (item = query tile x head x group) with stride
  // gridDim.x; softmax group g owns item 2 * pair + g from its first key block to its output rows -- its own Q tile, running
  // max / sum, O accumulator and P buffer -- so the two groups never merge, exchange or wait for each other (round 1 dealt
  // the key blocks of ONE item alternately to the two groups and merged two partial softmaxes per item: two named barriers,
  // a shared-memory exchange and cross reads of both accumulators per item, ~1000 cycles of a 4-block ViT item's ~6700).
  // An item's result depends on nothing but its own data: bit-identical wherever it sits in the batch.
  const int n_qt = (group_len + BQ - 1) / BQ;
  const int n_items = n_qt * heads * n_groups;
  const int n_pairs = (n_items + 1) / 2;
  constexpr bool SHARE_KV = EXP_FIRST;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_q);
    ptx::prefetch_tensormap(&tmap_kv);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int g = 0; g < 2; ++g) {
        ptx::mbar_init(q_full(g), 1);
        ptx::mbar_init(q_empty(g), 1);
      }
      for (int s = 0; s < K_STAGES; ++s) {
        ptx::mbar_init(k_full(s), 1);
        ptx::mbar_init(k_empty(s), 2);  // both groups' products when the pair shares its K / V tiles, one group twice otherwise
      }
      for (int s = 0; s < V_STAGES; ++s) {
        ptx::mbar_init(v_full(s), 1);
        ptx::mbar_init(v_empty(s), 2);
      }
      for (int g = 0; g < 2; ++g) {
        ptx::mbar_init(s_full(g), 1);
        ptx::mbar_init(s_free(g), NUM_SM_WARPS / 2);  // one arrival per warp of the group
        ptx::mbar_init(p_full(g), NUM_SM_WARPS / 2);
        ptx::mbar_init(o_full(g), 1);
      }
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
  tmem_base = ptx::warp_uniform(tmem_base);
  // register split (inside each role's branch, so that ptxas budgets the role's code accordingly): the four service warps
  // (TMA, S issue, two P V issuers) hand registers to the eight softmax warps
  const uint32_t tS = tmem_base + TM_S, tO = tmem_base + TM_O, tP = tmem_base + TM_P;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer of Q and K (whole warp, elected lane issues).
    // V tiles have their own producer (warp 3): a V stage is only released by the P V product, ~2000 cycles after its load; in one
    // in-order producer a full V ring also held back the K tiles the next score product was waiting for.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SERVICE_REGS));
    uint32_t kc = 0;  // K tiles loaded so far (all pairs): ring stage / phase follow from it
    int it = 0;
    for (int p = blockIdx.x; p < n_pairs; p += gridDim.x, ++it) {
      const int n_it = (2 * p + 1 < n_items) ? 2 : 1;
      // the two items of a pair are usually two query tiles of the SAME (group, head): then they share every K / V tile
      const int n_kv = (n_it == 2 && (!SHARE_KV || (2 * p) / n_qt != (2 * p + 1) / n_qt)) ? 2 : 1;
      int h2[2], row2[2];
      for (int g = 0; g < n_it; ++g) {
        const int w = 2 * p + g;
        const int qt = w % n_qt;
        h2[g] = (w / n_qt) % heads;
        row2[g] = (w / (n_qt * heads)) * group_len;
        ptx::mbar_wait(q_empty(g), (uint32_t)((it & 1) ^ 1));  // every S product of this group's previous item has read Q[g]
        ptx::mbar_arrive_expect_tx_elect(q_full(g), Q_BYTES);
        ptx::tma_load_2d_elect(sQ(g), &tmap_q, q_full(g), h2[g] * HD, row2[g] + qt * BQ);
      }
      for (int j = 0; j < n_blocks; ++j) {
        for (int g = 0; g < n_kv; ++g, ++kc) {
          const int ks = kc % K_STAGES;
          ptx::mbar_wait(k_empty(ks), (uint32_t)(((kc / K_STAGES) & 1) ^ 1));
          ptx::mbar_arrive_expect_tx_elect(k_full(ks), KV_TILE_BYTES);
          ptx::tma_load_2d_elect(sK(ks), &tmap_kv, k_full(ks), width + h2[g] * HD, row2[g] + j * BKV);
        }
      }
    }
  } else if (warp < 3) {
    // ------------------------------------------------------------ MMA issuer of softmax group g = warp - 1 (whole warp,
    // elected lane issues): for its group's item, S_j = Q K_j^T into the shared score buffer and O[g] (+)= P_{j-1} V_{j-1},
    // in that order.  One issuer per group: a single issuer for both groups spent ~850 cycles per block in its waits
    // (each mbarrier wait costs ~100 cycles even when the phase has long completed), commits and issues -- with two groups
    // to feed that was the pace of the whole kernel (traces in profiles/r02_attention_summary.md).
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SERVICE_REGS));
    const int g = warp - 1;
    const uint32_t tPg = tP + (uint32_t)(g * 32);
    const uint64_t dq = ptx::umma_desc_sw128_kmajor(sQ(g));
    uint32_t mine = 0;  // blocks of this group issued so far (all items): S hand-over / p_full phases follow from it
    uint32_t pv_done = 0;
    uint32_t kc0 = 0;     // ring position of the current pair's first K / V tile
    bool single = false;  // the current pair holds one item only (the grid's last pair)
    int it = 0;
    auto issue_pv = [&](int j, uint32_t kc0, int n_kv) {
      const uint32_t kcv = kc0 + (uint32_t)(n_kv == 2 ? 2 * j + g : j);
      const int vs = (int)(kcv % V_STAGES);
      const int nk = min(BKV, group_len - j * BKV);
      const int nk16 = (nk + 15) & ~15;
      if (it == ATTN_TRACE_ITEM) TRACE(j, 5);
      ptx::mbar_wait(v_full(vs), (kcv / V_STAGES) & 1u);  // landed long ago as a rule: checked while the group still computes P
      ptx::mbar_wait(p_full(g), pv_done & 1u);
      ptx::tc_fence_after();
      if (it == ATTN_TRACE_ITEM) TRACE(j, 6);
      const uint64_t dv = umma_desc_sw128_mnmajor(sV(vs));
      const uint32_t acc0 = j == 0 ? 0u : 1u;               // the item's first block starts a new O
#pragma unroll
      for (int k = 0; k < BKV / 16; ++k)  // 16 keys per step: P advances 8 TMEM columns (bf16 pairs), V advances 2 atoms
        if (k * 16 < nk16)
          umma_bf16_ts_elect(tO + (uint32_t)(g * 64), tPg + (uint32_t)(8 * k), dv + (uint64_t)(128 * k), idesc_pv(), k != 0 ? 1u : acc0);
      ptx::umma_commit_elect(o_full(g));
      ptx::umma_commit_elect(v_empty(vs));
      if (n_kv == 2 || single) ptx::umma_commit_elect(v_empty(vs));  // sole reader of this V tile: both arrivals
      if (it == ATTN_TRACE_ITEM) TRACE(j, 7);
      ++pv_done;
    };
    for (int p = blockIdx.x; p < n_pairs; p += gridDim.x, ++it) {
      const int n_it = (2 * p + 1 < n_items) ? 2 : 1;
      const int n_kv = (n_it == 2 && (!SHARE_KV || (2 * p) / n_qt != (2 * p + 1) / n_qt)) ? 2 : 1;  // tiles per key block: shared by the pair or one per item
      single = n_it == 1;
      if (g >= n_it) continue;  // (the last pair: nothing follows, kc0 needs no update)
      for (int j = 0; j < n_blocks; ++j, ++mine) {
        const uint32_t kc = kc0 + (uint32_t)(n_kv == 2 ? 2 * j + g : j);
        const int ks = (int)(kc % K_STAGES);
        const int nk = min(BKV, group_len - j * BKV);
        const int nk16 = (nk + 15) & ~15;
        if (it == ATTN_TRACE_ITEM) TRACE(j, 0);
        // operands first (they have usually landed long ago), THEN the hand-over of the score buffer
        if (j == 0) ptx::mbar_wait(q_full(g), (uint32_t)(it & 1));
        ptx::mbar_wait(k_full(ks), (kc / K_STAGES) & 1u);
        if (it == ATTN_TRACE_ITEM) TRACE(j, 1);
        // the block before this one in the CTA's sequence (the other group's, or -- in a half pair -- this group's own
        // previous block) must have left S for its group's registers
        if (n_it == 2 || j == 0) {
          if (g == 1) ptx::mbar_wait(s_free(0), mine & 1u);                       // group 0's block with the same index
          else if (mine > 0) ptx::mbar_wait(s_free(1), (mine - 1) & 1u);          // group 1's previous block
        } else {
          ptx::mbar_wait(s_free(0), (mine - 1) & 1u);
        }
        ptx::tc_fence_after();
        const uint64_t dk = ptx::umma_desc_sw128_kmajor(sK(ks));
        const uint32_t id_s = idesc_qk(nk16);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          ptx::umma_bf16_elect(tS, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), id_s, k != 0);
        if (it == ATTN_TRACE_ITEM) TRACE(j, 2);
        ptx::umma_commit_elect(s_full(g));
        ptx::umma_commit_elect(k_empty(ks));  // K is dead once the product(s) reading it have been computed
        if (n_kv == 2 || single) ptx::umma_commit_elect(k_empty(ks));  // sole reader of this K tile: both arrivals
        if (j + 1 == n_blocks) ptx::umma_commit_elect(q_empty(g));  // ... and so is Q[g] after the item's last block
        if (it == ATTN_TRACE_ITEM) TRACE(j, 3);
        if (j > 0) issue_pv(j - 1, kc0, n_kv);
      }
      issue_pv(n_blocks - 1, kc0, n_kv);
      kc0 += (uint32_t)(n_kv * n_blocks);
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------ TMA producer of V
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SERVICE_REGS));
    uint32_t kc = 0;
    for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
      const int n_it = (2 * p + 1 < n_items) ? 2 : 1;
      const int n_kv = (n_it == 2 && (!SHARE_KV || (2 * p) / n_qt != (2 * p + 1) / n_qt)) ? 2 : 1;
      int h2[2], row2[2];
      for (int g = 0; g < n_it; ++g) {
        const int w = 2 * p + g;
        h2[g] = (w / n_qt) % heads;
        row2[g] = (w / (n_qt * heads)) * group_len;
      }
      for (int j = 0; j < n_blocks; ++j) {
        for (int g = 0; g < n_kv; ++g, ++kc) {
          const int vs = kc % V_STAGES;
          ptx::mbar_wait(v_empty(vs), (uint32_t)(((kc / V_STAGES) & 1) ^ 1));
          ptx::mbar_arrive_expect_tx_elect(v_full(vs), KV_TILE_BYTES);
          ptx::tma_load_2d_elect(sV(vs), &tmap_kv, v_full(vs), 2 * width + h2[g] * HD, row2[g] + j * BKV);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warpgroups: one thread per query row
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SOFTMAX_REGS));
    const int quarter = warp & 3;
    const int grp = (warp - 4) >> 2;    // this group's item of every pair: 2 * pair + grp
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t tSg = tS + lane_off;
    const uint32_t tOg = tO + (uint32_t)(grp * 64) + lane_off;
    const uint32_t tPg = tP + (uint32_t)(grp * 32) + lane_off;
    uint32_t done = 0;  // key blocks this group has processed so far (all items): phase of s_full / o_full
    int it = 0;
    float m_run = -INFINITY, l_run = 0.f;

    // One 64-key block.  TAIL = the last block of a group (masked keys); every other block takes the straight path.
    // first: the group's first block of this work item (nothing to wait for: O[grp] / P[grp] were released by the merge)
    auto block = [&](int j, bool first, auto tail_tag) {
      constexpr bool TAIL = decltype(tail_tag)::value;
      const int nk = TAIL ? group_len - j * BKV : BKV;
      float s[64];
      if (it == ATTN_TRACE_ITEM) TRACE(j, 0);
      {
        uint32_t v0[32], v1[32];
        ptx::tmem_ld_32x32b_x32(tSg, v0);
        ptx::tmem_ld_32x32b_x32(tSg + 32u, v1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s[i] = __uint_as_float(v0[i]);
          s[32 + i] = __uint_as_float(v1[i]);
        }
      }
      // S[grp] is in registers: the S issuer may overwrite it with this group's next block
      ptx::tc_fence_before();
      ptx::mbar_arrive_elect(s_free(grp));
      if (it == ATTN_TRACE_ITEM) TRACE(j, 1);
      if (TAIL) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= nk) s[i] = -INFINITY;
      }
      // row maximum as a tree (8 independent 3-input chains, then 3 levels): ~6 dependent steps instead of 32
      float m8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) m8[i] = fmaxf(s[i], fmaxf(s[8 + i], s[16 + i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) m8[i] = fmaxf(m8[i], fmaxf(s[24 + i], s[32 + i]));
#pragma unroll
      for (int i = 0; i < 8; ++i) m8[i] = fmaxf(m8[i], fmaxf(s[40 + i], fmaxf(s[48 + i], s[56 + i])));
      const float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
      const float m_blk = mx * scale_log2;  // scale > 0: max commutes with it
      // lazy running max: move it only when this block would push P above 2^RESCALE_THRESHOLD
      const bool need = m_blk > m_run + RESCALE_THRESHOLD;  // always true on the group's first block (m_run = -inf)
      const float alpha = (need && !first) ? ex2_approx(m_run - m_blk) : 1.0f;  // factor the old O / l have to take
      if (need) m_run = m_blk;
      if (it == ATTN_TRACE_ITEM) TRACE(j, 2);
      if constexpr (EXP_FIRST) {
      // (long groups; measured +8 % on the 1182-key shape, -6 % on the 197-key shape, whose items are too short to amortise
      // the extra register pressure)  The exponentials are computed BEFORE waiting for the previous product of this group: P stays packed in registers
      // (32 x bf16x2) until O[grp] / P[grp] are free, so the XU-bound part of the block overlaps the P V latency of the
      // block before it instead of queueing behind it.
      const float2 sc2 = make_float2(scale_log2, scale_log2);
      const float2 nm2 = make_float2(-m_run, -m_run);
      float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
      uint32_t pk[32];
      if (TAIL) {
#pragma unroll
        for (int i = 0; i < 32; ++i) pk[i] = 0u;
      }
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        if (TAIL && c8 * 8 >= ((nk + 15) & ~15)) break;  // chunks beyond the issued 16-key steps are never read
        float2 t0 = __ffma2_rn(make_float2(s[c8 * 8 + 0], s[c8 * 8 + 1]), sc2, nm2);
        float2 t1 = __ffma2_rn(make_float2(s[c8 * 8 + 2], s[c8 * 8 + 3]), sc2, nm2);
        float2 t2 = __ffma2_rn(make_float2(s[c8 * 8 + 4], s[c8 * 8 + 5]), sc2, nm2);
        float2 t3 = __ffma2_rn(make_float2(s[c8 * 8 + 6], s[c8 * 8 + 7]), sc2, nm2);
        // exp2 of 8 scores: ATTN_POLY_PAIRS of the 4 pairs go through the FMA-pipe polynomial, the rest through MUFU
        if (ATTN_POLY_PAIRS >= 4) t0 = ex2_poly2(t0); else { t0.x = ex2_approx(t0.x); t0.y = ex2_approx(t0.y); }
        if (ATTN_POLY_PAIRS >= 3) t1 = ex2_poly2(t1); else { t1.x = ex2_approx(t1.x); t1.y = ex2_approx(t1.y); }
        if (ATTN_POLY_PAIRS >= 2) t2 = ex2_poly2(t2); else { t2.x = ex2_approx(t2.x); t2.y = ex2_approx(t2.y); }
        if (ATTN_POLY_PAIRS >= 1) t3 = ex2_poly2(t3); else { t3.x = ex2_approx(t3.x); t3.y = ex2_approx(t3.y); }
        sum_a = __fadd2_rn(sum_a, t0);
        sum_b = __fadd2_rn(sum_b, t1);
        sum_a = __fadd2_rn(sum_a, t2);
        sum_b = __fadd2_rn(sum_b, t3);
        pk[c8 * 4 + 0] = pack_bf16(t0.x, t0.y);
        pk[c8 * 4 + 1] = pack_bf16(t1.x, t1.y);
        pk[c8 * 4 + 2] = pack_bf16(t2.x, t2.y);
        pk[c8 * 4 + 3] = pack_bf16(t3.x, t3.y);
      }
      sum_a = __fadd2_rn(sum_a, sum_b);
      l_run = fmaf(l_run, alpha, sum_a.x + sum_a.y);
      if (it == ATTN_TRACE_ITEM) TRACE(j, 3);
      if (!first) {
        // P[grp] / O[grp] are free once the previous product of this group has completed
        ptx::mbar_wait(o_full(grp), (done - 1) & 1u);
        ptx::tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {  // rare: the running max moved, O[grp] takes the factor in TMEM
#pragma unroll 1
          for (int c = 0; c < HD; c += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(tOg + (uint32_t)c, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32b_x16(tOg + (uint32_t)c, v);
          }
          tmem_st_wait();
        }
      }
      // P_j -> TMEM: register i of the row = keys (2i, 2i+1) = column i of P[grp] (columns beyond the tail's issued 16-key
      // steps are never read by the product)
      tmem_st_32x32b_x32(tPg, pk);
      } else {
      if (!first) {
        // P[grp] / O[grp] are free once the previous product of this group has completed
        ptx::mbar_wait(o_full(grp), (done - 1) & 1u);
        ptx::tc_fence_after();
        if (__any_sync(0xffffffffu, need)) {  // rare: the running max moved, O[grp] takes the factor in TMEM
#pragma unroll 1
          for (int c = 0; c < HD; c += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(tOg + (uint32_t)c, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st_32x32b_x16(tOg + (uint32_t)c, v);
          }
          tmem_st_wait();
        }
      }
      if (it == ATTN_TRACE_ITEM) TRACE(j, 3);
      const float2 sc2 = make_float2(scale_log2, scale_log2);
      const float2 nm2 = make_float2(-m_run, -m_run);
      float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
#pragma unroll
      for (int half = 0; half < 2; ++half) {   // 32 keys = 16 TMEM columns per tcgen05.st: only 16 packed words are live at a time
        if (TAIL && half * 32 >= ((nk + 15) & ~15)) break;
        uint32_t pk[16];
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int c8 = half * 4 + c4;
          if (TAIL && c8 * 8 >= ((nk + 15) & ~15)) {  // beyond the issued 16-key steps: never read
            pk[c4 * 4 + 0] = pk[c4 * 4 + 1] = pk[c4 * 4 + 2] = pk[c4 * 4 + 3] = 0u;
            continue;
          }
          float2 t0 = __ffma2_rn(make_float2(s[c8 * 8 + 0], s[c8 * 8 + 1]), sc2, nm2);
          float2 t1 = __ffma2_rn(make_float2(s[c8 * 8 + 2], s[c8 * 8 + 3]), sc2, nm2);
          float2 t2 = __ffma2_rn(make_float2(s[c8 * 8 + 4], s[c8 * 8 + 5]), sc2, nm2);
          float2 t3 = __ffma2_rn(make_float2(s[c8 * 8 + 6], s[c8 * 8 + 7]), sc2, nm2);
          if (ATTN_POLY_PAIRS >= 4) t0 = ex2_poly2(t0); else { t0.x = ex2_approx(t0.x); t0.y = ex2_approx(t0.y); }
          if (ATTN_POLY_PAIRS >= 3) t1 = ex2_poly2(t1); else { t1.x = ex2_approx(t1.x); t1.y = ex2_approx(t1.y); }
          if (ATTN_POLY_PAIRS >= 2) t2 = ex2_poly2(t2); else { t2.x = ex2_approx(t2.x); t2.y = ex2_approx(t2.y); }
          if (ATTN_POLY_PAIRS >= 1) t3 = ex2_poly2(t3); else { t3.x = ex2_approx(t3.x); t3.y = ex2_approx(t3.y); }
          sum_a = __fadd2_rn(sum_a, t0);
          sum_b = __fadd2_rn(sum_b, t1);
          sum_a = __fadd2_rn(sum_a, t2);
          sum_b = __fadd2_rn(sum_b, t3);
          pk[c4 * 4 + 0] = pack_bf16(t0.x, t0.y);
          pk[c4 * 4 + 1] = pack_bf16(t1.x, t1.y);
          pk[c4 * 4 + 2] = pack_bf16(t2.x, t2.y);
          pk[c4 * 4 + 3] = pack_bf16(t3.x, t3.y);
        }
        tmem_st_32x32b_x16(tPg + (uint32_t)(half * 16), pk);
      }
      sum_a = __fadd2_rn(sum_a, sum_b);
      l_run = fmaf(l_run, alpha, sum_a.x + sum_a.y);
      }
      if (it == ATTN_TRACE_ITEM) TRACE(j, 4);
    };

    for (int p = blockIdx.x; p < n_pairs; p += gridDim.x, ++it) {
      const int w = 2 * p + grp;
      if (w >= n_items) continue;  // the grid's last pair may hold one item only
      const int qt = w % n_qt, h = (w / n_qt) % heads, g_idx = w / (n_qt * heads);
      const int row0 = g_idx * group_len, q0 = qt * BQ;
      const bool warp_has_rows = q0 + quarter * 32 < group_len;  // warp-uniform
      m_run = -INFINITY;
      l_run = 0.f;
      for (int j = 0; j < n_blocks; ++j, ++done) {
        const bool first = j == 0;
        ptx::mbar_wait(s_full(grp), done & 1u);
        ptx::tc_fence_after();
        if (warp_has_rows) {
          if (j + 1 < n_blocks || group_len % BKV == 0)
            block(j, first, std::false_type{});
          else
            block(j, first, std::true_type{});
        } else {
          // no valid query rows in this warp: keep the barrier phases in step
          ptx::mbar_arrive_elect(s_free(grp));
          if (!first) ptx::mbar_wait(o_full(grp), (done - 1) & 1u);
        }
        tmem_st_wait();            // P (and a rescaled O) have landed in tensor memory ...
        ptx::tc_fence_before();    // ... and are ordered before the MMA that reads / accumulates them
        __syncwarp();
        ptx::mbar_arrive_elect(p_full(grp));  // one arrival per warp: the group's P V issuer takes it from here
        __syncwarp();
      }
      // ---- the item's output rows: O[grp] / l once the last product has completed
      if (it == ATTN_TRACE_ITEM) TRACE(31, 5);
      ptx::mbar_wait(o_full(grp), (done - 1) & 1u);
      ptx::tc_fence_after();
      if (it == ATTN_TRACE_ITEM) TRACE(31, 6);
      if (warp_has_rows) {
        // O[grp] / l -> bf16 -> this warp's staging slice (32 rows x 32 dims, 16-byte chunks XOR-swizzled so that both the row-wise
        // writes and the 8-rows-per-instruction reads are bank-conflict free) -> global memory with 4 lanes per 64-byte run
        // of a row.  (One 16-byte store per lane straight from the row's registers touched 32 different 128-byte lines per
        // instruction: 8 such instructions per thread kept the warps ~2500 cycles in the LSU replay queue per item.)
        const float inv = 1.f / l_run;
        const uint32_t stage = sStage + (uint32_t)((warp - 4) * 2048);
        const int rows_left = group_len - (q0 + quarter * 32);  // valid rows of this warp's slice (may exceed 32)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t a[32];
          ptx::tmem_ld_32x32b_x32(tOg + (uint32_t)(half * 32), a);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t u0 = pack_bf16(__uint_as_float(a[c * 8 + 0]) * inv, __uint_as_float(a[c * 8 + 1]) * inv);
            const uint32_t u1 = pack_bf16(__uint_as_float(a[c * 8 + 2]) * inv, __uint_as_float(a[c * 8 + 3]) * inv);
            const uint32_t u2 = pack_bf16(__uint_as_float(a[c * 8 + 4]) * inv, __uint_as_float(a[c * 8 + 5]) * inv);
            const uint32_t u3 = pack_bf16(__uint_as_float(a[c * 8 + 6]) * inv, __uint_as_float(a[c * 8 + 7]) * inv);
            const uint32_t addr = stage + (uint32_t)(lane * 64) + (uint32_t)((c ^ ((lane >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(u0), "r"(u1), "r"(u2), "r"(u3) : "memory");
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = i * 8 + (lane >> 2), c = lane & 3;
            uint4 u;
            const uint32_t addr = stage + (uint32_t)(row * 64) + (uint32_t)((c ^ ((row >> 1) & 3)) << 4);
            asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr) : "memory");
            if (row < rows_left)
              *reinterpret_cast<uint4*>(out + (size_t)(row0 + q0 + quarter * 32 + row) * ldo + h * HD + half * 32 + c * 8) = u;
          }
          __syncwarp();
        }
        if (it == ATTN_TRACE_ITEM) TRACE(31, 10);
      }
      // O[grp] is in registers / memory: the next item's first product (issued only after this warp's next p_full arrival,
      // which follows this fence in program order) may overwrite it
      ptx::tc_fence_before();
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
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
#ifdef ATTN_FORCE_POLY_PAIRS
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc_kernel<ATTN_FORCE_POLY_PAIRS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc_kernel<ATTN_FORCE_POLY_PAIRS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
#endif
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int rows = n_groups * group_len;
  CUtensorMap tq, tkv;
  if (!gemm_get_tensor_map(qkv, rows, 3 * heads * HD, ld_qkv, HD, BQ, &tq)) return cudaErrorInvalidValue;
  if (!gemm_get_tensor_map(qkv, rows, 3 * heads * HD, ld_qkv, HD, BKV, &tkv)) return cudaErrorInvalidValue;
  const int n_items = ((group_len + BQ - 1) / BQ) * heads * n_groups;
  const int n_pairs = (n_items + 1) / 2;
  int grid = 2 * gemm_sm_count();  // two CTAs per SM (TMEM: 2 x 256 columns), each walks pairs of items with stride `grid`
  if (n_pairs < grid) grid = n_pairs;
  const float sl2 = scale * 1.4426950408889634f;
#ifdef ATTN_FORCE_POLY_PAIRS  // tuning builds (tools/attn_bench.cu): the same exp2 split for every shape, the usual order of the phases
  if (group_len >= LONG_GROUP)
    attention_tc_kernel<ATTN_FORCE_POLY_PAIRS, true><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tkv, out, ldo, group_len, heads, n_groups, sl2);
  else
    attention_tc_kernel<ATTN_FORCE_POLY_PAIRS, false><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tkv, out, ldo, group_len, heads, n_groups, sl2);
#else
  if (group_len >= LONG_GROUP)
    attention_tc_kernel<1, true><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tkv, out, ldo, group_len, heads, n_groups, sl2);
  else
    attention_tc_kernel<0, false><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tq, tkv, out, ldo, group_len, heads, n_groups, sl2);
#endif
  note_launch();
  return cudaGetLastError();
}
