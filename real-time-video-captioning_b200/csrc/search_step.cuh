// Device-side body of one caption-search step (shared by search.cu's kernel and the persistent decode kernel, decode_mega.cu).
#pragma once
#include <math.h>

#include "common.cuh"
#include "kernels.h"

namespace search_dev {

constexpr int SS_THREADS = 256;
constexpr int MAX_CAND = 16;  // per_node_beam_size * num_beams upper bound
constexpr int MAX_NB = 8;

__device__ __forceinline__ bool better(float s1, int i1, float s2, int i2) { return s1 > s2 || (s1 == s2 && i1 < i2); }

// One search step of clip `clip`, executed by SS_THREADS threads (tid = 0 .. SS_THREADS-1) that meet in sync()
// (search_step_kernel: one CTA per clip, __syncthreads; the persistent single-clip decode kernel calls it from the first
// SS_THREADS threads of its CTA 0 with a named barrier).  COHERENT: the logits were written by other CTAs of the same launch.
template <bool COHERENT, class Sync>
__device__ __forceinline__ void search_step_device(const SearchState& st, const float* __restrict__ logits, int cur_len, int parity,
                                                   int clip, int tid, Sync sync) {
  __shared__ float red_f[SS_THREADS / 32];
  __shared__ int red_i[SS_THREADS / 32];
  __shared__ float row_lse_max[MAX_NB], row_lse[MAX_NB];
  __shared__ float cand_score[MAX_CAND];
  __shared__ int cand_idx[MAX_CAND];
  __shared__ int nxt_parent[MAX_NB], nxt_word[MAX_NB];
  __shared__ float nxt_score[MAX_NB];
  __shared__ float bc_f;
  __shared__ int bc_i;

  const int warp = tid >> 5, lane = tid & 31;
  auto ldx = [](const float* q) { return COHERENT ? __ldcg(q) : *q; };
  const int nb = st.nb, V = st.V, C = st.cand;
  const int* tok_in = parity ? st.tokens_tmp : st.tokens;
  int* tok_out = parity ? st.tokens : st.tokens_tmp;
  const int* anc_in = parity ? st.anc_tmp : st.anc;
  int* anc_out = parity ? st.anc : st.anc_tmp;

  // ---- 1. log-softmax statistics per beam row: max, then log(sum(exp(x - max)))   (model.py:557)
  for (int b = 0; b < nb; ++b) {
    const float* x = logits + (size_t)(clip * nb + b) * st.ldl;
    float m = -INFINITY;
    for (int i = tid; i < V; i += SS_THREADS) m = fmaxf(m, ldx(x + i));
    m = warp_max(m);
    if (lane == 0) red_f[warp] = m;
    sync();
    if (tid == 0) {
      float mm = red_f[0];
      for (int w = 1; w < SS_THREADS / 32; ++w) mm = fmaxf(mm, red_f[w]);
      bc_f = mm;
    }
    sync();
    m = bc_f;
    float s = 0.f;
    for (int i = tid; i < V; i += SS_THREADS) s += expf(ldx(x + i) - m);
    s = warp_sum(s);
    sync();
    if (lane == 0) red_f[warp] = s;
    sync();
    if (tid == 0) {
      float ss = 0.f;
      for (int w = 0; w < SS_THREADS / 32; ++w) ss += red_f[w];
      row_lse_max[b] = m;
      row_lse[b] = logf(ss);
    }
    sync();
  }

  // ---- 2. top-C of (log_softmax + beam_score) over nb * V candidates               (model.py:561-565)
  float ls[MAX_CAND];
  int li[MAX_CAND];
#pragma unroll
  for (int j = 0; j < MAX_CAND; ++j) {
    ls[j] = -INFINITY;
    li[j] = 0x7fffffff;
  }
  for (int b = 0; b < nb; ++b) {
    const float* x = logits + (size_t)(clip * nb + b) * st.ldl;
    const float m = row_lse_max[b], l = row_lse[b], bs = st.beam_scores[clip * nb + b];
    for (int i = tid; i < V; i += SS_THREADS) {
      const float sc = ((ldx(x + i) - m) - l) + bs;
      const int id = b * V + i;
      if (better(sc, id, ls[C - 1], li[C - 1])) {
        // insertion into the sorted local list (C is tiny)
        int j = C - 1;
        while (j > 0 && better(sc, id, ls[j - 1], li[j - 1])) {
          ls[j] = ls[j - 1];
          li[j] = li[j - 1];
          --j;
        }
        ls[j] = sc;
        li[j] = id;
      }
    }
  }
  int head = 0;  // next unconsumed entry of this thread's sorted list
  for (int c = 0; c < C; ++c) {
    float bs_ = head < C ? ls[head] : -INFINITY;
    int bi_ = head < C ? li[head] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float s2 = __shfl_xor_sync(0xffffffffu, bs_, o);
      const int i2 = __shfl_xor_sync(0xffffffffu, bi_, o);
      if (better(s2, i2, bs_, bi_)) {
        bs_ = s2;
        bi_ = i2;
      }
    }
    if (lane == 0) {
      red_f[warp] = bs_;
      red_i[warp] = bi_;
    }
    sync();
    if (tid == 0) {
      float s = red_f[0];
      int i = red_i[0];
      for (int w = 1; w < SS_THREADS / 32; ++w)
        if (better(red_f[w], red_i[w], s, i)) {
          s = red_f[w];
          i = red_i[w];
        }
      cand_score[c] = s;
      cand_idx[c] = i;
      bc_i = i;
    }
    sync();
    if (head < C && li[head] == bc_i) ++head;  // candidate ids are unique, exactly one thread pops
    sync();
  }

  // ---- 3. candidate walk (thread 0), exactly the reference's control flow           (model.py:573-611)
  if (tid == 0) {
    const int max_len = st.max_len;
    const int nk = st.n_keep;
    double* hs = st.hyp_score + (size_t)clip * (nk + 1);
    int* hl = st.hyp_len + (size_t)clip * (nk + 1);
    int* ht = st.hyp_tok + (size_t)clip * (nk + 1) * max_len;
    int cnt = st.hyp_count[clip];
    double worst = st.worst[clip];
    int is_done = st.done[clip];
    if (!is_done) {
      // BeamHypotheses.is_done(best_sum_logprobs) with early_stopping=False
      if (cnt >= nk) is_done = worst >= (double)cand_score[0] / pow((double)(max_len - 1), (double)st.length_penalty);
    }
    int n_next = 0;
    if (!is_done) {
      for (int c = 0; c < C; ++c) {
        const int beam_id = cand_idx[c] / V, word = cand_idx[c] % V;
        const int prow = clip * nb + beam_id;
        if (word == st.eos || cur_len + 1 == max_len) {
          // BeamHypotheses.add(input_ids[prow, :cur_len], score)
          const double score = (double)cand_score[c] / pow((double)cur_len, (double)st.length_penalty);
          if (cnt < nk || score > worst) {
            hs[cnt] = score;
            hl[cnt] = cur_len;
            for (int t = 0; t < cur_len; ++t) ht[(size_t)cnt * max_len + t] = tok_in[(size_t)prow * max_len + t];
            ++cnt;
            if (cnt > nk) {
              // drop the worst (first minimum in (score, index) order), keep insertion order of the rest
              int w0 = 0;
              for (int k = 1; k < cnt; ++k)
                if (hs[k] < hs[w0]) w0 = k;
              for (int k = w0; k + 1 < cnt; ++k) {
                hs[k] = hs[k + 1];
                hl[k] = hl[k + 1];
                for (int t = 0; t < max_len; ++t) ht[(size_t)k * max_len + t] = ht[(size_t)(k + 1) * max_len + t];
              }
              --cnt;
              double w1 = hs[0];
              for (int k = 1; k < cnt; ++k) w1 = fmin(w1, hs[k]);
              worst = w1;
            } else {
              worst = fmin(score, worst);
            }
          }
        } else {
          nxt_score[n_next] = cand_score[c];
          nxt_word[n_next] = word;
          nxt_parent[n_next] = prow;
          ++n_next;
        }
        if (n_next == nb) break;
      }
    }
    if (n_next != nb) {
      // done clip, or the final step: "pad the batch" (model.py:578, :609).  The reference pads with global
      // row 0 as parent; the padded rows are never read again, so the clip's own row is used instead.
      for (int b = 0; b < nb; ++b) {
        nxt_score[b] = 0.f;
        nxt_word[b] = st.eos;
        nxt_parent[b] = clip * nb + b;
      }
    }
    st.hyp_count[clip] = cnt;
    st.worst[clip] = worst;
    if (is_done && !st.done[clip]) atomicAdd(st.done_count, 1);  // the host polls this to leave the step loop early
    st.done[clip] = is_done;
  }
  sync();

  // ---- 4. re-order token rows (model.py:615-621) and, optionally, the text-KV ancestor table
  for (int b = 0; b < nb; ++b) {
    const int row = clip * nb + b, prow = nxt_parent[b];
    for (int t = tid; t < cur_len; t += SS_THREADS) tok_out[(size_t)row * st.max_len + t] = tok_in[(size_t)prow * st.max_len + t];
    if (st.reorder_cache) {
      // ancestors of text positions 0 .. cur_len-2 are inherited; position cur_len-1 was computed by the parent itself
      for (int t = tid; t < cur_len - 1; t += SS_THREADS) anc_out[(size_t)row * st.max_len + t] = anc_in[(size_t)prow * st.max_len + t];
      if (tid == 0) anc_out[(size_t)row * st.max_len + cur_len - 1] = prow;
    }
    if (tid == 0) {
      tok_out[(size_t)row * st.max_len + cur_len] = nxt_word[b];
      st.cur_tok[row] = nxt_word[b];
      st.beam_scores[row] = nxt_score[b];
    }
  }
}


}  // namespace search_dev
