// Device-resident caption search: the reference's GeneratorWithBeamSearchV2.search
// (/root/reference/src/models/model.py:479-678, greedy-beam branch) and the upstream BeamHypotheses it
// uses, executed by one CTA per clip so that no logits, scores or indices ever cross PCIe per step.
//
//   step kernel:  log-softmax of every beam row (block reduce with warp shuffles), + running beam score,
//                 exact top-(per_node_beam_size * num_beams) over the clip's num_beams * V candidates,
//                 then thread 0 walks the candidates exactly like model.py:573-611 (EOS / last-step
//                 hypotheses, next beam content) and all threads re-order token rows and the text-KV
//                 ancestor table by the chosen parents.
#include <math.h>

#include "common.cuh"
#include "kernels.h"

#include "search_step.cuh"

namespace {

using search_dev::MAX_CAND;
using search_dev::MAX_NB;
using search_dev::SS_THREADS;

__global__ void __launch_bounds__(SS_THREADS) search_step_kernel(SearchState st, const float* __restrict__ logits, int cur_len, int parity) {
  __shared__ search_dev::SearchSmem sh;
  search_dev::search_step_device<false>(st, logits, cur_len, parity, blockIdx.x, threadIdx.x, [] { __syncthreads(); }, sh);
}

// Beam search at batch size: one CTA per beam ROW for the log-softmax statistics and the row's own top candidates (the one-CTA-
// per-clip kernel walks a clip's rows one after the other: 244 us per step at 256 clips x 4 beams, 1.7 waves of 4-pass CTAs),
// then one CTA per clip merges its rows' sorted lists and does the candidate walk.  The top `cand` of a clip lie inside the union
// of its rows' top `cand`, and the order `better` is total: same candidates in the same order, bit for bit.
__global__ void __launch_bounds__(SS_THREADS) search_rows_kernel(SearchState st, const float* __restrict__ logits) {
  __shared__ search_dev::SearchSmem sh;
  const int b = blockIdx.x, clip = blockIdx.y, nb = st.nb, ldl = st.ldl, C = st.cand;
  auto load = [=](int bb, int i) { return logits[(size_t)(clip * nb + bb) * ldl + i]; };
  search_dev::search_topc_rows(st, load, b, b + 1, clip, threadIdx.x, [] { __syncthreads(); }, sh);
  if (threadIdx.x < C) {
    st.row_cand_score[(size_t)(clip * nb + b) * C + threadIdx.x] = sh.cand_score[threadIdx.x];
    st.row_cand_idx[(size_t)(clip * nb + b) * C + threadIdx.x] = sh.cand_idx[threadIdx.x];
  }
}

__global__ void __launch_bounds__(SS_THREADS) search_merge_walk_kernel(SearchState st, int cur_len, int parity) {
  __shared__ search_dev::SearchSmem sh;
  __shared__ float all_s[search_dev::MAX_NB * search_dev::MAX_CAND];
  __shared__ int all_i[search_dev::MAX_NB * search_dev::MAX_CAND];
  const int clip = blockIdx.x, tid = threadIdx.x, nb = st.nb, C = st.cand;
  if (tid < nb * C) {
    all_s[tid] = st.row_cand_score[(size_t)clip * nb * C + tid];
    all_i[tid] = st.row_cand_idx[(size_t)clip * nb * C + tid];
  }
  __syncthreads();
  if (tid == 0) {
    int head[search_dev::MAX_NB];
    for (int r = 0; r < nb; ++r) head[r] = 0;
    for (int c = 0; c < C; ++c) {  // nb sorted lists of C candidates -> the best C of all, in `better` order
      int br = -1, bi = 0;
      float bs = 0.f;
      for (int r = 0; r < nb; ++r) {
        if (head[r] >= C) continue;
        const float s2 = all_s[r * C + head[r]];
        const int i2 = all_i[r * C + head[r]];
        if (br < 0 || search_dev::better(s2, i2, bs, bi)) {
          br = r;
          bs = s2;
          bi = i2;
        }
      }
      sh.cand_score[c] = bs;
      sh.cand_idx[c] = bi;
      ++head[br];
    }
  }
  __syncthreads();
  search_dev::search_walk_reorder(st, cur_len, parity, clip, tid, [] { __syncthreads(); }, sh);
}

__global__ void search_init_kernel(SearchState st, int sos) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int n_rows = st.n_clips * st.nb;
  if (row < n_rows) {
    st.tokens[(size_t)row * st.max_len] = sos;
    st.cur_tok[row] = sos;
    st.beam_scores[row] = (row % st.nb == 0) ? 0.f : -1e9f;  // model.py:508-510
    for (int t = 0; t < st.max_len; ++t) {
      st.anc[(size_t)row * st.max_len + t] = row;
      st.anc_tmp[(size_t)row * st.max_len + t] = row;
    }
  }
  if (row == 0) *st.done_count = 0;
  if (row < st.n_clips) {
    st.done[row] = 0;
    st.hyp_count[row] = 0;
    st.worst[row] = 1e9;
  }
}

// model.py:653-677: best n_keep hypotheses per clip, eos padded, + one eos appended
__global__ void search_finalize_kernel(SearchState st, int* decoded, float* logprobs) {
  const int clip = blockIdx.x * blockDim.x + threadIdx.x;
  if (clip >= st.n_clips) return;
  const int nk = st.n_keep, max_len = st.max_len;
  const double* hs = st.hyp_score + (size_t)clip * (nk + 1);
  const int* hl = st.hyp_len + (size_t)clip * (nk + 1);
  const int* ht = st.hyp_tok + (size_t)clip * (nk + 1) * max_len;
  const int cnt = st.hyp_count[clip];
  bool used[MAX_CAND];
  for (int k = 0; k < cnt && k < MAX_CAND; ++k) used[k] = false;
  for (int k = 0; k < nk; ++k) {
    int* dst = decoded + ((size_t)clip * nk + k) * max_len;
    for (int t = 0; t < max_len; ++t) dst[t] = st.eos;
    logprobs[(size_t)clip * nk + k] = -1e5f;
    if (k >= cnt) continue;
    int best = -1;  // torch.topk over float32 scores: largest first, earliest index on ties
    for (int j = 0; j < cnt; ++j)
      if (!used[j] && (best < 0 || (float)hs[j] > (float)hs[best])) best = j;
    used[best] = true;
    for (int t = 0; t < hl[best]; ++t) dst[t] = ht[(size_t)best * max_len + t];
    logprobs[(size_t)clip * nk + k] = (float)hs[best];
  }
}

__global__ void anc_reorder_kernel(const int* __restrict__ in, int* __restrict__ out, const int* __restrict__ beam_idx,
                                   int rows, int ld, int pos) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  const int p = beam_idx[r];
  for (int s = threadIdx.x; s < pos; s += blockDim.x) out[(size_t)r * ld + s] = in ? in[(size_t)p * ld + s] : p;
  if (threadIdx.x == 0) out[(size_t)r * ld + pos] = p;
}

}  // namespace

cudaError_t anc_reorder(const int* anc_in, int* anc_out, const int* beam_idx, int rows, int ld, int pos, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  anc_reorder_kernel<<<rows, 64, 0, stream>>>(anc_in, anc_out, beam_idx, rows, ld, pos);
  note_launch();
  return cudaGetLastError();
}

cudaError_t search_init(const SearchState& s, int sos, cudaStream_t stream) {
  const int n = s.n_clips * s.nb;
  if (n <= 0) return cudaSuccess;
  if (s.nb > MAX_NB || s.cand > MAX_CAND || s.n_keep + 1 > MAX_CAND) return cudaErrorInvalidValue;
  search_init_kernel<<<(n + 127) / 128, 128, 0, stream>>>(s, sos);
  note_launch();
  return cudaGetLastError();
}

cudaError_t search_step(const SearchState& s, const float* logits, int cur_len, int parity, cudaStream_t stream) {
  if (s.n_clips <= 0) return cudaSuccess;
  if (s.nb > MAX_NB || s.cand > MAX_CAND || cur_len + 1 > s.max_len) return cudaErrorInvalidValue;
  if (s.nb > 1 && s.row_cand_score != nullptr && s.row_cand_idx != nullptr && s.n_clips <= 65535) {
    search_rows_kernel<<<dim3(s.nb, s.n_clips), SS_THREADS, 0, stream>>>(s, logits);
    note_launch();
    search_merge_walk_kernel<<<s.n_clips, SS_THREADS, 0, stream>>>(s, cur_len, parity);
  } else {
    search_step_kernel<<<s.n_clips, SS_THREADS, 0, stream>>>(s, logits, cur_len, parity);
  }
  note_launch();
  return cudaGetLastError();
}

cudaError_t search_finalize(const SearchState& s, int* decoded, float* logprobs, cudaStream_t stream) {
  if (s.n_clips <= 0) return cudaSuccess;
  search_finalize_kernel<<<(s.n_clips + 63) / 64, 64, 0, stream>>>(s, decoded, logprobs);
  note_launch();
  return cudaGetLastError();
}
