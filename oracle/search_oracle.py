"""CPU oracle for the caption search loop -- TEST INFRASTRUCTURE ONLY (see oracle/git_oracle.py header).

Restates ``GeneratorWithBeamSearchV2.search`` (/root/reference/src/models/model.py:479-678): the
greedy-beam branch the reference uses (do_sample=False), the sampling branch (:532-554) with the upstream
``top_k_top_p_filtering`` (:537), and the upstream ``BeamHypotheses`` helper it constructs at model.py:503 (legacy HuggingFace beam hypotheses container; recalled -- SURVEY 3.3).
Plus ``infer`` (model.py:426-462) and the teacher's caption / logit post-processing (model.py:762-793).

PARITY STATUS: ``search`` and ``infer`` are PINNED TO THE REFERENCE'S OWN CODE: oracle/make_reference_golden.py imports
/root/reference/src/models/model.py unmodified (import stubs for the packages this image lacks) and runs its
``GeneratorWithBeamSearchV2.search`` / ``GenerativeImageTextModel.infer`` on seeded inputs; tests/test_reference_golden.py
replays the frozen outputs (tests/golden/ref_search.npz, ref_git_glue.npz) through this file: tokens, log-probabilities (bit
for bit), the step at which the loop breaks and the re-ordered beam histories agree.  The reference holds no test of its own
for search output.  NOT pinned that way: ``BeamHypotheses`` and ``top_k_top_p_filtering`` (upstream generativeimage2text code,
absent here: the published legacy-HuggingFace algorithm, exercised by the hand-worked cases in tests/test_cpu_oracle_and_host.py;
the reference's search ran ON TOP of these restatements when the fixtures were made).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .git_oracle import DecodingState, GitConfig, encode_clip


class BeamHypotheses:
    """Upstream generativeimage2text BeamHypotheses (constructed at model.py:503)."""

    def __init__(self, n_hyp: int, max_length: int, length_penalty: float, early_stopping: bool):
        self.max_length = max_length - 1  # ignoring bos_token
        self.length_penalty = length_penalty
        self.early_stopping = early_stopping
        self.n_hyp = n_hyp
        self.hyp: List[Tuple[float, torch.Tensor]] = []
        self.worst_score = 1e9

    def __len__(self):
        return len(self.hyp)

    def add(self, hyp: torch.Tensor, sum_logprobs: float):
        score = sum_logprobs / len(hyp) ** self.length_penalty
        if len(self) < self.n_hyp or score > self.worst_score:
            self.hyp.append((score, hyp))
            if len(self) > self.n_hyp:
                sorted_scores = sorted([(s, idx) for idx, (s, _) in enumerate(self.hyp)])
                del self.hyp[sorted_scores[0][1]]
                self.worst_score = sorted_scores[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs: float) -> bool:
        if len(self) < self.n_hyp:
            return False
        elif self.early_stopping:
            return True
        else:
            return self.worst_score >= best_sum_logprobs / self.max_length ** self.length_penalty


def top_k_top_p_filtering(logits: torch.Tensor, top_k=0, top_p=1.0, filter_value=-float("Inf"), min_tokens_to_keep=1):
    """Upstream generativeimage2text ``top_k_top_p_filtering`` (called at model.py:537; the legacy HuggingFace sampling
    filter, recalled): keep the top_k highest logits and/or the smallest prefix of the sorted distribution whose
    cumulative probability exceeds top_p (never fewer than min_tokens_to_keep); everything else becomes filter_value.
    Like upstream, ``top_k=None`` / ``top_p=None`` (the defaults ``search`` forwards) raise TypeError on the comparisons."""
    if top_k > 0:
        top_k = min(max(top_k, min_tokens_to_keep), logits.size(-1))
        kth = torch.topk(logits, top_k)[0][..., -1, None]
        logits[logits < kth] = filter_value
    if top_p < 1.0:
        sorted_logits, sorted_indices = torch.sort(logits, descending=True)
        cumulative = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
        remove_sorted = cumulative > top_p
        if min_tokens_to_keep > 1:
            remove_sorted[..., :min_tokens_to_keep] = 0
        remove_sorted[..., 1:] = remove_sorted[..., :-1].clone()  # shift right: the token that crosses top_p is kept
        remove_sorted[..., 0] = 0
        remove = remove_sorted.scatter(1, sorted_indices, remove_sorted)
        logits[remove] = filter_value
    return logits


def search(input_ids: torch.Tensor, step: Callable[[torch.Tensor], torch.Tensor], *, eos_index: int, max_steps: int,
           beam_size: int, length_penalty: float, per_node_beam_size: int = 2, num_keep_best: int = 1,
           on_reorder: Optional[Callable[[torch.Tensor], None]] = None, save_logits: bool = True,
           do_sample: bool = False, top_k=None, top_p=None, num_return_sequences: int = 1,
           repetition_penalty: float = 1.0, temperature: float = 1.0):
    """model.py:479-678: the greedy-beam branch the reference uses and the sampling branch (:532-554).
    Returns (decoded, logprobs, saved_logits).

    ``on_reorder(beam_idx)`` is called after every beam re-order; the reference never re-indexes the
    model-side cache (model.py:623-634 is commented out), so passing None reproduces it exactly.
    """
    if num_return_sequences != 1:                                                       # :481-484
        input_ids = input_ids[:, None, :].expand(input_ids.shape[0], num_return_sequences, input_ids.shape[1])
        input_ids = input_ids.reshape(-1, input_ids.shape[-1])
    batch_size, cur_len = input_ids.shape
    num_beams = beam_size
    pad_token_id = eos_index
    eos_token_ids = [eos_index]

    input_ids = input_ids.unsqueeze(1).expand(batch_size, num_beams, cur_len)           # :494
    input_ids = input_ids.contiguous().view(batch_size * num_beams, cur_len)            # :495
    max_length = max_steps                                                              # :500
    generated_hyps = [BeamHypotheses(num_keep_best, max_length, length_penalty, early_stopping=False)
                      for _ in range(batch_size)]                                       # :502-505
    beam_scores = torch.zeros((batch_size, num_beams), dtype=torch.float)               # :508
    beam_scores[:, 1:] = -1e9                                                           # :509
    beam_scores = beam_scores.view(-1)                                                  # :510
    done = [False for _ in range(batch_size)]                                           # :516
    saved_logits = []
    while cur_len < max_length:                                                         # :518
        scores = step(input_ids)                                                        # :519
        vocab_size = scores.shape[-1]
        if save_logits:
            saved_logits.append([i.detach().cpu().numpy() for i in scores])             # :521
        if repetition_penalty != 1.0:                                                   # :524-531
            for i in range(batch_size * num_beams):
                for previous_token in set(input_ids[i].tolist()):
                    if scores[i, previous_token] < 0:
                        scores[i, previous_token] *= repetition_penalty
                    else:
                        scores[i, previous_token] /= repetition_penalty
        if do_sample:                                                                   # :532
            if temperature != 1.0:
                scores = scores / temperature                                           # :535
            scores = top_k_top_p_filtering(scores, top_k=top_k, top_p=top_p, min_tokens_to_keep=2)   # :537
            next_words = torch.multinomial(F.softmax(scores, dim=-1), num_samples=per_node_beam_size)  # :540
            _scores = F.log_softmax(scores, dim=-1)                                     # :543
            _scores = torch.gather(_scores, -1, next_words)                             # :544
            next_scores = _scores + beam_scores[:, None].expand_as(_scores)             # :545
            beam_indices = torch.arange(num_beams, device=next_words.device) * vocab_size   # :548
            beam_indices = beam_indices.repeat(batch_size, per_node_beam_size)          # :549
            next_words = next_words.view(batch_size, per_node_beam_size * num_beams)     # :550
            next_words = next_words + beam_indices                                      # :552
            next_scores = next_scores.view(batch_size, per_node_beam_size * num_beams)  # :553
        else:
            scores = F.log_softmax(scores, dim=-1)                                      # :557
            assert scores.size() == (batch_size * num_beams, vocab_size)
            _scores = scores + beam_scores[:, None].expand_as(scores)                   # :561
            _scores = _scores.view(batch_size, num_beams * vocab_size)                  # :563
            next_scores, next_words = torch.topk(_scores, per_node_beam_size * num_beams, dim=1, largest=True,
                                                 sorted=True)                           # :564
        next_batch_beam = []
        for batch_ex in range(batch_size):                                              # :573
            done[batch_ex] = done[batch_ex] or generated_hyps[batch_ex].is_done(next_scores[batch_ex].max().item())
            if done[batch_ex]:
                next_batch_beam.extend([(0, pad_token_id, 0)] * num_beams)              # :578
                continue
            next_sent_beam = []
            for idx, score in zip(next_words[batch_ex], next_scores[batch_ex]):         # :585
                beam_id = idx // vocab_size
                word_id = idx % vocab_size
                if word_id.item() in eos_token_ids or cur_len + 1 == max_length:        # :592
                    generated_hyps[batch_ex].add(input_ids[batch_ex * num_beams + beam_id, :cur_len].clone(),
                                                 score.item())                          # :593
                else:
                    next_sent_beam.append((score, word_id, batch_ex * num_beams + beam_id))
                if len(next_sent_beam) == num_beams:                                    # :599
                    break
            if cur_len + 1 == max_length:
                assert len(next_sent_beam) == 0
            else:
                assert len(next_sent_beam) == num_beams
            if len(next_sent_beam) == 0:
                next_sent_beam = [(0, pad_token_id, 0)] * num_beams                     # :609
            next_batch_beam.extend(next_sent_beam)
        assert len(next_batch_beam) == batch_size * num_beams
        beam_scores = beam_scores.new([x[0] for x in next_batch_beam])                  # :615
        beam_words = input_ids.new([x[1] for x in next_batch_beam])                     # :616
        beam_idx = input_ids.new([x[2] for x in next_batch_beam])                       # :617
        input_ids = input_ids[beam_idx, :]                                              # :620
        input_ids = torch.cat([input_ids, beam_words.unsqueeze(1)], dim=-1)             # :621
        if on_reorder is not None:
            on_reorder(beam_idx)
        cur_len = cur_len + 1                                                           # :637
        if all(done):                                                                   # :640
            break

    tgt_len = torch.ones(batch_size, num_keep_best, dtype=torch.long)                   # :653
    logprobs = torch.zeros(batch_size, num_keep_best, dtype=torch.float).fill_(-1e5)    # :654
    all_best = []
    for i, hypotheses in enumerate(generated_hyps):                                     # :658
        best = []
        hyp_scores = torch.tensor([x[0] for x in hypotheses.hyp])
        _, best_indices = torch.topk(hyp_scores, min(num_keep_best, len(hyp_scores)), largest=True)
        for best_idx, hyp_idx in enumerate(best_indices):
            conf, best_hyp = hypotheses.hyp[hyp_idx]
            best.append(best_hyp)
            logprobs[i, best_idx] = conf
            tgt_len[i, best_idx] = len(best_hyp) + 1
        all_best.append(best)
    decoded = input_ids.new(batch_size, num_keep_best, max_length).fill_(pad_token_id)  # :671
    for batch_idx, best in enumerate(all_best):
        for best_idx, hypo in enumerate(best):
            decoded[batch_idx, best_idx, : tgt_len[batch_idx, best_idx] - 1] = hypo
            decoded[batch_idx, best_idx, tgt_len[batch_idx, best_idx] - 1] = eos_token_ids[0]
    if num_keep_best == 1:
        decoded = decoded.squeeze(dim=1)
    return decoded, logprobs, saved_logits


def infer(sd, cfg: GitConfig, visual_features: torch.Tensor, *, beam_size: Optional[int] = None,
          max_steps: Optional[int] = None, reorder_cache: bool = False, num_keep_best: int = 1,
          save_logits: bool = True):
    """GenerativeImageTextModel.infer (model.py:426-462) for a batch of visual features [B, Nv, Dv]."""
    B = visual_features.size(0)
    start = torch.full((B, 1), cfg.sos_index, dtype=torch.long)                         # :429-431
    state = DecodingState(sd, cfg, visual_features, reorder_cache=reorder_cache)        # :439-445
    decoded, logprobs, saved = search(
        start, state, eos_index=cfg.eos_index, max_steps=max_steps or cfg.max_steps,
        beam_size=beam_size or cfg.beam_size, length_penalty=cfg.length_penalty,
        per_node_beam_size=cfg.per_node_beam_size, num_keep_best=num_keep_best,
        on_reorder=state.reorder if reorder_cache else None, save_logits=save_logits)
    return {"predictions": decoded, "logprobs": logprobs, "logits_dict": saved, "visual_features": visual_features}


def caption_clip(sd, cfg: GitConfig, frames: torch.Tensor, *, per_frame_calls: bool = True, **kw):
    """``self.model({'image': imgs})`` for one clip in eval mode (model.py:768): F separate ViT calls
    (CaptioningModel.forward_one), temporal embeddings, concat, then infer."""
    vf = encode_clip(sd, cfg, frames, per_frame_calls=per_frame_calls)
    return infer(sd, cfg, vf, **kw)


def teacher_postprocess(result: dict, detok: Callable[[List[int]], str], num_beams: int = 4) -> dict:
    """GenerativeImageTextTeacher.forward post-processing (model.py:771-790) for one clip."""
    cap = detok(result["predictions"][0].tolist())                                      # :771
    n = min(len(cap.split(" ")), len(result["logits_dict"]))                            # :772
    dist = torch.from_numpy(np.array(result["logits_dict"][:n]))                        # :776  [n, beams, V]
    word_tokens = result["predictions"][0, 1:n + 1]                                     # :780
    word_tokens = word_tokens[:, None, None].expand(-1, num_beams, -1)                  # :781
    indices = torch.gather(dist, dim=2, index=word_tokens).squeeze().argmax(dim=1)      # :784
    indices_expanded = indices[:, None, None].expand(-1, -1, dist.shape[-1])            # :786
    res = torch.gather(dist, dim=1, index=indices_expanded).squeeze()[None, ...]        # :787
    out = dict(result)
    out["output"] = res
    out["cap"] = cap
    return out
