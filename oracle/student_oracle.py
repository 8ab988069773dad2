"""CPU oracle for the STUDENT DECODER -- TEST INFRASTRUCTURE ONLY (only tests/, __graft_entry__.smoke() and bench.py may import it).

Restates the decoder half of ``StudentCandidateV1`` (/root/reference/src/models/model.py:50-187), SURVEY 8(f) rank 3:
``forward_decoder`` (:135-154) and ``greedy_decode`` (:156-187) on a given ``memory`` tensor.

PARITY STATUS: **pinned to the reference's own code and by PyTorch itself.**  oracle/make_reference_golden.py imports
/root/reference/src/models/model.py unmodified and runs its ``StudentCandidateV1.forward_decoder / greedy_decode / beam_search``,
``PositionalEncoding``, the masking helpers and ``DistillationTrainer.training_step`` (loss + ``loss.backward()`` gradients) on
seeded inputs; tests/test_reference_golden.py replays the frozen outputs (tests/golden/ref_student.npz, ref_training_step.npz)
through this file.  Besides, the reference builds this decoder from stock ``torch.nn`` modules
(``nn.TransformerDecoderLayer(d_model, nhead, dim_feedforward, dropout, batch_first=True)`` stacked by
``nn.TransformerDecoder`` :73-76, ``nn.Embedding`` :78, ``nn.Linear`` :80) and this file instantiates exactly those
modules with the same arguments -- the layer arithmetic below IS the reference's arithmetic, not a recollection of it.
What is restated by hand (and checked against the stock modules in tests/test_student.py) is only the glue the
reference writes itself: ``PositionalEncoding`` (:320-340), ``create_padding_mask`` / ``create_casual_mask``
(src/utils/masking.py:4-26), the ``(embed + pe) / sqrt(d)`` scaling order (:144-148) and the greedy loop with its
"stop only when EVERY row emits SEP in the same step" rule (:184).

Not covered (cannot be restated here): the TinyViT image encoder (``timm.create_model(...)``, :38-40) -- timm is not
installed and its weights/architecture are not in /root/reference; the decoder therefore takes ``memory``
([B, F, d_model], the spatial mean of the last TinyViT stage per frame, :128) as an input.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch
import torch.nn as nn


@dataclass
class StudentConfig:
    """cfg['MODEL']['StudentCandidateV1'] (/root/reference/config.py:76-84) + tokenizer ids (model.py:82-83)."""
    d_model: int = 576
    n_head: int = 8
    d_ffn: int = 1024
    dropout: float = 0.3
    num_decoder_layers: int = 2
    vocab_length: int = 30522
    cls_token_id: int = 101
    sep_token_id: int = 102
    pad_token_id: int = 0      # create_padding_mask default (masking.py:4)
    max_len: int = 500         # PositionalEncoding max_len (model.py:324)


def positional_encoding(d_model: int, max_len: int = 500) -> torch.Tensor:
    """model.py:324-335: the sinusoidal table [max_len, d_model]."""
    pe = torch.zeros(max_len, d_model)
    position = torch.arange(0, max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * -(torch.log(torch.tensor(10000.0)) / d_model))
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


class StudentDecoderOracle(nn.Module):
    """The decoder-side modules of StudentCandidateV1 under the reference's attribute names (model.py:73-80)."""

    def __init__(self, cfg: StudentConfig):
        super().__init__()
        self.cfg = cfg
        self.decoder_layer = nn.TransformerDecoderLayer(d_model=cfg.d_model, nhead=cfg.n_head, dim_feedforward=cfg.d_ffn,
                                                        dropout=cfg.dropout, batch_first=True)            # :73-75
        self.decoder = nn.TransformerDecoder(self.decoder_layer, cfg.num_decoder_layers)                  # :76
        self.embed = nn.Embedding(cfg.vocab_length, cfg.d_model)                                          # :78
        self.linear = nn.Linear(cfg.d_model, cfg.vocab_length)                                            # :80
        self.register_buffer("pe", positional_encoding(cfg.d_model, cfg.max_len).unsqueeze(0))            # :334-335
        self.eval()

    @torch.no_grad()
    def forward_decoder(self, y: torch.Tensor, memory: torch.Tensor) -> torch.Tensor:
        """model.py:135-154: y int64 [B, L], memory [B, F, d] -> logits [B, L, V]."""
        pad_mask = y == self.cfg.pad_token_id                                                             # :140, masking.py:14
        tgt_mask = torch.triu(torch.ones(y.shape[1], y.shape[1]), diagonal=1).bool()                      # :142, masking.py:26
        tgt_embed = self.embed(y)                                                                         # :144
        tgt_embed = tgt_embed + self.pe[:, : y.size(1)]                                                   # :146, :338-339
        tgt_embed = tgt_embed / torch.sqrt(torch.tensor(self.embed.embedding_dim))                        # :148
        out = self.decoder(tgt=tgt_embed, memory=memory, tgt_mask=tgt_mask, tgt_key_padding_mask=pad_mask,
                           tgt_is_causal=True)                                                            # :150-151
        return self.linear(out)                                                                           # :152

    @torch.no_grad()
    def greedy_decode_from_memory(self, memory: torch.Tensor, max_len: int = 10) -> torch.Tensor:
        """model.py:156-187 after ``forward_image_enc``: full re-decode of the growing sequence every step."""
        batch_size = memory.size(0)
        tgt = torch.tensor([self.cfg.cls_token_id] * batch_size, dtype=torch.long).unsqueeze(1)           # :171
        for _ in range(max_len):                                                                          # :173
            output = self.forward_decoder(tgt, memory)                                                    # :176
            output = torch.argmax(output, dim=-1)                                                         # :178
            last_tokens = output[:, -1].unsqueeze(-1)                                                     # :180
            tgt = torch.cat((tgt, last_tokens), dim=1)                                                    # :182
            if torch.all(last_tokens.squeeze(-1) == self.cfg.sep_token_id):                               # :184
                break
        return tgt


def forward_decoder_train(m: StudentDecoderOracle, y: torch.Tensor, memory: torch.Tensor) -> torch.Tensor:
    """``forward_decoder`` (model.py:135-154) with autograd enabled -- the same statements; dropout is off (module in eval
    mode): the CUDA training step does not apply dropout either."""
    cfg = m.cfg
    pad_mask = y == cfg.pad_token_id
    tgt_mask = torch.triu(torch.ones(y.shape[1], y.shape[1]), diagonal=1).bool()
    tgt_embed = m.embed(y)
    tgt_embed = tgt_embed + m.pe[:, : y.size(1)]
    tgt_embed = tgt_embed / torch.sqrt(torch.tensor(m.embed.embedding_dim))
    out = m.decoder(tgt=tgt_embed, memory=memory, tgt_mask=tgt_mask, tgt_key_padding_mask=pad_mask, tgt_is_causal=True)
    return m.linear(out)


def distillation_loss(student_logits: torch.Tensor, teacher_logits: torch.Tensor, y: torch.Tensor, temperature: float = 1.0):
    """DistillationTrainer.training_step's active losses (model.py:919-935, :983): (kl + ce, kl, ce)."""
    kl = nn.KLDivLoss(reduction="batchmean")((student_logits / temperature).log_softmax(dim=-1),
                                             (teacher_logits / temperature).softmax(dim=-1)) * temperature ** 2   # :922-928
    y_target = y[:, 1:].reshape(-1)                                                                                 # :931-932
    y_pred = student_logits[:, :-1].reshape(-1, student_logits.size(-1))                                            # :933-934
    ce = nn.CrossEntropyLoss(ignore_index=0)(y_pred, y_target)                                                      # :935
    return kl + ce, kl, ce                                                                                          # :983


def distillation_step(m: StudentDecoderOracle, y: torch.Tensor, memory: torch.Tensor, teacher_logits: torch.Tensor,
                      temperature: float = 1.0):
    """One training step's loss and gradients by torch.autograd on the stock modules: ({'loss','kl','ce'}, grads by
    reference state-dict key, d loss / d memory)."""
    for p in m.parameters():
        p.grad = None
    memory = memory.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        logits = forward_decoder_train(m, y, memory)
        loss, kl, ce = distillation_loss(logits, teacher_logits, y, temperature)
        loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters() if not k.startswith("decoder_layer.") and p.grad is not None}
    return {"loss": loss.item(), "kl": kl.item(), "ce": ce.item(), "logits": logits.detach()}, grads, memory.grad.detach().clone()


def beam_search_from_memory(forward_decoder, memory: torch.Tensor, cls_token_id: int, max_len: int = 10, k: int = 3) -> torch.Tensor:
    """StudentCandidateV1.beam_search (model.py:189-316) after ``forward_image_enc``, statement by statement, with
    ``forward_decoder`` passed in (the oracle module's, or a scripted stand-in in the host-logic tests).  Quirks kept:
    no end-of-sequence handling (the EOS block is commented out in the reference), every beam's whole sequence is
    re-decoded every step, k*k candidates are ranked with a full sort, the result is the highest-scoring of the k
    final beams and always has ``max_len`` tokens."""
    import torch.nn.functional as F
    batch_size = memory.size(0)
    tgt = torch.full((batch_size, 1), cls_token_id, dtype=torch.long)                                     # :200
    sequences = tgt.unsqueeze(1).expand(-1, k, -1)                                                        # :205
    all_candidates = torch.empty(batch_size, k * k, 3)                                                    # :207
    decoder_output = forward_decoder(tgt, memory)                                                         # :223
    log_probs = F.log_softmax(decoder_output[:, -1, :], dim=-1)                                           # :225
    scores, top_indices = log_probs.topk(k, dim=-1)                                                       # :227
    sequences = torch.cat([sequences, top_indices.unsqueeze(-1)], dim=-1)                                 # :229
    for step in range(2, max_len):                                                                        # :232
        for i in range(k):                                                                                # :234
            tgt = sequences[:, i]                                                                         # :236
            decoder_output = forward_decoder(tgt, memory)                                                 # :239
            log_probs = F.log_softmax(decoder_output[:, -1, :], dim=-1)                                   # :241
            top_scores, top_indices = log_probs.topk(k, dim=-1)                                           # :243
            local_scores = scores[:, i].unsqueeze(-1) + top_scores                                        # :246
            offset = i * k                                                                                # :248
            all_candidates[:, offset:offset + k, 0] = local_scores                                        # :249
            all_candidates[:, offset:offset + k, 1] = i                                                   # :250
            all_candidates[:, offset:offset + k, 2] = top_indices                                         # :251
        scores_to_sort = all_candidates[:, :, 0].view(batch_size, -1)                                     # :254
        _, sorted_indices = scores_to_sort.sort(dim=1, descending=True)                                   # :256
        topk_indices = sorted_indices[:, :k]                                                              # :258
        new_sequences = torch.zeros(batch_size, k, step + 1, dtype=torch.long)                            # :260
        for b in range(batch_size):                                                                       # :262
            for idx in range(k):
                global_idx = topk_indices[b, idx]                                                         # :266
                beam_idx = all_candidates[b, global_idx, 1].long()                                        # :269
                token_idx = all_candidates[b, global_idx, 2].long()                                       # :270
                new_sequences[b, idx, :-1] = sequences[b, beam_idx, :]                                    # :273
                new_sequences[b, idx, -1] = token_idx                                                     # :274
                scores[b, idx] = all_candidates[b, global_idx, 0]                                         # :277
        sequences = new_sequences                                                                         # :296
    return sequences[torch.arange(batch_size), scores.argmax(dim=-1)]                                     # :315


def init_student(cfg: StudentConfig, seed: int = 0, logit_gain: float = 1.0) -> StudentDecoderOracle:
    """Seeded random initialisation (PyTorch defaults for every module, then biases / LayerNorm affines perturbed so that
    no term of the arithmetic is silently zero or one)."""
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    m = StudentDecoderOracle(cfg)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if name.endswith("bias"):
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
        m.linear.weight.mul_(logit_gain)
    return m


def state_dict_of(m: StudentDecoderOracle) -> Dict[str, torch.Tensor]:
    """Parameters under the reference's StudentCandidateV1 key names (decoder.layers.N.*, embed.weight, linear.*)."""
    return {k: v.detach().clone() for k, v in m.state_dict().items() if not k.startswith("decoder_layer.") and k != "pe"}
