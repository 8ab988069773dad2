"""CPU oracle for the GIT captioning hot path -- TEST INFRASTRUCTURE ONLY.

Plain PyTorch fp32 restatement of the arithmetic the reference reaches through
``/root/reference/src/models/model.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this module; the product
package never does (it fails loudly when the CUDA library is missing).

PARITY STATUS: two halves.
(1) The GLUE the reference wrote itself -- ``forward_one_custom`` (frame features + temporal embeddings in zip order,
truncation, concat, hidden-state stacking) and ``infer`` -- is PINNED TO THE REFERENCE'S OWN CODE: oracle/make_reference_golden.py
imports /root/reference/src/models/model.py unmodified and runs those methods around THIS file's layers;
tests/test_reference_golden.py replays the frozen outputs (tests/golden/ref_git_glue.npz) through ``forward_one_custom`` /
``search_oracle.infer`` here.
(2) The LAYER ARITHMETIC is *unpinned by the reference*: it lives in the un-vendored, un-pinned dependency
``generativeimage2text`` (microsoft/GenerativeImage2Text, ``/root/reference/requirements.txt:19``), absent from this image, and
the reference's only test (``/root/reference/tests/test_metrics.py:7-22``) pins no number on this path, so there are no golden
vectors to inherit.  What anchors the layers instead:

  * the in-tree call sites and hyper-parameters: ``model.py:681-718`` (get_git_model),
    ``:371-424`` (forward_one_custom), ``:426-462`` (infer), ``:479-678`` (search, restated in
    ``oracle/search_oracle.py``), ``:747-793`` (teacher wrapper);
  * the published GIT / CLIP / BERT layer definitions (restated below, each citing the upstream
    module it follows -- recalled);
  * an independent implementation available in this container: ``transformers.GitForCausalLM``;
    ``tests/test_oracle_vs_hf.py`` loads the same weights into both and checks visual features and
    text-row logits to ~1e-4 (fp32).

State dicts use the upstream key names so that a real GIT checkpoint
(``ckpt['model']``, model.py:736-738) loads unchanged.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

VIT_CONFIGS = {
    # upstream get_image_encoder(): CLIP VisionTransformer(width, layers, heads=width//64, patch)
    "CLIPViT_B_16": dict(width=768, layers=12, heads=12, patch=16),
    "CLIPViT_L_14": dict(width=1024, layers=24, heads=16, patch=14),
}


@dataclass
class GitConfig:
    """Hyper-parameters fixed by get_git_model (model.py:681-708) plus the `param` dict keys."""

    image_encoder_type: str = "CLIPViT_B_16"      # param.get('image_encoder_type', 'CLIPViT_B_16')  model.py:683
    resolution: int = 224                          # param.get('test_crop_size', 224)                 model.py:684
    visual_feature_size: int = 768                 # param.get('visual_feature_size', 768)            model.py:688
    num_image_with_embedding: int = 6              # param.get('num_image_with_embedding')            model.py:368
    vocab_size: int = 30522                        # model.py:689
    hidden_size: int = 768                         # model.py:690
    num_layers: int = 6                            # model.py:691
    attention_heads: int = 12                      # model.py:692
    feedforward_size: int = 3072                   # model.py:693
    max_caption_length: int = 1024                 # model.py:694
    sos_index: int = 101                           # tokenizer.cls_token_id (bert-base-uncased)      model.py:363
    eos_index: int = 102                           # tokenizer.sep_token_id                           model.py:364
    beam_size: int = 4                             # model.py:706
    max_steps: int = 15                            # model.py:704
    length_penalty: float = 0.6                    # model.py:707
    per_node_beam_size: int = 2                    # upstream GeneratorWithBeamSearch default
    embedding_ln_eps: float = 1e-8                 # upstream WordAndPositionalEmbedding LayerNorm eps (recalled)
    bert_ln_eps: float = 1e-12                     # upstream BertConfig.layer_norm_eps
    vit_ln_eps: float = 1e-5                       # CLIP LayerNorm default
    proj_ln_eps: float = 1e-5                      # nn.LayerNorm default in the 'linearLn' projection
    tie_output: bool = True                        # upstream ties textual.output.weight to embedding.words.weight

    @property
    def vit(self) -> dict:
        return VIT_CONFIGS[self.image_encoder_type]

    @property
    def tokens_per_frame(self) -> int:
        return (self.resolution // self.vit["patch"]) ** 2 + 1

    @classmethod
    def from_param(cls, param: dict, **kw) -> "GitConfig":
        return cls(
            image_encoder_type=param.get("image_encoder_type", "CLIPViT_B_16"),
            resolution=param.get("test_crop_size", 224),
            visual_feature_size=param.get("visual_feature_size", 768),
            num_image_with_embedding=param.get("num_image_with_embedding") or 0,
            **kw,
        )


# --------------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------------
def init_state_dict(cfg: GitConfig, seed: int = 0, temporal_std: float = 0.02, perturb: bool = True,
                    dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded random-init weights under the upstream key names.

    Distributions follow the upstream constructors where that is cheap (CLIP: cls/pos ~ N(0, width^-0.5),
    BERT: N(0, 0.02)); ``perturb`` additionally randomises every bias / LayerNorm affine so that a
    kernel which drops a bias or swaps gamma/beta cannot pass parity.  ``temporal_std > 0`` exercises the
    temporal embeddings (upstream initialises them to zero).
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=dtype) * std

    def ln(prefix, n):
        sd[prefix + ".weight"] = 1.0 + (randn(n, std=0.1) if perturb else torch.zeros(n))
        sd[prefix + ".bias"] = randn(n, std=0.1) if perturb else torch.zeros(n)

    def bias(n, std=0.02):
        return randn(n, std=std) if perturb else torch.zeros(n)

    v = cfg.vit
    W, P = v["width"], v["patch"]
    T = cfg.tokens_per_frame
    pre = "image_encoder."
    sd[pre + "conv1.weight"] = randn(W, 3, P, P, std=(3 * P * P) ** -0.5)
    sd[pre + "class_embedding"] = randn(W, std=W ** -0.5)
    sd[pre + "positional_embedding"] = randn(T, W, std=W ** -0.5)
    ln(pre + "ln_pre", W)
    for i in range(v["layers"]):
        b = f"{pre}transformer.resblocks.{i}."
        ln(b + "ln_1", W)
        sd[b + "attn.in_proj_weight"] = randn(3 * W, W, std=W ** -0.5)
        sd[b + "attn.in_proj_bias"] = bias(3 * W)
        sd[b + "attn.out_proj.weight"] = randn(W, W, std=W ** -0.5)
        sd[b + "attn.out_proj.bias"] = bias(W)
        ln(b + "ln_2", W)
        sd[b + "mlp.c_fc.weight"] = randn(4 * W, W, std=W ** -0.5)
        sd[b + "mlp.c_fc.bias"] = bias(4 * W)
        sd[b + "mlp.c_proj.weight"] = randn(W, 4 * W, std=(4 * W) ** -0.5)
        sd[b + "mlp.c_proj.bias"] = bias(W)
    ln(pre + "ln_post", W)

    H, Fd, V = cfg.hidden_size, cfg.feedforward_size, cfg.vocab_size
    t = "textual."
    sd[t + "visual_projection.0.weight"] = randn(H, cfg.visual_feature_size, std=cfg.visual_feature_size ** -0.5)
    sd[t + "visual_projection.0.bias"] = bias(H)
    ln(t + "visual_projection.1", H)
    sd[t + "embedding.words.weight"] = randn(V, H, std=0.02)
    sd[t + "embedding.positions.weight"] = randn(cfg.max_caption_length, H, std=0.02)
    ln(t + "embedding.layer_norm", H)
    for i in range(cfg.num_layers):
        b = f"{t}transformer.encoder.layer.{i}."
        for nm in ("query", "key", "value"):
            sd[b + f"attention.self.{nm}.weight"] = randn(H, H, std=0.02)
            sd[b + f"attention.self.{nm}.bias"] = bias(H)
        sd[b + "attention.output.dense.weight"] = randn(H, H, std=0.02)
        sd[b + "attention.output.dense.bias"] = bias(H)
        ln(b + "attention.output.LayerNorm", H)
        sd[b + "intermediate.dense.weight"] = randn(Fd, H, std=0.02)
        sd[b + "intermediate.dense.bias"] = bias(Fd)
        sd[b + "output.dense.weight"] = randn(H, Fd, std=0.02)
        sd[b + "output.dense.bias"] = bias(H)
        ln(b + "output.LayerNorm", H)
    if cfg.tie_output:
        sd[t + "output.weight"] = sd[t + "embedding.words.weight"]
    else:
        sd[t + "output.weight"] = randn(V, H, std=0.02)
    sd[t + "output.bias"] = bias(V)
    for i in range(cfg.num_image_with_embedding):
        sd[f"img_temperal_embedding.{i}"] = randn(1, 1, cfg.visual_feature_size, std=temporal_std) if temporal_std > 0 \
            else torch.zeros(1, 1, cfg.visual_feature_size)
    return sd


# --------------------------------------------------------------------------------------------
# CLIP vision tower  (upstream CLIP VisionTransformer with output_grid=True, grid_after_ln=True)
# --------------------------------------------------------------------------------------------
def _ln(x, sd, prefix, eps):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def vit_resblock(x: torch.Tensor, sd, prefix: str, heads: int, eps: float) -> torch.Tensor:
    """ResidualAttentionBlock: x + MHA(ln_1 x); x + c_proj(QuickGELU(c_fc(ln_2 x))).  x: [N, T, W]."""
    N, T, W = x.shape
    hd = W // heads
    h = _ln(x, sd, prefix + "ln_1", eps)
    qkv = F.linear(h, sd[prefix + "attn.in_proj_weight"], sd[prefix + "attn.in_proj_bias"])
    q, k, v = qkv.split(W, dim=-1)
    q = q.view(N, T, heads, hd).transpose(1, 2) * (hd ** -0.5)  # nn.MultiheadAttention scales q
    k = k.view(N, T, heads, hd).transpose(1, 2)
    v = v.view(N, T, heads, hd).transpose(1, 2)
    p = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    a = (p @ v).transpose(1, 2).reshape(N, T, W)
    x = x + F.linear(a, sd[prefix + "attn.out_proj.weight"], sd[prefix + "attn.out_proj.bias"])
    h = _ln(x, sd, prefix + "ln_2", eps)
    h = F.linear(h, sd[prefix + "mlp.c_fc.weight"], sd[prefix + "mlp.c_fc.bias"])
    h = h * torch.sigmoid(1.702 * h)  # QuickGELU
    x = x + F.linear(h, sd[prefix + "mlp.c_proj.weight"], sd[prefix + "mlp.c_proj.bias"])
    return x


def vit_forward(sd, cfg: GitConfig, images: torch.Tensor, return_blocks: bool = False):
    """images [N, 3, R, R] -> [N, T, W]: all tokens after ln_post, no projection (model.py:378)."""
    v = cfg.vit
    pre = "image_encoder."
    x = F.conv2d(images, sd[pre + "conv1.weight"], stride=v["patch"])
    N, W = x.shape[0], x.shape[1]
    x = x.reshape(N, W, -1).permute(0, 2, 1)
    cls = sd[pre + "class_embedding"].to(x.dtype).expand(N, 1, W)
    x = torch.cat([cls, x], dim=1) + sd[pre + "positional_embedding"]
    x = _ln(x, sd, pre + "ln_pre", cfg.vit_ln_eps)
    blocks = []
    for i in range(v["layers"]):
        x = vit_resblock(x, sd, f"{pre}transformer.resblocks.{i}.", v["heads"], cfg.vit_ln_eps)
        if return_blocks:
            blocks.append(x)
    x = _ln(x, sd, pre + "ln_post", cfg.vit_ln_eps)
    return (x, blocks) if return_blocks else x


def encode_clip(sd, cfg: GitConfig, frames: torch.Tensor, per_frame_calls: bool = False) -> torch.Tensor:
    """frames [F, 3, R, R] of ONE clip -> visual_features [1, F'*T, Dv] (model.py:378-382).

    ``zip`` with the temporal-embedding list silently drops frames beyond num_image_with_embedding
    (model.py:380).  per_frame_calls=True reproduces CaptioningModel.forward_one's
    one-ViT-call-per-frame cost structure (used by the CPU baseline).
    """
    if per_frame_calls:
        feats = [vit_forward(sd, cfg, f[None])[0] for f in frames]
    else:
        feats = list(vit_forward(sd, cfg, frames))
    if cfg.num_image_with_embedding:
        embs = [sd[f"img_temperal_embedding.{i}"] for i in range(cfg.num_image_with_embedding)]
        feats = [f[None] + e for f, e in zip(feats, embs)]
    else:
        feats = [f[None] for f in feats]
    return torch.cat(feats, dim=1)


# --------------------------------------------------------------------------------------------
# text decoder (upstream TransformerDecoderTextualHead / BertEncoderAsDecoder)
# --------------------------------------------------------------------------------------------
def project_visual(sd, cfg, visual_features):
    h = F.linear(visual_features, sd["textual.visual_projection.0.weight"], sd["textual.visual_projection.0.bias"])
    return _ln(h, sd, "textual.visual_projection.1", cfg.proj_ln_eps)


def embed_text(sd, cfg, tokens):
    L = tokens.shape[1]
    e = sd["textual.embedding.words.weight"][tokens] + sd["textual.embedding.positions.weight"][:L][None]
    return _ln(e, sd, "textual.embedding.layer_norm", cfg.embedding_ln_eps)


def prefix_lm_mask(nv: int, nt: int, dtype=torch.float32) -> torch.Tensor:
    """[[0, -inf], [0, causal]] additive mask over [visual; text] (upstream BertEncoderAsDecoder)."""
    m = torch.zeros(nv + nt, nv + nt, dtype=dtype)
    m[:nv, nv:] = float("-inf")
    m[nv:, nv:] = torch.triu(torch.full((nt, nt), float("-inf"), dtype=dtype), diagonal=1)
    return m


def bert_layer(x_q, x_kv, mask, sd, prefix, heads, eps):
    """Post-LN BERT layer.  Queries come from x_q [B, Lq, H]; keys/values are projected from x_kv
    [B, Lk, H] (x_kv = cat(history, x_q) on the cached path: the upstream cache stores layer INPUTS
    and re-projects K and V every step -- SURVEY Appendix A.6)."""
    B, Lq, H = x_q.shape
    hd = H // heads
    p = prefix + "attention.self."
    q = F.linear(x_q, sd[p + "query.weight"], sd[p + "query.bias"]).view(B, Lq, heads, hd).transpose(1, 2)
    k = F.linear(x_kv, sd[p + "key.weight"], sd[p + "key.bias"]).view(B, -1, heads, hd).transpose(1, 2)
    v = F.linear(x_kv, sd[p + "value.weight"], sd[p + "value.bias"]).view(B, -1, heads, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / math.sqrt(hd)
    if mask is not None:
        s = s + mask
    a = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, Lq, H)
    a = F.linear(a, sd[prefix + "attention.output.dense.weight"], sd[prefix + "attention.output.dense.bias"])
    a = _ln(a + x_q, sd, prefix + "attention.output.LayerNorm", eps)
    f = F.linear(a, sd[prefix + "intermediate.dense.weight"], sd[prefix + "intermediate.dense.bias"])
    f = F.gelu(f)  # exact erf GELU (hidden_act="gelu")
    f = F.linear(f, sd[prefix + "output.dense.weight"], sd[prefix + "output.dense.bias"])
    return _ln(f + a, sd, prefix + "output.LayerNorm", eps)


def textual_forward(sd, cfg: GitConfig, visual_features: torch.Tensor, tokens: torch.Tensor,
                    history: Optional[List[torch.Tensor]] = None):
    """``self.textual(visual_features, caption_tokens, ...)`` (model.py:412-418).

    Returns (logits over the text rows [B, L or 1, V], list of 7 hidden-state tensors).  With
    ``history`` (the per-layer inputs of all earlier positions) only the LAST token is run through the
    layers, as upstream does when ``encoder_history_states`` is given.
    """
    B, L = tokens.shape
    text = embed_text(sd, cfg, tokens)
    if history is None:
        vis = project_visual(sd, cfg, visual_features)
        nv = vis.shape[1]
        x = torch.cat([vis, text], dim=1)
        mask = prefix_lm_mask(nv, L, x.dtype)
        hidden = [x]
        for i in range(cfg.num_layers):
            x = bert_layer(x, x, mask, sd, f"textual.transformer.encoder.layer.{i}.", cfg.attention_heads, cfg.bert_ln_eps)
            hidden.append(x)
        out_rows = x[:, nv:]
    else:
        x = text[:, -1:]
        hidden = [x]
        for i in range(cfg.num_layers):
            kv_in = torch.cat([history[i], x], dim=1)
            x = bert_layer(x, kv_in, None, sd, f"textual.transformer.encoder.layer.{i}.", cfg.attention_heads, cfg.bert_ln_eps)
            hidden.append(x)
        out_rows = x
    logits = F.linear(out_rows, sd["textual.output.weight"], sd["textual.output.bias"])
    return logits, hidden


def forward_one_custom(sd, cfg: GitConfig, frames: torch.Tensor, caption_tokens: torch.Tensor):
    """model.py:371-424 for one clip: (logits [1,L,V], visual_features [1,Nv,Dv], hidden_states [7,Nv+L,H])."""
    vf = encode_clip(sd, cfg, frames)
    logits, hidden = textual_forward(sd, cfg, vf, caption_tokens)
    return logits, vf, torch.stack([h.squeeze(0) for h in hidden], dim=0)


class DecodingState:
    """``CaptioningModel.decoding_step`` with ``use_history_for_infer=True`` (model.py:366, :439-445)."""

    def __init__(self, sd, cfg: GitConfig, visual_features: torch.Tensor, reorder_cache: bool = False):
        self.sd, self.cfg, self.vf = sd, cfg, visual_features
        self.prev_encoded_layers: Optional[List[torch.Tensor]] = None
        self.reorder_cache = reorder_cache  # False == reference (model.py:623-634 is commented out)
        self.nv = None

    def reorder(self, beam_idx: torch.Tensor):
        if self.reorder_cache and self.prev_encoded_layers is not None:
            self.prev_encoded_layers = [h[beam_idx] for h in self.prev_encoded_layers]

    def __call__(self, partial_captions: torch.Tensor) -> torch.Tensor:
        vf = self.vf
        B = vf.shape[0]
        rows = partial_captions.shape[0]
        if rows > B:  # repeat visual features for every beam (upstream decoding_step)
            nb = rows // B
            vf = vf.unsqueeze(1).expand(B, nb, *vf.shape[1:]).reshape(rows, *vf.shape[1:])
        logits, hidden = textual_forward(self.sd, self.cfg, vf, partial_captions, self.prev_encoded_layers)
        if self.prev_encoded_layers is None:
            self.prev_encoded_layers = hidden
        else:
            self.prev_encoded_layers = [torch.cat((p, c), dim=1) for p, c in zip(self.prev_encoded_layers, hidden)]
        return logits[:, -1, :].float()
