"""Golden vectors produced by the REFERENCE'S OWN in-tree code, executed in this container (TEST INFRASTRUCTURE).

    python -m oracle.make_reference_golden        # needs /root/reference; writes tests/golden/ref_*.npz

`import src.models.model` fails here because lightning, timm, nltk, pycocoevalcap and generativeimage2text are not installed
(DESIGN.md section 2).  The code the reference itself wrote for the hot path does not need them, though: this script puts
import stubs for the missing packages into sys.modules, imports /root/reference/src/models/model.py UNMODIFIED, and runs

  * ``GeneratorWithBeamSearchV2.search``            (model.py:479-678: the whole caption search loop, a9)
  * ``GenerativeImageTextModel.forward_one_custom`` (model.py:372-428: frame features + temporal embeddings, concat, a4/a6)
  * ``GenerativeImageTextModel.infer``              (model.py:430-463: start tokens, search call, result dict, a7)
  * ``get_git_model`` / ``GenerativeImageTextModel.__init__`` (model.py:681-718, :350-369: the hyper-parameters it passes, a1/a2)
  * ``GenerativeImageTextTeacher.forward / forward_output_logits`` (model.py:747-793: per-clip loop, caption, 'output', a11)
  * ``StudentCandidateV1.forward_decoder / greedy_decode / beam_search`` (model.py:135-316, rank f3)
  * ``DistillationTrainer.training_step``           (model.py:880-1004: KL + CE loss; gradients by its ``loss.backward()``)
  * ``PositionalEncoding`` (model.py:320-341), ``create_padding_mask`` / ``create_casual_mask`` (src/utils/masking.py)

on seeded inputs, and freezes inputs + outputs as small fixtures.  `tests/test_reference_golden.py` replays them through the
oracle (CPU) and through the CUDA path (GPU): that is what pins the oracle's restatement of these functions to the reference.

What the stubs stand in for -- code that is NOT in /root/reference and therefore cannot be executed:
  * ``generativeimage2text.layers.decoder.BeamHypotheses`` / ``top_k_top_p_filtering``: bound to the oracle's restatement of the
    published implementation (oracle/search_oracle.py); ``GeneratorWithBeamSearch`` / ``CaptioningModel``: attribute holders
    with upstream's constructor arguments (``CaptioningModel`` creates ``img_temperal_embedding`` as upstream does and takes
    ``decoding_step`` from the oracle: upstream code);
  * the image encoder / text head handed to ``GenerativeImageTextModel`` are the ORACLE's layers (their arithmetic is pinned
    separately against transformers.GitForCausalLM): the fixtures pin the reference's GLUE around them;
  * ``timm.create_model``: the student's TinyViT is replaced by a module that returns the given feature map (out of scope,
    DESIGN.md section 8); lightning / nltk / pycocoevalcap / evaluate / wandb: empty modules (never called on these paths).
Nothing in `tests/`, `smoke()` or `bench.py` reads /root/reference at run time: only this generator does."""
import importlib
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

from . import git_oracle as go
from . import search_oracle as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


class _AnyModule(types.ModuleType):
    """A module whose every attribute is a harmless placeholder (class / callable that is never used on the tested paths)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = f"{self.__name__}.{name}"
        if sub in sys.modules:
            return sys.modules[sub]
        ph = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})
        setattr(self, name, ph)
        return ph


def _stub(name, **attrs):
    m = _AnyModule(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []  # lets "import a.b" resolve through sys.modules
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _GeneratorWithBeamSearch:
    """Upstream's attribute holder (generativeimage2text/layers/decoder.py GeneratorWithBeamSearch.__init__ arguments)."""

    def __init__(self, eos_index, max_steps, beam_size, per_node_beam_size=2, length_penalty=1.0, repetition_penalty=1.0,
                 temperature=1.0):
        self._eos_index = eos_index
        self.max_steps = max_steps
        self.beam_size = beam_size
        self.per_node_beam_size = per_node_beam_size
        self.length_penalty = length_penalty
        self.repetition_penalty = repetition_penalty
        self.temperature = temperature


class _CaptioningModel(nn.Module):
    """Upstream's base class reduced to what GenerativeImageTextModel's in-tree methods touch."""

    def __init__(self, visual, textual, sos_index=1, eos_index=2, decoder=None, loss_type=None, context_not_share_embedding=False,
                 scst=False, tokenizer=None, scst_temperature=1., use_history_for_infer=False, pooling_images=None,
                 num_image_with_embedding=0):
        super().__init__()
        self.image_encoder = visual
        self.textual = textual
        self.sos_index = sos_index
        self.eos_index = eos_index
        self.decoder = decoder
        self.tokenizer = tokenizer
        self.use_history_for_infer = use_history_for_infer
        self.pooling_images = pooling_images
        self.context_not_share_embedding = context_not_share_embedding
        self.num_image_with_embedding = num_image_with_embedding
        if num_image_with_embedding:
            self.img_temperal_embedding = nn.ParameterList(
                nn.Parameter(torch.zeros(1, 1, textual.visual_feature_size)) for _ in range(num_image_with_embedding))

    def forward(self, batch):
        # upstream CaptioningModel.forward in eval mode (not in /root/reference): features through the reference's own glue
        # (forward_one_custom's first half, a dummy caption feeds its text-head call), then the reference's infer with the
        # default search parameters
        dummy = torch.full((1, 1), self.sos_index, dtype=torch.long)
        _, visual_features, _ = self.forward_one_custom({"image": batch["image"], "caption_tokens": dummy})
        return self.infer(batch, visual_features, None, search_param=None)

    def decoding_step(self, visual_features, visual_features_valid, bi_valid_mask_caption, partial_captions):
        # upstream code, not in /root/reference: the oracle's restatement (hidden-state history, visual features
        # repeated per beam).  infer() resets self.prev_encoded_layers = None before every search (model.py:445).
        if self.prev_encoded_layers is None:
            self._state = go.DecodingState(self.textual.sd, self.textual.cfg, visual_features)
            self.prev_encoded_layers = "held by the oracle's DecodingState"
        return self._state(partial_captions)


def import_reference():
    """The reference's model module, imported unmodified under stubs for the packages this image lacks."""
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not present: the fixtures can only be regenerated where the reference is")
    from transformers import BertTokenizer  # noqa: F401  (model.py:19; resolved before the stubs confuse transformers' package probes)
    _stub("lightning", LightningModule=nn.Module)
    _stub("timm", create_model=lambda *a, **k: nn.Identity())
    for n in ("pycocotools", "pycocotools.coco", "pycocoevalcap", "pycocoevalcap.eval", "nltk", "nltk.translate",
              "nltk.translate.bleu_score", "nltk.translate.meteor_score", "evaluate", "wandb"):
        _stub(n)
    _stub("generativeimage2text")
    _stub("generativeimage2text.layers")
    _stub("generativeimage2text.layers.decoder", CaptioningModel=_CaptioningModel, BeamHypotheses=so.BeamHypotheses,
          top_k_top_p_filtering=so.top_k_top_p_filtering, GeneratorWithBeamSearch=_GeneratorWithBeamSearch,
          convert2valid=lambda *a, **k: None)
    _stub("generativeimage2text.model")
    _stub("generativeimage2text.torch_common")
    _stub("generativeimage2text.tsv_io")
    sys.path.insert(0, REF)
    try:
        return importlib.import_module("src.models.model")
    finally:
        sys.path.remove(REF)


# ------------------------------------------------------------------ search (a9)
SEARCH_CASES = [  # name, clips, beam, per-node, keep, max_steps, vocab, eos boost, length penalty
    ("greedy", 3, 1, 2, 1, 8, 61, 0.0, 1.0),
    ("greedy_eos", 4, 1, 2, 1, 12, 61, 2.5, 1.0),
    ("beam4", 3, 4, 2, 1, 10, 97, 0.0, 1.0),
    ("beam4_eos_keep3", 4, 4, 2, 3, 12, 97, 3.0, 1.0),
    ("beam3_keep2_lp", 2, 3, 2, 2, 9, 53, 2.0, 0.6),
    ("beam4_max20", 2, 4, 2, 1, 20, 211, 1.5, 1.0),
]


def search_logits(name, rows, steps, vocab, eos, boost):
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    x = torch.randn(steps, rows, vocab, generator=g) * 2.0
    x[..., eos] += boost
    return x


def run_search(ref, name, clips, nb, pn, keep, max_steps, vocab, boost, lp):
    eos, sos = vocab - 1, vocab - 2
    logits = search_logits(name, clips * nb, max_steps - 1, vocab, eos, boost)
    dec = ref.GeneratorWithBeamSearchV2(eos, max_steps, nb, per_node_beam_size=pn, length_penalty=lp)
    calls = []

    def step(input_ids):
        calls.append(input_ids.clone())
        return logits[len(calls) - 1].clone()

    start = torch.full((clips, 1), sos, dtype=torch.long)
    decoded, logprobs, saved = dec.search(start, step, num_keep_best=keep)
    return dict(logits=logits.numpy(), decoded=decoded.reshape(clips, keep, max_steps).numpy(), logprobs=logprobs.numpy(),
                steps_run=np.int64(len(calls)), last_input_ids=calls[-1].numpy(),
                spec=np.array([clips, nb, pn, keep, max_steps, vocab, eos, sos], dtype=np.int64), length_penalty=np.float64(lp))


# ------------------------------------------------------------------ GIT glue (a4 / a6 / a7)
class _OracleImageEncoder(nn.Module):
    def __init__(self, sd, cfg):
        super().__init__()
        self.sd, self.cfg = sd, cfg

    def forward(self, images):  # [F, 3, H, W] -> [F, T, D]: the list comprehension at model.py:380 zips over dim 0
        return go.vit_forward(self.sd, self.cfg, images)


class _OracleTextHead(nn.Module):
    def __init__(self, sd, cfg):
        super().__init__()
        self.sd, self.cfg = sd, cfg
        self.visual_feature_size = cfg.visual_feature_size

    def forward(self, visual_features, caption_tokens, hidden_valid_mask=None, bi_valid_mask_caption=None):
        return go.textual_forward(self.sd, self.cfg, visual_features, caption_tokens)  # (logits, [7 hidden states])


GLUE_SEEDS = dict(weights=31, frames=32)
SUB = 8  # the fixture keeps every 8th token row / feature column (inputs are regenerated from the seeds)


def small_git(n_frames, n_embed):
    """GIT-base layers with a 2-layer text decoder; weights and frames are functions of GLUE_SEEDS only."""
    cfg = go.GitConfig(num_image_with_embedding=n_embed, num_layers=2)
    sd = go.init_state_dict(cfg, seed=GLUE_SEEDS["weights"], temporal_std=0.02, perturb=True)
    frames = torch.randn(n_frames, 3, 224, 224, generator=torch.Generator().manual_seed(GLUE_SEEDS["frames"]))
    return cfg, sd, frames


def run_glue(ref, n_frames, n_embed):
    cfg, sd, frames = small_git(n_frames, n_embed)

    class Tok:
        cls_token_id, sep_token_id = cfg.sos_index, cfg.eos_index

    dec = ref.GeneratorWithBeamSearchV2(cfg.eos_index, 6, 2, per_node_beam_size=cfg.per_node_beam_size, length_penalty=cfg.length_penalty)
    m = ref.GenerativeImageTextModel(_OracleImageEncoder(sd, cfg), _OracleTextHead(sd, cfg), dec, Tok, {"num_image_with_embedding": n_embed})
    with torch.no_grad():
        for i, p in enumerate(m.img_temperal_embedding):
            p.copy_(sd[f"img_temperal_embedding.{i}"])
    tokens = torch.tensor([[cfg.sos_index, 2023, 2003, 1037, 3231]], dtype=torch.long)
    batch = {"image": [f[None] for f in frames], "caption_tokens": tokens}
    logits, vf, hidden = m.forward_one_custom(batch)
    res = m.infer(batch, vf, None, search_param={"num_keep_best": 1})
    return dict(tokens=tokens.numpy(), logits=logits[..., ::SUB].numpy(), visual_features=vf[:, ::SUB, ::SUB].numpy(),
                visual_features_shape=np.array(vf.shape), hidden_states=hidden[:, ::SUB, ::SUB].numpy(), hidden_states_shape=np.array(hidden.shape),
                predictions=res["predictions"].numpy(), logprobs=res["logprobs"].numpy(),
                infer_visual_features_is_input=np.bool_(res["visual_features"] is vf), result_keys=np.array(sorted(res.keys())),
                spec=np.array([n_frames, n_embed, 2, 2, 6, 2], dtype=np.int64))


# ------------------------------------------------------------------ get_git_model (a1 / a2, model.py:681-718, :350-369)
def run_get_git_model(ref):
    """The reference's own get_git_model with recording stand-ins for the two upstream constructors: which hyper-parameters the
    REFERENCE passes (they size every kernel of the path), for the default and for the shipped GIT-large parameter dict."""
    import json
    rec = {}

    class Enc(nn.Module):
        pass

    class Head(nn.Module):
        def __init__(self, **kw):
            super().__init__()
            rec["text_decoder"] = dict(kw)
            self.visual_feature_size = kw["visual_feature_size"]

    def get_image_encoder(name, **kw):
        rec["image_encoder"] = dict(name=name, **kw)
        return Enc()

    ref.get_image_encoder, ref.TransformerDecoderTextualHead = get_image_encoder, Head

    class Tok:
        cls_token_id, sep_token_id = 101, 102

    out = {}
    for label, param in (("default", {"num_image_with_embedding": 6}),
                         ("large", {"image_encoder_type": "CLIPViT_L_14", "test_crop_size": 224, "visual_feature_size": 1024,
                                    "num_image_with_embedding": 6})):
        rec.clear()
        m = ref.get_git_model(Tok, param)
        d = m.decoder
        out[label] = dict(image_encoder=rec["image_encoder"], text_decoder=rec["text_decoder"],
                          decoder=dict(eos_index=d._eos_index, max_steps=d.max_steps, beam_size=d.beam_size, length_penalty=d.length_penalty,
                                       per_node_beam_size=d.per_node_beam_size, repetition_penalty=d.repetition_penalty,
                                       temperature=d.temperature),
                          model=dict(sos_index=m.sos_index, eos_index=m.eos_index, use_history_for_infer=m.use_history_for_infer,
                                     num_image_with_embedding=m.num_image_with_embedding,
                                     n_temporal_embeddings=len(m.img_temperal_embedding),
                                     temporal_embedding_shape=list(m.img_temperal_embedding[0].shape)))
    return dict(json=np.array(json.dumps(out, sort_keys=True)))


# ------------------------------------------------------------------ teacher wrapper (a11, model.py:747-793)
def detok(ids, skip_special_tokens=True, sos=101, eos=102):
    """Stand-in for BertTokenizer.decode (bert-base-uncased is not in this image): one pseudo-word per non-special id."""
    return " ".join(f"w{i}" for i in ids if not (skip_special_tokens and i in (0, sos, eos)))


def run_teacher(ref):
    n_frames = 2
    cfg = go.GitConfig(num_image_with_embedding=n_frames, num_layers=2, tie_output=False)
    sd = go.init_state_dict(cfg, seed=GLUE_SEEDS["weights"] + 1, temporal_std=0.02, perturb=True)
    x = torch.randn(2, n_frames, 3, 224, 224, generator=torch.Generator().manual_seed(GLUE_SEEDS["frames"] + 1))

    class Tok:
        cls_token_id, sep_token_id = cfg.sos_index, cfg.eos_index
        decode = staticmethod(detok)

    dec = ref.GeneratorWithBeamSearchV2(cfg.eos_index, 7, 4, per_node_beam_size=cfg.per_node_beam_size, length_penalty=cfg.length_penalty)
    m = ref.GenerativeImageTextModel(_OracleImageEncoder(sd, cfg), _OracleTextHead(sd, cfg), dec, Tok, {"num_image_with_embedding": n_frames})
    with torch.no_grad():
        for i, p in enumerate(m.img_temperal_embedding):
            p.copy_(sd[f"img_temperal_embedding.{i}"])
    m.eval()

    class Self:
        model, tokenizer = m, Tok

    out = ref.GenerativeImageTextTeacher.forward(Self(), x)
    y = torch.tensor([[cfg.sos_index, 2023, 2003, 1037], [cfg.sos_index, 1996, 4937, 102]], dtype=torch.long)
    with torch.no_grad():
        logits, vfs, hiddens = ref.GenerativeImageTextTeacher.forward_output_logits(Self(), x, y)
    d = dict(spec=np.array([n_frames, 2, 7, 4], dtype=np.int64), y=y.numpy(), n_clips=np.int64(len(out)))
    for i, o in enumerate(out):
        d[f"clip{i}.predictions"] = o["predictions"].numpy()
        d[f"clip{i}.logprobs"] = o["logprobs"].numpy()
        d[f"clip{i}.cap"] = np.array(o["cap"])
        d[f"clip{i}.output"] = o["output"][..., ::SUB].numpy()
        d[f"clip{i}.output_shape"] = np.array(o["output"].shape)
        d[f"clip{i}.n_saved_steps"] = np.int64(len(o["logits_dict"]))
        d[f"clip{i}.fol_logits"] = logits[i][..., ::SUB].numpy()
        d[f"clip{i}.fol_visual_features"] = vfs[i][:, ::SUB, ::SUB].numpy()
        d[f"clip{i}.fol_hidden_states"] = hiddens[i][:, ::SUB, ::SUB].numpy()
    d["result_keys"] = np.array(sorted(out[0].keys()))
    return d


# ------------------------------------------------------------------ student (f3)
def run_student(ref):
    torch.manual_seed(41)
    d_model, n_head, d_ffn, layers, vocab = 64, 4, 96, 2, 101
    m = ref.StudentCandidateV1("unused", d_model, n_head, d_ffn, 0.0, layers, vocab, cls_token_id=vocab - 2, sep_token_id=vocab - 1)
    m.eval()
    g = torch.Generator().manual_seed(42)
    memory = torch.randn(3, 6, d_model, generator=g)
    y = torch.randint(1, vocab - 2, (3, 7), generator=g)
    y[:, 0] = vocab - 2
    y[1, 5:] = 0
    y[2, 3:] = 0  # padded tails (create_padding_mask)

    class FeatureMap(nn.Module):  # TinyViT stand-in: forward_image_enc averages the last map spatially -> memory
        def forward(self, x):
            return [memory.reshape(18, d_model, 1, 1)]

    m.image_encoder = FeatureMap()
    src = torch.zeros(3, 6, 3, 8, 8)
    with torch.no_grad():
        logits = m.forward_decoder(y, memory)
        greedy = m.greedy_decode(src, max_len=9)
        beam = m.beam_search(src, max_len=8, k=3)
    sd = {k: v.numpy() for k, v in m.state_dict().items() if k.startswith(("decoder.", "embed.", "linear."))}
    pe = ref.PositionalEncoding(d_model=d_model).pe[0, :16].numpy()
    from src.utils import masking  # imported with the reference module
    return dict(memory=memory.numpy(), y=y.numpy(), logits=logits.numpy(), greedy=greedy.numpy(), beam=beam.numpy(), pe=pe,
                pad_mask=masking.create_padding_mask(y).numpy(), causal_mask=masking.create_casual_mask(7).numpy(),
                spec=np.array([d_model, n_head, d_ffn, layers, vocab], dtype=np.int64), **{"sd." + k: v for k, v in sd.items()})


# ------------------------------------------------------------------ distillation step (model.py:880-1004, configs[4])
def run_training_step(ref):
    """``DistillationTrainer.training_step`` executed on a duck-typed trainer: the constructor wires Lightning, hooks into the
    teacher's layers and a log file (out of scope); the step itself needs only the attributes set below.  Student = the
    reference's StudentCandidateV1 (TinyViT replaced by fixed feature maps), teacher = given logits.  Returns the loss the
    reference computed AND the gradients its ``loss.backward()`` leaves on the decoder-side parameters."""
    torch.manual_seed(51)
    d_model, n_head, d_ffn, layers, vocab, B, Fr, Lc = 64, 4, 96, 2, 101, 3, 6, 7
    m = ref.StudentCandidateV1("unused", d_model, n_head, d_ffn, 0.0, layers, vocab, cls_token_id=vocab - 2, sep_token_id=vocab - 1)
    m.train()  # Lightning trains in train mode; dropout is 0.0 here so the step is deterministic
    g = torch.Generator().manual_seed(52)
    fmaps = [torch.randn(B * Fr, c, 2, 2, generator=g) for c in (8, 16, 32, d_model)]
    memory = torch.mean(fmaps[-1], dim=[2, 3]).view(B, Fr, -1)

    class FeatureMaps(nn.Module):
        def forward(self, x):
            return fmaps

    m.image_encoder = FeatureMaps()
    y = torch.randint(1, vocab - 2, (B, Lc), generator=g)
    y[:, 0] = vocab - 2
    y[1, 5:] = 0
    teacher_logits = torch.randn(B, Lc, vocab, generator=g) * 2.0

    class Teacher:
        def eval(self):
            return self

        def forward_output_logits(self, x, yy):
            return [t[None] for t in teacher_logits], None, None

    class Self:
        pass

    me = Self()
    me.teacher, me.student = Teacher(), m
    me.fmap_distill_loss = nn.MSELoss()
    me.kl_div_loss = nn.KLDivLoss(reduction="batchmean")   # model.py:819
    me.ce_loss = nn.CrossEntropyLoss(ignore_index=0)       # model.py:821
    me.student_decoder_activations = {0: [torch.zeros(1, 1)]}
    me.teacher_encoder_activations = {i: [torch.randn(3, B * Fr, 1024, generator=g)] for i in range(4)}
    me.teacher_decoder_activations = {}
    logged = {}
    me.log = lambda name, value, **kw: logged.__setitem__(name, float(value))
    batch = {"frames": torch.zeros(B, Fr, 3, 8, 8), "caption": y, "caption-id": None, "vid-id": None}
    loss = ref.DistillationTrainer.training_step(me, batch, 0)
    loss.backward()
    grads = {k: p.grad.numpy() for k, p in m.named_parameters()
             if p.grad is not None and k.startswith(("decoder.", "embed.", "linear."))}
    sd = {k: v.detach().numpy() for k, v in m.state_dict().items() if k.startswith(("decoder.", "embed.", "linear."))}
    return dict(memory=memory.numpy(), y=y.numpy(), teacher_logits=teacher_logits.numpy(), loss=np.float64(loss.item()),
                kl=np.float64(logged["train_kl_loss"]), ce=np.float64(logged["ce_loss"]),
                spec=np.array([d_model, n_head, d_ffn, layers, vocab], dtype=np.int64),
                **{"sd." + k: v for k, v in sd.items()}, **{"grad." + k: v for k, v in grads.items()})


def main():
    ref = import_reference()
    os.makedirs(OUT, exist_ok=True)
    search = {}
    for case in SEARCH_CASES:
        for k, v in run_search(ref, *case).items():
            search[f"{case[0]}.{k}"] = v
    np.savez_compressed(os.path.join(OUT, "ref_search.npz"), **search)
    glue = {}
    for name, n_frames, n_embed in (("f2", 2, 2), ("zip_truncation", 3, 2)):
        for k, v in run_glue(ref, n_frames, n_embed).items():
            glue[f"{name}.{k}"] = v
    np.savez_compressed(os.path.join(OUT, "ref_git_glue.npz"), **glue)
    np.savez_compressed(os.path.join(OUT, "ref_teacher.npz"), **run_teacher(ref))
    np.savez_compressed(os.path.join(OUT, "ref_get_git_model.npz"), **run_get_git_model(ref))
    np.savez_compressed(os.path.join(OUT, "ref_student.npz"), **run_student(ref))
    np.savez_compressed(os.path.join(OUT, "ref_training_step.npz"), **run_training_step(ref))
    for f in ("ref_search.npz", "ref_git_glue.npz", "ref_teacher.npz", "ref_get_git_model.npz", "ref_student.npz", "ref_training_step.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
