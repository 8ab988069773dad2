"""Fixtures produced by the REFERENCE'S OWN in-tree code (oracle/make_reference_golden.py ran /root/reference/src/models/model.py
unmodified, under import stubs for the packages this image lacks) replayed through

  * the oracle's restatements (CPU): this is what pins `oracle/search_oracle.py`, the glue of `oracle/git_oracle.py` and
    `oracle/student_oracle.py` to the reference instead of to themselves;
  * the package's host-side mirror of the reference interface (CPU);
  * the CUDA path (`-m gpu`): device search, encoder + text head + caption, student decoder.

The fixtures travel with the repo; /root/reference is never read here."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import git_oracle as go  # noqa: E402
from oracle import make_reference_golden as mk  # noqa: E402  (case table + seeds only; its generator needs /root/reference)
from oracle import search_oracle as so  # noqa: E402
from oracle import student_oracle as sto  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def search_case(z, name):
    clips, nb, pn, keep, max_steps, vocab, eos, sos = (int(v) for v in z[f"{name}.spec"])
    return dict(clips=clips, nb=nb, pn=pn, keep=keep, max_steps=max_steps, vocab=vocab, eos=eos, sos=sos,
                lp=float(z[f"{name}.length_penalty"]), logits=torch.from_numpy(z[f"{name}.logits"]),
                decoded=torch.from_numpy(z[f"{name}.decoded"]), logprobs=torch.from_numpy(z[f"{name}.logprobs"]),
                steps_run=int(z[f"{name}.steps_run"]), last_input_ids=torch.from_numpy(z[f"{name}.last_input_ids"]))


SEARCH_NAMES = [c[0] for c in mk.SEARCH_CASES]


def scripted_step(logits, calls):
    def step(input_ids):
        calls.append(input_ids.clone())
        return logits[len(calls) - 1].clone()
    return step


# ------------------------------------------------------------------------------------------ search (model.py:479-678)
@pytest.mark.parametrize("name", SEARCH_NAMES)
def test_oracle_search_equals_the_reference_search(name):
    c = search_case(load("ref_search.npz"), name)
    calls = []
    dec, lp, _ = so.search(torch.full((c["clips"], 1), c["sos"], dtype=torch.long), scripted_step(c["logits"], calls),
                           eos_index=c["eos"], max_steps=c["max_steps"], beam_size=c["nb"], length_penalty=c["lp"],
                           per_node_beam_size=c["pn"], num_keep_best=c["keep"], save_logits=False)
    assert torch.equal(dec.reshape(c["clips"], c["keep"], c["max_steps"]), c["decoded"])
    assert torch.equal(lp, c["logprobs"])                       # same float operations in the same order
    assert len(calls) == c["steps_run"]                         # `if all(done): break` (model.py:640) at the same step
    assert torch.equal(calls[-1], c["last_input_ids"])          # the beams' re-ordered histories (model.py:615-621)


@pytest.mark.parametrize("name", SEARCH_NAMES)
def test_package_host_search_equals_the_reference_search(name):
    gm = importlib.import_module("real-time-video-captioning_b200.model")
    c = search_case(load("ref_search.npz"), name)
    calls = []
    dec = gm.GeneratorWithBeamSearchV2(c["eos"], c["max_steps"], c["nb"], c["lp"], per_node_beam_size=c["pn"])
    out, lp, _ = dec.search(torch.full((c["clips"], 1), c["sos"], dtype=torch.long), scripted_step(c["logits"], calls),
                            num_keep_best=c["keep"])
    assert torch.equal(out.reshape(c["clips"], c["keep"], c["max_steps"]), c["decoded"])
    assert torch.allclose(lp, c["logprobs"], atol=1e-6, rtol=0)
    assert len(calls) == c["steps_run"] and torch.equal(calls[-1], c["last_input_ids"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", SEARCH_NAMES)
def test_device_search_equals_the_reference_search(name):
    g = importlib.import_module("real-time-video-captioning_b200")
    eng = importlib.import_module("real-time-video-captioning_b200.engine")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    c = search_case(load("ref_search.npz"), name)
    ld = (c["vocab"] + 7) // 8 * 8
    logits = torch.zeros(c["max_steps"] - 1, c["clips"] * c["nb"], ld)
    logits[..., : c["vocab"]] = c["logits"]
    sp = g.SearchConfig(beam_size=c["nb"], max_steps=c["max_steps"], length_penalty=c["lp"], per_node_beam_size=c["pn"],
                        num_keep_best=c["keep"])
    tok, lp = eng.op_search(logits.cuda(), c["vocab"], c["clips"], c["sos"], c["eos"], sp)
    assert torch.equal(tok.cpu().long(), c["decoded"]), (tok.cpu()[0], c["decoded"][0])
    assert torch.allclose(lp.cpu(), c["logprobs"], atol=2e-5, rtol=1e-5)


# ------------------------------------------------------------------------------------------ GIT glue (model.py:372-463)
def glue_case(z, name):
    n_frames, n_embed, _, _, max_steps, beam = (int(v) for v in z[f"{name}.spec"])
    cfg, sd, frames = mk.small_git(n_frames, n_embed)
    return cfg, sd, frames, max_steps, beam, {k.split(".", 1)[1]: z[k] for k in z.files if k.startswith(name + ".")}


@pytest.mark.parametrize("name", ["f2", "zip_truncation"])
def test_oracle_glue_equals_the_reference_glue(name):
    """forward_one_custom (frame features + temporal embeddings in zip order, frames beyond the embedding list dropped,
    concat, text head, stacked hidden states) and infer (start tokens, search, result dict), executed by the reference around
    the oracle's layers, against the oracle's own restatement of that glue."""
    cfg, sd, frames, max_steps, beam, want = glue_case(load("ref_git_glue.npz"), name)
    tokens = torch.from_numpy(want["tokens"])
    with torch.no_grad():
        logits, vf, hidden = go.forward_one_custom(sd, cfg, frames, tokens)
        res = so.infer(sd, cfg, vf, beam_size=beam, max_steps=max_steps, save_logits=False)
    S = mk.SUB
    assert list(vf.shape) == list(want["visual_features_shape"]) and list(hidden.shape) == list(want["hidden_states_shape"])
    assert vf.shape[1] == min(frames.shape[0], cfg.num_image_with_embedding) * cfg.tokens_per_frame   # zip() truncation, model.py:380
    assert torch.allclose(vf[:, ::S, ::S], torch.from_numpy(want["visual_features"]), atol=1e-5, rtol=1e-5)
    assert torch.allclose(hidden[:, ::S, ::S], torch.from_numpy(want["hidden_states"]), atol=2e-4, rtol=1e-4)
    assert torch.allclose(logits[..., ::S], torch.from_numpy(want["logits"]), atol=2e-4, rtol=1e-4)
    assert torch.equal(res["predictions"], torch.from_numpy(want["predictions"]))
    assert torch.allclose(res["logprobs"], torch.from_numpy(want["logprobs"]), atol=1e-5, rtol=1e-5)
    assert bool(want["infer_visual_features_is_input"]) and sorted(res.keys()) == list(want["result_keys"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["f2", "zip_truncation"])
def test_cuda_path_equals_the_reference_glue(name):
    g = importlib.import_module("real-time-video-captioning_b200")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    cfg, sd, frames, max_steps, beam, want = glue_case(load("ref_git_glue.npz"), name)
    eng = g.Engine(g.make_config({"num_image_with_embedding": cfg.num_image_with_embedding}, cfg.sos_index, cfg.eos_index,
                                 layers=cfg.num_layers), 0)
    eng.load_state_dict(sd)
    S = mk.SUB
    tokens = torch.from_numpy(want["tokens"])
    logits, vf, hidden = eng.forward_logits(frames[None].cuda(), tokens.cuda())      # forward_one_custom, batched form
    vf, logits, hidden = vf.cpu(), logits.cpu(), hidden.cpu()
    ref_vf = torch.from_numpy(want["visual_features"])
    assert list(vf.shape) == list(want["visual_features_shape"])          # 3 frames, 2 temporal embeddings -> 2 frames
    assert ((vf[:, ::S, ::S] - ref_vf).norm() / ref_vf.norm()).item() < 2e-2
    ref_logits = torch.from_numpy(want["logits"])
    err = (logits[..., ::S] - ref_logits).abs()
    assert err.max().item() < 0.15 * ref_logits.std().item() and err.mean().item() < 0.03 * ref_logits.std().item()
    ref_h = torch.from_numpy(want["hidden_states"])
    h = hidden.reshape(want["hidden_states_shape"].tolist())              # [1, 3, Nv+L, H] -> the reference's [3, Nv+L, H]
    assert ((h[:, ::S, ::S] - ref_h).norm() / ref_h.norm()).item() < 3e-2
    sp = g.SearchConfig(beam_size=beam, max_steps=max_steps, length_penalty=cfg.length_penalty, per_node_beam_size=cfg.per_node_beam_size)
    tok, lp, _ = eng.caption(frames[None].cuda(), sp)
    assert torch.equal(tok[:, 0].cpu().long(), torch.from_numpy(want["predictions"]))   # tied head: 9 sigma copy margin
    assert torch.allclose(lp.cpu(), torch.from_numpy(want["logprobs"]), rtol=2e-2, atol=5e-3)


# ------------------------------------------------------------------------------------------ student (model.py:50-341)
def student_case():
    z = load("ref_student.npz")
    d_model, n_head, d_ffn, layers, vocab = (int(v) for v in z["spec"])
    cfg = sto.StudentConfig(d_model=d_model, n_head=n_head, d_ffn=d_ffn, dropout=0.0, num_decoder_layers=layers, vocab_length=vocab,
                            cls_token_id=vocab - 2, sep_token_id=vocab - 1)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    return z, cfg, sd


def test_oracle_student_equals_the_reference_student():
    z, cfg, sd = student_case()
    m = sto.StudentDecoderOracle(cfg)
    missing = m.load_state_dict(sd, strict=False)
    # the reference's own parameter names (the fixture leaves out the template `decoder_layer` TransformerDecoder deep-copies)
    assert all(k == "pe" or k.startswith("decoder_layer.") for k in missing.missing_keys) and not missing.unexpected_keys
    memory, y = torch.from_numpy(z["memory"]), torch.from_numpy(z["y"])
    assert torch.equal(m.pe[0, :16], torch.from_numpy(z["pe"]))                       # PositionalEncoding, model.py:320-341
    assert torch.equal(y == cfg.pad_token_id, torch.from_numpy(z["pad_mask"]))        # create_padding_mask
    assert torch.equal(torch.triu(torch.ones(7, 7), diagonal=1).bool(), torch.from_numpy(z["causal_mask"]))
    want = torch.from_numpy(z["logits"])
    got = m.forward_decoder(y, memory)
    valid = (y != cfg.pad_token_id)                     # rows behind a padded tail are NaN-free but unspecified upstream
    assert torch.allclose(got[valid], want[valid], atol=1e-5, rtol=1e-5)
    assert torch.equal(m.greedy_decode_from_memory(memory, max_len=9), torch.from_numpy(z["greedy"]))
    assert torch.equal(sto.beam_search_from_memory(m.forward_decoder, memory, cfg.cls_token_id, max_len=8, k=3), torch.from_numpy(z["beam"]))


def test_package_student_host_logic_equals_the_reference_student():
    """The package's batched k x k beam loop, driven by a decoder that carries the REFERENCE's weights, picks the sequences the
    reference's own loop picked; the reference's parameter names load into the package's student without renaming."""
    g = importlib.import_module("real-time-video-captioning_b200")
    st = importlib.import_module("real-time-video-captioning_b200.student")
    z, cfg, sd = student_case()
    assert torch.equal(st.PositionalEncoding(d_model=cfg.d_model).pe[0, :16], torch.from_numpy(z["pe"]))
    m = sto.StudentDecoderOracle(cfg)
    m.load_state_dict(sd, strict=False)
    memory = torch.from_numpy(z["memory"])
    s = g.StudentCandidateV1(None, cfg.d_model, cfg.n_head, cfg.d_ffn, 0.0, cfg.num_decoder_layers, cfg.vocab_length, cfg.cls_token_id,
                             cfg.sep_token_id)
    res = s.load_state_dict(sd)
    assert not res.unexpected_keys and set(res.missing_keys) <= {"pos_enc.pe"}, (res.missing_keys, res.unexpected_keys)
    out = s.beam_search_from_memory(memory, max_len=8, k=3, forward_decoder=m.forward_decoder)
    assert torch.equal(out, torch.from_numpy(z["beam"]))


@pytest.mark.gpu
def test_cuda_student_decoder_equals_the_reference_student():
    """The CUDA student decoder carrying the reference's weights (its own state-dict names) against the logits / greedy tokens
    the reference's forward_decoder / greedy_decode produced."""
    g = importlib.import_module("real-time-video-captioning_b200")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    z, cfg, sd = student_case()
    s = g.StudentCandidateV1(None, cfg.d_model, cfg.n_head, cfg.d_ffn, 0.0, cfg.num_decoder_layers, cfg.vocab_length, cfg.cls_token_id,
                             cfg.sep_token_id)
    s.load_state_dict(sd)
    s = s.to("cuda")
    memory, y = torch.from_numpy(z["memory"]), torch.from_numpy(z["y"])
    want = torch.from_numpy(z["logits"])
    out = s.forward_decoder(y, memory).cpu()
    valid = y != cfg.pad_token_id
    sigma = want[valid].std().item()
    err = (out[valid] - want[valid]).abs()
    assert err.max().item() < 0.15 * sigma and err.mean().item() < 0.03 * sigma, (err.max().item() / sigma, err.mean().item() / sigma)
    ref_greedy = torch.from_numpy(z["greedy"])
    got = s.greedy_decode_from_memory(memory, max_len=9).cpu()
    m = sto.StudentDecoderOracle(cfg)
    m.load_state_dict(sd, strict=False)
    lo = m.forward_decoder(ref_greedy[:, :-1], memory)      # the reference sequence's own margins
    top2 = lo.topk(2, dim=-1).values
    clear = (top2[..., 0] - top2[..., 1]) >= 0.15 * lo.std().item()
    if clear.all():
        assert torch.equal(got, ref_greedy)
    else:  # up to the first near-tie the runs must agree
        first = int((~clear).float().argmax(dim=1).min().item())
        assert torch.equal(got[:, : first + 1], ref_greedy[:, : first + 1])


# ------------------------------------------------------------------------------------------ distillation step (model.py:880-1004)
def training_case():
    z = load("ref_training_step.npz")
    d_model, n_head, d_ffn, layers, vocab = (int(v) for v in z["spec"])
    cfg = sto.StudentConfig(d_model=d_model, n_head=n_head, d_ffn=d_ffn, dropout=0.0, num_decoder_layers=layers, vocab_length=vocab,
                            cls_token_id=vocab - 2, sep_token_id=vocab - 1)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad.")}
    return z, cfg, sd, grads


def test_oracle_distillation_step_equals_the_reference_training_step():
    """KL(batchmean) * T^2 + CE(ignore_index=0) and the gradients of `loss.backward()`, as the reference's own
    DistillationTrainer.training_step computed them, against the oracle's restatement (the checker of the CUDA training step)."""
    z, cfg, sd, want = training_case()
    m = sto.StudentDecoderOracle(cfg)
    m.load_state_dict(sd, strict=False)
    y, memory, teacher = torch.from_numpy(z["y"]), torch.from_numpy(z["memory"]), torch.from_numpy(z["teacher_logits"])
    info, grads, _ = sto.distillation_step(m, y, memory, teacher, 1.0)
    assert abs(info["loss"] - float(z["loss"])) < 1e-5 and abs(info["kl"] - float(z["kl"])) < 1e-5 and abs(info["ce"] - float(z["ce"])) < 1e-5
    assert set(grads) == set(want)
    for k, ref in want.items():
        assert torch.allclose(grads[k], ref, atol=1e-6, rtol=1e-4), k


@pytest.mark.gpu
def test_cuda_distillation_step_equals_the_reference_training_step():
    g = importlib.import_module("real-time-video-captioning_b200")
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    z, cfg, sd, want = training_case()
    s = g.StudentCandidateV1(None, cfg.d_model, cfg.n_head, cfg.d_ffn, 0.0, cfg.num_decoder_layers, cfg.vocab_length, cfg.cls_token_id,
                             cfg.sep_token_id)
    s.load_state_dict(sd)
    s = s.to("cuda")
    s.enable_training(lr=1e-4)
    y, memory, teacher = torch.from_numpy(z["y"]), torch.from_numpy(z["memory"]), torch.from_numpy(z["teacher_logits"])
    out = s.distillation_step(y, memory, teacher, temperature=1.0, apply=False)
    kl, ce = out["kl"].item(), out["ce"].item()
    assert abs(kl - float(z["kl"])) < 2e-2 * float(z["kl"]) and abs(ce - float(z["ce"])) < 2e-2 * float(z["ce"]), (kl, ce)
    grads = {k: v.cpu() for k, v in s.gradients().items()}
    assert set(grads) == set(want)
    num = sum((grads[k].double() - want[k].double()).pow(2).sum().item() for k in want)
    den = sum(want[k].double().pow(2).sum().item() for k in want)
    assert (num / den) ** 0.5 < 3e-2, (num / den) ** 0.5


# ------------------------------------------------------------------------------------------ teacher wrapper (model.py:747-793)
def test_oracle_teacher_wrapper_equals_the_reference_teacher():
    """GenerativeImageTextTeacher.forward (per-clip loop, caption, n = min(words, saved steps), beam pick by the word's logit,
    'output') and forward_output_logits, executed by the reference, against the oracle's restatement (teacher_postprocess is the
    checker of the package's batched teacher in tests/test_gpu_baseline_geometry.py)."""
    z = load("ref_teacher.npz")
    n_frames, layers, max_steps, beam = (int(v) for v in z["spec"])
    cfg = go.GitConfig(num_image_with_embedding=n_frames, num_layers=layers, tie_output=False)
    sd = go.init_state_dict(cfg, seed=mk.GLUE_SEEDS["weights"] + 1, temporal_std=0.02, perturb=True)
    x = torch.randn(2, n_frames, 3, 224, 224, generator=torch.Generator().manual_seed(mk.GLUE_SEEDS["frames"] + 1))
    y = torch.from_numpy(z["y"])
    S = mk.SUB
    assert list(z["result_keys"]) == ["cap", "logits_dict", "logprobs", "output", "predictions", "visual_features"]
    for i in range(int(z["n_clips"])):
        with torch.no_grad():
            vf = go.encode_clip(sd, cfg, x[i])
            res = so.infer(sd, cfg, vf, beam_size=beam, max_steps=max_steps)
            post = so.teacher_postprocess(res, mk.detok, num_beams=beam)
            logits, vf2, hidden = go.forward_one_custom(sd, cfg, x[i], y[i:i + 1])
        assert torch.equal(post["predictions"], torch.from_numpy(z[f"clip{i}.predictions"]))
        assert torch.allclose(post["logprobs"], torch.from_numpy(z[f"clip{i}.logprobs"]), atol=1e-5)
        assert post["cap"] == str(z[f"clip{i}.cap"]) and len(post["cap"]) > 0
        assert len(res["logits_dict"]) == int(z[f"clip{i}.n_saved_steps"])
        assert list(post["output"].shape) == list(z[f"clip{i}.output_shape"])
        assert torch.allclose(post["output"][..., ::S], torch.from_numpy(z[f"clip{i}.output"]), atol=2e-4, rtol=1e-4)
        assert torch.allclose(logits[..., ::S], torch.from_numpy(z[f"clip{i}.fol_logits"]), atol=2e-4, rtol=1e-4)
        assert torch.allclose(vf2[:, ::S, ::S], torch.from_numpy(z[f"clip{i}.fol_visual_features"]), atol=1e-5, rtol=1e-5)
        assert torch.allclose(hidden[:, ::S, ::S], torch.from_numpy(z[f"clip{i}.fol_hidden_states"]), atol=2e-4, rtol=1e-4)


# ------------------------------------------------------------------------------------------ get_git_model (model.py:681-718)
@pytest.mark.parametrize("label,param", [("default", {"num_image_with_embedding": 6}),
                                         ("large", {"image_encoder_type": "CLIPViT_L_14", "test_crop_size": 224,
                                                    "visual_feature_size": 1024, "num_image_with_embedding": 6})])
def test_hyper_parameters_equal_what_the_reference_get_git_model_passes(label, param):
    """What the reference's OWN get_git_model handed to its constructors (recorded by the fixture generator) sizes every kernel:
    the oracle's GitConfig and the package's get_git_model must carry the same numbers."""
    import json
    want = json.loads(str(load("ref_get_git_model.npz")["json"]))[label]
    td, dec, mod, enc = want["text_decoder"], want["decoder"], want["model"], want["image_encoder"]
    cfg = go.GitConfig.from_param(param)
    assert (cfg.image_encoder_type, cfg.resolution) == (enc["name"], enc["input_resolution"])
    assert (cfg.visual_feature_size, cfg.vocab_size, cfg.hidden_size, cfg.num_layers, cfg.attention_heads, cfg.feedforward_size,
            cfg.max_caption_length) == (td["visual_feature_size"], td["vocab_size"], td["hidden_size"], td["num_layers"],
                                        td["attention_heads"], td["feedforward_size"], td["max_caption_length"])
    assert (cfg.beam_size, cfg.max_steps, cfg.length_penalty, cfg.eos_index, cfg.sos_index) == \
        (dec["beam_size"], dec["max_steps"], dec["length_penalty"], dec["eos_index"], mod["sos_index"])
    assert cfg.num_image_with_embedding == mod["num_image_with_embedding"] == mod["n_temporal_embeddings"]
    gm = importlib.import_module("real-time-video-captioning_b200.model")
    m = gm.get_git_model(gm.SyntheticTokenizer(), param)
    d = m.decoder
    assert (d._eos_index, d.max_steps, d.beam_size, d.length_penalty, d.repetition_penalty, d.temperature) == \
        (dec["eos_index"], dec["max_steps"], dec["beam_size"], dec["length_penalty"], dec["repetition_penalty"], dec["temperature"])
    assert (m.sos_index, m.eos_index) == (mod["sos_index"], mod["eos_index"])
    assert len(m.img_temperal_embedding) == mod["n_temporal_embeddings"]
    assert list(m.img_temperal_embedding[0].shape) == mod["temporal_embedding_shape"]
    sd = m.state_dict()
    assert sd["textual.embedding.words.weight"].shape == (td["vocab_size"], td["hidden_size"])
    assert sd["textual.visual_projection.0.weight"].shape == (td["hidden_size"], td["visual_feature_size"])
    assert f"textual.transformer.encoder.layer.{td['num_layers'] - 1}.output.dense.weight" in sd
    assert f"textual.transformer.encoder.layer.{td['num_layers']}.output.dense.weight" not in sd
    assert sd["textual.transformer.encoder.layer.0.intermediate.dense.weight"].shape == (td["feedforward_size"], td["hidden_size"])
