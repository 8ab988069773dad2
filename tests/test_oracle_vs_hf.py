"""Pin the oracle's layer arithmetic against an independent implementation that exists in this image:
transformers.GitForCausalLM (HF port of GIT).  Same weights are loaded into both (upstream -> HF key
map from SURVEY.md section 8b); visual features and text-row logits must agree to fp32 round-off.
The reference itself cannot be imported here, so this is the strongest external
anchor available for the un-vendored generativeimage2text arithmetic."""
import pytest
import torch

from oracle import git_oracle as go

transformers = pytest.importorskip("transformers")


def to_hf_state_dict(sd, cfg):
    out = {}
    v = cfg.vit
    W = v["width"]
    vm = "git.image_encoder.vision_model."
    out[vm + "embeddings.patch_embedding.weight"] = sd["image_encoder.conv1.weight"]
    out[vm + "embeddings.class_embedding"] = sd["image_encoder.class_embedding"]
    out[vm + "embeddings.position_embedding.weight"] = sd["image_encoder.positional_embedding"]
    for a, b in (("ln_pre", "pre_layrnorm"), ("ln_post", "post_layernorm")):
        out[vm + b + ".weight"] = sd[f"image_encoder.{a}.weight"]
        out[vm + b + ".bias"] = sd[f"image_encoder.{a}.bias"]
    for i in range(v["layers"]):
        s = f"image_encoder.transformer.resblocks.{i}."
        d = f"{vm}encoder.layers.{i}."
        for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
            out[d + f"self_attn.{nm}.weight"] = sd[s + "attn.in_proj_weight"][j * W:(j + 1) * W]
            out[d + f"self_attn.{nm}.bias"] = sd[s + "attn.in_proj_bias"][j * W:(j + 1) * W]
        for a, b in (("attn.out_proj", "self_attn.out_proj"), ("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"),
                     ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
            out[d + b + ".weight"] = sd[s + a + ".weight"]
            out[d + b + ".bias"] = sd[s + a + ".bias"]
    for k, val in sd.items():
        if k.startswith("textual.visual_projection."):
            out["git.visual_projection.visual_projection." + k[len("textual.visual_projection."):]] = val
        elif k.startswith("textual.transformer.encoder."):
            out["git.encoder." + k[len("textual.transformer.encoder."):]] = val
        elif k.startswith("img_temperal_embedding."):
            out["git.img_temporal_embedding." + k.split(".")[-1]] = val
    out["git.embeddings.word_embeddings.weight"] = sd["textual.embedding.words.weight"]
    out["git.embeddings.position_embeddings.weight"] = sd["textual.embedding.positions.weight"]
    out["git.embeddings.LayerNorm.weight"] = sd["textual.embedding.layer_norm.weight"]
    out["git.embeddings.LayerNorm.bias"] = sd["textual.embedding.layer_norm.bias"]
    out["output.weight"] = sd["textual.output.weight"]
    out["output.bias"] = sd["textual.output.bias"]
    return out


@pytest.mark.parametrize("encoder", ["CLIPViT_B_16", "CLIPViT_L_14"])
def test_oracle_matches_hf_git_forward(encoder):
    """Both vision towers the reference can be configured with (model.py:682-685): ViT-B/16 (GIT-base, BASELINE configs
    1-3) and ViT-L/14 (the shipped teacher, data/teacher_configs/GIT_LARGE_MSRVTT/parameter.yaml; BASELINE configs 4-5)."""
    from transformers import GitConfig as HFGitConfig, GitForCausalLM

    torch.manual_seed(0)
    F_, L = 2, 5
    # HF uses one LayerNorm eps (1e-12) for embeddings too: align the oracle for this comparison only.
    large = encoder == "CLIPViT_L_14"
    cfg = go.GitConfig(num_image_with_embedding=F_, embedding_ln_eps=1e-12, image_encoder_type=encoder,
                       visual_feature_size=1024 if large else 768)
    sd = go.init_state_dict(cfg, seed=3, temporal_std=0.02, perturb=True)
    vision = dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16, patch_size=14,
                  image_size=224) if large else {}
    hf_cfg = HFGitConfig(num_image_with_embedding=F_, layer_norm_eps=1e-12, tie_word_embeddings=False, vision_config=vision or None)
    hf = GitForCausalLM(hf_cfg).eval()
    missing, unexpected = hf.load_state_dict(to_hf_state_dict(sd, cfg), strict=False)
    missing = [m for m in missing if "position_ids" not in m]
    assert not missing and not unexpected, (missing, unexpected)

    g = torch.Generator().manual_seed(1)
    frames = torch.randn(F_, 3, 224, 224, generator=g)
    tokens = torch.tensor([[101, 2023, 2003, 1037, 3231]])
    with torch.no_grad():
        logits, vf, hidden = go.forward_one_custom(sd, cfg, frames, tokens)
        hf_out = hf(input_ids=tokens, pixel_values=frames[None], output_hidden_states=False)
        hf_vis = torch.cat([hf.git.image_encoder(frames[i:i + 1]).last_hidden_state + hf.git.img_temporal_embedding[i]
                            for i in range(F_)], dim=1)
    nv = vf.shape[1]
    assert nv == F_ * (257 if large else 197)
    assert torch.allclose(vf, hf_vis, atol=2e-4, rtol=1e-4), (vf - hf_vis).abs().max()
    hf_logits = hf_out.logits
    if hf_logits.shape[1] == nv + L:  # HF returns every row; upstream slices the text rows
        hf_logits = hf_logits[:, nv:]
    assert hf_logits.shape == logits.shape
    err = (logits - hf_logits).abs().max().item()
    assert err < 2e-3 * max(1.0, logits.abs().max().item()), err
    assert hidden.shape == (7, nv + L, 768)


def test_history_cache_equals_full_forward():
    """The upstream hidden-state history path must be arithmetically identical to a full re-forward."""
    cfg = go.GitConfig(num_image_with_embedding=1, resolution=32, image_encoder_type="CLIPViT_B_16")
    sd = go.init_state_dict(cfg, seed=5)
    g = torch.Generator().manual_seed(2)
    vf = go.encode_clip(sd, cfg, torch.randn(1, 3, 32, 32, generator=g))
    toks = torch.tensor([[101, 7, 9, 11]])
    with torch.no_grad():
        full, _ = go.textual_forward(sd, cfg, vf, toks)
        st = go.DecodingState(sd, cfg, vf)
        outs = [st(toks[:, :t + 1]) for t in range(toks.shape[1])]
    for t, o in enumerate(outs):
        assert torch.allclose(o, full[:, t], atol=1e-4), (t, (o - full[:, t]).abs().max())


def test_oracle_greedy_search_matches_hf_generate_on_a_single_image():
    """An independent END-TO-END check of encode + decoding + greedy search: HuggingFace's ``GitForCausalLM.generate`` (its own
    generation loop) against the oracle's ``infer`` (the reference's ``GeneratorWithBeamSearchV2.search`` restated line by line,
    beam_size 1, over the upstream hidden-state history cache) on the single-image branch (model.py:387-388; HF's generation
    cannot take videos: SURVEY Appendix C).  Untied random head, so the tokens are not the trivial copy.  ``use_cache=False``:
    in transformers 5.5 HF's CACHED GIT generation disagrees with HF's own full forward from the second step on
    ([101, 18861, 24035, ...] cached vs [101, 18861, 29428, ...] for both its un-cached loop and its teacher-forced forward on
    this input); the un-cached loop re-runs the full forward the other test pins the oracle against, inside HF's loop.
    The reference's loop scores ``max_steps - 1`` steps, keeps the first ``max_steps - 2`` words and appends EOS (model.py:585-596,
    :653-677): those words must be HF's greedy continuation."""
    from transformers import GitConfig as HFGitConfig, GitForCausalLM
    from oracle import search_oracle as so

    torch.manual_seed(0)
    cfg = go.GitConfig(num_image_with_embedding=0, embedding_ln_eps=1e-12, tie_output=False)
    sd = go.init_state_dict(cfg, seed=9, temporal_std=0.0, perturb=True)
    hf_cfg = HFGitConfig(num_image_with_embedding=None, layer_norm_eps=1e-12, tie_word_embeddings=False,
                         bos_token_id=101, eos_token_id=102, pad_token_id=0)
    hf = GitForCausalLM(hf_cfg).eval()
    missing, unexpected = hf.load_state_dict(to_hf_state_dict(sd, cfg), strict=False)
    assert not [m for m in missing if "position_ids" not in m] and not unexpected
    image = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(4))
    max_steps = 9
    with torch.no_grad():
        vf = go.vit_forward(sd, cfg, image)                                   # [1, 197, 768]: no temporal embedding
        ref = so.infer(sd, cfg, vf, beam_size=1, max_steps=max_steps, save_logits=False)["predictions"][0]
        out = hf.generate(pixel_values=image, input_ids=torch.tensor([[101]]), max_length=max_steps - 1, do_sample=False, num_beams=1,
                          use_cache=False)[0]
    words = ref[1:max_steps - 1]
    assert ref[0].item() == 101 and ref[max_steps - 1].item() == 102
    assert 102 not in words.tolist()                                          # random head: nobody emits EOS in 7 steps
    assert out[0].item() == 101 and torch.equal(out[1:], words), (out, ref)
