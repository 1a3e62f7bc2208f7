"""Pins the oracle's FLUX block math — restated from diffusers, which is absent here ("parity unpinned" in
oracle/flux.py) — against an INDEPENDENT implementation that does exist in this image: the original
black-forest-labs/FLUX blocks vendored by torchtitan (torchtitan/experiments/flux/model/layers.py: "imported from
black-forest-labs/FLUX"). diffusers' FluxTransformer2DModel is the same network with q / k / v (/ MLP-in) split into
separate Linear layers; the weight mapping below is the one of diffusers' conversion script. fp32, small width, CPU.
Checked: adaLN chunk order, LayerNorm / modulation, per-head RMSNorm on q and k, rotary convention, [text, image]
concatenation order, gating, GELU-tanh MLPs, the single block's fused in / out projections, sinusoidal time embedding."""
import pytest
import torch

from oracle import flux as of
from oracle import region_ops as ro

layers = pytest.importorskip("torchtitan.experiments.flux.model.layers")

D, H, T, G = 256, 2, 7, 4          # width (2 heads of 128), text tokens, image grid G x G


def _weights(seed):
    g = torch.Generator().manual_seed(seed)
    w = {}

    def lin(name, n, k, scale=0.05):
        w[name + ".weight"] = torch.randn(n, k, generator=g) * scale
        w[name + ".bias"] = torch.randn(n, generator=g) * 0.05

    p = "transformer_blocks.0."
    lin(p + "norm1.linear", 6 * D, D)
    lin(p + "norm1_context.linear", 6 * D, D)
    for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
        lin(p + "attn." + n, D, D)
    for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
        w[p + f"attn.{n}.weight"] = 1 + 0.1 * torch.randn(128, generator=g)
    lin(p + "ff.net.0.proj", 4 * D, D); lin(p + "ff.net.2", D, 4 * D)
    lin(p + "ff_context.net.0.proj", 4 * D, D); lin(p + "ff_context.net.2", D, 4 * D)
    s = "single_transformer_blocks.0."
    lin(s + "norm.linear", 3 * D, D)
    for n in ("to_q", "to_k", "to_v"):
        lin(s + "attn." + n, D, D)
    for n in ("norm_q", "norm_k"):
        w[s + f"attn.{n}.weight"] = 1 + 0.1 * torch.randn(128, generator=g)
    lin(s + "proj_mlp", 4 * D, D); lin(s + "proj_out", D, 5 * D)
    return w, g


def _state(ids):
    st = ro.RegionState()
    st.set_parameters(dict(num_inference_steps=28, warmup_step=6, post_step=2, refresh_step="16", threshold=0.9,
                           cache_threshold=0.04, erosion_dilation=True))
    st.refresh(torch.zeros(1, G * G, 4), torch.zeros(1, G * G, 4), ids, torch.zeros(T, 3), 16 * G, 16 * G)
    return st                       # current_step 0: the processor's plain (no cache) branch, inplace.py:717


def _cat(w, names, suffix):
    return torch.cat([w[n + suffix] for n in names], 0)


def test_double_and_single_block_match_bfl_flux():
    w, g = _weights(0)
    img_ids = torch.zeros(G * G, 3)
    img_ids[:, 1] = torch.arange(G * G) // G
    img_ids[:, 2] = torch.arange(G * G) % G
    ids = torch.cat([torch.zeros(T, 3), img_ids])
    img = torch.randn(1, G * G, D, generator=g)
    txt = torch.randn(1, T, D, generator=g)
    vec = torch.randn(1, D, generator=g)
    oracle = of.FluxOracle(w, H, 1, 1, True)
    st = _state(img_ids)
    rope = of.rope_cos_sin(ids)
    pe = layers.EmbedND(128, 10000, [16, 56, 56])(ids[None])

    # ---- double block
    blk = layers.DoubleStreamBlock(D, H, 4.0, qkv_bias=True)
    blk.load_state_dict(_double_sd(w), strict=True)
    with torch.no_grad():
        ref_img, ref_txt = blk(img, txt, vec, pe)
        got_txt, got_img = oracle.double_block(0, img, txt, vec, rope, rope, st)
    assert torch.allclose(got_img, ref_img, rtol=2e-4, atol=2e-4), float((got_img - ref_img).abs().max())
    assert torch.allclose(got_txt, ref_txt, rtol=2e-4, atol=2e-4), float((got_txt - ref_txt).abs().max())

    # ---- single block on [text; image]
    sb = layers.SingleStreamBlock(D, H, 4.0)
    sb.load_state_dict(_single_sd(w), strict=True)
    with torch.no_grad():
        ref = sb(torch.cat([txt, img], 1), vec, pe)
        got_txt, got_img = oracle.single_block(0, img, txt, vec, rope, rope, st)
    got = torch.cat([got_txt, got_img], 1)
    assert torch.allclose(got, ref, rtol=2e-4, atol=2e-4), float((got - ref).abs().max())


def _double_sd(w):
    """diffusers FluxTransformerBlock parameters -> BFL DoubleStreamBlock state dict (convert_flux_to_diffusers, reversed)."""
    p = "transformer_blocks.0."
    return {
        "img_mod.lin.weight": w[p + "norm1.linear.weight"], "img_mod.lin.bias": w[p + "norm1.linear.bias"],
        "txt_mod.lin.weight": w[p + "norm1_context.linear.weight"], "txt_mod.lin.bias": w[p + "norm1_context.linear.bias"],
        "img_attn.qkv.weight": _cat(w, [p + "attn.to_q", p + "attn.to_k", p + "attn.to_v"], ".weight"),
        "img_attn.qkv.bias": _cat(w, [p + "attn.to_q", p + "attn.to_k", p + "attn.to_v"], ".bias"),
        "txt_attn.qkv.weight": _cat(w, [p + "attn.add_q_proj", p + "attn.add_k_proj", p + "attn.add_v_proj"], ".weight"),
        "txt_attn.qkv.bias": _cat(w, [p + "attn.add_q_proj", p + "attn.add_k_proj", p + "attn.add_v_proj"], ".bias"),
        "img_attn.norm.query_norm.weight": w[p + "attn.norm_q.weight"],
        "img_attn.norm.key_norm.weight": w[p + "attn.norm_k.weight"],
        "txt_attn.norm.query_norm.weight": w[p + "attn.norm_added_q.weight"],
        "txt_attn.norm.key_norm.weight": w[p + "attn.norm_added_k.weight"],
        "img_attn.proj.weight": w[p + "attn.to_out.0.weight"], "img_attn.proj.bias": w[p + "attn.to_out.0.bias"],
        "txt_attn.proj.weight": w[p + "attn.to_add_out.weight"], "txt_attn.proj.bias": w[p + "attn.to_add_out.bias"],
        "img_mlp.0.weight": w[p + "ff.net.0.proj.weight"], "img_mlp.0.bias": w[p + "ff.net.0.proj.bias"],
        "img_mlp.2.weight": w[p + "ff.net.2.weight"], "img_mlp.2.bias": w[p + "ff.net.2.bias"],
        "txt_mlp.0.weight": w[p + "ff_context.net.0.proj.weight"], "txt_mlp.0.bias": w[p + "ff_context.net.0.proj.bias"],
        "txt_mlp.2.weight": w[p + "ff_context.net.2.weight"], "txt_mlp.2.bias": w[p + "ff_context.net.2.bias"],
    }


def _single_sd(w):
    s = "single_transformer_blocks.0."
    names = [s + "attn.to_q", s + "attn.to_k", s + "attn.to_v", s + "proj_mlp"]
    return {
        "modulation.lin.weight": w[s + "norm.linear.weight"], "modulation.lin.bias": w[s + "norm.linear.bias"],
        "linear1.weight": _cat(w, names, ".weight"), "linear1.bias": _cat(w, names, ".bias"),
        "linear2.weight": w[s + "proj_out.weight"], "linear2.bias": w[s + "proj_out.bias"],
        "norm.query_norm.weight": w[s + "attn.norm_q.weight"], "norm.key_norm.weight": w[s + "attn.norm_k.weight"],
    }


def test_whole_forward_matches_bfl_flux_model():
    """Front and back of the DiT too: x_embedder / context_embedder, time + pooled-text embedding (no guidance
    embedder in this BFL variant), norm_out (diffusers chunks (scale, shift), BFL (shift, scale): the conversion
    script swaps the halves of the Linear) and proj_out, around one double and one single block."""
    model_mod = pytest.importorskip("torchtitan.experiments.flux.model.model")
    from torchtitan.experiments.flux.model.args import FluxModelArgs
    w, g = _weights(3)
    CTX, POOL, CH = 32, 16, 64

    def lin(name, n, k):
        w[name + ".weight"] = torch.randn(n, k, generator=g) * 0.05
        w[name + ".bias"] = torch.randn(n, generator=g) * 0.05

    lin("x_embedder", D, CH); lin("context_embedder", D, CTX)
    lin("time_text_embed.timestep_embedder.linear_1", D, 256); lin("time_text_embed.timestep_embedder.linear_2", D, D)
    lin("time_text_embed.text_embedder.linear_1", D, POOL); lin("time_text_embed.text_embedder.linear_2", D, D)
    lin("norm_out.linear", 2 * D, D); lin("proj_out", CH, D)
    args = FluxModelArgs(in_channels=CH, out_channels=CH, vec_in_dim=POOL, context_in_dim=CTX, hidden_size=D,
                         num_heads=H, depth=1, depth_single_blocks=1)
    bfl = model_mod.FluxModel(args)
    swap = lambda t: torch.cat([t[D:], t[:D]], 0)                              # noqa: E731  (scale, shift) -> (shift, scale)
    sd = {"img_in.weight": w["x_embedder.weight"], "img_in.bias": w["x_embedder.bias"],
          "txt_in.weight": w["context_embedder.weight"], "txt_in.bias": w["context_embedder.bias"],
          "time_in.in_layer.weight": w["time_text_embed.timestep_embedder.linear_1.weight"],
          "time_in.in_layer.bias": w["time_text_embed.timestep_embedder.linear_1.bias"],
          "time_in.out_layer.weight": w["time_text_embed.timestep_embedder.linear_2.weight"],
          "time_in.out_layer.bias": w["time_text_embed.timestep_embedder.linear_2.bias"],
          "vector_in.in_layer.weight": w["time_text_embed.text_embedder.linear_1.weight"],
          "vector_in.in_layer.bias": w["time_text_embed.text_embedder.linear_1.bias"],
          "vector_in.out_layer.weight": w["time_text_embed.text_embedder.linear_2.weight"],
          "vector_in.out_layer.bias": w["time_text_embed.text_embedder.linear_2.bias"],
          "final_layer.adaLN_modulation.1.weight": swap(w["norm_out.linear.weight"]),
          "final_layer.adaLN_modulation.1.bias": swap(w["norm_out.linear.bias"]),
          "final_layer.linear.weight": w["proj_out.weight"], "final_layer.linear.bias": w["proj_out.bias"]}
    sd.update({"double_blocks.0." + k: v for k, v in _double_sd(w).items()})
    sd.update({"single_blocks.0." + k: v for k, v in _single_sd(w).items()})
    bfl.load_state_dict(sd, strict=True)
    img_ids = torch.zeros(G * G, 3)
    img_ids[:, 1] = torch.arange(G * G) // G
    img_ids[:, 2] = torch.arange(G * G) % G
    txt_ids = torch.zeros(T, 3)
    x = torch.randn(1, G * G, CH, generator=g)
    ctx = torch.randn(1, T, CTX, generator=g)
    pooled = torch.randn(1, POOL, generator=g)
    t = torch.tensor([0.9356])
    with torch.no_grad():
        ref = bfl(x, img_ids[None], ctx, txt_ids[None], t, pooled)
        got = of.FluxOracle(w, H, 1, 1, guidance_embeds=False).forward(_state(img_ids), x, ctx, pooled, t, img_ids, txt_ids, None)
    assert torch.allclose(got, ref, rtol=3e-4, atol=3e-4), float((got - ref).abs().max())


def test_rotary_table_and_time_embedding_match_bfl_flux():
    ids = torch.zeros(9, 3)
    ids[:, 0] = torch.tensor([0, 0, 0, 0, 1, 1, 1, 1, 1.0])
    ids[:, 1] = torch.arange(9) * 3.0
    ids[:, 2] = torch.arange(9) % 4
    cos, sin = of.rope_cos_sin(ids)                                     # [S, 128], pairs repeated
    pe = layers.EmbedND(128, 10000, [16, 56, 56])(ids[None])[0, 0]      # [S, 64, 2, 2] rotation matrices
    assert torch.allclose(cos[:, 0::2], pe[:, :, 0, 0], atol=1e-6) and torch.allclose(sin[:, 0::2], pe[:, :, 1, 0], atol=1e-6)
    x = torch.randn(1, 2, 9, 128)
    from torchtitan.experiments.flux.model.math import apply_rope
    ref, _ = apply_rope(x, x, pe[None, None])
    assert torch.allclose(of.apply_rope(x, (cos, sin)), ref, atol=1e-5)
    t = torch.tensor([0.9356, 0.1047])
    assert torch.allclose(of.timestep_projection(t * 1000), layers.timestep_embedding(t, 256), atol=1e-5)


def test_sigma_schedule_matches_bfl_flux_sampling():
    """The dynamic exponential time shift of diffusers' FlowMatchEulerDiscreteScheduler.set_timesteps (restated in
    oracle/schedule.py, mirrored by the stand-in scheduler) against BFL's own sampler schedule (get_schedule:
    linspace(1, 0, N + 1) -> time_shift(mu(seq_len), 1, t)); mu = the reference's calculate_shift."""
    sampling = pytest.importorskip("torchtitan.experiments.flux.sampling")
    import numpy as np
    from oracle.schedule import calculate_shift, flow_match_sigmas
    from regione_b200 import schedule
    from standins.diffusers_like import FlowMatchEulerDiscreteScheduler
    for n_tokens in (4096, 4050, 2304, 1024):
        ref = torch.tensor(sampling.get_schedule(28, n_tokens), dtype=torch.float32)          # 29 values, last 0
        sig, ts = flow_match_sigmas(28, n_tokens)
        assert sig.shape == ref.shape and float(sig[-1]) == 0.0
        assert torch.allclose(sig, ref, atol=2e-6, rtol=0)
        assert torch.allclose(ts, ref[:-1] * 1000, atol=2e-3, rtol=0)
        assert calculate_shift(n_tokens) == pytest.approx(sampling.get_lin_function()(n_tokens), abs=1e-12)
        sch = FlowMatchEulerDiscreteScheduler()
        schedule.retrieve_timesteps(sch, 28, "cpu", sigmas=np.linspace(1.0, 1 / 28, 28), mu=schedule.calculate_shift(n_tokens))
        assert torch.equal(sch.sigmas, sig)
