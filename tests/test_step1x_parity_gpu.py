"""Step1X-Edit hot path (cond + uncond stacked on the batch axis -> two K/V cache sets, external temb / per-step
context through rge_dit_step_ex, norm-processed CFG) through RegionEHelper against the CPU oracle."""
import pytest
import torch

from oracle.step1x import Step1XOracle, run_regione_step1x

pytestmark = pytest.mark.gpu
TOL = 1e-2


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def _run(grid, txt_len, rho, params, cfg_scale, seed=7):
    from regione_b200 import RegionEHelper
    from standins import step1x as sx
    from standins import synthetic as syn
    from regione_b200.schedule import latent_image_ids

    gh, gw = grid
    arch = dict(dim=256, heads=2, n_double=2, n_single=2, mlp_ratio=4, in_channels=64, ctx_dim=128, vec_dim=64)
    tr = sx.Step1XEditTransformer2DModel(**arch).init_synthetic(110, "cpu")
    with torch.no_grad():
        tr.proj_out.weight.mul_(0.3 / (0.02 * 16))
        tr.proj_out.bias.mul_(0.3 / (0.02 * 16))
    weights = {k: v.detach().clone() for k, v in tr.state_dict().items()}
    inp = syn.make_inputs(seed, gh, gw, txt_len, arch["ctx_dim"], 64, rho=rho)
    g = torch.Generator().manual_seed(seed + 1)
    neg = (0.1 * torch.randn(1, txt_len, arch["ctx_dim"], generator=g)).bfloat16() if cfg_scale > 1 else None
    mask = torch.ones(1, txt_len, dtype=torch.long)
    mask[0, txt_len - 5:] = 0                                  # padded prompt tail
    ids = torch.cat([latent_image_ids(gh, gw, 0.0), latent_image_ids(gh, gw, 1.0)])
    with torch.no_grad():
        ref, ref_tr = run_regione_step1x(
            Step1XOracle(weights, arch["heads"], arch["n_double"], arch["n_single"], tr),
            dict(num_inference_steps=28, **params), inp["latents"], inp["image_latents"], ids, torch.zeros(txt_len, 3),
            inp["prompt_embeds"], mask, neg, mask, cfg_scale, sx.Step1XEditPipeline.process_diff_norm, inp["height"],
            inp["width"], record=True)
    pipe = sx.Step1XEditPipeline(tr.to("cuda"))
    helper = RegionEHelper(pipe)
    helper.set_params(**params)
    helper.enable()
    pipe = helper.pipeline
    pipe.regione_record = True
    out = pipe(latents=inp["latents"].cuda(), image_latents=inp["image_latents"].cuda(),
               prompt_embeds=inp["prompt_embeds"].cuda(), prompt_embeds_mask=mask.cuda(),
               negative_prompt_embeds=None if neg is None else neg.cuda(),
               negative_prompt_embeds_mask=None if neg is None else mask.cuda(), true_cfg_scale=cfg_scale,
               height=inp["height"], width=inp["width"], num_inference_steps=28, output_type="latent",
               return_dict=False)[0]
    torch.cuda.synchronize()
    tr_cu = pipe.regione_trace
    helper.disable()
    assert tr_cu["modes"] == ref_tr["modes"]
    assert torch.equal(tr_cu["edited_ids"].cpu(), ref_tr["edited_ids"].squeeze(0).to(torch.int32))
    # the gate is on LATENTS (north_star). The guided velocity neg + s (pos - neg) amplifies the independent bf16
    # rounding of the two passes by ~s, so its own tolerance scales with the guidance scale.
    v_tol = TOL * max(1.0, cfg_scale)
    for i, (a, b) in enumerate(zip(tr_cu["noise_pred"], ref_tr["noise_pred"])):
        if ref_tr["modes"][i] != "SKIP":
            assert rel_l2(a, b[0]) <= v_tol, f"step {i} ({ref_tr['modes'][i]}): velocity rel-L2 {rel_l2(a, b[0]):.3e}"
    for i, (a, b) in enumerate(zip(tr_cu["latents"], ref_tr["latents"])):
        assert rel_l2(a, b[0]) <= TOL, f"step {i}: latent rel-L2 {rel_l2(a, b[0]):.3e}"
    assert rel_l2(out, ref) <= TOL


STEP1X = dict(warmup_step=6, post_step=2, refresh_step="16", threshold=0.88, cache_threshold=0.02, erosion_dilation=True)


def test_step1x_cfg_on_batch_axis():
    _run((16, 16), 32, 0.25, STEP1X, 6.0)


def test_step1x_without_cfg():
    _run((12, 20), 40, 0.4, dict(STEP1X, refresh_step="10,18"), 1.0, seed=11)


def test_step1x_cfg_kernels_bit_exact():
    from regione_b200 import ops
    from standins.step1x import Step1XEditPipeline
    g = torch.Generator(device="cuda").manual_seed(3)
    pos = torch.randn(1, 4096, 64, device="cuda", generator=g).bfloat16()
    neg = torch.randn(1, 4096, 64, device="cuda", generator=g).bfloat16()
    diff_norm = torch.norm(pos - neg, dim=2, keepdim=True)
    got_norm = ops.cfg_diff_norm(pos[0], neg[0])
    assert float((got_norm == diff_norm.reshape(-1)).float().mean()) > 0.98      # fp32 reduction order may differ
    denom = Step1XEditPipeline.process_diff_norm(diff_norm, k=0.4)
    ref = neg + 6.0 * (pos - neg) / denom
    assert torch.equal(ops.cfg_combine(pos[0], neg[0], 6.0, denom.reshape(-1)), ref[0])
    assert torch.equal(ops.cfg_combine(pos[0], neg[0], 6.0), (neg + 6.0 * (pos - neg))[0])
