"""Step1X-Edit v1p2 hot path: two tagged forwards per step with their own K/V caches AND their own text lengths
(rge_set_pass_text_len), text_token_mapping front end, norm-processed CFG; against the CPU oracle."""
from types import SimpleNamespace

import pytest
import torch

from oracle.step1x import Step1XV1P2Oracle, run_regione_step1x_v1p2

pytestmark = pytest.mark.gpu
TOL = 1e-2


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def _embeds(T, ctx, text_dim, seed, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    mask = torch.ones(1, T, dtype=torch.long)
    mask[0, T - 3:] = 0
    tm = torch.ones(1, T)
    tm[0, T // 2:] = 0
    e = SimpleNamespace(embedding=(0.1 * torch.randn(1, T, ctx, generator=g)).bfloat16(), mask=mask,
                        txt_ids=torch.zeros(T, 3), text_embeds=(0.1 * torch.randn(1, T, text_dim, generator=g)).bfloat16(),
                        text_masks=tm.bfloat16())
    return e


def _to(e, dev):
    return SimpleNamespace(**{k: v.to(dev) for k, v in vars(e).items()})


def test_step1x_v1p2_two_text_lengths():
    from regione_b200 import RegionEHelper, params
    from standins import step1x as sx
    from standins import synthetic as syn
    from regione_b200.schedule import latent_image_ids

    gh, gw, Tc, Tu = 16, 16, 40, 24          # cond prompt longer than the uncond prompt
    arch = dict(dim=256, heads=2, n_double=2, n_single=2, mlp_ratio=4, in_channels=64, ctx_dim=128, vec_dim=64,
                text_dim=96)
    tr = sx.Step1XEditV1P2Transformer2DModel(**arch).init_synthetic(110, "cpu")
    with torch.no_grad():
        tr.proj_out.weight.mul_(0.3 / (0.02 * 16))
        tr.proj_out.bias.mul_(0.3 / (0.02 * 16))
    weights = {k: v.detach().clone() for k, v in tr.state_dict().items()}
    inp = syn.make_inputs(7, gh, gw, Tc, arch["ctx_dim"], 64, rho=0.25)
    pe, ne = _embeds(Tc, arch["ctx_dim"], 96, 1), _embeds(Tu, arch["ctx_dim"], 96, 2)
    ids = torch.cat([latent_image_ids(gh, gw, 0.0), latent_image_ids(gh, gw, 1.0)])
    p = dict(warmup_step=6, post_step=2, refresh_step="16", threshold=0.88, cache_threshold=0.02, erosion_dilation=True)
    with torch.no_grad():
        ref, ref_tr = run_regione_step1x_v1p2(
            Step1XV1P2Oracle(weights, arch["heads"], 2, 2, tr), dict(num_inference_steps=28, **p),
            params.GAMMA["Step1XEditPipelineV1P2"], inp["latents"], inp["image_latents"], ids, pe, ne, 6.0,
            sx.Step1XEditPipeline.process_diff_norm, inp["height"], inp["width"], record=True)
    pipe = sx.Step1XEditPipelineV1P2(tr.to("cuda"))
    helper = RegionEHelper(pipe)
    helper.set_params(**p)
    helper.enable()
    pipe = helper.pipeline
    pipe.regione_record = True
    out = pipe(latents=inp["latents"].cuda(), image_latents=inp["image_latents"].cuda(), prompt_embeds=_to(pe, "cuda"),
               negative_prompt_embeds=_to(ne, "cuda"), true_cfg_scale=6.0, height=inp["height"], width=inp["width"],
               num_inference_steps=28, output_type="latent", return_dict=False)[0]
    torch.cuda.synchronize()
    t = pipe.regione_trace
    helper.disable()
    assert t["modes"] == ref_tr["modes"]
    assert torch.equal(t["edited_ids"].cpu(), ref_tr["edited_ids"].squeeze(0).to(torch.int32))
    for i, (a, b) in enumerate(zip(t["noise_pred"], ref_tr["noise_pred"])):
        if ref_tr["modes"][i] != "SKIP":   # guided velocity: tolerance scales with the guidance scale (see test_step1x_*)
            assert rel_l2(a, b[0]) <= TOL * 6.0, f"step {i}: velocity rel-L2 {rel_l2(a, b[0]):.3e}"
    for i, (a, b) in enumerate(zip(t["latents"], ref_tr["latents"])):
        assert rel_l2(a, b[0]) <= TOL, f"step {i}: latent rel-L2 {rel_l2(a, b[0]):.3e}"
    assert rel_l2(out, ref) <= TOL
