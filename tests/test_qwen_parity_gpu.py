"""Qwen-Image-Edit hot path (two passes with separate K/V caches, complex RoPE table from the pipeline's pos_embed,
norm-rescaled CFG) through RegionEHelper / the C ABI against the CPU oracle. Gate: rel-L2 <= 1e-2, masks bit-exact."""
import pytest
import torch

from oracle.qwen import QwenOracle, run_regione_qwen

pytestmark = pytest.mark.gpu
TOL = 1e-2


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def _run(grid, txt_len, rho, params, cfg_scale, seed=7, n_blocks=3):
    from regione_b200 import RegionEHelper
    from standins import standin
    from standins import synthetic as syn

    gh, gw = grid
    arch = dict(dim=256, heads=2, n_blocks=n_blocks, mlp_ratio=4, in_channels=64, ctx_dim=128)
    tr = standin.QwenImageTransformer2DModel(**arch).init_synthetic(110, "cpu")
    with torch.no_grad():
        tr.proj_out.weight.mul_(0.3 / (0.02 * 16))
        tr.proj_out.bias.mul_(0.3 / (0.02 * 16))
    weights = {k: v.detach().clone() for k, v in tr.state_dict().items()}
    inp = syn.make_inputs(seed, gh, gw, txt_len, arch["ctx_dim"], 64, rho=rho)
    g = torch.Generator().manual_seed(seed + 1)
    neg = (0.1 * torch.randn(1, txt_len, arch["ctx_dim"], generator=g)).bfloat16() if cfg_scale > 1 else None
    img_f, txt_f = tr.pos_embed([[(1, gh, gw), (1, gh, gw)]], [txt_len])
    ref, ref_tr = run_regione_qwen(QwenOracle(weights, arch["heads"], n_blocks), dict(num_inference_steps=28, **params),
                                   inp["latents"], inp["image_latents"], inp["prompt_embeds"], neg, cfg_scale, img_f,
                                   txt_f, txt_f, inp["height"], inp["width"], record=True)
    pipe = standin.QwenImageEditPipeline(tr.to("cuda"))
    helper = RegionEHelper(pipe)
    helper.set_params(**params)
    helper.enable()
    pipe = helper.pipeline
    pipe.regione_record = True
    out = pipe(latents=inp["latents"].cuda(), image_latents=inp["image_latents"].cuda(),
               prompt_embeds=inp["prompt_embeds"].cuda(), negative_prompt_embeds=None if neg is None else neg.cuda(),
               true_cfg_scale=cfg_scale, height=inp["height"], width=inp["width"], num_inference_steps=28,
               output_type="latent", return_dict=False)[0]
    torch.cuda.synchronize()
    tr_cu = pipe.regione_trace
    helper.disable()
    assert tr_cu["modes"] == ref_tr["modes"]
    assert torch.equal(tr_cu["edited_ids"].cpu(), ref_tr["edited_ids"].squeeze(0).to(torch.int32))
    v_tol = TOL * max(1.0, cfg_scale)   # guided velocity amplifies the two passes' bf16 rounding; the gate is on latents
    for i, (a, b) in enumerate(zip(tr_cu["noise_pred"], ref_tr["noise_pred"])):
        if ref_tr["modes"][i] != "SKIP":
            assert rel_l2(a, b[0]) <= v_tol, f"step {i} ({ref_tr['modes'][i]}): velocity rel-L2 {rel_l2(a, b[0]):.3e}"
    for i, (a, b) in enumerate(zip(tr_cu["latents"], ref_tr["latents"])):
        assert rel_l2(a, b[0]) <= TOL, f"step {i}: latent rel-L2 {rel_l2(a, b[0]):.3e}"
    assert rel_l2(out, ref) <= TOL
    return rel_l2(out, ref)


QWEN = dict(warmup_step=6, post_step=2, refresh_step="16", threshold=0.80, cache_threshold=0.03, erosion_dilation=True)


def test_qwen_true_cfg_two_caches():
    err = _run((16, 16), 24, 0.25, QWEN, 4.0)
    print(f"qwen cfg: final rel-L2 {err:.3e}")


def test_qwen_without_cfg_ragged():
    _run((12, 20), 40, 0.4, dict(QWEN, refresh_step="12,20", cache_threshold=0.02), 1.0, seed=11, n_blocks=2)


def test_rmsnorm_and_cfg_rescale_kernels():
    from regione_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(77, 3584, device="cuda", generator=g).bfloat16()
    w = (1 + 0.1 * torch.randn(3584, device="cuda", generator=g)).bfloat16()
    var = x.float().pow(2).mean(-1, keepdim=True)
    ref = (x * torch.rsqrt(var + 1e-6)).to(torch.bfloat16) * w
    got = ops.rmsnorm(x, w)
    assert rel_l2(got, ref) <= 1e-3 and float((got == ref).float().mean()) > 0.99
    pos = torch.randn(4096, 64, device="cuda", generator=g).bfloat16()
    neg = torch.randn(4096, 64, device="cuda", generator=g).bfloat16()
    comb = neg + 4.0 * (pos - neg)
    ref = comb * (torch.norm(pos, dim=-1, keepdim=True) / torch.norm(comb, dim=-1, keepdim=True))
    got = ops.cfg_rescale(pos, neg, 4.0)
    assert rel_l2(got, ref) <= 4e-3
    assert float((got == ref).float().mean()) > 0.9
