"""Parity of the CUDA hot path (through RegionEHelper / the C ABI) against the CPU oracle on identical seeded inputs.

Gate (BASELINE.json north_star): relative L2 <= 1e-2 on bf16 latents, region masks bit-exact.
"""
import pytest
import torch

from oracle.flux import FluxOracle
from oracle.loop import run_regione
from oracle.schedule import GAMMA

pytestmark = pytest.mark.gpu

TOL = 1e-2


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def _run_both(arch, grid, txt_len, rho, params, seed=7, true_cfg_scale=1.0):
    from regione_b200 import RegionEHelper
    from standins import synthetic as syn
    from regione_b200.schedule import latent_image_ids

    gh, gw = grid
    pipe = syn.build_pipeline(arch, seed=110, device="cpu")
    weights = {k: v.detach().clone() for k, v in pipe.transformer.state_dict().items()}
    inp = syn.make_inputs(seed, gh, gw, txt_len, arch["ctx_dim"], arch["pooled_dim"], rho=rho)
    # oracle on the CPU
    model = FluxOracle(weights, arch["heads"], arch["n_double"], arch["n_single"], arch["guidance_embeds"])
    ids = torch.cat([latent_image_ids(gh, gw, 0.0), latent_image_ids(gh, gw, 1.0)])
    o_params = dict(num_inference_steps=28, **params)
    neg_kw, negative = {}, None
    if true_cfg_scale > 1:
        g = torch.Generator().manual_seed(seed + 100)
        neg_e = (0.1 * torch.randn(1, txt_len, arch["ctx_dim"], generator=g)).bfloat16()
        neg_p = torch.randn(1, arch["pooled_dim"], generator=g).bfloat16()
        negative = (neg_e, neg_p, true_cfg_scale)
        neg_kw = dict(negative_prompt_embeds=neg_e.cuda(), negative_pooled_prompt_embeds=neg_p.cuda(),
                      true_cfg_scale=true_cfg_scale)
    ref, ref_tr = run_regione(model, o_params, GAMMA["FluxKontext"], inp["latents"], inp["image_latents"], ids,
                              torch.zeros(txt_len, 3), inp["prompt_embeds"], inp["pooled_prompt_embeds"], 2.5,
                              inp["height"], inp["width"], record=True, negative=negative)
    # CUDA path through the plugin surface
    pipe.transformer.to("cuda")
    helper = RegionEHelper(pipe)
    helper.set_params(**params)
    helper.enable()
    pipe = helper.pipeline
    pipe.regione_record = True
    cu = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in inp.items() if k != "intended_mask"}
    out = pipe(guidance_scale=2.5, num_inference_steps=28, output_type="latent", return_dict=False, **cu, **neg_kw)[0]
    torch.cuda.synchronize()
    tr = pipe.regione_trace
    helper.disable()
    return ref, ref_tr, out, tr


def _check(ref, ref_tr, out, tr, v_tol=TOL):
    assert tr["modes"] == ref_tr["modes"]
    e_ref = ref_tr["edited_ids"].squeeze(0).to(torch.int32)
    u_ref = ref_tr["unedited_ids"].squeeze(0).to(torch.int32)
    margin = float((ref_tr["similarity"] - 0.0).abs().min()) if "similarity" in ref_tr else float("nan")
    assert torch.equal(tr["edited_ids"].cpu(), e_ref), f"edited ids differ (similarity margin info {margin})"
    assert torch.equal(tr["unedited_ids"].cpu(), u_ref)
    worst = 0.0
    for i, (a, b) in enumerate(zip(tr["noise_pred"], ref_tr["noise_pred"])):
        if ref_tr["modes"][i] == "SKIP":
            continue
        err = rel_l2(a, b[0])
        worst = max(worst, err)
        assert err <= v_tol, f"step {i} ({ref_tr['modes'][i]}): velocity rel-L2 {err:.3e} > {v_tol}"
    for i, (a, b) in enumerate(zip(tr["latents"], ref_tr["latents"])):
        err = rel_l2(a, b[0])
        assert err <= TOL, f"step {i}: latent rel-L2 {err:.3e} > {TOL}"
    final = rel_l2(out, ref)
    assert final <= TOL, f"final latent rel-L2 {final:.3e}"
    return worst, final


DEFAULT = dict(warmup_step=6, post_step=2, refresh_step="16", threshold=0.88, cache_threshold=0.04,
               erosion_dilation=True)


def test_tiny_flux_default_schedule():
    from standins import synthetic as syn
    worst, final = _check(*_run_both(syn.TINY, (16, 16), 32, 0.25, DEFAULT))
    print(f"tiny: worst velocity rel-L2 {worst:.3e}, final latent rel-L2 {final:.3e}")


def test_ragged_grid_and_text_length():
    """Non-square grid, token counts that are not tile multiples, several refresh steps, no morphology."""
    from standins import synthetic as syn
    params = dict(warmup_step=4, post_step=3, refresh_step="10,18", threshold=0.88, cache_threshold=0.01,
                  erosion_dilation=False)
    _check(*_run_both(syn.TINY, (12, 20), 40, 0.4, params, seed=11))


def test_all_tokens_edited():
    """rho = 1: empty unedited set."""
    from standins import synthetic as syn
    ref, ref_tr, out, tr = _run_both(syn.TINY, (16, 16), 32, 1.0, DEFAULT, seed=3)
    assert tr["unedited_ids"].numel() == 0
    _check(ref, ref_tr, out, tr)


def test_no_token_edited():
    """rho = 0 without salt noise would still leave stray pixels; with erosion they vanish -> empty edited set."""
    from standins import synthetic as syn
    ref, ref_tr, out, tr = _run_both(syn.TINY, (16, 16), 32, 0.0, DEFAULT, seed=5)
    assert tr["edited_ids"].numel() == ref_tr["edited_ids"].numel()
    _check(ref, ref_tr, out, tr)


def test_true_cfg_second_forward_shares_the_cache():
    """true_cfg_scale > 1 with a negative prompt (inplace.py:349-364): two forwards per step over ONE K/V cache set,
    like the reference's single per-processor cache. The gate is on latents; the guided velocity amplifies the two
    forwards' independent bf16 rounding by ~the scale, so its own bound scales with it."""
    from standins import synthetic as syn
    ref, ref_tr, out, tr = _run_both(syn.TINY, (16, 16), 32, 0.25, DEFAULT, seed=17, true_cfg_scale=3.0)
    worst, final = _check(ref, ref_tr, out, tr, v_tol=3.0 * TOL)
    print(f"true-CFG 3.0: worst velocity rel-L2 {worst:.3e}, final latent rel-L2 {final:.3e}")


def test_wider_model_three_heads():
    from standins import synthetic as syn
    arch = dict(syn.TINY, dim=768, heads=6, n_double=1, n_single=2, ctx_dim=256)
    _check(*_run_both(arch, (16, 16), 64, 0.25, DEFAULT, seed=13))


@pytest.mark.parametrize("env", [{"RGE_GROUPED": "1"}, {"RGE_GROUPED": "1", "RGE_FILL_ATTN_TAIL": "0"},
                                 {"RGE_FILL_ATTN_TAIL": "0"}, {"RGE_NO_FANOUT": "1"}, {"RGE_GROUP_QKV": "1"},
                                 {"RGE_GROUP_QKV": "1", "RGE_FILL_ATTN_TAIL": "0"}])
def test_launch_schedule_variants_give_the_same_image(env, monkeypatch):
    """The engine's launch-schedule knobs (read at rge_create): one grouped launch per stage, attention-tail fill off,
    no side streams. They only reorder independent launches, so every one must pass the same parity gate."""
    from standins import synthetic as syn
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    _check(*_run_both(syn.TINY, (16, 16), 32, 0.25, DEFAULT))
    _check(*_run_both(syn.TINY, (12, 20), 40, 0.0, DEFAULT, seed=5))      # empty edited set through the grouped path


def test_callback_on_step_end_and_interrupt():
    """inplace.py:377-385 / :321-322: the callback sees every step's latents at the step's own token count, may
    replace them, and an identity callback leaves the image bit-identical; `interrupt` skips a computed step exactly
    like the reference's `continue` (which then trips the reference's own step-counter assertion)."""
    from regione_b200 import RegionEHelper
    from standins import synthetic as syn
    pipe = syn.build_pipeline(syn.TINY, seed=110, device="cuda")
    helper = RegionEHelper(pipe)
    helper.set_params(**DEFAULT)
    helper.enable()
    pipe = helper.pipeline
    inp = syn.make_inputs(7, 16, 16, 32, syn.TINY["ctx_dim"], syn.TINY["pooled_dim"], rho=0.25, device="cuda")
    kw = {k: v for k, v in inp.items() if k != "intended_mask"}
    call = dict(guidance_scale=2.5, num_inference_steps=28, output_type="latent", return_dict=False, **kw)
    base = pipe(**call)[0]
    seen = []

    def identity(p, i, t, tensors):
        seen.append((i, float(t), tuple(tensors["latents"].shape)))
        return {}

    same = pipe(callback_on_step_end=identity, **call)[0]
    assert torch.equal(same, base)
    assert [s[0] for s in seen] == list(range(28))
    n_e = pipe.regione_trace["edited_ids"].numel()
    assert seen[0][2] == (1, 256, 64) and seen[10][2] == (1, n_e, 64) and seen[-1][2] == (1, 256, 64)

    def halve_at_3(p, i, t, tensors):
        return {"latents": tensors["latents"] * 0.5} if i == 3 else {}

    changed = pipe(callback_on_step_end=halve_at_3, **call)[0]
    assert not torch.equal(changed, base) and torch.isfinite(changed.float()).all()

    def stop_at_5(p, i, t, tensors):
        if i == 5:
            p._interrupt = True
        return {}

    with pytest.raises(AssertionError):       # the reference's `continue` skips MANAGER.step: its :293 assert fires
        pipe(callback_on_step_end=stop_at_5, **call)
    helper.disable()
