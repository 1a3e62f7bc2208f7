"""Full-WIDTH parity of the Step1X-Edit path at BASELINE configs[0]'s geometry (assets/demo_0.png is
900x1440 -> 800x1296 px ~ 1024^2 area -> a RAGGED 50 x 81 token grid, L = C = 4050; T = 640 prompt tokens; D = 3072,
24 heads; cond + uncond stacked on the batch axis = two K/V cache sets), depth-reduced to 1 double + 1 single block so
the oracle finishes in seconds on the box's device: one FULL step writing both caches and one REGION step against
them, through the patched transformer forward (RegionE/Step1XEdit/inplace.py:514-571) and the C ABI.
Tolerance (north_star): rel-L2 <= 1e-2 on the bf16 velocity of each batch row."""
import pytest
import torch

from oracle import region_ops as ro
from oracle.step1x import Step1XOracle

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def test_step1x_full_and_region_step_at_demo0_shapes():
    from regione_b200 import RegionEHelper
    from standins import step1x as sx
    from regione_b200 import step1x_edit as s1
    from standins import synthetic as syn
    from regione_b200.schedule import latent_image_ids

    dev = "cuda"
    gh, gw, T = 50, 81, 640
    L = gh * gw
    arch = dict(dim=3072, heads=24, n_double=1, n_single=1, mlp_ratio=4, in_channels=64, ctx_dim=4096, vec_dim=768)
    tr = sx.Step1XEditTransformer2DModel(**arch).init_synthetic(110, dev)
    w = {k: v.detach() for k, v in tr.state_dict().items()}
    inp = syn.make_inputs(110, gh, gw, T, arch["ctx_dim"], 64, rho=0.25, device=dev)
    g = torch.Generator().manual_seed(5)
    neg = (0.1 * torch.randn(1, T, arch["ctx_dim"], generator=g)).bfloat16().to(dev)
    mask = torch.ones(1, T, dtype=torch.long, device=dev)
    mask[0, T - 37:] = 0
    embeds = torch.cat([inp["prompt_embeds"], neg], dim=0)
    masks = torch.cat([mask, mask], dim=0)
    ids = torch.cat([latent_image_ids(gh, gw, 0.0, dev), latent_image_ids(gh, gw, 1.0, dev)])
    txt_ids = torch.zeros(T, 3, device=dev)
    t_in = (torch.tensor([935.6], device=dev).bfloat16() / 1000).expand(2)       # Step1XEdit/inplace.py:379-384

    params = dict(num_inference_steps=28, warmup_step=6, post_step=2, refresh_step="16", threshold=0.88,
                  cache_threshold=0.02, erosion_dilation=True)
    st = ro.RegionState()
    st.set_parameters(params)
    st.refresh(inp["latents"], inp["image_latents"], ids, txt_ids, gh * 16, gw * 16)
    model = Step1XOracle(w, arch["heads"], 1, 1, tr)

    pipe = sx.Step1XEditPipeline(tr)
    helper = RegionEHelper(pipe)
    saved = dict(helper.config)
    helper.set_params(**{k: v for k, v in params.items() if k != "num_inference_steps"})
    helper.enable()
    errs = {}
    try:
        M = s1.MANAGER
        engine = s1._get_engine(tr, T, L, L)
        tr.__dict__["_regione_b200_engine"] = engine
        M.refresh(inp["latents"][0], inp["image_latents"][0], ids, txt_ids, 2, 8, gh * 16, gw * 16)
        cos, sin = tr.pos_embed(torch.cat((txt_ids, ids), dim=0))
        for b in range(2):
            engine.begin_image_rope(cos, sin, b)
        # FULL step + cache write
        st.current_step = st.warmup_step - 1
        x_full = torch.cat([inp["latents"], inp["image_latents"]], dim=1).expand(2, -1, -1).contiguous()
        with torch.no_grad():
            ref = model.forward(st, x_full, embeds, t_in, masks, ids, txt_ids)[:, :L]
            got = tr(hidden_states=x_full, timestep=t_in, encoder_hidden_states=embeds, prompt_embeds_mask=masks,
                     txt_ids=txt_ids, img_ids=ids, return_dict=False)[0][:, :L]
        torch.cuda.synchronize()
        errs["FULL cond"], errs["FULL uncond"] = rel_l2(got[0], ref[0]), rel_l2(got[1], ref[1])
        # REGION step: 1003 edited tokens of the ragged grid
        gsel = torch.Generator().manual_seed(3)
        edited = torch.randperm(L, generator=gsel)[:1003].sort().values.to(dev)
        st.edited_ids = edited.unsqueeze(0)
        st.current_step = st.warmup_step
        M.edited_ids = edited.int()
        x_reg = (inp["latents"][:, edited] + 0.05).expand(2, -1, -1).contiguous()
        with torch.no_grad():
            ref = model.forward(st, x_reg, embeds, t_in, masks, ids[edited], txt_ids)
            got = tr(hidden_states=x_reg, timestep=t_in, encoder_hidden_states=embeds, prompt_embeds_mask=masks,
                     txt_ids=txt_ids, img_ids=ids[edited], return_dict=False)[0]
        torch.cuda.synchronize()
        errs["REGION cond"], errs["REGION uncond"] = rel_l2(got[0], ref[0]), rel_l2(got[1], ref[1])
    finally:
        helper.disable()
        helper.config.clear()
        helper.config.update(saved)
    print("step1x demo_0 shapes: " + ", ".join(f"{k} {v:.3e}" for k, v in errs.items()))
    for k, v in errs.items():
        assert v <= 1e-2, f"{k}: rel-L2 {v:.3e}"
