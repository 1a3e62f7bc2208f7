"""BASELINE-size parity (FLUX.1-Kontext 1024^2 shapes: T=512, L=C=4096, D=3072, 24 heads) of one FULL and one REGION
transformer step, depth-reduced to 1 double + 1 single block so the oracle (run on the GPU box's device as the checker)
finishes in seconds. Tolerance: relative L2 <= 1e-2 on the bf16 velocity (north_star)."""
import pytest
import torch

from oracle import region_ops as ro
from oracle.flux import FluxOracle

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def test_full_and_region_step_at_baseline_shapes():
    from standins import synthetic as syn
    from regione_b200.engine import FluxEngine
    from regione_b200.schedule import latent_image_ids

    dev = "cuda"
    arch = dict(syn.FLUX_KONTEXT, n_double=1, n_single=1)
    G, T = 64, 512
    L = G * G
    pipe = syn.build_pipeline(arch, seed=110, device=dev)
    w = {k: v.detach() for k, v in pipe.transformer.state_dict().items()}
    inp = syn.make_inputs(110, G, G, T, arch["ctx_dim"], arch["pooled_dim"], rho=0.25, device=dev)
    ids = torch.cat([latent_image_ids(G, G, 0.0, dev), latent_image_ids(G, G, 1.0, dev)])
    txt_ids = torch.zeros(T, 3, device=dev)
    guidance = torch.full([1], 2.5, device=dev)
    t_in = torch.tensor([935.6], device=dev).bfloat16() / 1000            # inplace.py:334-338
    t_x1000 = float((t_in * 1000)[0])
    g_x1000 = float(torch.tensor(2.5).bfloat16() * 1000)

    st = ro.RegionState()
    st.set_parameters(dict(num_inference_steps=28, warmup_step=6, post_step=2, refresh_step="16", threshold=0.88,
                           cache_threshold=0.04, erosion_dilation=True))
    st.refresh(inp["latents"], inp["image_latents"], ids, txt_ids, 1024, 1024)
    model = FluxOracle(w, arch["heads"], 1, 1, True)
    eng = FluxEngine(pipe.transformer, T, L, L)
    try:
        eng.begin_image(txt_ids, ids, inp["prompt_embeds"][0], inp["pooled_prompt_embeds"][0], g_x1000)
        # FULL step that also writes the cache (current_step == warmup-1)
        st.current_step = st.warmup_step - 1
        x_full = torch.cat([inp["latents"], inp["image_latents"]], dim=1)
        with torch.no_grad():
            ref_full = model.forward(st, x_full, inp["prompt_embeds"], inp["pooled_prompt_embeds"], t_in, ids, txt_ids,
                                     guidance)[0, :L]
        got_full = eng.step(x_full[0], None, t_x1000, L)
        torch.cuda.synchronize()
        e_full = rel_l2(got_full, ref_full)
        # REGION step on ~1000 edited tokens against the cache of the FULL step
        gsel = torch.Generator().manual_seed(3)
        edited = torch.randperm(L, generator=gsel)[:1000].sort().values.to(dev)
        st.edited_ids = edited.unsqueeze(0)
        st.current_step = st.warmup_step
        x_reg = inp["latents"][:, edited] + 0.05            # the edited tokens moved since the cache was written
        with torch.no_grad():
            ref_reg = model.forward(st, x_reg, inp["prompt_embeds"], inp["pooled_prompt_embeds"], t_in,
                                    ids[edited], txt_ids, guidance)[0]
        got_reg = eng.step(x_reg[0], edited.int(), t_x1000, edited.numel())
        torch.cuda.synchronize()
        e_reg = rel_l2(got_reg, ref_reg)
    finally:
        eng.close()
    print(f"baseline shapes: FULL rel-L2 {e_full:.3e}, REGION rel-L2 {e_reg:.3e}")
    assert e_full <= 1e-2 and e_reg <= 1e-2
