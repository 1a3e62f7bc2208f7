"""The last block of the stack computes queries, attention, MLP and output projections only for the rows whose result
survives (`norm_out` / `proj_out` read the noise rows; the reference discards text and instruction-image rows after the
last block, inplace.py:347, :566-567). Rows are independent in each of those ops, so the velocity must be BIT-IDENTICAL
with the trimming on and off - for a FULL step, a REGION step, the single-stream last block (FLUX / Step1X) and the
dual-stream last block (Qwen: no single blocks)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _both(run):
    from regione_b200 import ops
    outs = []
    for on in (1, 0):
        ops.set_option("trim_last", on)
        try:
            outs.append(run())
            torch.cuda.synchronize()
        finally:
            ops.set_option("trim_last", 1)
    return outs


def test_flux_last_single_block_trimmed_bit_identical():
    from standins import synthetic as syn
    from regione_b200.engine import FluxEngine
    from regione_b200.schedule import latent_image_ids
    dev = "cuda"
    arch = dict(dim=512, heads=4, n_double=1, n_single=2, mlp_ratio=4, in_channels=64, ctx_dim=128, pooled_dim=64,
                guidance_embeds=True)
    G, T = 20, 40
    L = G * G
    pipe = syn.build_pipeline(arch, seed=3, device=dev)
    inp = syn.make_inputs(3, G, G, T, arch["ctx_dim"], arch["pooled_dim"], rho=0.3, device=dev)
    ids = torch.cat([latent_image_ids(G, G, 0.0, dev), latent_image_ids(G, G, 1.0, dev)])
    eng = FluxEngine(pipe.transformer, T, L, L)
    try:
        eng.begin_image(torch.zeros(T, 3, device=dev), ids, inp["prompt_embeds"][0], inp["pooled_prompt_embeds"][0],
                        2500.0)
        x, c = inp["latents"][0], inp["image_latents"][0]
        a, b = _both(lambda: eng.step(x, None, 936.0, L, x_cond=c))
        assert torch.equal(a, b) and torch.isfinite(a.float()).all()
        # the two-pointer entry equals the concatenated one
        cat = eng.step(torch.cat([x, c]), None, 936.0, L)
        assert torch.equal(cat, a)
        ed = torch.randperm(L, device=dev)[:137].sort().values.int()
        xr = x[ed.long()] + 0.05
        a, b = _both(lambda: eng.step(xr, ed, 920.0, ed.numel()))
        assert torch.equal(a, b) and torch.isfinite(a.float()).all()
    finally:
        eng.close()


def test_qwen_last_dual_block_trimmed_bit_identical():
    from standins import standin
    from standins import synthetic as syn
    from regione_b200.engine_qwen import QwenEngine
    dev = "cuda"
    arch = dict(dim=512, heads=4, n_blocks=2, mlp_ratio=4, in_channels=64, ctx_dim=128)
    G, T = 18, 24
    L = G * G
    tr = standin.QwenImageTransformer2DModel(**arch).init_synthetic(5, dev)
    inp = syn.make_inputs(5, G, G, T, arch["ctx_dim"], 64, rho=0.3, device=dev)
    img_f, txt_f = tr.pos_embed([[(1, G, G), (1, G, G)]], [T], device=dev)
    eng = QwenEngine(tr, T, L, L, n_pass=2)
    try:
        for p in range(2):
            eng.begin_image_qwen(img_f, txt_f[:T], inp["prompt_embeds"][0], p)
        x, c = inp["latents"][0], inp["image_latents"][0]
        a, b = _both(lambda: eng.step(x, None, 936.0, L, pass_id=1, x_cond=c))
        assert torch.equal(a, b) and torch.isfinite(a.float()).all()
        ed = torch.randperm(L, device=dev)[:101].sort().values.int()
        xr = x[ed.long()] + 0.05
        a, b = _both(lambda: eng.step(xr, ed, 920.0, ed.numel(), pass_id=1))
        assert torch.equal(a, b) and torch.isfinite(a.float()).all()
    finally:
        eng.close()
