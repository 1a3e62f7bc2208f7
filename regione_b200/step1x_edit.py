"""Step1X-Edit (v1p1) variant of the plugin: host side of RegionE/Step1XEdit/inplace.py over the CUDA library.

The block stack is FLUX's; what differs (all from the reference):
  * classifier-free guidance runs cond + uncond stacked on the batch axis in ONE forward (inplace.py:381-399); the
    Triton scatter-GEMM shares its index across the batch (fused_kernels.py:77-78), i.e. each batch row has its own
    K/V cache — here batch row b = pass b of the library handle (two cache sets);
  * the front end is the fork's: `connector(embeds, timestep, mask)` -> (tokens, y), `temb = time_embed(time_proj(t *
    1000)) + vec_embed(y)` (:514-520). Those modules stay the pipeline's own (they are tiny and outside the block
    stack); their results enter the library through `rge_dit_step_ex` (external temb, per-step embedded context);
  * norm-processed CFG (:388-400): `rge_cfg_diff_norm` + the pipeline's `process_diff_norm` + `rge_cfg_combine`;
  * `scheduler.set_begin_index(0)` (:336) and `t.item()` (:401) — no device syncs here, the schedule lives on the host.
"""
from __future__ import annotations

import ctypes as C
import types

import numpy as np
import torch

from . import _lib, ops
from ._lib import G as GS
from ._lib import check, ptr, stream_ptr
from .engine import FluxEngine, cached_engine
from .flux_kontext import LATENT_SPACE_ONLY, RegionEB200AttnProcessor, RegionESchedulerMixin, calculate_shift, retrieve_timesteps
from .manager import RegionManager, plan_steps
from .params import GAMMA

gamma = GAMMA["Step1XEditPipeline"]           # Step1XEdit/inplace.py:47-49
MANAGER = RegionManager()                     # :50


class Step1XEngine(FluxEngine):
    def __init__(self, transformer, txt_len, lat_len, cond_len, n_pass=2):
        self.lib = _lib.load()
        tr = transformer
        blocks, singles = list(tr.transformer_blocks), list(tr.single_transformer_blocks)
        dim, in_ch = tr.x_embedder.weight.shape
        mlp_dim = blocks[0].ff.net[0].proj.weight.shape[0] if blocks else singles[0].proj_mlp.weight.shape[0]
        self.guidance_embeds = False
        cfg = _lib.Config(dim=dim, heads=(blocks[0] if blocks else singles[0]).attn.heads, n_double=len(blocks),
                          n_single=len(singles), mlp_ratio=mlp_dim // dim, in_channels=in_ch, ctx_dim=0, pooled_dim=0,
                          txt_len=txt_len, lat_len=lat_len, cond_len=cond_len, guidance_embeds=0, n_pass=n_pass,
                          device=tr.x_embedder.weight.device.index or 0, external_embed=3)
        self.cfg, self.key, self.in_channels, self.transformer = cfg, (txt_len, lat_len, cond_len, n_pass), in_ch, tr
        self._keep = []
        self._src = []
        self._h = C.c_void_p()
        check(self.lib.rge_create(C.byref(cfg), C.byref(self._h)), "rge_create")
        try:
            g = _lib.BLK_GLOBAL
            self._lin(g, 0, GS, "X_EMBED", tr.x_embedder, "x_embedder")
            self._lin(g, 0, GS, "NORM_OUT", tr.norm_out.linear, "norm_out.linear")
            self._lin(g, 0, GS, "PROJ_OUT", tr.proj_out, "proj_out")
            self._register_blocks(blocks, singles)
            check(self.lib.rge_finalize_weights(self._h), "rge_finalize_weights")
        except Exception:
            self.close()
            raise

    def begin_image_rope(self, cos, sin, pass_id):
        """(cos, sin) fp32 [S,128] of the pipeline's pos_embed over [text; noise; condition] ids (:495-499)."""
        cs = torch.stack([cos[:, 0::2], sin[:, 0::2]], dim=-1).to(torch.float32).contiguous()
        check(self.lib.rge_begin_image_ex(self._h, pass_id, ptr(cs), None, stream_ptr()), "rge_begin_image_ex")

    def step_ex(self, x_in, sel, temb, ctx, n_out, pass_id, x_cond=None):
        x, n_x = self._rows(x_in, "latents")
        xc, n_c = self._rows(x_cond, "condition latents")
        out = torch.empty(n_out, self.in_channels, dtype=torch.bfloat16, device=x.device)
        sel_ptr = ptr(sel)
        if sel is not None and sel.numel() == 0:
            sel_ptr = ptr(self._dummy_sel(x.device))
        temb, ctx = temb.contiguous(), ctx.contiguous()
        check(self.lib.rge_dit_step_ex(self._h, pass_id, ptr(x) if n_x else None, n_x, ptr(xc) if n_c else None, n_c,
                                       sel_ptr, ptr(temb), ptr(ctx), ptr(out) if n_out else None, n_out, stream_ptr()),
              "rge_dit_step_ex")
        return out


def _get_engine(transformer, T, L, C) -> Step1XEngine:
    return cached_engine(transformer, (T, L, C, 2), lambda: Step1XEngine(transformer, T, L, C))


def RegionEStep1XEditTransformer2DModelforward(self, hidden_states, encoder_hidden_states=None, timestep=None,
                                               prompt_embeds_mask=None, img_ids=None, txt_ids=None, guidance=None,
                                               joint_attention_kwargs=None, return_dict=True, condition_latents=None,
                                               **unused):
    """Signature of the reference's patched forward (Step1XEdit/inplace.py:459-475) plus `condition_latents` ([1,C,64]
    or [B,C,64], FULL steps: read in place instead of concatenated, :377-378); batch row b runs as pass b."""
    engine = self.__dict__.get("_regione_b200_engine")
    if engine is None:
        raise RuntimeError("regione_b200: no image in flight — the pipeline loop begins the image first")
    M = MANAGER
    B = hidden_states.shape[0]
    if B > 2:
        raise NotImplementedError("regione_b200: at most cond + uncond on the batch axis")
    dev = hidden_states.device
    ts = timestep.to(dev)
    enc, y = self.connector(encoder_hidden_states, ts, prompt_embeds_mask)                        # :514-516
    temb = self.time_embed(self.time_proj(ts * 1000).to(ts)) + self.vec_embed(y)                  # :519-520
    n_c = 0 if condition_latents is None else condition_latents.shape[1]
    full = hidden_states.shape[1] + n_c == M.latent_length + M.condition_length
    if full:
        sel, n_out = None, M.latent_length
    else:
        sel, n_out = M.edited_ids, hidden_states.shape[1]
    outs = []
    cw, cb = self.context_embedder.weight.detach(), self.context_embedder.bias.detach()
    for b in range(B):
        ctx = ops.gemm(enc[b].contiguous(), cw, cb)                                               # :521
        xc = None if condition_latents is None else condition_latents[min(b, condition_latents.shape[0] - 1)]
        outs.append(engine.step_ex(hidden_states[b], sel, temb[b], ctx, n_out, b, x_cond=xc))
    out = torch.stack(outs, dim=0)
    if not return_dict:
        return (out,)
    return types.SimpleNamespace(sample=out)


class RegionEStep1XEditPipelineMixin:
    """`RegionEStep1XEditPipeline.__call__` (Step1XEdit/inplace.py:76-455), latent-space entry."""

    @torch.no_grad()
    def __call__(self, image=None, prompt=None, negative_prompt=None, true_cfg_scale=6.0, height=None, width=None,
                 num_inference_steps=28, latents=None, prompt_embeds=None, prompt_embeds_mask=None,
                 negative_prompt_embeds=None, negative_prompt_embeds_mask=None, output_type="pil", return_dict=True,
                 joint_attention_kwargs=None, image_latents=None, timesteps_truncate=0.93, process_norm_power=0.4,
                 **unused):
        assert num_inference_steps == MANAGER.inference_step, "num_inference_steps should be equal to 28"
        if image_latents is None or latents is None or prompt_embeds is None or output_type != "latent":
            raise NotImplementedError(LATENT_SPACE_ONLY)
        if height is None or width is None:
            raise ValueError("height and width are required with packed latents")
        from .schedule import latent_image_ids
        device = self._execution_device
        gh, gw = height // (self.vae_scale_factor * 2), width // (self.vae_scale_factor * 2)
        assert latents.shape[1] == gh * gw and image_latents.shape[1] == gh * gw, "latents do not match H x W"
        text_ids = torch.zeros(prompt_embeds.shape[1], 3, device=device)
        latent_ids = torch.cat([latent_image_ids(gh, gw, 0.0, device), latent_image_ids(gh, gw, 1.0, device)])
        sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
        cfg = self.scheduler.config
        mu = calculate_shift(latents.shape[1], cfg.get("base_image_seq_len", 256), cfg.get("max_image_seq_len", 4096),
                             cfg.get("base_shift", 0.5), cfg.get("max_shift", 1.15))
        retrieve_timesteps(self.scheduler, num_inference_steps, device, sigmas=sigmas, mu=mu)
        self.scheduler.set_begin_index(0)                                                         # :336
        self.scheduler._step_index = 0
        do_true_cfg = true_cfg_scale > 1 and negative_prompt_embeds is not None
        if prompt_embeds_mask is None:
            prompt_embeds_mask = torch.ones(prompt_embeds.shape[:2], device=device, dtype=torch.long)
        if do_true_cfg and negative_prompt_embeds_mask is None:
            negative_prompt_embeds_mask = torch.ones(negative_prompt_embeds.shape[:2], device=device, dtype=torch.long)
        out = self.regione_denoise(latents, image_latents, latent_ids, text_ids, prompt_embeds, prompt_embeds_mask,
                                   negative_prompt_embeds if do_true_cfg else None, negative_prompt_embeds_mask,
                                   true_cfg_scale, timesteps_truncate, process_norm_power, height, width)
        if not return_dict:
            return (out,)
        return types.SimpleNamespace(images=out)

    def regione_denoise(self, latents, image_latents, latent_ids, text_ids, prompt_embeds, prompt_embeds_mask,
                        negative_prompt_embeds, negative_prompt_embeds_mask, true_cfg_scale, timesteps_truncate,
                        process_norm_power, height, width):
        """The hot loop, Step1XEdit/inplace.py:331-438."""
        M = MANAGER
        N = M.inference_step
        sch, tr = self.scheduler, self.transformer
        ts_host = sch.timesteps.detach().to("cpu", torch.float32)
        x, cond = latents[0], image_latents[0]
        L, Cn, T = x.shape[0], cond.shape[0], text_ids.shape[0]
        do_cfg = negative_prompt_embeds is not None
        engine = _get_engine(tr, T, L, Cn)
        tr.__dict__["_regione_b200_engine"] = engine
        M.refresh(x, cond, latent_ids, text_ids, 2, self.vae_scale_factor, height, width)          # :331
        cos, sin = tr.pos_embed(torch.cat((text_ids, latent_ids), dim=0))                         # :523-528 (full ids)
        for b in range(2 if do_cfg else 1):
            engine.begin_image_rope(cos, sin, b)
        if do_cfg:                                                                                # :383-386
            embeds = torch.cat((prompt_embeds, negative_prompt_embeds), dim=0)
            masks = torch.cat((prompt_embeds_mask, negative_prompt_embeds_mask), dim=0)
        else:
            embeds, masks = prompt_embeds, prompt_embeds_mask
        plan = plan_steps(ts_host, gamma, M)                                                      # :342-360
        cache = None
        record = bool(getattr(self, "regione_record", False))
        self.regione_trace = {"modes": [], "latents": [], "noise_pred": []}
        for i in range(N):
            assert i == M.current_step                                                            # :340
            t = ts_host[i]
            skip, ratio = plan[i]
            if skip:                                                                              # :362-366
                if cache.shape[0] != x.shape[0]:
                    cache = ops.gather_rows(cache, M.edited_ids)
                x = sch.step(cache, t, x, return_dict=False, reuse_ratio=ratio)[0]
                self.regione_trace["modes"].append("SKIP")
            else:
                cur = M.current_step
                full = cur <= M.warmup_step - 1 or cur > N - M.post_step - 1 or cur == M.prev_refresh_step   # :377
                timestep = t.expand(1).to(x.dtype)                                                # :379
                if do_cfg:                                                                        # :381-399
                    x_b = x[None].expand(2, -1, -1)      # both batch rows read the same latent (a view, no copy)
                    timestep = torch.cat((timestep, timestep), dim=0)
                else:
                    x_b = x[None]
                pred = self.transformer(hidden_states=x_b, timestep=timestep / 1000, guidance=None,
                                        encoder_hidden_states=embeds, prompt_embeds_mask=masks, txt_ids=text_ids,
                                        img_ids=latent_ids, joint_attention_kwargs=None, return_dict=False,
                                        condition_latents=image_latents if full else None)[0]   # :377-378
                pred = pred[:, : x.shape[0]]
                if do_cfg:
                    pos, neg = pred[0], pred[1]
                    if float(t) > timesteps_truncate:                                             # :401-407
                        diff_norm = ops.cfg_diff_norm(pos, neg)
                        denom = self.process_diff_norm(diff_norm.reshape(1, -1, 1), k=process_norm_power)
                        noise_pred = ops.cfg_combine(pos, neg, true_cfg_scale, denom.reshape(-1).to(pos.dtype))
                    else:                                                                         # :409-410
                        noise_pred = ops.cfg_combine(pos, neg, true_cfg_scale)
                else:
                    noise_pred = pred[0]
                cache = noise_pred
                x = sch.step(noise_pred, t, x, return_dict=False)[0]
                self.regione_trace["modes"].append("FULL" if full else "REGION")
            x, latent_ids = M.step(x, latent_ids)                                                 # :438
            if record:
                self.regione_trace["latents"].append(x.clone())
                self.regione_trace["noise_pred"].append(cache.clone())
        self.regione_trace["edited_ids"] = M.edited_ids
        self.regione_trace["unedited_ids"] = M.unedited_ids
        return x[None]


def warp_modules(pipeline, **args):
    """Step1XEdit/inplace.py:52-61."""
    if "_regione_b200_saved" in pipeline.__dict__:
        unwarp_modules(pipeline)
    MANAGER.set_parameters(args)
    tr = pipeline.transformer
    blocks = list(tr.transformer_blocks) + list(tr.single_transformer_blocks)
    saved = {"cls": pipeline.__class__, "scheduler": pipeline.scheduler, "forward": tr.__dict__.get("forward"),
             "processors": [getattr(b.attn, "processor", None) for b in blocks]}
    pipeline.__dict__["_regione_b200_saved"] = saved
    pipeline.__class__ = type("RegionEStep1XEditPipeline", (RegionEStep1XEditPipelineMixin, saved["cls"]), {})
    sch_cls = type("RegionEFlowMatchEulerDiscreteScheduler", (RegionESchedulerMixin, saved["scheduler"].__class__), {})
    pipeline.scheduler = sch_cls.from_config(saved["scheduler"].config)
    pipeline.scheduler._regione_manager = MANAGER
    tr.forward = types.MethodType(RegionEStep1XEditTransformer2DModelforward, tr)
    for b in tr.transformer_blocks:
        b.attn.set_processor(RegionEB200AttnProcessor(False))
    for b in tr.single_transformer_blocks:
        b.attn.set_processor(RegionEB200AttnProcessor(True))
    return pipeline


def unwarp_modules(pipeline):
    """Step1XEdit/inplace.py:64-72."""
    saved = pipeline.__dict__.pop("_regione_b200_saved", None)
    if saved is None:
        return pipeline
    tr = pipeline.transformer
    pipeline.__class__ = saved["cls"]
    pipeline.scheduler = saved["scheduler"].__class__.from_config(saved["scheduler"].config)
    if saved["forward"] is None:
        tr.__dict__.pop("forward", None)
    else:
        tr.forward = saved["forward"]
    for b, p in zip(list(tr.transformer_blocks) + list(tr.single_transformer_blocks), saved["processors"]):
        b.attn.set_processor(p)
    for eng in tr.__dict__.pop("_regione_b200_engines", {}).values():
        eng.close()
    tr.__dict__.pop("_regione_b200_engine", None)
    return pipeline
