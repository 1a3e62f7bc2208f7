"""Step1X-Edit v1p2 variant of the plugin: host side of the DiT / loop part of RegionE/Step1XEditV1P2/inplace.py.

Relative to v1p1 (step1x_edit.py), from the reference: cond and uncond are TWO forwards per step, tagged 'cond' /
'uncond' through `joint_attention_kwargs` (inplace.py:398, :416); the processor keeps two caches (`k/v_cache_even` for
cond, `_odd` for uncond, :800-803, :819-890) — pass 0 / pass 1 here; the two prompts have their own lengths
(`txt_length` / `neg_txt_length`, utils.py:444-445, inplace.py:833, :868) — `rge_set_pass_text_len`; an optional
`text_token_mapping(text_embeddings) * text_mask` is added to the connector output (:606-609) — front end, stays the
pipeline's own module. The thinking / reflection loop around the denoise (:192-212, :470-486) is LLM prompt rewriting,
not denoising, and is out of scope.
"""
from __future__ import annotations

import types

import numpy as np
import torch

from . import ops
from .flux_kontext import LATENT_SPACE_ONLY, RegionEB200AttnProcessor, RegionESchedulerMixin, calculate_shift, retrieve_timesteps
from .manager import RegionManager, plan_steps
from .params import GAMMA
from .engine import cached_engine
from .step1x_edit import Step1XEngine
from ._lib import check

gamma = GAMMA["Step1XEditPipelineV1P2"]       # Step1XEditV1P2/inplace.py:48-50
MANAGER = RegionManager()
MANAGER.neg_txt_length = None                 # utils.py:445


def _get_engine(transformer, T, L, C) -> Step1XEngine:
    return cached_engine(transformer, (T, L, C, 2), lambda: Step1XEngine(transformer, T, L, C))


def RegionEStep1XEditV1P2Transformer2DModelforward(self, hidden_states, encoder_hidden_states=None, timestep=None,
                                                   prompt_embeds_mask=None, img_ids=None, txt_ids=None, guidance=None,
                                                   text_embeddings=None, text_mask=None, joint_attention_kwargs=None,
                                                   return_dict=True, condition_latents=None, **unused):
    """Signature of the reference's patched forward (Step1XEditV1P2/inplace.py:540-560); the tag selects the pass."""
    engine = self.__dict__.get("_regione_b200_engine")
    if engine is None:
        raise RuntimeError("regione_b200: no image in flight — the pipeline loop begins the image first")
    if hidden_states.shape[0] != 1:
        raise NotImplementedError("regione_b200: batch size must be 1 per tagged forward")
    tag = (joint_attention_kwargs or {}).get("tag", "cond")
    if tag not in ("cond", "uncond"):
        raise NotImplementedError(f"Error tag: {tag}")
    M = MANAGER
    dev = hidden_states.device
    ts = timestep.to(dev)
    enc, y = self.connector(encoder_hidden_states, ts, prompt_embeds_mask)                        # :601-603
    if getattr(self, "text_token_mapping", None) is not None and text_embeddings is not None:    # :606-609
        enc = enc + self.text_token_mapping(text_embeddings) * text_mask[:, :, None].to(enc.dtype)
    temb = self.time_embed(self.time_proj(ts * 1000).to(ts)) + self.vec_embed(y)                  # :613-614
    ctx = ops.gemm(enc[0].contiguous(), self.context_embedder.weight.detach(), self.context_embedder.bias.detach())
    n_c = 0 if condition_latents is None else condition_latents.shape[1]
    full = hidden_states.shape[1] + n_c == M.latent_length + M.condition_length
    sel, n_out = (None, M.latent_length) if full else (M.edited_ids, hidden_states.shape[1])
    out = engine.step_ex(hidden_states[0], sel, temb[0], ctx, n_out, 0 if tag == "cond" else 1,
                         x_cond=None if condition_latents is None else condition_latents[0])[None]
    if not return_dict:
        return (out,)
    return types.SimpleNamespace(sample=out)


class RegionEStep1XEditV1P2PipelineMixin:
    """Denoising part of `RegionEStep1XEditPipelineV1P2.__call__` (inplace.py:347-458), latent-space entry.
    `prompt_embeds` / `negative_prompt_embeds` are objects with `.embedding [1,T,ctx]`, `.mask [1,T]`, `.txt_ids
    [T,3]`, `.text_embeds`, `.text_masks` like the reference's (:391-396)."""

    @torch.no_grad()
    def __call__(self, image=None, prompt=None, true_cfg_scale=6.0, height=None, width=None, num_inference_steps=28,
                 latents=None, prompt_embeds=None, negative_prompt_embeds=None, output_type="pil", return_dict=True,
                 joint_attention_kwargs=None, image_latents=None, timesteps_truncate=0.93, process_norm_power=0.4,
                 **unused):
        assert num_inference_steps == MANAGER.inference_step, "num_inference_steps should be equal to 28"
        if (image_latents is None or latents is None or prompt_embeds is None or negative_prompt_embeds is None
                or output_type != "latent"):
            raise NotImplementedError(LATENT_SPACE_ONLY)
        if height is None or width is None:
            raise ValueError("height and width are required with packed latents")
        from .schedule import latent_image_ids
        device = self._execution_device
        self._joint_attention_kwargs = joint_attention_kwargs or {}
        gh, gw = height // (self.vae_scale_factor * 2), width // (self.vae_scale_factor * 2)
        assert latents.shape[1] == gh * gw and image_latents.shape[1] == gh * gw, "latents do not match H x W"
        latent_ids = torch.cat([latent_image_ids(gh, gw, 0.0, device), latent_image_ids(gh, gw, 1.0, device)])
        sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
        cfg = self.scheduler.config
        mu = calculate_shift(latents.shape[1], cfg.get("base_image_seq_len", 256), cfg.get("max_image_seq_len", 4096),
                             cfg.get("base_shift", 0.5), cfg.get("max_shift", 1.15))
        retrieve_timesteps(self.scheduler, num_inference_steps, device, sigmas=sigmas, mu=mu)
        self.scheduler.set_begin_index(0)
        self.scheduler._step_index = 0
        out = self.regione_denoise(latents, image_latents, latent_ids, prompt_embeds, negative_prompt_embeds,
                                   true_cfg_scale, timesteps_truncate, process_norm_power, height, width)
        if not return_dict:
            return (out,)
        return types.SimpleNamespace(images=out)

    def regione_denoise(self, latents, image_latents, latent_ids, pe, ne, true_cfg_scale, timesteps_truncate,
                        process_norm_power, height, width):
        M = MANAGER
        N = M.inference_step
        sch, tr = self.scheduler, self.transformer
        ts_host = sch.timesteps.detach().to("cpu", torch.float32)
        x, cond = latents[0], image_latents[0]
        L, Cn = x.shape[0], cond.shape[0]
        Tc, Tu = pe.txt_ids.shape[0], ne.txt_ids.shape[0]
        engine = _get_engine(tr, max(Tc, Tu), L, Cn)
        tr.__dict__["_regione_b200_engine"] = engine
        M.refresh(x, cond, latent_ids, pe.txt_ids, 2, self.vae_scale_factor, height, width)       # utils.py:437-465
        M.neg_txt_length = Tu                                                                    # utils.py:445
        for b, emb in enumerate((pe, ne)):
            check(engine.lib.rge_set_pass_text_len(engine._h, b, emb.txt_ids.shape[0]), "rge_set_pass_text_len")
            cos, sin = tr.pos_embed(torch.cat((emb.txt_ids, latent_ids), dim=0))                 # :617-622
            engine.begin_image_rope(cos, sin, b)
        plan = plan_steps(ts_host, gamma, M)
        cache = None
        record = bool(getattr(self, "regione_record", False))
        self.regione_trace = {"modes": [], "latents": [], "noise_pred": []}
        for i in range(N):
            assert i == M.current_step                                                           # :349
            t = ts_host[i]
            skip, ratio = plan[i]
            if skip:                                                                             # :371-375
                if cache.shape[0] != x.shape[0]:
                    cache = ops.gather_rows(cache, M.edited_ids)
                x = sch.step(cache, t, x, return_dict=False, reuse_ratio=ratio)[0]
                self.regione_trace["modes"].append("SKIP")
            else:
                cur = M.current_step
                full = cur <= M.warmup_step - 1 or cur > N - M.post_step - 1 or cur == M.prev_refresh_step
                timestep = t.expand(1).to(x.dtype)                                               # :386

                def forward(e, tag):
                    return self.transformer(hidden_states=x[None], timestep=timestep / 1000, guidance=None,
                                            encoder_hidden_states=e.embedding, prompt_embeds_mask=e.mask,
                                            txt_ids=e.txt_ids, img_ids=latent_ids, text_embeddings=e.text_embeds,
                                            text_mask=e.text_masks,
                                            joint_attention_kwargs={**self._joint_attention_kwargs, "tag": tag},
                                            return_dict=False,
                                            condition_latents=image_latents if full else None)[0][0, : x.shape[0]]
                pos = forward(pe, "cond")                                                        # :388-401
                neg = forward(ne, "uncond")                                                      # :403-419
                if float(t) > timesteps_truncate:                                                # :421-427
                    diff_norm = ops.cfg_diff_norm(pos, neg)
                    denom = self.process_diff_norm(diff_norm.reshape(1, -1, 1), k=process_norm_power)
                    noise_pred = ops.cfg_combine(pos, neg, true_cfg_scale, denom.reshape(-1).to(pos.dtype))
                else:
                    noise_pred = ops.cfg_combine(pos, neg, true_cfg_scale)
                cache = noise_pred                                                               # :431
                x = sch.step(noise_pred, t, x, return_dict=False)[0]
                self.regione_trace["modes"].append("FULL" if full else "REGION")
            x, latent_ids = M.step(x, latent_ids)                                                # :458
            if record:
                self.regione_trace["latents"].append(x.clone())
                self.regione_trace["noise_pred"].append(cache.clone())
        self.regione_trace["edited_ids"] = M.edited_ids
        self.regione_trace["unedited_ids"] = M.unedited_ids
        return x[None]


def warp_modules(pipeline, **args):
    """Step1XEditV1P2/inplace.py:53-62."""
    if "_regione_b200_saved" in pipeline.__dict__:
        unwarp_modules(pipeline)
    MANAGER.set_parameters(args)
    tr = pipeline.transformer
    blocks = list(tr.transformer_blocks) + list(tr.single_transformer_blocks)
    saved = {"cls": pipeline.__class__, "scheduler": pipeline.scheduler, "forward": tr.__dict__.get("forward"),
             "processors": [getattr(b.attn, "processor", None) for b in blocks]}
    pipeline.__dict__["_regione_b200_saved"] = saved
    pipeline.__class__ = type("RegionEStep1XEditPipelineV1P2", (RegionEStep1XEditV1P2PipelineMixin, saved["cls"]), {})
    sch_cls = type("RegionEFlowMatchEulerDiscreteScheduler", (RegionESchedulerMixin, saved["scheduler"].__class__), {})
    pipeline.scheduler = sch_cls.from_config(saved["scheduler"].config)
    pipeline.scheduler._regione_manager = MANAGER
    tr.forward = types.MethodType(RegionEStep1XEditV1P2Transformer2DModelforward, tr)
    for b in tr.transformer_blocks:
        b.attn.set_processor(RegionEB200AttnProcessor(False))
    for b in tr.single_transformer_blocks:
        b.attn.set_processor(RegionEB200AttnProcessor(True))
    return pipeline


def unwarp_modules(pipeline):
    """Step1XEditV1P2/inplace.py:65-73."""
    from .step1x_edit import unwarp_modules as _unwarp
    return _unwarp(pipeline)
