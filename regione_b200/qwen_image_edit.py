"""Qwen-Image-Edit variant of the plugin (also serves QwenImageEditPlusPipeline's denoising loop): host side of
RegionE/QwenImageEdit/inplace.py over the CUDA library.

Differences to the FLUX variant, all taken from the reference: 60 dual-stream blocks only (inplace.py:58-59); two
transformer passes per step, tagged 'cond' / 'uncond', each with its own per-layer K/V cache (processor :747-815 —
`k_cache_even` / `k_cache_odd`; here pass 0 / pass 1 of the library handle); norm-rescaled classifier-free guidance
(:396-405, `rge_cfg_rescale`); rotary frequencies from the pipeline's own `pos_embed(img_shapes, txt_seq_lens)` with the
image-query rows gathered through `latent_ids` (:530-531, :851) — here the edited-id selection; 1-D `latent_ids =
arange(L + C)` (:322). Scheduler step, partition, AVDC planning and the split / merge state machine are shared with
flux_kontext.py.
"""
from __future__ import annotations

import types

import numpy as np
import torch

from . import ops
from .engine import cached_engine
from .engine_qwen import QwenEngine
from .flux_kontext import LATENT_SPACE_ONLY, RegionEB200AttnProcessor, RegionESchedulerMixin, calculate_shift, retrieve_timesteps
from .manager import RegionManager, plan_steps
from .params import GAMMA

gamma = GAMMA["QwenImageEditPipeline"]        # QwenImageEdit/inplace.py:47-50
MANAGER = RegionManager()                     # :51


def _get_engine(transformer, T, L, C) -> QwenEngine:
    return cached_engine(transformer, (T, L, C, 2), lambda: QwenEngine(transformer, T, L, C, n_pass=2))


def RegionEQwenImageTransformer2DModelforward(self, hidden_states, encoder_hidden_states=None,
                                              encoder_hidden_states_mask=None, timestep=None, img_shapes=None,
                                              txt_seq_lens=None, guidance=None, attention_kwargs=None,
                                              controlnet_block_samples=None, latent_ids=None, return_dict=True,
                                              condition_latents=None):
    """Signature of the reference's patched forward (QwenImageEdit/inplace.py:462-475) plus `condition_latents` (FULL
    steps: the instruction-image latent, read in place instead of concatenated, :366-367). `attention_kwargs['tag']`
    selects the pass ('cond' -> cache set 0, 'uncond' -> cache set 1), as in the processor (:747, :784)."""
    if controlnet_block_samples is not None:
        raise NotImplementedError("regione_b200: ControlNet residuals are outside the hot path")
    engine = self.__dict__.get("_regione_b200_engine")
    if engine is None:
        raise RuntimeError("regione_b200: no image in flight — the pipeline loop begins the image first")
    if hidden_states.shape[0] != 1:
        raise NotImplementedError("regione_b200: batch size must be 1")
    tag = (attention_kwargs or {}).get("tag", "cond")
    if tag not in ("cond", "uncond"):
        raise NotImplementedError(f"Error tag: {tag}")
    M = MANAGER
    # time_proj has scale 1000 (diffusers Timesteps(scale=1000)); the bf16 timestep is what the reference feeds (:517)
    t_x1000 = float(timestep.to(hidden_states.dtype).float().reshape(-1)[0]) * 1000.0
    x = hidden_states[0]
    x_cond = None if condition_latents is None else condition_latents[0]
    full = x.shape[0] + (0 if x_cond is None else x_cond.shape[0]) == M.latent_length + M.condition_length
    if full:
        sel, n_out = None, M.latent_length
    else:
        if x_cond is not None or M.edited_ids is None or x.shape[0] != M.edited_ids.numel():
            raise RuntimeError("regione_b200: region step without a matching edited-token selection")
        sel, n_out = M.edited_ids, x.shape[0]
    out = engine.step(x, sel, t_x1000, n_out, pass_id=0 if tag == "cond" else 1, x_cond=x_cond)[None]
    if not return_dict:
        return (out,)
    return types.SimpleNamespace(sample=out)


class RegionEQwenImageEditPipelineMixin:
    """`RegionEQwenImageEditPipeline.__call__` (QwenImageEdit/inplace.py:76-449), latent-space entry: packed `latents`
    [1,L,64] and `image_latents` [1,C,64], `prompt_embeds` [1,T,ctx] (+ `negative_prompt_embeds` for true CFG)."""

    @torch.no_grad()
    def __call__(self, image=None, prompt=None, negative_prompt=None, true_cfg_scale=4.0, height=None, width=None,
                 num_inference_steps=28, guidance_scale=None, latents=None, prompt_embeds=None,
                 prompt_embeds_mask=None, negative_prompt_embeds=None, negative_prompt_embeds_mask=None,
                 output_type="pil", return_dict=True, attention_kwargs=None, image_latents=None,
                 condition_height=None, condition_width=None, **unused):
        assert num_inference_steps == MANAGER.inference_step, "num_inference_steps should be equal to 28"
        if image_latents is None or latents is None or prompt_embeds is None or output_type != "latent":
            raise NotImplementedError(LATENT_SPACE_ONLY)
        if height is None or width is None:
            raise ValueError("height and width are required with packed latents")
        device = self._execution_device
        self._attention_kwargs = attention_kwargs or {}
        do_true_cfg = true_cfg_scale > 1 and negative_prompt_embeds is not None                      # :208-211
        ch, cw = condition_height or height, condition_width or width
        sf = self.vae_scale_factor
        img_shapes = [[(1, height // sf // 2, width // sf // 2), (1, ch // sf // 2, cw // sf // 2)]]   # :273-278
        sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)                      # :281-296
        cfg = self.scheduler.config
        mu = calculate_shift(latents.shape[1], cfg.get("base_image_seq_len", 256), cfg.get("max_image_seq_len", 4096),
                             cfg.get("base_shift", 0.5), cfg.get("max_shift", 1.15))
        retrieve_timesteps(self.scheduler, num_inference_steps, device, sigmas=sigmas, mu=mu)
        self.scheduler.set_begin_index(0)                                                            # :326
        self.scheduler._step_index = 0
        txt_lens = [prompt_embeds.shape[1]] if prompt_embeds_mask is None else prompt_embeds_mask.sum(dim=1).tolist()
        neg_lens = None
        if do_true_cfg:
            neg_lens = ([negative_prompt_embeds.shape[1]] if negative_prompt_embeds_mask is None
                        else negative_prompt_embeds_mask.sum(dim=1).tolist())
        out = self.regione_denoise(latents, image_latents, prompt_embeds, negative_prompt_embeds if do_true_cfg else None,
                                   true_cfg_scale, img_shapes, txt_lens, neg_lens, height, width)
        if not return_dict:
            return (out,)
        return types.SimpleNamespace(images=out)

    def regione_denoise(self, latents, image_latents, prompt_embeds, negative_prompt_embeds, true_cfg_scale,
                        img_shapes, txt_seq_lens, negative_txt_seq_lens, height, width):
        """The hot loop, QwenImageEdit/inplace.py:322-433."""
        M = MANAGER
        N = M.inference_step
        sch = self.scheduler
        tr = self.transformer
        ts_host = sch.timesteps.detach().to("cpu", torch.float32)
        x, cond = latents[0], image_latents[0]
        L, C = x.shape[0], cond.shape[0]
        do_cfg = negative_prompt_embeds is not None
        T = prompt_embeds.shape[1]
        if do_cfg and negative_prompt_embeds.shape[1] != T:
            raise NotImplementedError("regione_b200: cond / uncond prompts must be padded to one length")
        engine = _get_engine(tr, T, L, C)
        tr.__dict__["_regione_b200_engine"] = engine
        latent_ids = torch.arange(L + C, device=x.device)                                            # :322
        M.refresh(x, cond, latent_ids, torch.empty(T, 0), 2, self.vae_scale_factor, height, width)   # :323
        img_freqs, txt_freqs = tr.pos_embed(img_shapes, txt_seq_lens, device=x.device)               # :530
        engine.begin_image_qwen(img_freqs, txt_freqs[:T], prompt_embeds[0], 0)
        if do_cfg:
            nimg, ntxt = tr.pos_embed(img_shapes, negative_txt_seq_lens, device=x.device)
            engine.begin_image_qwen(nimg, ntxt[:T], negative_prompt_embeds[0], 1)
        plan = plan_steps(ts_host, gamma, M)                                                         # :334-350
        cache = None
        record = bool(getattr(self, "regione_record", False))
        self.regione_trace = {"modes": [], "latents": [], "noise_pred": []}
        for i in range(N):
            assert i == M.current_step                                                               # :332
            t = ts_host[i]
            skip, ratio = plan[i]
            if skip:                                                                                 # :352-356
                if cache.shape[0] != x.shape[0]:
                    cache = ops.gather_rows(cache, M.edited_ids)
                x = sch.step(cache, t, x, return_dict=False, reuse_ratio=ratio)[0]
                self.regione_trace["modes"].append("SKIP")
            else:
                cur = M.current_step
                full = cur <= M.warmup_step - 1 or cur > N - M.post_step - 1 or cur == M.prev_refresh_step   # :365
                timestep = t.expand(1).to(x.dtype)                                                   # :369

                def forward(embeds, lens, tag):
                    return self.transformer(hidden_states=x[None], timestep=timestep / 1000, guidance=None,
                                            encoder_hidden_states_mask=None, encoder_hidden_states=embeds,
                                            img_shapes=img_shapes, txt_seq_lens=lens, latent_ids=latent_ids,
                                            attention_kwargs={**self._attention_kwargs, "tag": tag},
                                            return_dict=False,
                                            condition_latents=image_latents if full else None)[0][0, : x.shape[0]]
                noise_pred = forward(prompt_embeds, txt_seq_lens, "cond")                            # :371-384
                if do_cfg:                                                                           # :386-405
                    neg = forward(negative_prompt_embeds, negative_txt_seq_lens, "uncond")
                    noise_pred = ops.cfg_rescale(noise_pred, neg, true_cfg_scale)
                cache = noise_pred                                                                   # :406
                x = sch.step(noise_pred, t, x, return_dict=False)[0]
                self.regione_trace["modes"].append("FULL" if full else "REGION")
            x, latent_ids = M.step(x, latent_ids)                                                    # :433
            if record:
                self.regione_trace["latents"].append(x.clone())
                self.regione_trace["noise_pred"].append(cache.clone())
        self.regione_trace["edited_ids"] = M.edited_ids
        self.regione_trace["unedited_ids"] = M.unedited_ids
        return x[None]


def warp_modules(pipeline, **args):
    """QwenImageEdit/inplace.py:53-61."""
    if "_regione_b200_saved" in pipeline.__dict__:
        unwarp_modules(pipeline)
    MANAGER.set_parameters(args)
    global gamma   # QwenImageEditPlusPipeline shares this loop but has its own fitted table
    gamma = GAMMA.get(pipeline.__class__.__name__, GAMMA["QwenImageEditPipeline"])
    tr = pipeline.transformer
    saved = {"cls": pipeline.__class__, "scheduler": pipeline.scheduler, "forward": tr.__dict__.get("forward"),
             "processors": [getattr(b.attn, "processor", None) for b in tr.transformer_blocks]}
    pipeline.__dict__["_regione_b200_saved"] = saved
    pipeline.__class__ = type("RegionEQwenImageEditPipeline", (RegionEQwenImageEditPipelineMixin, saved["cls"]), {})
    sch_cls = type("RegionEFlowMatchEulerDiscreteScheduler", (RegionESchedulerMixin, saved["scheduler"].__class__), {})
    pipeline.scheduler = sch_cls.from_config(saved["scheduler"].config)
    pipeline.scheduler._regione_manager = MANAGER
    tr.forward = types.MethodType(RegionEQwenImageTransformer2DModelforward, tr)
    for block in tr.transformer_blocks:
        block.attn.set_processor(RegionEB200AttnProcessor(False))
    return pipeline


def unwarp_modules(pipeline):
    """QwenImageEdit/inplace.py:64-71."""
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
    for b, p in zip(tr.transformer_blocks, saved["processors"]):
        b.attn.set_processor(p)
    for eng in tr.__dict__.pop("_regione_b200_engines", {}).values():
        eng.close()
    tr.__dict__.pop("_regione_b200_engine", None)
    return pipeline
