"""Oracle restatement of Step1X-Edit (v1p1) under RegionE: patched forward (RegionE/Step1XEdit/inplace.py:514-571) and
loop with classifier-free guidance stacked on the batch axis (:338-438). TEST INFRASTRUCTURE.

The block stack and the attention processor are FLUX's (same code as oracle/flux.py; the processor :686-816 differs from
FluxKontext's only in names). The front end (`connector`, `time_proj`, `time_embed`, `vec_embed`) and
`process_diff_norm` belong to the fork `Peyton-Chen/diffusers@step1xedit_v1p2`, which is not available: they are taken
as callables from the pipeline under test, exactly where the reference calls them. PARITY UNPINNED for the block math.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import region_ops as ro
from .flux import FluxOracle, rope_cos_sin
from .loop import EulerState, scalar_times, scheduler_step
from .schedule import flow_match_sigmas

GAMMA_STEP1X = [0.9746, 0.9593, 1.0036, 1.0084, 1.0106, 1.0114, 1.0138, 1.0163, 1.0152, 1.0163, 1.0197, 1.0186, 1.0219,
                1.0218, 1.0223, 1.0266, 1.0272, 1.0305, 1.0311, 1.0362, 1.0385, 1.0423, 1.0500, 1.0536, 1.0671, 1.0866,
                1.1015]   # Step1XEdit/inplace.py:47-49


class Step1XOracle(FluxOracle):
    def __init__(self, weights, heads, n_double, n_single, front):
        """front: object with connector / time_proj / time_embed / vec_embed callables (CPU modules)."""
        super().__init__(weights, heads, n_double, n_single, guidance_embeds=False)
        self.front = front

    def forward(self, st, hidden_states, encoder_hidden_states, timestep, prompt_embeds_mask, img_ids, txt_ids):
        f = self.front
        enc, y = f.connector(encoder_hidden_states, timestep, prompt_embeds_mask)          # :514-516
        h = self.lin("x_embedder", hidden_states)                                          # :517
        temb = f.time_embed(f.time_proj(timestep * 1000).to(timestep))                     # :519
        temb = temb + f.vec_embed(y)                                                       # :520
        enc = self.lin("context_embedder", enc)                                            # :521
        rope_q = rope_cos_sin(torch.cat((txt_ids, img_ids), dim=0))                        # :523-524
        rope_k = rope_cos_sin(torch.cat((txt_ids, st.latent_ids), dim=0))                  # :527
        for i in range(self.n_double):
            enc, h = self.double_block(i, h, enc, temb, rope_q, rope_k, st)
        for i in range(self.n_single):
            enc, h = self.single_block(i, h, enc, temb, rope_q, rope_k, st)
        scale, shift = self.lin("norm_out.linear", F.silu(temb).to(h.dtype)).chunk(2, dim=1)
        h = self._ln(h) * (1 + scale)[:, None, :] + shift[:, None, :]
        return self.lin("proj_out", h)                                                     # :566-567


def cfg_norm_processed(pos, neg, true_cfg_scale, t, timesteps_truncate, process_diff_norm, process_norm_power):
    """Step1X classifier-free guidance, RegionE/Step1XEdit/inplace.py:401-410 (same in Step1XEditV1P2): above the
    truncation timestep the guidance term is divided by the fork's `process_diff_norm` of the per-token norm of
    (pos - neg). Pinned by tests/golden/cfg.pt (the reference's own lines exec'd)."""
    if t.item() > timesteps_truncate:
        diff = pos - neg
        diff_norm = torch.norm(diff, dim=(2), keepdim=True)
        return neg + true_cfg_scale * (pos - neg) / process_diff_norm(diff_norm, k=process_norm_power)
    return neg + true_cfg_scale * (pos - neg)


def run_regione_step1x(model, params, latents, image_latents, latent_ids, text_ids, prompt_embeds, prompt_mask,
                       negative_prompt_embeds, negative_mask, true_cfg_scale, process_diff_norm, height, width,
                       timesteps_truncate=0.93, process_norm_power=0.4, record=False):
    """Step1XEdit/inplace.py:331-438 for output_type='latent'."""
    st = ro.RegionState()
    st.set_parameters(params)
    sigmas, timesteps = flow_match_sigmas(params["num_inference_steps"], latents.shape[1])
    sch = EulerState(sigmas, timesteps)
    g = torch.tensor(GAMMA_STEP1X, dtype=torch.float16)
    st.refresh(latents, image_latents, latent_ids, text_ids, height, width)               # :331
    do_cfg = negative_prompt_embeds is not None
    cache, accumulate = None, 1
    trace = {"modes": [], "latents": [], "noise_pred": []}
    for i, t in enumerate(timesteps):
        assert i == st.current_step
        cur, N = st.current_step, st.inference_step
        if cur <= st.warmup_step or cur > N - st.post_step - 1 or cur == st.prev_refresh_step:   # :342-360
            should_cache, accumulate = False, 1
        else:
            ratio = g[i - 1] * (1 + (t - timesteps[i - 1]) / 1000)
            if ratio >= 1:
                should_cache, accumulate = False, 1
            else:
                accumulate = accumulate * ratio
                if 1 - accumulate > st.cache_threshold:
                    should_cache, accumulate = False, 1
                else:
                    should_cache = True
        if should_cache:                                                                   # :362-366
            if cache.shape[1] != latents.shape[1]:
                cache = ro.gather_rows(cache, st.edited_ids)
            noise_pred = scalar_times(ratio, cache)
            trace["modes"].append("SKIP")
        else:
            x_in = latents
            full = cur <= st.warmup_step - 1 or cur > N - st.post_step - 1 or cur == st.prev_refresh_step
            if full:                                                                       # :377-378
                x_in = torch.cat([latents, image_latents], dim=1)
            timestep = t.expand(latents.shape[0]).to(latents.dtype)                        # :379
            if do_cfg:                                                                     # :381-399
                x_in = torch.cat((x_in, x_in), dim=0)
                timestep = torch.cat((timestep, timestep), dim=0)
                embeds = torch.cat((prompt_embeds, negative_prompt_embeds), dim=0)
                masks = torch.cat((prompt_mask, negative_mask), dim=0)
            else:
                embeds, masks = prompt_embeds, prompt_mask
            pred = model.forward(st, x_in, embeds, timestep / 1000, masks, latent_ids, text_ids)
            pred = pred[:, : latents.size(1)]
            if do_cfg:
                noise_pred, neg = pred.chunk(2)
                noise_pred = cfg_norm_processed(noise_pred, neg, true_cfg_scale, t, timesteps_truncate,
                                                process_diff_norm, process_norm_power)
            else:
                noise_pred = pred
            cache = noise_pred
            trace["modes"].append("FULL" if full else "REGION")
        latents = scheduler_step(sch, st, noise_pred, latents, trace)
        latents, latent_ids = st.step(latents, latent_ids)                                 # :438
        if record:
            trace["latents"].append(latents.clone())
            trace["noise_pred"].append(noise_pred.clone())
    trace["edited_ids"], trace["unedited_ids"] = st.edited_ids, st.unedited_ids
    return latents, trace


class Step1XV1P2Oracle(Step1XOracle):
    """Step1XEditV1P2/inplace.py:596-660 forward + :800-890 processor: per-tag caches and text lengths."""

    def __init__(self, weights, heads, n_double, n_single, front):
        super().__init__(weights, heads, n_double, n_single, front)
        self._kc = {"cond": {}, "uncond": {}}
        self._vc = {"cond": {}, "uncond": {}}

    def forward_tag(self, st, hidden_states, e, timestep, img_ids, tag):
        self.k_cache, self.v_cache = self._kc[tag], self._vc[tag]      # k/v_cache_even (cond) / _odd (uncond)
        st.txt_length = e.txt_ids.shape[0]                             # txt_length / neg_txt_length (:833, :868)
        f = self.front
        enc, y = f.connector(e.embedding, timestep, e.mask)
        if getattr(f, "text_token_mapping", None) is not None:         # :606-609
            enc = enc + f.text_token_mapping(e.text_embeds) * e.text_masks[:, :, None].to(enc.dtype)
        h = self.lin("x_embedder", hidden_states)
        temb = f.time_embed(f.time_proj(timestep * 1000).to(timestep)) + f.vec_embed(y)
        enc = self.lin("context_embedder", enc)
        rope_q = rope_cos_sin(torch.cat((e.txt_ids, img_ids), dim=0))
        rope_k = rope_cos_sin(torch.cat((e.txt_ids, st.latent_ids), dim=0))
        for i in range(self.n_double):
            enc, h = self.double_block(i, h, enc, temb, rope_q, rope_k, st)
        for i in range(self.n_single):
            enc, h = self.single_block(i, h, enc, temb, rope_q, rope_k, st)
        scale, shift = self.lin("norm_out.linear", F.silu(temb).to(h.dtype)).chunk(2, dim=1)
        h = self._ln(h) * (1 + scale)[:, None, :] + shift[:, None, :]
        return self.lin("proj_out", h)


def run_regione_step1x_v1p2(model, params, gamma, latents, image_latents, latent_ids, pe, ne, true_cfg_scale,
                            process_diff_norm, height, width, timesteps_truncate=0.93, process_norm_power=0.4,
                            record=False):
    """Step1XEditV1P2/inplace.py:347-458 (denoising part) for output_type='latent'."""
    st = ro.RegionState()
    st.set_parameters(params)
    sigmas, timesteps = flow_match_sigmas(params["num_inference_steps"], latents.shape[1])
    sch = EulerState(sigmas, timesteps)
    g = torch.tensor(gamma, dtype=torch.float16)
    st.refresh(latents, image_latents, latent_ids, pe.txt_ids, height, width)
    cache, accumulate = None, 1
    trace = {"modes": [], "latents": [], "noise_pred": []}
    for i, t in enumerate(timesteps):
        assert i == st.current_step
        cur, N = st.current_step, st.inference_step
        if cur <= st.warmup_step or cur > N - st.post_step - 1 or cur == st.prev_refresh_step:
            should_cache, accumulate = False, 1
        else:
            ratio = g[i - 1] * (1 + (t - timesteps[i - 1]) / 1000)
            if ratio >= 1:
                should_cache, accumulate = False, 1
            else:
                accumulate = accumulate * ratio
                if 1 - accumulate > st.cache_threshold:
                    should_cache, accumulate = False, 1
                else:
                    should_cache = True
        if should_cache:
            if cache.shape[1] != latents.shape[1]:
                cache = ro.gather_rows(cache, st.edited_ids)
            noise_pred = scalar_times(ratio, cache)
            trace["modes"].append("SKIP")
        else:
            x_in = latents
            full = cur <= st.warmup_step - 1 or cur > N - st.post_step - 1 or cur == st.prev_refresh_step
            if full:
                x_in = torch.cat([latents, image_latents], dim=1)
            timestep = t.expand(latents.shape[0]).to(latents.dtype)
            pos = model.forward_tag(st, x_in, pe, timestep / 1000, latent_ids, "cond")[:, : latents.size(1)]
            neg = model.forward_tag(st, x_in, ne, timestep / 1000, latent_ids, "uncond")[:, : latents.size(1)]
            noise_pred = cfg_norm_processed(pos, neg, true_cfg_scale, t, timesteps_truncate, process_diff_norm,
                                            process_norm_power)
            cache = noise_pred
            trace["modes"].append("FULL" if full else "REGION")
        latents = scheduler_step(sch, st, noise_pred, latents, trace)
        latents, latent_ids = st.step(latents, latent_ids)
        if record:
            trace["latents"].append(latents.clone())
            trace["noise_pred"].append(noise_pred.clone())
    trace["edited_ids"], trace["unedited_ids"] = st.edited_ids, st.unedited_ids
    return latents, trace
