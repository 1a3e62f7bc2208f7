"""Oracle restatement of RegionE's denoising loop and patched scheduler step. TEST INFRASTRUCTURE.

Follows RegionE/FluxKontext/inplace.py: loop body :287-392 (AVDC :295-318), scheduler.step :594-691.

Scalar rounding: the reference multiplies 0-dim fp32 CUDA tensors (dt, dt_final, ratio) with bf16 tensors. On CUDA the
0-dim operand is first cast to the common dtype bf16 (TensorIterator dynamic cast), on CPU the outcome depends on the
operand order. The reference runs on CUDA, so `scalar_times` reproduces the CUDA behaviour explicitly
(measured with tools/scalar_semantics.py on the B200 box).
"""
from __future__ import annotations

import torch

from . import region_ops as ro
from .schedule import flow_match_sigmas

SCALAR_ROUNDS_TO_TENSOR_DTYPE = True


def scalar_times(s: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """`s * v` for a 0-dim fp32 tensor s and a bf16 tensor v, with CUDA semantics; result has v's dtype."""
    if SCALAR_ROUNDS_TO_TENSOR_DTYPE:
        return s.to(v.dtype) * v
    return (v.float() * s).to(v.dtype)


class EulerState:
    """The scheduler fields the reference reads."""

    def __init__(self, sigmas, timesteps):
        self.sigmas, self.timesteps, self.step_index = sigmas, timesteps, 0


def scheduler_step(sch: EulerState, st: ro.RegionState, model_output, sample, trace=None):
    """inplace.py:594-691 (non-stochastic branch)."""
    sample = sample.to(torch.float32)                                                    # :610
    i = sch.step_index
    sigma, sigma_next = sch.sigmas[i], sch.sigmas[i + 1]
    if st.current_step == st.warmup_step - 1:                                            # :630-634
        st.prev_refresh_step = st.refresh_step_real_time.pop(0) - 1
        dt_final = sch.sigmas[-1] - sigma
        dt_direct = sch.sigmas[st.prev_refresh_step] - sigma
    elif st.prev_refresh_step is not None and st.current_step == st.prev_refresh_step and st.refresh_step_real_time:
        st.next_refresh_step = st.refresh_step_real_time.pop(0) - 1                       # :636-639
        dt_direct = sch.sigmas[st.next_refresh_step] - sigma
    dt = sigma_next - sigma                                                              # :641
    two_speed = False
    if st.current_step == st.warmup_step - 1:                                            # :648-651
        estimate = sample + scalar_times(dt_final, model_output)
        st.edited_ids, st.unedited_ids, raw, final, sim = ro.select_tokens(
            estimate, st.condition_latent, st.threshold, st.height // 16, st.width // 16, st.erosion_dilation)
        if trace is not None:
            trace.update(raw_mask=raw, final_mask=final, similarity=sim, estimate=estimate)
        two_speed = True
    elif st.prev_refresh_step is not None and st.current_step == st.prev_refresh_step:   # :665
        two_speed = True
    if two_speed:                                                                        # :653-663 / :667-677
        e = ro.gather_rows(sample, st.edited_ids) + scalar_times(dt, ro.gather_rows(model_output, st.edited_ids))
        u = ro.gather_rows(sample, st.unedited_ids) + scalar_times(dt_direct,
                                                                   ro.gather_rows(model_output, st.unedited_ids))
        prev = torch.zeros_like(sample)
        ro.scatter_rows(e, st.edited_ids, prev)
        ro.scatter_rows(u, st.unedited_ids, prev)
    else:
        prev = sample + scalar_times(dt, model_output)                                   # :680
    sch.step_index += 1
    return prev.to(model_output.dtype)                                                   # :686


def run_regione(model, params: dict, gamma, latents, image_latents, latent_ids, text_ids, prompt_embeds, pooled,
                guidance_scale: float, height: int, width: int, record: bool = False, negative=None):
    """inplace.py:287-392 for output_type='latent'. Returns (final latents [1,L,64], trace dict).
    negative = (negative_prompt_embeds, negative_pooled_prompt_embeds, true_cfg_scale): the second forward of true-CFG
    (:349-364). The model's processors own ONE k/v cache (:700-702), so both forwards patch and read the same cache -
    the negative forward runs second and its rows are what a later REGION step of the prompt forward finds for the
    unedited / instruction tokens."""
    st = ro.RegionState()
    st.set_parameters(params)
    n_steps = params["num_inference_steps"]
    sigmas, timesteps = flow_match_sigmas(n_steps, latents.shape[1])
    sigmas, timesteps = sigmas.to(latents.device), timesteps.to(latents.device)
    sch = EulerState(sigmas, timesteps)
    g = torch.tensor(gamma, dtype=torch.float16, device=latents.device)
    guidance = torch.full([1], guidance_scale, dtype=torch.float32, device=latents.device).expand(latents.shape[0])
    st.refresh(latents, image_latents, latent_ids, text_ids, height, width)              # :287
    cache, accumulate = None, 1
    trace = {"modes": [], "latents": [], "noise_pred": []}
    for i, t in enumerate(timesteps):
        assert i == st.current_step                                                      # :293
        cur, N = st.current_step, st.inference_step
        if cur <= st.warmup_step or cur > N - st.post_step - 1 or cur == st.prev_refresh_step:   # :295-313
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
        if should_cache:                                                                 # :315-318
            if cache.shape[1] != latents.shape[1]:
                cache = ro.gather_rows(cache, st.edited_ids)
            noise_pred = scalar_times(ratio, cache)
            trace["modes"].append("SKIP")
        else:
            x_in = latents
            full = cur <= st.warmup_step - 1 or cur > N - st.post_step - 1 or cur == st.prev_refresh_step
            if full:                                                                     # :331-332
                x_in = torch.cat([latents, image_latents], dim=1)
            timestep = t.expand(latents.shape[0]).to(latents.dtype)                      # :334 (bf16-rounded)
            noise_pred = model.forward(st, x_in, prompt_embeds, pooled, timestep / 1000, latent_ids, text_ids,
                                       guidance)[:, : latents.size(1)]                   # :336-347
            if negative is not None:                                                     # :349-364
                neg = model.forward(st, x_in, negative[0], negative[1], timestep / 1000, latent_ids, text_ids,
                                    guidance)[:, : latents.size(1)]
                noise_pred = neg + negative[2] * (noise_pred - neg)
            cache = noise_pred                                                           # :365
            trace["modes"].append("FULL" if full else "REGION")
        latents = scheduler_step(sch, st, noise_pred, latents, trace)                    # :369
        latents, latent_ids = st.step(latents, latent_ids)                               # :392
        if record:
            trace["latents"].append(latents.clone())
            trace["noise_pred"].append(noise_pred.clone())
    trace["edited_ids"], trace["unedited_ids"] = st.edited_ids, st.unedited_ids
    return latents, trace
