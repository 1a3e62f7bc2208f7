"""FLUX.1-Kontext variant of the plugin: `warp_modules` / `unwarp_modules`, the patched pipeline loop, the patched
transformer forward and the patched scheduler step — the host side of RegionE/FluxKontext/inplace.py, with every
tensor operation delegated to the CUDA library (regione_b200/csrc) through the C ABI.

Host-side departures from the reference, none of which change results:
  * AVDC decisions are planned before the loop from the host copy of the schedule (manager.plan_steps), removing the
    1-2 device->host syncs per step of inplace.py:301-308; the only sync left per image is the edited-token count
    after the partition (the reference syncs there too, utils.py:347);
  * the K/V cache, attention processors' state and all workspaces live inside the library handle (engine.FluxEngine);
  * latents are handled as 2-D [tokens, channels] views of the reference's [1, tokens, channels].
"""
from __future__ import annotations

import types

import numpy as np
import torch

from . import ops
from .engine import FluxEngine, cached_engine
from .manager import RegionManager, plan_steps
from .params import GAMMA, SCALAR_ROUNDS_TO_BF16
from .schedule import calculate_shift, retrieve_timesteps

gamma = GAMMA["FluxKontextPipeline"]          # inplace.py:47-50

LATENT_SPACE_ONLY = (
    "regione_b200: the patched pipeline of this family is LATENT-SPACE ONLY - it replaces the denoising loop and the "
    "transformer forward, not the image processor / text encoders / VAE. Call it with packed `latents`, `image_latents` "
    "and prompt embeds and `output_type='latent'`, encoding and decoding with the un-patched pipeline's own methods; "
    "only the FluxKontext variant wires the host pipeline's encode_prompt / prepare_latents / VAE decode "
    "(pipe(image=..., prompt=...)).")
MANAGER = RegionManager()                     # inplace.py:51 (module-global singleton, one pipeline per process)


_MASK_ALLGATHER = False


def enable_mask_allgather(on: bool = True) -> None:
    """Multi-GPU (one image stream per rank): after the partition kernel every rank all-gathers the raw
    mask bytes over NCCL so each holds the [world, L] batch mask (`MANAGER.batch_masks`); the only collective on the
    path. Requires torch.distributed to be initialised with the nccl backend (gloo in the CPU tests)."""
    global _MASK_ALLGATHER
    _MASK_ALLGATHER = bool(on)


def _world() -> int:
    import torch.distributed as dist
    if not (_MASK_ALLGATHER and dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size()


def allgather_masks(mask: torch.Tensor):
    """[L] uint8 per rank -> [world, L] on every rank (no-op without an initialised process group). Blocking form."""
    import torch.distributed as dist
    if _world() == 1:
        return mask[None]
    out = torch.empty(dist.get_world_size() * mask.numel(), dtype=mask.dtype, device=mask.device)
    dist.all_gather_into_tensor(out, mask.contiguous())
    return out.view(dist.get_world_size(), mask.numel())


class PendingMasks:
    """The [world, L] batch mask while its all-gather is still in flight on a side stream. `get()` makes the CURRENT
    stream wait for the collective (no host block) and returns the tensor."""

    def __init__(self, out, work, side):
        self.out, self.work, self.side = out, work, side

    def get(self):
        if self.work is not None:
            self.work.wait()
            if self.side is not None:
                torch.cuda.current_stream().wait_stream(self.side)
            self.work = None
        return self.out


_SIDE_STREAM = {}


def allgather_masks_async(mask: torch.Tensor) -> PendingMasks:
    """Starts the all-gather of this rank's [L] partition mask on a side stream ordered after the work already queued
    on the current stream (the kernel that produced the mask), so that the collective overlaps the two-speed Euler
    update and the host's wait for the token counts. CPU tensors (gloo, tests) gather asynchronously too."""
    import torch.distributed as dist
    if _world() == 1:
        return PendingMasks(mask[None], None, None)
    world = dist.get_world_size()
    out = torch.empty(world * mask.numel(), dtype=mask.dtype, device=mask.device)
    if not mask.is_cuda:
        work = dist.all_gather_into_tensor(out, mask.contiguous(), async_op=True)
        return PendingMasks(out.view(world, mask.numel()), work, None)
    side = _SIDE_STREAM.get(mask.device)
    if side is None:
        side = _SIDE_STREAM[mask.device] = torch.cuda.Stream(device=mask.device)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        work = dist.all_gather_into_tensor(out, mask.contiguous(), async_op=True)
    mask.record_stream(side)
    out.record_stream(side)
    return PendingMasks(out.view(world, mask.numel()), work, side)


def _scalar(s: torch.Tensor) -> float:
    """A 0-dim fp32 schedule scalar as the CUDA reference would apply it to a bf16 tensor (params.py)."""
    return float(s.to(torch.bfloat16)) if SCALAR_ROUNDS_TO_BF16 else float(s)


class RegionEB200AttnProcessor:
    """Placeholder installed on every attention module by `warp_modules` (the reference installs
    RegoionEFluxAttnProcessor2_0 there, inplace.py:58-61). The per-layer K/V cache that processor owned now lives in
    the library handle, and attention runs inside `rge_dit_step`, so calling the processor is an error."""

    def __init__(self, single: bool):
        self.single = single

    def __call__(self, *a, **k):
        raise RuntimeError("regione_b200: attention runs inside the CUDA library (rge_dit_step), not per module")


# ---------------------------------------------------------------------------------------------- patched scheduler
class RegionESchedulerMixin:
    """`RegionEFlowMatchEulerDiscreteScheduler.step` (inplace.py:581-691) over the library kernels."""

    def _host_sigmas(self):
        hs = getattr(self, "_regione_host_sigmas", None)
        if hs is None or hs.shape != self.sigmas.shape:
            hs = self.sigmas.detach().to("cpu", torch.float32)
            self._regione_host_sigmas = hs
        return hs

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None, timesteps=None):
        # explicit signature: retrieve_timesteps inspects it for `sigmas` / `timesteps` (utils.py:84-99)
        kw = {} if timesteps is None else {"timesteps": timesteps}
        super().set_timesteps(num_inference_steps, device=device, sigmas=sigmas, mu=mu, **kw)
        self._regione_host_sigmas = None

    def step(self, model_output, timestep, sample, *args, return_dict=True, reuse_ratio=None, **kwargs):
        if isinstance(timestep, int):
            raise ValueError("pass one of scheduler.timesteps, not an integer index")            # :594-605
        if self.step_index is None:
            self._init_step_index(timestep)                                                      # :606-607
        M = getattr(self, "_regione_manager", MANAGER)   # each family module owns its singleton
        sig = self._host_sigmas()
        idx = self.step_index
        sigma, sigma_next = sig[idx], sig[idx + 1]
        dt_final = dt_direct = None
        if M.current_step == M.warmup_step - 1:                                                  # :630-634
            M.prev_refresh_step = M.refresh_step_real_time.pop(0) - 1
            dt_final = sig[-1] - sigma
            dt_direct = sig[M.prev_refresh_step] - sigma
        elif M.prev_refresh_step is not None and M.current_step == M.prev_refresh_step and M.refresh_step_real_time:
            M.next_refresh_step = M.refresh_step_real_time.pop(0) - 1                             # :636-639
            dt_direct = sig[M.next_refresh_step] - sigma
        dt = sigma_next - sigma                                                                  # :641
        shape = sample.shape
        ch = shape[-1]
        x, v = sample.reshape(-1, ch), model_output.reshape(-1, ch)
        ratio = None if reuse_ratio is None else _scalar(reuse_ratio)
        if M.current_step == M.warmup_step - 1:                                                  # :648-663
            raw = ops.partition(x, v, M.condition_latent.reshape(-1, ch), _scalar(dt_final), float(M.threshold))
            gh = M.height // (M.patch_size * M.vae_scale_factor)
            gw = M.width // (M.patch_size * M.vae_scale_factor)
            box = {}

            def after_mask(final_mask):
                # needs the mask but not the token counts: enqueued BEFORE the host waits for the counts. The
                # all-gather of the partition (one image per rank -> the [world, L] batch partition, the path's only
                # collective) runs on a side stream beside the Euler update.
                box["prev"] = ops.euler(x, v, _scalar(dt), _scalar(dt_direct), edited_mask=final_mask)
                box["masks"] = allgather_masks_async(final_mask)

            M.edited_mask, M.edited_ids, M.unedited_ids = ops.compact(raw, gh, gw, bool(M.erosion_dilation),
                                                                      before_sync=after_mask)
            M.batch_masks_pending = box["masks"]       # [world, L] once `M.batch_masks` is read
            prev = box["prev"]
        elif M.prev_refresh_step is not None and M.current_step == M.prev_refresh_step:          # :665-677
            prev = ops.euler(x, v, _scalar(dt), _scalar(dt_direct) if dt_direct is not None else 0.0,
                             edited_mask=M.edited_mask)
        else:                                                                                    # :680, :318
            prev = ops.euler(x, v, _scalar(dt), reuse_ratio=ratio)
        self._step_index += 1                                                                    # :683
        prev = prev.reshape(shape)
        if not return_dict:
            return (prev,)
        return types.SimpleNamespace(prev_sample=prev)


# ---------------------------------------------------------------------------------------------- patched forward
def _get_engine(transformer, T, L, C, n_pass=1) -> FluxEngine:
    # n_pass = 2 is true-CFG: two text contexts over ONE K/V cache set, like the reference's single processor cache
    return cached_engine(transformer, (T, L, C, n_pass),
                         lambda: FluxEngine(transformer, T, L, C, n_pass, shared_cache=n_pass > 1))


def RegionEFluxTransformer2DModelforward(self, hidden_states, encoder_hidden_states=None, pooled_projections=None,
                                         timestep=None, img_ids=None, txt_ids=None, guidance=None,
                                         joint_attention_kwargs=None, controlnet_block_samples=None,
                                         controlnet_single_block_samples=None, return_dict=True,
                                         controlnet_blocks_repeat=False, condition_latents=None, regione_pass=0):
    """Same signature as the reference's patched forward (inplace.py:413-427) plus `condition_latents`: on FULL steps
    the loop hands the instruction-image latent separately ([1,C,64]) and the library reads the two row ranges in
    place, instead of `torch.cat([latents, image_latents], dim=1)` (inplace.py:332; a concatenated `hidden_states` is
    still accepted), and `regione_pass` (1 = the negative-prompt forward of true-CFG, inplace.py:349-364: its own text
    context, the SAME K/V cache). Mode selection follows the processor's rule (inplace.py:717-732): all L+C image tokens present
    -> FULL (cache rows of every token are rewritten); fewer -> REGION with selection = MANAGER.edited_ids."""
    if controlnet_block_samples is not None or controlnet_single_block_samples is not None:
        raise NotImplementedError("regione_b200: ControlNet residuals are outside the hot path")
    if joint_attention_kwargs and "ip_adapter_image_embeds" in joint_attention_kwargs:
        raise NotImplementedError("regione_b200: IP-Adapter is outside the hot path")
    engine = self.__dict__.get("_regione_b200_engine")
    if engine is None:
        raise RuntimeError("regione_b200: no image in flight — the pipeline loop calls begin_image first")
    if hidden_states.shape[0] != 1:
        raise NotImplementedError("regione_b200: batch size must be 1 (the reference's partition is B=1, SURVEY C-10)")
    M = MANAGER
    t_x1000 = float((timestep.to(hidden_states.dtype) * 1000).reshape(-1)[0])                    # :471
    x = hidden_states[0]
    x_cond = None if condition_latents is None else condition_latents[0]
    full = x.shape[0] + (0 if x_cond is None else x_cond.shape[0]) == M.latent_length + M.condition_length
    if full:
        sel, n_out = None, M.latent_length
    else:
        if x_cond is not None or M.edited_ids is None or x.shape[0] != M.edited_ids.numel():
            raise RuntimeError("regione_b200: region step without a matching edited-token selection")
        sel, n_out = M.edited_ids, x.shape[0]
    out = engine.step(x, sel, t_x1000, n_out, pass_id=int(regione_pass), x_cond=x_cond)[None]
    if not return_dict:
        return (out,)
    return types.SimpleNamespace(sample=out)


# ---------------------------------------------------------------------------------------------- patched pipeline
class RegionEFluxKontextPipelineMixin:
    """`RegionEFluxKontextPipeline.__call__` (inplace.py:76-410). Latent-space entry: pass packed `latents` [1,L,64],
    packed `image_latents` [1,C,64], `prompt_embeds` [1,T,ctx] and `pooled_prompt_embeds` [1,pooled] (bf16, CUDA) with
    `output_type="latent"`; pixel-space pre/post-processing (image processor, text encoders, VAE) stays the host
    pipeline's own code and is used when the pipeline provides it."""

    @torch.no_grad()
    def __call__(self, image=None, prompt=None, prompt_2=None, height=None, width=None, num_inference_steps=28,
                 guidance_scale=3.5, num_images_per_prompt=1, generator=None, latents=None, prompt_embeds=None,
                 pooled_prompt_embeds=None, output_type="pil", return_dict=True, joint_attention_kwargs=None,
                 max_sequence_length=512, max_area=1024 ** 2, _auto_resize=True, image_latents=None,
                 true_cfg_scale=1.0, **unused):
        assert num_inference_steps == MANAGER.inference_step, "num_inference_steps should be equal to 28"   # :112
        # arguments of the reference's __call__ that this path does not implement are rejected, not swallowed
        negative_prompt, negative_prompt_2 = unused.get("negative_prompt"), unused.get("negative_prompt_2")
        negative_prompt_embeds = unused.get("negative_prompt_embeds")
        negative_pooled_prompt_embeds = unused.get("negative_pooled_prompt_embeds")
        has_neg_prompt = negative_prompt is not None or (                                                   # :176-178
            negative_prompt_embeds is not None and negative_pooled_prompt_embeds is not None)
        do_true_cfg = true_cfg_scale > 1 and has_neg_prompt                                                 # :180
        for k in ("ip_adapter_image", "ip_adapter_image_embeds", "negative_ip_adapter_image",
                  "negative_ip_adapter_image_embeds", "sigmas"):
            if unused.get(k) is not None:
                raise NotImplementedError(f"regione_b200: `{k}` is outside the hot path and not supported")
        known = {"negative_prompt", "negative_prompt_2", "negative_prompt_embeds", "negative_pooled_prompt_embeds",
                 "ip_adapter_image", "ip_adapter_image_embeds", "negative_ip_adapter_image",
                 "negative_ip_adapter_image_embeds", "callback_on_step_end", "callback_on_step_end_tensor_inputs",
                 "sigmas"}
        extra = set(unused) - known
        if extra:
            raise TypeError(f"__call__() got unexpected keyword arguments {sorted(extra)}")
        device = self._execution_device
        multiple_of = self.vae_scale_factor * 2
        self._guidance_scale = guidance_scale
        self._joint_attention_kwargs = joint_attention_kwargs
        self._interrupt = False
        if image_latents is None:
            # pixel-space path: the pipeline's own preprocessing / encoders, same calls as inplace.py:113-226
            if not hasattr(self, "encode_prompt") or not hasattr(self, "prepare_latents"):
                raise RuntimeError("this pipeline has no encoders/VAE: pass latents, image_latents and prompt embeds")
            if image is not None and not (isinstance(image, torch.Tensor) and image.size(1) == self.latent_channels):
                img = image[0] if isinstance(image, list) else image
                ih, iw = self.image_processor.get_default_height_width(img)
                if _auto_resize:
                    from .resolutions import nearest_kontext_resolution
                    iw, ih = nearest_kontext_resolution(iw / ih)
                iw, ih = iw // multiple_of * multiple_of, ih // multiple_of * multiple_of
                image = self.image_processor.resize(image, ih, iw)
                image = self.image_processor.preprocess(image, ih, iw)
                height, width = image.shape[-2], image.shape[-1]
            else:                                                                                  # :133-144
                height = height or self.default_sample_size * self.vae_scale_factor
                width = width or self.default_sample_size * self.vae_scale_factor
                aspect_ratio = width / height
                width = round((max_area * aspect_ratio) ** 0.5) // multiple_of * multiple_of
                height = round((max_area / aspect_ratio) ** 0.5) // multiple_of * multiple_of
            prompt_embeds, pooled_prompt_embeds, text_ids = self.encode_prompt(
                prompt=prompt, prompt_2=prompt_2, prompt_embeds=prompt_embeds,
                pooled_prompt_embeds=pooled_prompt_embeds, device=device, num_images_per_prompt=num_images_per_prompt,
                max_sequence_length=max_sequence_length, lora_scale=None)
            if do_true_cfg:                                                                       # :193-209
                negative_prompt_embeds, negative_pooled_prompt_embeds, _ = self.encode_prompt(
                    prompt=negative_prompt, prompt_2=negative_prompt_2, prompt_embeds=negative_prompt_embeds,
                    pooled_prompt_embeds=negative_pooled_prompt_embeds, device=device,
                    num_images_per_prompt=num_images_per_prompt, max_sequence_length=max_sequence_length,
                    lora_scale=None)
            nch = self.transformer.config.in_channels // 4
            latents, image_latents, latent_ids, image_ids = self.prepare_latents(
                image, 1, nch, height, width, prompt_embeds.dtype, device, generator, latents)
            if image_ids is not None:
                latent_ids = torch.cat([latent_ids, image_ids], dim=0)
        else:
            from .schedule import latent_image_ids
            if height is None or width is None:
                raise ValueError("height and width are required with packed latents")
            gh, gw = height // multiple_of, width // multiple_of
            assert latents.shape[1] == gh * gw and image_latents.shape[1] == gh * gw, "latents do not match H x W"
            text_ids = torch.zeros(prompt_embeds.shape[1], 3, device=device)
            latent_ids = torch.cat([latent_image_ids(gh, gw, 0.0, device), latent_image_ids(gh, gw, 1.0, device)])
        # timesteps (inplace.py:229-244)
        sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
        cfg = self.scheduler.config
        mu = calculate_shift(latents.shape[1], cfg.get("base_image_seq_len", 256), cfg.get("max_image_seq_len", 4096),
                             cfg.get("base_shift", 0.5), cfg.get("max_shift", 1.15))
        retrieve_timesteps(self.scheduler, num_inference_steps, device, sigmas=sigmas, mu=mu)     # :238-244
        self.scheduler.set_begin_index(0)   # known start: spares _init_step_index's device lookup (:606-607)
        self.scheduler._step_index = 0
        negative = (negative_prompt_embeds, negative_pooled_prompt_embeds, float(true_cfg_scale)) if do_true_cfg else None
        latents = self.regione_denoise(latents, image_latents, latent_ids, text_ids, prompt_embeds,
                                       pooled_prompt_embeds, guidance_scale, height, width, negative=negative,
                                       callback_on_step_end=unused.get("callback_on_step_end"),
                                       callback_on_step_end_tensor_inputs=unused.get(
                                           "callback_on_step_end_tensor_inputs") or ["latents"])
        if output_type == "latent":
            image = latents
        else:
            x = self._unpack_latents(latents, height, width, self.vae_scale_factor)
            x = (x / self.vae.config.scaling_factor) + self.vae.config.shift_factor
            image = self.image_processor.postprocess(self.vae.decode(x, return_dict=False)[0], output_type=output_type)
        if not return_dict:
            return (image,)
        return types.SimpleNamespace(images=image)

    def regione_denoise(self, latents, image_latents, latent_ids, text_ids, prompt_embeds, pooled_prompt_embeds,
                        guidance_scale, height, width, negative=None, callback_on_step_end=None,
                        callback_on_step_end_tensor_inputs=("latents",)):
        """The hot loop, inplace.py:287-392. `negative` = (negative_prompt_embeds, negative_pooled_prompt_embeds,
        true_cfg_scale) enables the second forward of true-CFG (:349-364). `callback_on_step_end(pipe, i, t, kwargs)`
        runs after every scheduler step and may replace `latents` (:377-385); `self._interrupt` skips the rest of a
        computed step exactly as the reference's `continue` does (:321-322)."""
        M = MANAGER
        N = M.inference_step
        sch = self.scheduler
        ts_host = sch.timesteps.detach().to("cpu", torch.float32)
        x = latents[0]
        cond = image_latents[0]
        L, C, T = x.shape[0], cond.shape[0], text_ids.shape[0]
        if negative is not None and negative[0].shape[1] != T:
            raise NotImplementedError("regione_b200: prompt and negative prompt must be padded to one length "
                                      "(encode_prompt pads both to max_sequence_length)")
        engine = _get_engine(self.transformer, T, L, C, n_pass=2 if negative is not None else 1)
        self.transformer.__dict__["_regione_b200_engine"] = engine
        M.refresh(x, cond, latent_ids, text_ids, 2, self.vae_scale_factor, height, width)          # :287
        g_x1000 = float(torch.tensor(guidance_scale, dtype=torch.float32).to(x.dtype) * 1000)     # :250, :473
        engine.begin_image(text_ids, latent_ids, prompt_embeds[0], pooled_prompt_embeds[0], g_x1000)
        if negative is not None:
            engine.begin_image(text_ids, latent_ids, negative[0][0], negative[1][0], g_x1000, pass_id=1)
        guidance = torch.full([1], guidance_scale, dtype=torch.float32)
        plan = plan_steps(ts_host, gamma, M)                                                      # :295-313
        cache = None
        record = bool(getattr(self, "regione_record", False))   # tests: keep per-step tensors
        self.regione_trace = {"modes": [], "latents": [], "noise_pred": []}
        marks = [] if getattr(self, "regione_time_steps", False) else None   # diagnostics: CUDA event per step
        for i in range(N):
            assert i == M.current_step                                                            # :293
            if marks is not None:
                marks.append(torch.cuda.Event(enable_timing=True))
                marks[-1].record()
            t = ts_host[i]
            skip, ratio = plan[i]
            if skip:                                                                              # :315-318
                if cache.shape[0] != x.shape[0]:
                    cache = ops.gather_rows(cache, M.edited_ids)
                x = sch.step(cache, t, x, return_dict=False, reuse_ratio=ratio)[0]
                self.regione_trace["modes"].append("SKIP")
            else:
                if getattr(self, "_interrupt", False):                                            # :321-322
                    continue
                cur = M.current_step
                full = cur <= M.warmup_step - 1 or cur > N - M.post_step - 1 or cur == M.prev_refresh_step   # :331
                timestep = t.expand(1).to(x.dtype)                                                # :334
                noise_pred = self.transformer(hidden_states=x[None], timestep=timestep / 1000, guidance=guidance,
                                              pooled_projections=pooled_prompt_embeds,
                                              encoder_hidden_states=prompt_embeds, txt_ids=text_ids,
                                              img_ids=latent_ids, joint_attention_kwargs=None, return_dict=False,
                                              condition_latents=image_latents if full else None)[0]   # :331-332
                noise_pred = noise_pred[0, : x.shape[0]]                                          # :347
                if negative is not None:                                                          # :349-364
                    neg = self.transformer(hidden_states=x[None], timestep=timestep / 1000, guidance=guidance,
                                           pooled_projections=negative[1], encoder_hidden_states=negative[0],
                                           txt_ids=text_ids, img_ids=latent_ids, joint_attention_kwargs=None,
                                           return_dict=False, condition_latents=image_latents if full else None,
                                           regione_pass=1)[0][0, : x.shape[0]]
                    noise_pred = ops.cfg_combine(noise_pred, neg, negative[2])   # neg + scale * (pos - neg), :364
                cache = noise_pred                                                                # :365
                x = sch.step(noise_pred, t, x, return_dict=False)[0]                               # :369
                self.regione_trace["modes"].append("FULL" if full else "REGION")
            if callback_on_step_end is not None:                                                  # :377-385
                kw = {}
                for name in callback_on_step_end_tensor_inputs:
                    if name == "latents":
                        kw[name] = x[None]
                    elif name == "prompt_embeds":
                        kw[name] = prompt_embeds
                    else:
                        raise KeyError(f"regione_b200: callback tensor input `{name}` is not available on this path")
                out = callback_on_step_end(self, i, ts_host[i], kw) or {}
                new_x = out.pop("latents", None)
                if new_x is not None and new_x is not kw.get("latents"):
                    x = new_x[0].contiguous()
                if out.get("prompt_embeds", prompt_embeds) is not prompt_embeds:
                    raise NotImplementedError("regione_b200: the prompt embeddings are registered once per image; "
                                              "a callback cannot replace them mid-loop")
            x, latent_ids = M.step(x, latent_ids)                                                 # :392
            if record:
                self.regione_trace["latents"].append(x.clone())
                self.regione_trace["noise_pred"].append(cache.clone())
        if marks is not None:
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
            marks[-1].synchronize()
            self.regione_trace["step_ms"] = [a.elapsed_time(b) for a, b in zip(marks[:-1], marks[1:])]
        self.regione_trace["edited_ids"] = M.edited_ids
        self.regione_trace["unedited_ids"] = M.unedited_ids
        return x[None]


def warp_modules(pipeline, **args):
    """inplace.py:53-62: install the RegionE loop, scheduler, forward and processors on a live pipeline object."""
    if "_regione_b200_saved" in pipeline.__dict__:
        unwarp_modules(pipeline)
    MANAGER.set_parameters(args)
    tr = pipeline.transformer
    saved = {
        "cls": pipeline.__class__,
        "scheduler": pipeline.scheduler,
        "forward": tr.__dict__.get("forward"),
        "processors": [getattr(b.attn, "processor", None)
                       for b in list(tr.transformer_blocks) + list(tr.single_transformer_blocks)],
    }
    pipeline.__dict__["_regione_b200_saved"] = saved
    pipeline.__class__ = type("RegionEFluxKontextPipeline", (RegionEFluxKontextPipelineMixin, saved["cls"]), {})
    sch_cls = type("RegionEFlowMatchEulerDiscreteScheduler", (RegionESchedulerMixin, saved["scheduler"].__class__), {})
    pipeline.scheduler = sch_cls.from_config(saved["scheduler"].config)
    pipeline.scheduler._regione_manager = MANAGER
    tr.forward = types.MethodType(RegionEFluxTransformer2DModelforward, tr)
    for block in tr.transformer_blocks:
        block.attn.set_processor(RegionEB200AttnProcessor(False))
    for block in tr.single_transformer_blocks:
        block.attn.set_processor(RegionEB200AttnProcessor(True))
    return pipeline


def unwarp_modules(pipeline):
    """inplace.py:65-73: restore the vanilla class, scheduler, forward and processors; frees the library handle."""
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
    blocks = list(tr.transformer_blocks) + list(tr.single_transformer_blocks)
    for b, p in zip(blocks, saved["processors"]):
        b.attn.set_processor(p)
    for eng in tr.__dict__.pop("_regione_b200_engines", {}).values():
        eng.close()
    tr.__dict__.pop("_regione_b200_engine", None)
    return pipeline
