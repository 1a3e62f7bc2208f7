"""Host-side region state (the reference's module-global MANAGER) and the zero-sync step planner.

Mirrors RegionE/FluxKontext/utils.py:357-465 (`FluxKontextManager`: set_parameters / refresh / step) with the same
field names, so host code written against the reference's manager reads the same here. Differences, all host-only:
ids are int32 device tensors of shape [n] (batch is 1), latents are [M, C] views, and row moves go through the
library's gather/scatter kernels.
"""
from __future__ import annotations

import torch

from . import ops


class RegionManager:
    def __init__(self) -> None:
        # model config
        self.patch_size = 2
        self.vae_scale_factor = 8
        self.inference_step = 28
        self.txt_length = None
        self.height = None
        self.width = None
        self.latent_length = 0
        self.condition_latent = None
        self.condition_length = 0
        self.latent_ids = None
        # regione config
        self.warmup_step = 8
        self.post_step = 0
        self.erosion_dilation = False
        self.threshold = None
        self.cache_threshold = 0
        self.refresh_step = []
        self.gamma = None            # caller-supplied table (None: the family's fitted 27-entry table)
        # realtime data
        self.current_step = 0
        self.edited_ids = None       # int32 [n_e], ascending
        self.unedited_ids = None     # int32 [L - n_e]
        self.edited_mask = None      # uint8 [L]
        self.unedited_latent = None
        self.prev_refresh_step = None
        self.next_refresh_step = None
        self.refresh_step_real_time = []

    @property
    def batch_masks(self):
        """[world, L] uint8: the partition (after morphology) of every image of the data-parallel batch, one image per
        rank, all-gathered at step warmup-1 (flux_kontext.allgather_masks_async); this rank's image is row `rank`.
        Reading it orders the current stream after the collective. None before the partition step."""
        p = getattr(self, "batch_masks_pending", None)
        return None if p is None else p.get()

    def batch_edited_counts(self):
        """Edited tokens of every image of the batch (host list; one small device->host read)."""
        m = self.batch_masks
        return None if m is None else [int(c) for c in m.sum(dim=1).tolist()]

    def set_parameters(self, args) -> None:
        """utils.py:390-402, same validation and the same appended sentinel."""
        user_gamma = args.get("gamma")   # extension: a caller-supplied table lifts the 28-step restriction
        assert args["warmup_step"] >= 1 and (args["num_inference_steps"] == 28 or user_gamma is not None), \
            "Changing the inference step requires fitting a new gamma"
        assert user_gamma is None or len(user_gamma) == args["num_inference_steps"] - 1, \
            "gamma needs num_inference_steps - 1 entries"
        self.gamma = None if user_gamma is None else [float(g) for g in user_gamma]
        self.inference_step = args["num_inference_steps"]
        self.warmup_step = args["warmup_step"]
        self.post_step = args["post_step"]
        self.threshold = args["threshold"]
        self.cache_threshold = args["cache_threshold"]
        self.erosion_dilation = args["erosion_dilation"]
        self.refresh_step = sorted(int(item) for item in str(args["refresh_step"]).split(","))
        assert min(self.refresh_step) > self.warmup_step + 1 and \
            max(self.refresh_step) <= self.inference_step - self.post_step - 1
        assert not any(b - a == 1 for a, b in zip(self.refresh_step, self.refresh_step[1:])), \
            "Refresh steps must not be adjacent."
        self.refresh_step.append(self.inference_step - self.post_step + 1)

    def refresh(self, latents, image_latents, latent_ids, text_ids, patch_size=2, vae_scale_factor=8, height=None,
                width=None) -> None:
        """utils.py:437-465: per-image reset."""
        self.width, self.height = width, height
        self.patch_size, self.vae_scale_factor = patch_size, vae_scale_factor
        self.latent_length = latents.size(-2)
        self.txt_length = text_ids.size(0)
        self.condition_latent = image_latents
        self.condition_length = image_latents.size(-2) if image_latents is not None else 0
        self.current_step = 0
        self.prev_refresh_step = None
        self.next_refresh_step = None
        self.edited_ids = self.unedited_ids = self.edited_mask = None
        self.batch_masks_pending = None
        self.unedited_latent = None
        self.latent_ids = latent_ids
        self.refresh_step_real_time = list(self.refresh_step)

    # -- split / merge of the latent rows (utils.py:404-435)
    def _split(self, latent, latent_ids):
        """utils.py:407-410 / :429-432. The reference also gathers `latent_ids` down to the edited rows; the only
        consumer of those ids is the patched forward's rotary lookup, which here goes through `edited_ids` itself
        (the library looks query rows up in the full table through the selection), so no id tensor is gathered and
        the loop carries None until the next merge."""
        self.unedited_latent = ops.gather_rows(latent, self.unedited_ids)
        return ops.gather_rows(latent, self.edited_ids), None

    def _merge(self, latent):
        """utils.py:412-418 / :421-426: edited and unedited ids partition [0, L), so every row of the result is
        written by exactly one of the two scatters and no zero fill is needed."""
        full = torch.empty(self.latent_length, latent.shape[1], dtype=latent.dtype, device=latent.device)
        ops.scatter_rows(latent, self.edited_ids, full)
        ops.scatter_rows(self.unedited_latent, self.unedited_ids, full)
        return full, self.latent_ids

    def step(self, latent, latent_ids):
        """latent [M, C] (2-D view of the reference's [1, M, C])."""
        self.current_step += 1
        s = self.current_step
        if s == self.warmup_step:
            latent, latent_ids = self._split(latent, latent_ids)
        elif s == self.inference_step - self.post_step:
            latent, latent_ids = self._merge(latent)
            self.prev_refresh_step = None
        elif self.prev_refresh_step is not None and s == self.prev_refresh_step:
            latent, latent_ids = self._merge(latent)
        elif self.prev_refresh_step is not None and s == self.prev_refresh_step + 1:
            latent, latent_ids = self._split(latent, latent_ids)
            self.prev_refresh_step = self.next_refresh_step
        return latent, latent_ids


def plan_steps(timesteps_host: torch.Tensor, gamma, manager: RegionManager):
    """AVDC decisions for the whole image, computed on the host before the loop starts.

    The rule (RegionE/FluxKontext/inplace.py:295-313) reads only the schedule — gamma (fp16), the timesteps (fp32) and
    the refresh bookkeeping — never the data, so evaluating it here with the reference's exact tensor arithmetic
    (fp16 x fp32 0-dim CPU tensors) removes its 1-2 device->host syncs per step. Returns a list of
    (skip: bool, ratio: float | None).
    """
    N, warm, post = manager.inference_step, manager.warmup_step, manager.post_step
    if manager.gamma is not None:      # caller-supplied table (set_params(gamma=...)); else the family's fitted one
        gamma = manager.gamma
    assert len(gamma) == N - 1, "gamma needs num_inference_steps - 1 entries"
    g = torch.tensor(gamma, dtype=torch.float16)
    ts = timesteps_host.detach().to("cpu", torch.float32)
    rt = list(manager.refresh_step)
    prev_refresh = next_refresh = None
    accumulate = 1
    plan = []
    for i in range(N):
        ratio = None
        if i <= warm or i > N - post - 1 or i == prev_refresh:
            skip, accumulate = False, 1
        else:
            ratio = g[i - 1] * (1 + (ts[i] - ts[i - 1]) / 1000)
            if ratio >= 1:
                skip, accumulate = False, 1
            else:
                accumulate = accumulate * ratio
                if 1 - accumulate > manager.cache_threshold:
                    skip, accumulate = False, 1
                else:
                    skip = True
        plan.append((skip, None if ratio is None else ratio.clone()))
        if i == warm - 1:                                         # scheduler bookkeeping, inplace.py:630-639
            prev_refresh = rt.pop(0) - 1
        elif prev_refresh is not None and i == prev_refresh and rt:
            next_refresh = rt.pop(0) - 1
        if i + 1 == N - post:                                     # MANAGER.step, utils.py:412-433
            prev_refresh = None
        elif prev_refresh is not None and i + 1 == prev_refresh + 1:
            prev_refresh = next_refresh
    return plan
