"""Oracle restatement of the reference's region state + token ops (RegionE/FluxKontext/utils.py). TEST INFRASTRUCTURE.

Each function cites the reference lines it follows. Pinned against outputs of the reference's own code
(tests/golden/region_ops_*.pt, produced by oracle/make_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def cosine_similarity_map(estimate: torch.Tensor, condition: torch.Tensor) -> torch.Tensor:
    """utils.py:308-312 — each tensor is normalised in ITS OWN dtype (estimate fp32, condition bf16), the product and
    the channel sum run in fp32. Shapes [B, L, ch] -> [B, L]."""
    a = F.normalize(estimate, dim=-1)
    b = F.normalize(condition, dim=-1)
    return torch.sum(a * b, dim=-1)


def erode_cross3(mask: torch.Tensor) -> torch.Tensor:
    """utils.py:150-181 with the 3x3 cross of :138-143: a cell survives iff it and its 4 neighbours are set; zero
    padding, so border cells never survive."""
    k = torch.zeros(1, 1, 3, 3, device=mask.device)
    k[0, 0, 1, :] = 1
    k[0, 0, :, 1] = 1
    hits = F.conv2d(mask.float()[None, None], k, padding=1)
    return (hits == k.sum()).float()[0, 0]


def dilate_square5(mask: torch.Tensor) -> torch.Tensor:
    """utils.py:184-212 with the 5x5 all-ones element (:229): a cell is set iff any cell of its 5x5 window is."""
    hits = F.conv2d(mask.float()[None, None], torch.ones(1, 1, 5, 5, device=mask.device), padding=2)
    return (hits > 0).float()[0, 0]


def clean_mask(mask2d: torch.Tensor) -> torch.Tensor:
    """utils.py:215-237 remove_scattered_points: erosion then dilation (the kernel_size argument is ignored there)."""
    return dilate_square5(erode_cross3(mask2d))


def select_tokens(estimate, condition, threshold, grid_h, grid_w, erosion_dilation=True):
    """utils.py:282-354 token_selector(similarity_type='cosine'): -> (edited_ids [1,n], unedited_ids [1,L-n]) int64,
    ascending; also returns the raw and final masks and the similarity for margin reporting."""
    sim = cosine_similarity_map(estimate, condition)
    raw = sim <= threshold                                            # :333
    final = raw
    if erosion_dilation:
        grid = raw.float().squeeze().reshape(grid_h, grid_w)          # :337-340 (row-major token grid)
        final = clean_mask(grid).bool().flatten().unsqueeze(0)        # :342-343
    L = estimate.shape[1]
    all_ids = torch.arange(L, device=estimate.device).unsqueeze(0)
    edited = all_ids[final].unsqueeze(0)                              # :346-347
    unedited = all_ids[~final].view(1, L - edited.shape[1])           # :351-352
    return edited, unedited, raw, final, sim


def gather_rows(latent: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """utils.py:260-279 ids_gather: latent [B,S,D], ids [B,K] -> [B,K,D]."""
    b = torch.arange(ids.shape[0]).unsqueeze(1).expand(-1, ids.shape[1])
    return latent[b, ids, :]


def scatter_rows(rows: torch.Tensor, ids: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """utils.py:240-257 ids_scatter (in place on dst, which is returned)."""
    dst[torch.arange(rows.shape[0]).unsqueeze(1), ids] = rows
    return dst


class RegionState:
    """utils.py:357-465 FluxKontextManager: hyper-parameters, per-image refresh and the split/merge state machine."""

    def __init__(self):
        self.inference_step = 28
        self.current_step = 0
        self.edited_ids = self.unedited_ids = self.unedited_latent = None
        self.prev_refresh_step = self.next_refresh_step = None
        self.refresh_step = []
        self.refresh_step_real_time = []

    def set_parameters(self, args: dict) -> None:  # :390-402
        # the reference asserts == 28 (its gamma tables have 27 entries); a caller-supplied table ("gamma") lifts it,
        # mirroring the product's extension for BASELINE configs[2] - such runs are unpinned w.r.t. the reference
        assert args["warmup_step"] >= 1 and (args["num_inference_steps"] == 28 or args.get("gamma") is not None)
        self.inference_step = args["num_inference_steps"]
        self.warmup_step = args["warmup_step"]
        self.post_step = args["post_step"]
        self.threshold = args["threshold"]
        self.cache_threshold = args["cache_threshold"]
        self.erosion_dilation = args["erosion_dilation"]
        steps = sorted(int(s) for s in args["refresh_step"].split(","))
        assert min(steps) > self.warmup_step + 1 and max(steps) <= self.inference_step - self.post_step - 1
        assert not any(b - a == 1 for a, b in zip(steps, steps[1:])), "Refresh steps must not be adjacent."
        self.refresh_step = steps + [self.inference_step - self.post_step + 1]

    def refresh(self, latents, image_latents, latent_ids, text_ids, height, width) -> None:  # :437-465
        self.height, self.width = height, width
        self.patch_size, self.vae_scale_factor = 2, 8
        self.latent_length = latents.size(1)
        self.txt_length = text_ids.size(0)
        self.condition_latent = image_latents
        self.condition_length = image_latents.size(1)
        self.current_step = 0
        self.prev_refresh_step = self.next_refresh_step = None
        self.edited_ids = self.unedited_ids = self.unedited_latent = None
        self.latent_ids = latent_ids
        self.refresh_step_real_time = list(self.refresh_step)

    def _split(self, latent, latent_ids):
        self.unedited_latent = gather_rows(latent, self.unedited_ids)
        if latent_ids.dim() == 1:   # Qwen keeps 1-D position indices (QwenImageEdit/inplace.py:322)
            ids = latent_ids[self.edited_ids.squeeze(0)]
        else:
            ids = gather_rows(latent_ids.unsqueeze(0), self.edited_ids).squeeze(0)
        return gather_rows(latent, self.edited_ids), ids

    def _merge(self, latent):
        full = torch.zeros_like(self.condition_latent)
        scatter_rows(latent, self.edited_ids, full)
        scatter_rows(self.unedited_latent, self.unedited_ids, full)
        return full, self.latent_ids

    def step(self, latent, latent_ids):  # :404-435
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
