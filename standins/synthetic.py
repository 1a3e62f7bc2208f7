"""Synthetic weights / inputs for tests and benchmarks (no datasets or checkpoints exist offline).

The condition latent is built so that the adaptive region partition is CONTROLLABLE with the reference's own
threshold (0.88): outside a chosen blob region the condition latent is a noisy copy of the initial noise latent (the
one-step x0 estimate stays close to it because the synthetic velocity is small), inside the blobs it is independent
noise. The edited fraction rho is then set by the blob area; salt noise exercises the erosion/dilation.
"""
from __future__ import annotations

import math

import torch

from .diffusers_like import FlowMatchEulerDiscreteScheduler, FluxKontextPipeline, FluxTransformer2DModel

FLUX_KONTEXT = dict(dim=3072, heads=24, n_double=19, n_single=38, mlp_ratio=4, in_channels=64, ctx_dim=4096,
                    pooled_dim=768, guidance_embeds=True)
TINY = dict(dim=256, heads=2, n_double=2, n_single=2, mlp_ratio=4, in_channels=64, ctx_dim=128, pooled_dim=64,
            guidance_embeds=True)


def build_pipeline(arch: dict, seed: int = 110, device="cpu", velocity_scale: float | None = None):
    """Stand-in FluxKontextPipeline with seeded synthetic weights. `velocity_scale` multiplies proj_out so that
    |dt_final * v| stays well below |x| (defaults to 0.3 / (0.02 * sqrt(dim)), i.e. v has std ~0.3)."""
    tr = FluxTransformer2DModel(**arch).init_synthetic(seed, device)
    if velocity_scale is None:
        velocity_scale = 0.3 / (0.02 * math.sqrt(arch["dim"]))
    with torch.no_grad():
        tr.proj_out.weight.mul_(velocity_scale)
        tr.proj_out.bias.mul_(velocity_scale)
    return FluxKontextPipeline(tr, FlowMatchEulerDiscreteScheduler())


def blob_mask(grid_h: int, grid_w: int, rho: float, gen: torch.Generator, salt: float = 0.01) -> torch.Tensor:
    """Boolean [grid_h*grid_w]: a few discs covering ~rho of the grid plus salt noise."""
    yy, xx = torch.meshgrid(torch.arange(grid_h), torch.arange(grid_w), indexing="ij")
    m = torch.zeros(grid_h, grid_w, dtype=torch.bool)
    if rho >= 1.0:
        return torch.ones(grid_h * grid_w, dtype=torch.bool)
    if rho > 0:
        n = 3
        r = math.sqrt(rho * grid_h * grid_w / n / math.pi)
        for _ in range(n):
            cy = float(torch.rand(1, generator=gen)) * (grid_h - 2 * r - 4) + r + 2
            cx = float(torch.rand(1, generator=gen)) * (grid_w - 2 * r - 4) + r + 2
            m |= ((yy - cy) ** 2 + (xx - cx) ** 2) <= r * r
    if salt > 0:
        m |= torch.rand(grid_h, grid_w, generator=gen) < salt
    return m.flatten()


def make_inputs(seed: int, grid_h: int, grid_w: int, txt_len: int, ctx_dim: int, pooled_dim: int, rho: float = 0.25,
                device="cpu", channels: int = 64):
    """All tensors are created with a CPU generator (identical for the oracle and the CUDA path), bf16."""
    g = torch.Generator().manual_seed(seed)
    L = grid_h * grid_w
    latents = torch.randn(1, L, channels, generator=g)
    edit = blob_mask(grid_h, grid_w, rho, g)
    cond = latents + 0.1 * torch.randn(1, L, channels, generator=g)
    fresh = 0.6 * torch.randn(1, L, channels, generator=g)
    cond[0, edit] = fresh[0, edit]
    prompt = 0.1 * torch.randn(1, txt_len, ctx_dim, generator=g)
    pooled = torch.randn(1, pooled_dim, generator=g)
    to = dict(device=device, dtype=torch.bfloat16)
    return dict(latents=latents.to(**to), image_latents=cond.to(**to), prompt_embeds=prompt.to(**to),
                pooled_prompt_embeds=pooled.to(**to), height=grid_h * 16, width=grid_w * 16, intended_mask=edit)
