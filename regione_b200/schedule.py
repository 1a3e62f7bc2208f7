"""Front of the denoising loop: the schedule helpers the reference keeps next to its pipelines
(RegionE/FluxKontext/utils.py:38-107, identical in every family) — same names, arguments and error behaviour."""
from __future__ import annotations

import torch

import inspect


def calculate_shift(image_seq_len, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                    max_shift: float = 1.15):
    """Schedule shift mu, linear in the token count (utils.py:38-48). The operation order is the reference's: the
    AVDC decision downstream flips on 1-ulp changes of the timesteps, so `m * len + b` it is."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    mu = image_seq_len * m + b
    return mu


def retrieve_timesteps(scheduler, num_inference_steps=None, device=None, timesteps=None, sigmas=None, **kwargs):
    """Calls `scheduler.set_timesteps` and returns `(scheduler.timesteps, num_inference_steps)`; custom `timesteps`
    XOR custom `sigmas` are forwarded when the scheduler accepts them (utils.py:51-107)."""
    if timesteps is not None and sigmas is not None:
        raise ValueError("Only one of `timesteps` or `sigmas` can be passed. Please choose one to set custom values")
    accepted = set(inspect.signature(scheduler.set_timesteps).parameters.keys())
    if timesteps is not None:
        if "timesteps" not in accepted:
            raise ValueError(f"The current scheduler class {scheduler.__class__}'s `set_timesteps` does not support "
                             f"custom timestep schedules. Please check whether you are using the correct scheduler.")
        scheduler.set_timesteps(timesteps=timesteps, device=device, **kwargs)
        timesteps = scheduler.timesteps
        num_inference_steps = len(timesteps)
    elif sigmas is not None:
        if "sigmas" not in accepted:
            raise ValueError(f"The current scheduler class {scheduler.__class__}'s `set_timesteps` does not support "
                             f"custom sigmas schedules. Please check whether you are using the correct scheduler.")
        scheduler.set_timesteps(sigmas=sigmas, device=device, **kwargs)
        timesteps = scheduler.timesteps
        num_inference_steps = len(timesteps)
    else:
        scheduler.set_timesteps(num_inference_steps, device=device, **kwargs)
        timesteps = scheduler.timesteps
    return timesteps, num_inference_steps


def latent_image_ids(grid_h: int, grid_w: int, first: float = 0.0, device="cpu", dtype=torch.float32):
    """diffusers FluxKontextPipeline._prepare_latent_image_ids: rows (first, r, c), row-major."""
    ids = torch.zeros(grid_h, grid_w, 3)
    ids[..., 0] = first
    ids[..., 1] = torch.arange(grid_h)[:, None]
    ids[..., 2] = torch.arange(grid_w)[None, :]
    return ids.reshape(grid_h * grid_w, 3).to(device=device, dtype=dtype)
