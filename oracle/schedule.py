"""Oracle restatement of the step schedule: sigmas/timesteps and the AVDC decision rule. TEST INFRASTRUCTURE.

Reference: RegionE/FluxKontext/inplace.py:229-244 (sigmas, mu), :295-313 (AVDC), utils.py:38-48 (calculate_shift).
`FlowMatchEulerDiscreteScheduler.set_timesteps` is diffusers code (not in /root/reference, not installed): restated
from its published algorithm (dynamic exponential time shift); unpinned w.r.t. diffusers, cross-pinned against BFL's own
sampler schedule (tests/test_oracle_vs_bfl_flux.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch

# fitted per-step velocity decay factors, fp16 tensors in the reference (inplace.py:47-50 of each family)
GAMMA = {
    "FluxKontext": [0.8352, 0.9986, 1.0090, 1.0097, 1.0161, 1.0152, 1.0160, 1.0173, 1.0177, 1.0199, 1.0213, 1.0203,
                    1.0257, 1.0236, 1.0235, 1.0278, 1.0302, 1.0311, 1.0352, 1.0371, 1.0391, 1.0459, 1.0498, 1.0581,
                    1.0693, 1.0866, 1.1090],
    "Step1XEdit": [0.9746, 0.9593, 1.0036, 1.0084, 1.0106, 1.0114, 1.0138, 1.0163, 1.0152, 1.0163, 1.0197, 1.0186,
                   1.0219, 1.0218, 1.0223, 1.0266, 1.0272, 1.0305, 1.0311, 1.0362, 1.0385, 1.0423, 1.0500, 1.0536,
                   1.0671, 1.0866, 1.1015],
}


def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=4096, base_shift=0.5, max_shift=1.15):
    """utils.py:38-48."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    return image_seq_len * m + (base_shift - m * base_seq_len)


def flow_match_sigmas(num_steps: int, image_seq_len: int):
    """inplace.py:229-244 + diffusers set_timesteps(sigmas=..., mu=...) with use_dynamic_shifting (exponential).
    Returns (sigmas fp32 [N+1] with trailing 0, timesteps fp32 [N])."""
    sig = np.linspace(1.0, 1 / num_steps, num_steps).astype(np.float32)
    mu = calculate_shift(image_seq_len)
    sig = math.exp(mu) / (math.exp(mu) + (1 / sig - 1) ** 1.0)
    sig = torch.from_numpy(np.asarray(sig)).to(torch.float32)
    return torch.cat([sig, torch.zeros(1)]), sig * 1000.0


def avdc_plan(timesteps: torch.Tensor, gamma, warmup_step=6, post_step=2, refresh_step="16", cache_threshold=0.04,
              inference_step=28):
    """inplace.py:295-313 together with the refresh bookkeeping of :630-639 and utils.py:404-435.
    Returns per step a dict(mode=FULL|REGION|SKIP, ratio=float|None, write_cache=bool)."""
    g = torch.tensor(gamma, dtype=torch.float16)
    refresh = sorted(int(s) for s in refresh_step.split(",")) + [inference_step - post_step + 1]
    rt = list(refresh)
    prev_refresh = next_refresh = None
    accumulate = 1
    plan = []
    for i in range(inference_step):
        t = timesteps[i]
        ratio = None
        forced = i <= warmup_step or i > inference_step - post_step - 1 or i == prev_refresh
        if forced:
            skip, accumulate = False, 1
        else:
            ratio = g[i - 1] * (1 + (t - timesteps[i - 1]) / 1000)
            if ratio >= 1:
                skip, accumulate = False, 1
            else:
                accumulate = accumulate * ratio
                if 1 - accumulate > cache_threshold:
                    skip, accumulate = False, 1
                else:
                    skip = True
        full = i <= warmup_step - 1 or i > inference_step - post_step - 1 or i == prev_refresh   # :331
        write = i == warmup_step - 1 or i == prev_refresh                                          # :721
        plan.append({"mode": "SKIP" if skip else ("FULL" if full else "REGION"),
                     "ratio": None if ratio is None else float(ratio), "write_cache": bool(write and not skip)})
        # scheduler.step bookkeeping (:630-639) happens after the forward of step i
        if i == warmup_step - 1:
            prev_refresh = rt.pop(0) - 1
        elif prev_refresh is not None and i == prev_refresh and rt:
            next_refresh = rt.pop(0) - 1
        # MANAGER.step (utils.py:404-435), evaluated with current_step = i + 1
        s = i + 1
        if s == inference_step - post_step:
            prev_refresh = None
        elif prev_refresh is not None and s == prev_refresh + 1:
            prev_refresh = next_refresh
    return plan
