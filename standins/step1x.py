"""Stand-ins for Step1X-Edit (v1p1): the block stack is FLUX's (19 double + 38 single blocks, same attention attribute
names, RegionE/Step1XEdit/inplace.py:720-808); the front end — `connector`, `time_proj`, `time_embed`, `vec_embed` — and
`process_diff_norm` live in the fork `Peyton-Chen/diffusers@step1xedit_v1p2` (README.md:76-77), which is not available
here, so these are small callable modules with the same call signatures the reference's forward uses
(:514-520, :407). They only have to be the SAME objects for the oracle and the CUDA path; after `enable()` the block
stack runs in the CUDA library and the front end stays the pipeline's own modules, as in the reference.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from .diffusers_like import (FlowMatchEulerDiscreteScheduler, FluxKontextPipeline, FluxTransformer2DModel, _AdaNorm,
                      _Config, _DoubleBlock, _SingleBlock)


class _MLP(nn.Module):
    def __init__(self, inp, dim):
        super().__init__()
        self.linear_1, self.linear_2 = nn.Linear(inp, dim), nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class _Connector(nn.Module):
    """(embeds [B,T,ctx], timestep [B], mask [B,T]) -> (tokens [B,T,ctx], pooled vector y [B,vec])."""

    def __init__(self, ctx_dim, vec_dim):
        super().__init__()
        self.proj = nn.Linear(ctx_dim, ctx_dim)
        self.pool = nn.Linear(ctx_dim, vec_dim)

    def forward(self, x, timestep, mask):
        m = mask.to(x.dtype)[..., None]
        h = F.silu(self.proj(x)) * m
        y = self.pool(h.sum(1) / m.sum(1).clamp_min(1)) * (1 + timestep.to(x.dtype)[:, None])
        return h, y


class Step1XEditTransformer2DModel(nn.Module):
    def __init__(self, dim=3072, heads=24, n_double=19, n_single=38, mlp_ratio=4, in_channels=64, ctx_dim=4096,
                 vec_dim=768):
        super().__init__()
        self.config = _Config(in_channels=in_channels, guidance_embeds=False, num_layers=n_double,
                              num_single_layers=n_single, attention_head_dim=dim // heads, num_attention_heads=heads)
        self.connector = _Connector(ctx_dim, vec_dim)
        self.x_embedder = nn.Linear(in_channels, dim)
        self.time_embed = _MLP(256, dim)
        self.vec_embed = _MLP(vec_dim, dim)
        self.context_embedder = nn.Linear(ctx_dim, dim)
        self.transformer_blocks = nn.ModuleList([_DoubleBlock(dim, heads, mlp_ratio) for _ in range(n_double)])
        self.single_transformer_blocks = nn.ModuleList([_SingleBlock(dim, heads, mlp_ratio) for _ in range(n_single)])
        self.norm_out = _AdaNorm(dim, 2)
        self.proj_out = nn.Linear(dim, in_channels)

    @staticmethod
    def time_proj(t):
        """Sinusoidal projection (256 channels, cos first) of `timestep * 1000`, fp32."""
        half = 128
        e = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
        a = t[:, None].float() * e[None]
        return torch.cat([torch.cos(a), torch.sin(a)], dim=-1)

    @staticmethod
    def pos_embed(ids):
        """FluxPosEmbed(theta 10000, axes (16, 56, 56)) -> (cos, sin) fp32 [S, 128]."""
        pos = ids.float()
        cos, sin = [], []
        for a, d in enumerate((16, 56, 56)):
            freqs = 1.0 / (10000.0 ** (torch.arange(0, d, 2, dtype=torch.float64, device=ids.device) / d))
            ang = torch.outer(pos[:, a].double(), freqs)
            cos.append(ang.cos().repeat_interleave(2, dim=1).float())
            sin.append(ang.sin().repeat_interleave(2, dim=1).float())
        return torch.cat(cos, dim=-1), torch.cat(sin, dim=-1)

    def forward(self, *a, **k):
        raise NotImplementedError("the vanilla Step1X transformer forward is fork code; enable RegionE first")

    init_synthetic = FluxTransformer2DModel.init_synthetic


class Step1XEditPipeline(FluxKontextPipeline):
    def __init__(self, transformer, scheduler=None):
        super().__init__(transformer, scheduler or FlowMatchEulerDiscreteScheduler())

    @staticmethod
    def process_diff_norm(diff_norm, k):
        """Fork's norm compression for CFG (Step1XEdit/inplace.py:407): norms above 1 are raised to the power k."""
        return torch.where(diff_norm > 1.0, torch.pow(diff_norm, k), torch.ones_like(diff_norm))


class Step1XEditV1P2Transformer2DModel(Step1XEditTransformer2DModel):
    """v1p2 adds `text_token_mapping` on an extra text-embedding stream (Step1XEditV1P2/inplace.py:606-609)."""

    def __init__(self, text_dim=96, ctx_dim=4096, **kw):
        super().__init__(ctx_dim=ctx_dim, **kw)
        self.text_token_mapping = nn.Linear(text_dim, ctx_dim)


class Step1XEditPipelineV1P2(Step1XEditPipeline):
    pass
