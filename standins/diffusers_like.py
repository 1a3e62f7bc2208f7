"""Duck-typed stand-ins for the diffusers objects RegionE patches (diffusers is not installed in this image and no
model weights exist offline): a weight container with diffusers' FluxTransformer2DModel attribute names, a
FlowMatchEulerDiscreteScheduler with the fields the reference reads, and a `FluxKontextPipeline` whose class NAME is
what `RegionEHelper` dispatches on (RegionE/tool/RegionE.py:12-13). They carry synthetic weights at real or reduced
shapes for tests and benchmarks. None of this is on the hot path: after `RegionEHelper.enable()` every forward runs
in the CUDA library; the un-patched (vanilla) forward belongs to diffusers and is deliberately not re-implemented.

TEST / BENCHMARK SCAFFOLDING: lives outside the product package (`regione_b200/` never imports it, except the CLI's
`--model_path synthetic` mode, lazily).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch
from torch import nn


class _Config(dict):
    __getattr__ = dict.get


class _RMSNorm(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.eps = 1e-6


class _AdaNorm(nn.Module):
    def __init__(self, dim, mult):
        super().__init__()
        self.linear = nn.Linear(dim, mult * dim)


class _GELUProj(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner)


class _FeedForward(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(dim, inner), nn.Identity(), nn.Linear(inner, dim)])


class _Attention(nn.Module):
    def __init__(self, dim, heads, context: bool):
        super().__init__()
        self.heads = heads
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.norm_q, self.norm_k = _RMSNorm(dim // heads), _RMSNorm(dim // heads)
        if context:
            self.add_q_proj, self.add_k_proj, self.add_v_proj = (nn.Linear(dim, dim), nn.Linear(dim, dim),
                                                                 nn.Linear(dim, dim))
            self.norm_added_q, self.norm_added_k = _RMSNorm(dim // heads), _RMSNorm(dim // heads)
            self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])
            self.to_add_out = nn.Linear(dim, dim)
        self.processor = None

    def set_processor(self, processor):
        self.processor = processor


class _DoubleBlock(nn.Module):
    def __init__(self, dim, heads, ratio):
        super().__init__()
        self.norm1, self.norm1_context = _AdaNorm(dim, 6), _AdaNorm(dim, 6)
        self.attn = _Attention(dim, heads, context=True)
        self.ff, self.ff_context = _FeedForward(dim, ratio * dim), _FeedForward(dim, ratio * dim)


class _SingleBlock(nn.Module):
    def __init__(self, dim, heads, ratio):
        super().__init__()
        self.norm = _AdaNorm(dim, 3)
        self.proj_mlp = nn.Linear(dim, ratio * dim)
        self.proj_out = nn.Linear(dim + ratio * dim, dim)
        self.attn = _Attention(dim, heads, context=False)


class _MLPEmbed(nn.Module):
    def __init__(self, inp, dim):
        super().__init__()
        self.linear_1, self.linear_2 = nn.Linear(inp, dim), nn.Linear(dim, dim)


class _TimeTextEmbed(nn.Module):
    def __init__(self, dim, pooled_dim, guidance):
        super().__init__()
        self.timestep_embedder = _MLPEmbed(256, dim)
        if guidance:
            self.guidance_embedder = _MLPEmbed(256, dim)
        self.text_embedder = _MLPEmbed(pooled_dim, dim)


class FluxTransformer2DModel(nn.Module):
    """Weight container with the module surface of diffusers' FluxTransformer2DModel."""

    def __init__(self, dim=3072, heads=24, n_double=19, n_single=38, mlp_ratio=4, in_channels=64, ctx_dim=4096,
                 pooled_dim=768, guidance_embeds=True):
        super().__init__()
        self.config = _Config(in_channels=in_channels, guidance_embeds=guidance_embeds, num_layers=n_double,
                              num_single_layers=n_single, attention_head_dim=dim // heads,
                              num_attention_heads=heads, joint_attention_dim=ctx_dim, pooled_projection_dim=pooled_dim)
        self.x_embedder = nn.Linear(in_channels, dim)
        self.context_embedder = nn.Linear(ctx_dim, dim)
        self.time_text_embed = _TimeTextEmbed(dim, pooled_dim, guidance_embeds)
        self.transformer_blocks = nn.ModuleList([_DoubleBlock(dim, heads, mlp_ratio) for _ in range(n_double)])
        self.single_transformer_blocks = nn.ModuleList([_SingleBlock(dim, heads, mlp_ratio) for _ in range(n_single)])
        self.norm_out = _AdaNorm(dim, 2)
        self.proj_out = nn.Linear(dim, in_channels)

    def forward(self, *args, **kwargs):
        raise NotImplementedError(
            "the vanilla FluxTransformer2DModel.forward is diffusers code and is not part of regione_b200; call "
            "RegionEHelper(pipeline).enable() to run the B200 hot path")

    @torch.no_grad()
    def init_synthetic(self, seed: int = 110, device="cpu"):
        """Linear ~ 0.02 N(0,1), biases 0.01 N(0,1), RMSNorm 1 + 0.1 N(0,1); one seeded
        generator, parameters visited in state_dict order. Ends in bf16 on `device`."""
        dev = torch.device(device)
        gen = torch.Generator(device=dev).manual_seed(seed)
        self.to(device=dev, dtype=torch.bfloat16)
        for name, p in self.named_parameters():
            r = torch.randn(p.shape, generator=gen, device=dev, dtype=torch.float32)
            if name.endswith("norm_q.weight") or name.endswith("norm_k.weight") or "norm_added" in name:
                p.copy_(1.0 + 0.1 * r)
            elif name.endswith(".bias"):
                p.copy_(0.01 * r)
            else:
                p.copy_(0.02 * r)
        return self


class FlowMatchEulerDiscreteScheduler:
    """The slice of diffusers' scheduler the reference relies on: dynamic exponential shift,
    `sigmas` (fp32, trailing 0), `timesteps`, `step_index`, `set_begin_index`, plain Euler `step`."""

    order = 1

    def __init__(self, num_train_timesteps=1000, shift=3.0, use_dynamic_shifting=True, base_shift=0.5, max_shift=1.15,
                 base_image_seq_len=256, max_image_seq_len=4096, stochastic_sampling=False, **extra):
        self.config = _Config(num_train_timesteps=num_train_timesteps, shift=shift,
                              use_dynamic_shifting=use_dynamic_shifting, base_shift=base_shift, max_shift=max_shift,
                              base_image_seq_len=base_image_seq_len, max_image_seq_len=max_image_seq_len,
                              stochastic_sampling=stochastic_sampling, **extra)
        self.sigmas = self.timesteps = None
        self._step_index = self._begin_index = None

    @classmethod
    def from_config(cls, config):
        return cls(**dict(config))

    @property
    def step_index(self):
        return self._step_index

    def set_begin_index(self, begin_index: int = 0):
        self._begin_index = begin_index

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None, timesteps=None):
        n_train = self.config.num_train_timesteps
        if sigmas is None:
            sigmas = np.linspace(1.0, 1 / n_train, num_inference_steps)
        sigmas = np.array(sigmas).astype(np.float32)
        if self.config.use_dynamic_shifting:
            sigmas = math.exp(mu) / (math.exp(mu) + (1 / sigmas - 1) ** 1.0)
        else:
            sigmas = self.config.shift * sigmas / (1 + (self.config.shift - 1) * sigmas)
        sig = torch.from_numpy(np.asarray(sigmas, dtype=np.float32)).to(device=device)
        self.timesteps = sig * n_train
        self.sigmas = torch.cat([sig, torch.zeros(1, device=sig.device)])
        self.num_inference_steps = len(sig)
        self._step_index = self._begin_index

    def _init_step_index(self, timestep):
        if self._begin_index is None:
            self._step_index = int((self.timesteps == timestep).nonzero()[0])
        else:
            self._step_index = self._begin_index

    def step(self, model_output, timestep, sample, return_dict=True, **kw):
        if self._step_index is None:
            self._init_step_index(timestep)
        dt = self.sigmas[self._step_index + 1] - self.sigmas[self._step_index]
        prev = (sample.to(torch.float32) + dt * model_output).to(model_output.dtype)
        self._step_index += 1
        return (prev,) if not return_dict else SimpleNamespace(prev_sample=prev)


from regione_b200.schedule import latent_image_ids  # noqa: E402,F401  (re-exported: part of the product)


class FluxKontextPipeline:
    """Minimal pipeline object: what `warp_modules` touches (class, scheduler, transformer) plus the latent-space
    entry used offline. Text encoders and VAE are out of scope (no weights), so callers pass
    `prompt_embeds`, `pooled_prompt_embeds`, packed `latents` and packed `image_latents`, and get latents back."""

    def __init__(self, transformer: FluxTransformer2DModel, scheduler: FlowMatchEulerDiscreteScheduler | None = None):
        self.transformer = transformer
        self.scheduler = scheduler or FlowMatchEulerDiscreteScheduler()
        self.vae_scale_factor = 8
        self.latent_channels = 16
        self._interrupt = False
        self._joint_attention_kwargs = None

    @property
    def _execution_device(self):
        return self.transformer.x_embedder.weight.device

    @property
    def interrupt(self):
        return self._interrupt

    @property
    def joint_attention_kwargs(self):
        return self._joint_attention_kwargs

    # latent <-> token layout either side of the loop (diffusers FluxKontextPipeline statics the reference calls through
    # prepare_latents and at inplace.py:398), on the CUDA library's pack kernels
    @staticmethod
    def _pack_latents(latents, batch_size=None, num_channels_latents=None, height=None, width=None):
        from regione_b200 import ops
        return ops.pack_latents(latents)

    @staticmethod
    def _unpack_latents(latents, height, width, vae_scale_factor):
        from regione_b200 import ops
        return ops.unpack_latents(latents, height, width, vae_scale_factor)

    @staticmethod
    def _prepare_latent_image_ids(batch_size, height, width, device, dtype):
        return latent_image_ids(height, width, 0.0, device, dtype)

    def __call__(self, *args, **kwargs):
        raise NotImplementedError(
            "vanilla FluxKontextPipeline.__call__ is diffusers code; enable RegionE first: "
            "RegionEHelper(pipeline).enable()")


class Step1XEditPipeline(FluxKontextPipeline):
    pass


# ------------------------------------------------------------------------------------------------ Qwen-Image-Edit
class _ModSeq(nn.Sequential):
    """diffusers `img_mod` / `txt_mod`: nn.Sequential(SiLU, Linear(dim, 6 dim)); index 1 is the Linear."""

    def __init__(self, dim):
        super().__init__(nn.SiLU(), nn.Linear(dim, 6 * dim))


class _QwenBlock(nn.Module):
    def __init__(self, dim, heads, ratio):
        super().__init__()
        self.img_mod, self.txt_mod = _ModSeq(dim), _ModSeq(dim)
        self.attn = _Attention(dim, heads, context=True)
        self.img_mlp, self.txt_mlp = _FeedForward(dim, ratio * dim), _FeedForward(dim, ratio * dim)


class _QwenTimeEmbed(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.timestep_embedder = _MLPEmbed(256, dim)


class QwenEmbedRope(nn.Module):
    """Rotary table of Qwen-Image (diffusers QwenEmbedRope, theta 10000, axes (16, 56, 56), scale_rope=True),
    restated: complex frequencies per (frame, row, column) with rows / columns centred around zero; text tokens
    continue after the largest image extent. Returns (img_freqs [N, 64] complex64, txt_freqs [T, 64] complex64)."""

    def __init__(self, theta=10000, axes_dim=(16, 56, 56)):
        super().__init__()
        self.theta, self.axes_dim = theta, axes_dim

    def _table(self, index, dim):
        freqs = torch.outer(index.float(), 1.0 / torch.pow(self.theta, torch.arange(0, dim, 2).float() / dim))
        return torch.polar(torch.ones_like(freqs), freqs)

    def forward(self, video_fhw, txt_seq_lens, device=None):
        pos = torch.cat([self._table(torch.arange(4096), d) for d in self.axes_dim], dim=1)
        neg = torch.cat([self._table(torch.arange(4096).flip(0) * -1 - 1, d) for d in self.axes_dim], dim=1)
        if isinstance(video_fhw, list) and isinstance(video_fhw[0], list):
            video_fhw = video_fhw[0]
        out, max_vid = [], 0
        a0, a1, a2 = [d // 2 for d in self.axes_dim]
        for idx, (f, h, w) in enumerate(video_fhw):
            fp = pos[:, :a0][idx:idx + f].view(f, 1, 1, -1).expand(f, h, w, -1)
            hh = torch.cat([neg[:, a0:a0 + a1][-(h - h // 2):], pos[:, a0:a0 + a1][:h // 2]], dim=0)
            ww = torch.cat([neg[:, a0 + a1:][-(w - w // 2):], pos[:, a0 + a1:][:w // 2]], dim=0)
            hp = hh.view(1, h, 1, -1).expand(f, h, w, -1)
            wp = ww.view(1, 1, w, -1).expand(f, h, w, -1)
            out.append(torch.cat([fp, hp, wp], dim=-1).reshape(f * h * w, -1))
            max_vid = max(max_vid, h // 2, w // 2)
        txt = pos[max_vid:max_vid + max(txt_seq_lens)]
        return torch.cat(out, dim=0).to(device), txt.to(device)


class QwenImageTransformer2DModel(nn.Module):
    """Weight container with the module surface of diffusers' QwenImageTransformer2DModel (60 dual-stream blocks)."""

    def __init__(self, dim=3072, heads=24, n_blocks=60, mlp_ratio=4, in_channels=64, ctx_dim=3584):
        super().__init__()
        self.config = _Config(in_channels=in_channels, guidance_embeds=False, num_layers=n_blocks,
                              attention_head_dim=dim // heads, num_attention_heads=heads, joint_attention_dim=ctx_dim)
        self.img_in = nn.Linear(in_channels, dim)
        self.txt_norm = _RMSNorm(ctx_dim)
        self.txt_in = nn.Linear(ctx_dim, dim)
        self.time_text_embed = _QwenTimeEmbed(dim)
        self.pos_embed = QwenEmbedRope()
        self.transformer_blocks = nn.ModuleList([_QwenBlock(dim, heads, mlp_ratio) for _ in range(n_blocks)])
        self.norm_out = _AdaNorm(dim, 2)
        self.proj_out = nn.Linear(dim, in_channels)

    def forward(self, *args, **kwargs):
        raise NotImplementedError("the vanilla QwenImageTransformer2DModel.forward is diffusers code; enable RegionE")

    init_synthetic = FluxTransformer2DModel.init_synthetic


class QwenImageEditPipeline(FluxKontextPipeline):
    """Stand-in with the class name `RegionEHelper` dispatches on (RegionE/tool/RegionE.py:5)."""

    def __init__(self, transformer, scheduler=None):
        super().__init__(transformer, scheduler)
        self._attention_kwargs = None

    @property
    def _execution_device(self):
        return self.transformer.img_in.weight.device

    @property
    def attention_kwargs(self):
        return self._attention_kwargs
