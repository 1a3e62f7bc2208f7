"""Oracle restatement of the FLUX.1-Kontext DiT under RegionE's patched forward. TEST INFRASTRUCTURE.

Block math: diffusers FluxTransformer2DModel / FluxTransformerBlock / FluxSingleTransformerBlock / FluxPosEmbed /
CombinedTimestepGuidanceTextProjEmbeddings — third-party code that is NOT in /root/reference and not installed here
(the reference pins only "latest of git+https://github.com/Peyton-Chen/diffusers.git@step1xedit_v1p2", README.md:76-77),
restated from the published architecture. PARITY UNPINNED for these parts (no reference-owned test or fixture exists).
RegionE-owned logic follows RegionE/FluxKontext/inplace.py: forward :469-567, attention processor :712-824,
`_partially_linear` fused_kernels.py:81-101 (including its fp32->fp16->bf16 store, :80).

Weights come as a flat dict with diffusers parameter names (e.g. "transformer_blocks.0.attn.to_q.weight").
Plain torch ops on bf16 tensors are used on purpose: rounding points then fall where the reference's do.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def timestep_projection(t: torch.Tensor, dim: int = 256) -> torch.Tensor:
    """diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): fp32 [B, dim], cos first."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device)
    exponent = exponent / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def rope_cos_sin(ids: torch.Tensor, axes=(16, 56, 56), theta: float = 10000.0):
    """FluxPosEmbed: per axis outer(pos, theta^(-2j/d)) in float64, cos/sin repeat-interleaved x2, fp32 [S,128]."""
    pos = ids.float()
    cos, sin = [], []
    for a, d in enumerate(axes):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float64, device=ids.device) / d))
        ang = torch.outer(pos[:, a].double(), freqs)
        cos.append(ang.cos().repeat_interleave(2, dim=1).float())
        sin.append(ang.sin().repeat_interleave(2, dim=1).float())
    return torch.cat(cos, dim=-1), torch.cat(sin, dim=-1)


def apply_rope(x: torch.Tensor, cs) -> torch.Tensor:
    """diffusers apply_rotary_emb(use_real=True, unbind_dim=-1) on [B,H,S,128]."""
    cos, sin = cs
    cos, sin = cos[None, None], sin[None, None]
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    return (x.float() * cos + rot.float() * sin).to(x.dtype)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """diffusers RMSNorm: fp32 variance, x * rsqrt (fp32), cast to the weight dtype, times weight."""
    var = x.float().pow(2).mean(-1, keepdim=True)
    y = x * torch.rsqrt(var + eps)
    return y.to(w.dtype) * w


def exact_attention(q, k, v) -> torch.Tensor:
    """Stand-in for flash_attn_func(q,k,v, causal=False) (inplace.py:796-801): fp32 softmax, [B,H,S,128] in,
    [B,Sq,H*128] out in the input dtype."""
    B, H, Sq, hd = q.shape
    out = torch.empty(B, Sq, H * hd, dtype=q.dtype, device=q.device)
    scale = hd ** -0.5
    for b in range(B):
        for h in range(H):
            p = torch.softmax(q[b, h].float() @ k[b, h].float().t() * scale, dim=-1)
            out[b, :, h * hd:(h + 1) * hd] = (p @ v[b, h].float()).to(q.dtype)
    return out


def partially_linear(x, w, b, index, cache) -> None:
    """fused_kernels.py:81-101: cache[:, index, :] = x @ w^T + b with fp32 accumulation, stored through fp16 (:80)."""
    y = x.float() @ w.float().t()
    if b is not None:
        y = y + b.float()
    cache[:, index, :] = y.to(torch.float16).to(cache.dtype)


class FluxOracle:
    def __init__(self, weights: dict, heads: int, n_double: int, n_single: int, guidance_embeds: bool = True):
        self.w = weights
        self.heads, self.n_double, self.n_single, self.guidance_embeds = heads, n_double, n_single, guidance_embeds
        self.k_cache = {}
        self.v_cache = {}

    # ---- helpers
    def lin(self, name, x):
        return F.linear(x, self.w[name + ".weight"], self.w.get(name + ".bias"))

    def time_text_embed(self, timestep, guidance, pooled):
        dt = pooled.dtype
        p = "time_text_embed."
        t = self.lin(p + "timestep_embedder.linear_2",
                     F.silu(self.lin(p + "timestep_embedder.linear_1", timestep_projection(timestep).to(dt))))
        if self.guidance_embeds:
            g = self.lin(p + "guidance_embedder.linear_2",
                         F.silu(self.lin(p + "guidance_embedder.linear_1", timestep_projection(guidance).to(dt))))
            t = t + g
        pe = self.lin(p + "text_embedder.linear_2", F.silu(self.lin(p + "text_embedder.linear_1", pooled)))
        return t + pe

    def _heads(self, x):
        B, S, D = x.shape   # explicit head_dim: the reference's view(B, -1, heads, head_dim) (:756) is ambiguous for S = 0
        return x.view(B, S, self.heads, D // self.heads).transpose(1, 2)

    # ---- attention processor (inplace.py:704-824)
    def attn(self, prefix, layer, single, hidden, enc, rope_q, rope_k, st):
        cur, N = st.current_step, st.inference_step
        q = self.lin(prefix + "to_q", hidden)
        if cur < st.warmup_step - 1 or cur > N - st.post_step - 1:                      # :717
            k = self.lin(prefix + "to_k", hidden)
            v = self.lin(prefix + "to_v", hidden)
        elif cur == st.warmup_step - 1 or cur == st.prev_refresh_step:                   # :721
            k = self.lin(prefix + "to_k", hidden)
            v = self.lin(prefix + "to_v", hidden)
            self.k_cache[layer], self.v_cache[layer] = k, v
        else:                                                                            # :727
            ed = st.edited_ids.squeeze(0)
            sel = torch.cat((torch.arange(st.txt_length).to(ed), ed + st.txt_length)) if single else ed
            partially_linear(hidden, self.w[prefix + "to_k.weight"], self.w[prefix + "to_k.bias"], sel,
                             self.k_cache[layer])
            partially_linear(hidden, self.w[prefix + "to_v.weight"], self.w[prefix + "to_v.bias"], sel,
                             self.v_cache[layer])
            k, v = self.k_cache[layer], self.v_cache[layer]
        q, k, v = self._heads(q), self._heads(k), self._heads(v)
        q = rms_norm(q, self.w[prefix + "norm_q.weight"])                                # :760-763 (whole cache)
        k = rms_norm(k, self.w[prefix + "norm_k.weight"])
        if enc is not None:                                                              # :766-790
            eq = rms_norm(self._heads(self.lin(prefix + "add_q_proj", enc)), self.w[prefix + "norm_added_q.weight"])
            ek = rms_norm(self._heads(self.lin(prefix + "add_k_proj", enc)), self.w[prefix + "norm_added_k.weight"])
            ev = self._heads(self.lin(prefix + "add_v_proj", enc))
            q, k, v = torch.cat([eq, q], 2), torch.cat([ek, k], 2), torch.cat([ev, v], 2)
        q = apply_rope(q, rope_q)                                                        # :792-794
        k = apply_rope(k, rope_k)
        o = exact_attention(q, k, v)
        if enc is not None:                                                              # :809-822
            T = enc.shape[1]
            return self.lin(prefix + "to_out.0", o[:, T:]), self.lin(prefix + "to_add_out", o[:, :T])
        return o

    # ---- blocks
    @staticmethod
    def _ln(x):
        return F.layer_norm(x, (x.shape[-1],), eps=1e-6)

    def double_block(self, i, hidden, enc, temb, rope_q, rope_k, st):
        p = f"transformer_blocks.{i}."
        sh, sc, g, sh2, sc2, g2 = self.lin(p + "norm1.linear", F.silu(temb)).chunk(6, dim=1)
        csh, csc, cg, csh2, csc2, cg2 = self.lin(p + "norm1_context.linear", F.silu(temb)).chunk(6, dim=1)
        n = self._ln(hidden) * (1 + sc[:, None]) + sh[:, None]
        nc = self._ln(enc) * (1 + csc[:, None]) + csh[:, None]
        a, ac = self.attn(p + "attn.", i, False, n, nc, rope_q, rope_k, st)
        hidden = hidden + g.unsqueeze(1) * a
        n = self._ln(hidden) * (1 + sc2[:, None]) + sh2[:, None]
        ff = self.lin(p + "ff.net.2", F.gelu(self.lin(p + "ff.net.0.proj", n), approximate="tanh"))
        hidden = hidden + g2.unsqueeze(1) * ff
        enc = enc + cg.unsqueeze(1) * ac
        nc = self._ln(enc) * (1 + csc2[:, None]) + csh2[:, None]
        ffc = self.lin(p + "ff_context.net.2", F.gelu(self.lin(p + "ff_context.net.0.proj", nc), approximate="tanh"))
        enc = enc + cg2.unsqueeze(1) * ffc
        return enc, hidden

    def single_block(self, i, hidden, enc, temb, rope_q, rope_k, st):
        p = f"single_transformer_blocks.{i}."
        T = enc.shape[1]
        h = torch.cat([enc, hidden], dim=1)
        res = h
        sh, sc, g = self.lin(p + "norm.linear", F.silu(temb)).chunk(3, dim=1)
        n = self._ln(h) * (1 + sc[:, None]) + sh[:, None]
        mlp = F.gelu(self.lin(p + "proj_mlp", n), approximate="tanh")
        a = self.attn(p + "attn.", self.n_double + i, True, n, None, rope_q, rope_k, st)
        h = res + g.unsqueeze(1) * self.lin(p + "proj_out", torch.cat([a, mlp], dim=2))
        return h[:, :T], h[:, T:]

    # ---- patched forward (inplace.py:413-576)
    def forward(self, st, hidden_states, encoder_hidden_states, pooled, timestep, img_ids, txt_ids, guidance):
        h = self.lin("x_embedder", hidden_states)                                        # :469
        t = timestep.to(h.dtype) * 1000                                                  # :471
        g = guidance.to(h.dtype) * 1000 if guidance is not None else None
        temb = self.time_text_embed(t, g, pooled)                                        # :475-479
        enc = self.lin("context_embedder", encoder_hidden_states)                        # :480
        rope_q = rope_cos_sin(torch.cat((txt_ids, img_ids), dim=0))                      # :495-496 (current ids)
        rope_k = rope_cos_sin(torch.cat((txt_ids, st.latent_ids), dim=0))                # :499 (always the full ids)
        for i in range(self.n_double):
            enc, h = self.double_block(i, h, enc, temb, rope_q, rope_k, st)
        for i in range(self.n_single):
            enc, h = self.single_block(i, h, enc, temb, rope_q, rope_k, st)
        scale, shift = self.lin("norm_out.linear", F.silu(temb).to(h.dtype)).chunk(2, dim=1)   # :566 (scale first)
        h = self._ln(h) * (1 + scale)[:, None, :] + shift[:, None, :]
        return self.lin("proj_out", h)                                                   # :567
