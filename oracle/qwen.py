"""Oracle restatement of Qwen-Image-Edit under RegionE's patched forward / processor / loop. TEST INFRASTRUCTURE.

RegionE-owned logic follows RegionE/QwenImageEdit/inplace.py: loop :322-433 (CFG norm rescale :396-405), forward
:515-571, processor :747-890. The block math (QwenImageTransformerBlock, QwenTimestepProjEmbeddings,
apply_rotary_emb_qwen) is diffusers code that is neither in /root/reference nor installed: restated from the published
architecture — PARITY UNPINNED for those parts. The rotary frequencies are an INPUT (the pipeline's
own pos_embed output), exactly as in the reference's forward (:530).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import region_ops as ro
from .flux import exact_attention, partially_linear, rms_norm
from .loop import EulerState, scalar_times, scheduler_step
from .schedule import flow_match_sigmas

GAMMA_QWEN = [1.0195, 1.0233, 1.0243, 1.0185, 1.0321, 1.0208, 1.0260, 1.0233, 1.0258, 1.0292, 1.0316, 1.0306, 1.0289,
              1.0347, 1.0329, 1.0402, 1.0378, 1.0384, 1.0413, 1.0444, 1.0526, 1.0400, 1.0555, 1.0439, 1.0357, 1.0118,
              0.7603]   # QwenImageEdit/inplace.py:47-50


def apply_rope_complex(x: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """diffusers apply_rotary_emb_qwen(use_real=False) on [B,S,H,D]: pairs (2i, 2i+1) as complex numbers times
    freqs [S, D/2]."""
    xc = torch.view_as_complex(x.float().reshape(*x.shape[:-1], -1, 2))
    out = torch.view_as_real(xc * freqs.unsqueeze(1)).flatten(3)
    return out.type_as(x)


class QwenOracle:
    def __init__(self, weights: dict, heads: int, n_blocks: int):
        self.w, self.heads, self.n_blocks = weights, heads, n_blocks
        self.cache = {}   # (tag, layer) -> (k, v): k_cache_even/odd, v_cache_even/odd (:733-736)

    def lin(self, name, x):
        return F.linear(x, self.w[name + ".weight"], self.w.get(name + ".bias"))

    def _heads(self, x):
        return x.unflatten(-1, (self.heads, -1))   # [B,S,H,D] (:825-827)

    def attn(self, p, layer, img, txt, img_freqs_q, img_freqs_k, txt_freqs, st, tag):
        cur, N = st.current_step, st.inference_step
        q = self.lin(p + "to_q", img)
        if cur < st.warmup_step - 1 or cur > N - st.post_step - 1:                       # :749-752
            k, v = self.lin(p + "to_k", img), self.lin(p + "to_v", img)
        elif cur == st.warmup_step - 1 or cur == st.prev_refresh_step:                    # :754-759
            k, v = self.lin(p + "to_k", img), self.lin(p + "to_v", img)
            self.cache[(tag, layer)] = (k, v)
        else:                                                                             # :761-782
            kc, vc = self.cache[(tag, layer)]
            sel = st.edited_ids.squeeze(0)
            partially_linear(img, self.w[p + "to_k.weight"], self.w[p + "to_k.bias"], sel, kc)
            partially_linear(img, self.w[p + "to_v.weight"], self.w[p + "to_v.bias"], sel, vc)
            k, v = kc, vc
        tq, tk, tv = self.lin(p + "add_q_proj", txt), self.lin(p + "add_k_proj", txt), self.lin(p + "add_v_proj", txt)
        q, k, v = self._heads(q), self._heads(k), self._heads(v)
        q = rms_norm(q, self.w[p + "norm_q.weight"])                                      # :832-835
        k = rms_norm(k, self.w[p + "norm_k.weight"])
        tq, tk, tv = self._heads(tq), self._heads(tk), self._heads(tv)
        tq = rms_norm(tq, self.w[p + "norm_added_q.weight"])
        tk = rms_norm(tk, self.w[p + "norm_added_k.weight"])
        if q.shape[1] != 0:                                                               # :850-852
            q = apply_rope_complex(q, img_freqs_q)
        k = apply_rope_complex(k, img_freqs_k)
        tq, tk = apply_rope_complex(tq, txt_freqs), apply_rope_complex(tk, txt_freqs)
        jq, jk, jv = torch.cat([tq, q], 1), torch.cat([tk, k], 1), torch.cat([tv, v], 1)  # :859-861 text first
        o = exact_attention(jq.transpose(1, 2), jk.transpose(1, 2), jv.transpose(1, 2))
        T = txt.shape[1]
        return self.lin(p + "to_out.0", o[:, T:]), self.lin(p + "to_add_out", o[:, :T])    # :878-887

    @staticmethod
    def _ln(x):
        return F.layer_norm(x, (x.shape[-1],), eps=1e-6)

    def block(self, i, img, txt, temb, fq, fk, ft, st, tag):
        """QwenImageTransformerBlock."""
        p = f"transformer_blocks.{i}."
        im = self.lin(p + "img_mod.1", F.silu(temb))
        tm = self.lin(p + "txt_mod.1", F.silu(temb))
        (sh1, sc1, g1), (sh2, sc2, g2) = [m.chunk(3, dim=-1) for m in im.chunk(2, dim=-1)]
        (tsh1, tsc1, tg1), (tsh2, tsc2, tg2) = [m.chunk(3, dim=-1) for m in tm.chunk(2, dim=-1)]
        n = self._ln(img) * (1 + sc1.unsqueeze(1)) + sh1.unsqueeze(1)
        nt = self._ln(txt) * (1 + tsc1.unsqueeze(1)) + tsh1.unsqueeze(1)
        a, at = self.attn(p + "attn.", i, n, nt, fq, fk, ft, st, tag)
        img = img + g1.unsqueeze(1) * a
        txt = txt + tg1.unsqueeze(1) * at
        n = self._ln(img) * (1 + sc2.unsqueeze(1)) + sh2.unsqueeze(1)
        img = img + g2.unsqueeze(1) * self.lin(p + "img_mlp.net.2",
                                               F.gelu(self.lin(p + "img_mlp.net.0.proj", n), approximate="tanh"))
        nt = self._ln(txt) * (1 + tsc2.unsqueeze(1)) + tsh2.unsqueeze(1)
        txt = txt + tg2.unsqueeze(1) * self.lin(p + "txt_mlp.net.2",
                                                F.gelu(self.lin(p + "txt_mlp.net.0.proj", nt), approximate="tanh"))
        return txt, img

    def forward(self, st, hidden_states, encoder_hidden_states, timestep, img_freqs, txt_freqs, latent_ids, tag):
        """:515-571. img_freqs [L+C,64] / txt_freqs [T,64] complex = pos_embed(img_shapes, txt_seq_lens)."""
        h = self.lin("img_in", hidden_states)                                              # :515
        t = timestep.to(h.dtype)                                                           # :517
        enc = self.lin("txt_in", rms_norm(encoder_hidden_states, self.w["txt_norm.weight"]))   # :518-519
        # diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0, scale=1000): angle = (t * f) * 1000
        ang = 1000 * (t[:, None].float() * self._freq().to(t.device)[None, :])
        tp = torch.cat([torch.cos(ang), torch.sin(ang)], dim=-1)
        temb = self.lin("time_text_embed.timestep_embedder.linear_2",
                        F.silu(self.lin("time_text_embed.timestep_embedder.linear_1", tp.to(h.dtype))))   # :524-528
        fq = img_freqs[latent_ids, :]                                                      # :531 query rows
        for i in range(self.n_blocks):
            enc, h = self.block(i, h, enc, temb, fq, img_freqs, txt_freqs, st, tag)
        scale, shift = self.lin("norm_out.linear", F.silu(temb).to(h.dtype)).chunk(2, dim=1)   # :561
        h = self._ln(h) * (1 + scale)[:, None, :] + shift[:, None, :]
        return self.lin("proj_out", h)                                                     # :562

    @staticmethod
    def _freq(half: int = 128):
        import math
        e = -math.log(10000) * torch.arange(half, dtype=torch.float32)
        return torch.exp(e / half)


def cfg_norm_rescaled(pos, neg, true_cfg_scale):
    """Qwen-Image classifier-free guidance with norm rescaling, RegionE/QwenImageEdit/inplace.py:401-405 (same in
    QwenImageEditPlus). Pinned by tests/golden/cfg.pt (the reference's own lines exec'd)."""
    comb = neg + true_cfg_scale * (pos - neg)
    cond_norm = torch.norm(pos, dim=-1, keepdim=True)
    noise_norm = torch.norm(comb, dim=-1, keepdim=True)
    return comb * (cond_norm / noise_norm)


def run_regione_qwen(model: QwenOracle, params: dict, latents, image_latents, prompt_embeds, negative_prompt_embeds,
                     true_cfg_scale, img_freqs, txt_freqs, neg_txt_freqs, height, width, record=False):
    """QwenImageEdit/inplace.py:322-433 for output_type='latent'."""
    st = ro.RegionState()
    st.set_parameters(params)
    n_steps = params["num_inference_steps"]
    sigmas, timesteps = flow_match_sigmas(n_steps, latents.shape[1])
    g = torch.tensor(params.get("gamma") or GAMMA_QWEN, dtype=torch.float16, device=latents.device)
    assert g.numel() == n_steps - 1
    sigmas, timesteps = sigmas.to(latents.device), timesteps.to(latents.device)
    sch = EulerState(sigmas, timesteps)
    latent_ids = torch.arange(latents.shape[1] + image_latents.shape[1], device=latents.device)   # :322
    st.refresh(latents, image_latents, latent_ids, torch.empty(prompt_embeds.shape[1], 0), height, width)
    cache, accumulate = None, 1
    do_cfg = negative_prompt_embeds is not None
    trace = {"modes": [], "latents": [], "noise_pred": []}
    for i, t in enumerate(timesteps):
        assert i == st.current_step
        cur, N = st.current_step, st.inference_step
        if cur <= st.warmup_step or cur > N - st.post_step - 1 or cur == st.prev_refresh_step:   # :334-350
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
        if should_cache:                                                                   # :352-356
            if cache.shape[1] != latents.shape[1]:
                cache = ro.gather_rows(cache, st.edited_ids)
            noise_pred = scalar_times(ratio, cache)
            trace["modes"].append("SKIP")
        else:
            x_in = latents
            full = cur <= st.warmup_step - 1 or cur > N - st.post_step - 1 or cur == st.prev_refresh_step
            if full:
                x_in = torch.cat([latents, image_latents], dim=1)
            timestep = t.expand(latents.shape[0]).to(latents.dtype)                        # :369
            noise_pred = model.forward(st, x_in, prompt_embeds, timestep / 1000, img_freqs, txt_freqs, latent_ids,
                                       "cond")[:, : latents.size(1)]
            if do_cfg:                                                                     # :386-405
                neg = model.forward(st, x_in, negative_prompt_embeds, timestep / 1000, img_freqs, neg_txt_freqs,
                                    latent_ids, "uncond")[:, : latents.size(1)]
                noise_pred = cfg_norm_rescaled(noise_pred, neg, true_cfg_scale)
            cache = noise_pred
            trace["modes"].append("FULL" if full else "REGION")
        latents = scheduler_step(sch, st, noise_pred, latents, trace)
        latents, latent_ids = st.step(latents, latent_ids)                                 # :433
        if record:
            trace["latents"].append(latents.clone())
            trace["noise_pred"].append(noise_pred.clone())
    trace["edited_ids"], trace["unedited_ids"] = st.edited_ids, st.unedited_ids
    return latents, trace
