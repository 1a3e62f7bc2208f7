"""`rge_handle` owner for Qwen-Image-Edit: 60 dual-stream blocks, no single-stream blocks, two passes (cond / uncond)
with their own K/V caches, rotary table and embedded context supplied by the pipeline's own pos_embed / txt_norm+txt_in
(external_embed bit 0), timestep embedding evaluated inside the library (no pooled / guidance term).

Module surface read off the transformer = what the reference's patched forward and processor touch
(RegionE/QwenImageEdit/inplace.py:515-571, 747-890).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, ops
from ._lib import D as DS
from ._lib import G as GS
from ._lib import check, ptr, stream_ptr
from .engine import FluxEngine


class QwenEngine(FluxEngine):
    def __init__(self, transformer, txt_len: int, lat_len: int, cond_len: int, n_pass: int = 2):
        self.lib = _lib.load()
        tr = transformer
        blocks = list(tr.transformer_blocks)
        dim, in_ch = tr.img_in.weight.shape
        mlp_dim = blocks[0].img_mlp.net[0].proj.weight.shape[0]
        self.guidance_embeds = False
        cfg = _lib.Config(
            dim=dim, heads=blocks[0].attn.heads, n_double=len(blocks), n_single=0, mlp_ratio=mlp_dim // dim,
            in_channels=in_ch, ctx_dim=0, pooled_dim=0, txt_len=txt_len, lat_len=lat_len, cond_len=cond_len,
            guidance_embeds=0, n_pass=n_pass, device=tr.img_in.weight.device.index or 0, external_embed=1)
        self.cfg = cfg
        self.key = (txt_len, lat_len, cond_len, n_pass)
        self.in_channels = in_ch
        self.transformer = tr
        self._keep = []
        self._src = []
        self._h = C.c_void_p()
        check(self.lib.rge_create(C.byref(cfg), C.byref(self._h)), "rge_create")
        try:
            self._register_qwen(tr, blocks)
            check(self.lib.rge_finalize_weights(self._h), "rge_finalize_weights")
        except Exception:
            self.close()
            raise

    def _register_qwen(self, tr, blocks):
        g, d = _lib.BLK_GLOBAL, _lib.BLK_DOUBLE
        te = tr.time_text_embed.timestep_embedder
        self._lin(g, 0, GS, "X_EMBED", tr.img_in, "img_in")
        self._lin(g, 0, GS, "TIME1", te.linear_1, "time_text_embed.timestep_embedder.linear_1")
        self._lin(g, 0, GS, "TIME2", te.linear_2, "time_text_embed.timestep_embedder.linear_2")
        self._lin(g, 0, GS, "NORM_OUT", tr.norm_out.linear, "norm_out.linear")
        self._lin(g, 0, GS, "PROJ_OUT", tr.proj_out, "proj_out")
        for i, b in enumerate(blocks):
            n = f"transformer_blocks.{i}."
            a = b.attn
            self._lin(d, i, DS, "MOD", b.img_mod[1], n + "img_mod.1")
            self._lin(d, i, DS, "MOD_CTX", b.txt_mod[1], n + "txt_mod.1")
            self._lin(d, i, DS, "Q", a.to_q, n + "attn.to_q")
            self._lin(d, i, DS, "K", a.to_k, n + "attn.to_k")
            self._lin(d, i, DS, "V", a.to_v, n + "attn.to_v")
            self._lin(d, i, DS, "ADD_Q", a.add_q_proj, n + "attn.add_q_proj")
            self._lin(d, i, DS, "ADD_K", a.add_k_proj, n + "attn.add_k_proj")
            self._lin(d, i, DS, "ADD_V", a.add_v_proj, n + "attn.add_v_proj")
            self._set(d, i, DS["NORM_Q"], a.norm_q.weight, n + "attn.norm_q.weight", a.norm_q, "weight")
            self._set(d, i, DS["NORM_K"], a.norm_k.weight, n + "attn.norm_k.weight", a.norm_k, "weight")
            self._set(d, i, DS["NORM_ADD_Q"], a.norm_added_q.weight, n + "attn.norm_added_q.weight", a.norm_added_q, "weight")
            self._set(d, i, DS["NORM_ADD_K"], a.norm_added_k.weight, n + "attn.norm_added_k.weight", a.norm_added_k, "weight")
            self._lin(d, i, DS, "OUT", a.to_out[0], n + "attn.to_out.0")
            self._lin(d, i, DS, "ADD_OUT", a.to_add_out, n + "attn.to_add_out")
            self._lin(d, i, DS, "FF_UP", b.img_mlp.net[0].proj, n + "img_mlp.net.0.proj")
            self._lin(d, i, DS, "FF_DOWN", b.img_mlp.net[2], n + "img_mlp.net.2")
            self._lin(d, i, DS, "FFC_UP", b.txt_mlp.net[0].proj, n + "txt_mlp.net.0.proj")
            self._lin(d, i, DS, "FFC_DOWN", b.txt_mlp.net[2], n + "txt_mlp.net.2")

    def begin_image_qwen(self, img_freqs, txt_freqs, prompt_embeds, pass_id: int):
        """img_freqs [L+C, 64] / txt_freqs [T, 64] complex (pipeline's pos_embed, QwenImageEdit/inplace.py:530);
        prompt_embeds [T, ctx] bf16: txt_in(txt_norm(.)) (:518-519) runs here on the library's kernels."""
        tr = self.transformer
        pe = prompt_embeds.reshape(-1, prompt_embeds.shape[-1]).contiguous()
        if pe.dtype != torch.bfloat16 or not pe.is_cuda:
            raise _lib.RegionEB200Error("prompt embeddings must be bf16 CUDA tensors")
        if pe.shape[0] != self.cfg.txt_len or img_freqs.shape[0] != self.cfg.lat_len + self.cfg.cond_len:
            raise _lib.RegionEB200Error("sequence lengths do not match the engine")
        ctx = ops.gemm(ops.rmsnorm(pe, tr.txt_norm.weight.detach(), float(getattr(tr.txt_norm, "eps", 1e-6))),
                       tr.txt_in.weight.detach(), tr.txt_in.bias.detach())
        cs = torch.view_as_real(torch.cat([txt_freqs, img_freqs], dim=0).to(torch.complex64)).contiguous()
        check(self.lib.rge_begin_image_ex(self._h, pass_id, ptr(cs), ptr(ctx), stream_ptr()), "rge_begin_image_ex")
