"""Python owner of one `rge_handle`: hands the pipeline's weight pointers to the C library once and runs one
transformer forward per call. Device memory, streams and lifetime management only — all math is in the CUDA library.

What is read off the transformer is exactly the module surface the reference's patched forward and processor touch
(RegionE/FluxKontext/inplace.py:469-567, 715-820).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import D as DS
from ._lib import G as GS
from ._lib import S as SS
from ._lib import check, ptr, stream_ptr


def _w(t: torch.Tensor, name: str) -> torch.Tensor:
    t = t.detach()
    if not t.is_cuda or t.dtype != torch.bfloat16 or not t.is_contiguous():
        raise _lib.RegionEB200Error(
            f"{name}: weights must be contiguous bf16 CUDA tensors (got {t.dtype}, {t.device}); there is no CPU path")
    return t


class FluxEngine:
    """FLUX-style DiT (double + single stream blocks; also Step1X-Edit's block stack)."""

    def __init__(self, transformer, txt_len: int, lat_len: int, cond_len: int, n_pass: int = 1,
                 shared_cache: bool = False):
        self.lib = _lib.load()
        tr = transformer
        blocks, singles = list(tr.transformer_blocks), list(tr.single_transformer_blocks)
        dim, in_ch = tr.x_embedder.weight.shape
        tte = tr.time_text_embed
        self.guidance_embeds = hasattr(tte, "guidance_embedder") and tte.guidance_embedder is not None
        ref_attn = (blocks[0] if blocks else singles[0]).attn
        mlp_dim = blocks[0].ff.net[0].proj.weight.shape[0] if blocks else singles[0].proj_mlp.weight.shape[0]
        cfg = _lib.Config(
            dim=dim, heads=ref_attn.heads, n_double=len(blocks), n_single=len(singles), mlp_ratio=mlp_dim // dim,
            in_channels=in_ch, ctx_dim=tr.context_embedder.weight.shape[1],
            pooled_dim=tte.text_embedder.linear_1.weight.shape[1], txt_len=txt_len, lat_len=lat_len,
            cond_len=cond_len, guidance_embeds=int(self.guidance_embeds), n_pass=n_pass,
            device=tr.x_embedder.weight.device.index or 0, shared_cache=int(shared_cache))
        self.cfg = cfg
        self.key = (txt_len, lat_len, cond_len, n_pass)
        self.in_channels = in_ch
        self._keep = []  # borrowed by the library: keep the tensors alive as long as the handle
        self._src = []   # (module, attribute, data_ptr, shape) of every registered weight, for weights_current()
        self._h = C.c_void_p()
        check(self.lib.rge_create(C.byref(cfg), C.byref(self._h)), "rge_create")
        try:
            self._register(tr, blocks, singles)
            check(self.lib.rge_finalize_weights(self._h), "rge_finalize_weights")
        except Exception:
            self.close()
            raise

    # ------------------------------------------------------------------ weights
    def _set(self, kind, index, slot, tensor, name, owner=None, attr=None):
        t = _w(tensor, name)
        self._keep.append(t)
        if owner is not None:
            self.__dict__.setdefault("_src", []).append((owner, attr, t.data_ptr(), tuple(t.shape)))
        check(self.lib.rge_set_weight(self._h, kind, index, slot, ptr(t)), f"rge_set_weight({name})")

    def _lin(self, kind, index, table, stem, mod, name):
        self._set(kind, index, table[stem + "_W"], mod.weight, name + ".weight", mod, "weight")
        self._set(kind, index, table[stem + "_B"], mod.bias, name + ".bias", mod, "bias")

    def weights_current(self) -> bool:
        """The handle BORROWS raw weight pointers (the reference reads the live module weights on every call). True
        iff every registered parameter still lives where it was registered, with the same shape, in bf16 on the GPU;
        `pipeline.to(...)`, LoRA fusing, `load_state_dict(assign=True)` or CPU offload re-allocate parameters and make
        this False, upon which the caller rebuilds the engine (`_get_engine`) instead of running on stale memory."""
        for owner, attr, p, shape in self.__dict__.get("_src", ()):
            t = getattr(owner, attr, None)
            if t is None or t.data_ptr() != p or tuple(t.shape) != shape or t.dtype != torch.bfloat16 or not t.is_cuda:
                return False
        return True

    def _register(self, tr, blocks, singles):
        g, d, s = _lib.BLK_GLOBAL, _lib.BLK_DOUBLE, _lib.BLK_SINGLE
        tte = tr.time_text_embed
        self._lin(g, 0, GS, "X_EMBED", tr.x_embedder, "x_embedder")
        self._lin(g, 0, GS, "CTX_EMBED", tr.context_embedder, "context_embedder")
        self._lin(g, 0, GS, "TIME1", tte.timestep_embedder.linear_1, "timestep_embedder.linear_1")
        self._lin(g, 0, GS, "TIME2", tte.timestep_embedder.linear_2, "timestep_embedder.linear_2")
        if self.guidance_embeds:
            self._lin(g, 0, GS, "GUID1", tte.guidance_embedder.linear_1, "guidance_embedder.linear_1")
            self._lin(g, 0, GS, "GUID2", tte.guidance_embedder.linear_2, "guidance_embedder.linear_2")
        self._lin(g, 0, GS, "POOL1", tte.text_embedder.linear_1, "text_embedder.linear_1")
        self._lin(g, 0, GS, "POOL2", tte.text_embedder.linear_2, "text_embedder.linear_2")
        self._lin(g, 0, GS, "NORM_OUT", tr.norm_out.linear, "norm_out.linear")
        self._lin(g, 0, GS, "PROJ_OUT", tr.proj_out, "proj_out")
        self._register_blocks(blocks, singles)

    def _register_blocks(self, blocks, singles):
        """FLUX-style double / single blocks (shared with Step1X-Edit, whose block stack is the same)."""
        d, s = _lib.BLK_DOUBLE, _lib.BLK_SINGLE
        for i, b in enumerate(blocks):
            n = f"transformer_blocks.{i}."
            a = b.attn
            self._lin(d, i, DS, "MOD", b.norm1.linear, n + "norm1.linear")
            self._lin(d, i, DS, "MOD_CTX", b.norm1_context.linear, n + "norm1_context.linear")
            self._lin(d, i, DS, "Q", a.to_q, n + "attn.to_q")
            self._lin(d, i, DS, "K", a.to_k, n + "attn.to_k")
            self._lin(d, i, DS, "V", a.to_v, n + "attn.to_v")
            self._lin(d, i, DS, "ADD_Q", a.add_q_proj, n + "attn.add_q_proj")
            self._lin(d, i, DS, "ADD_K", a.add_k_proj, n + "attn.add_k_proj")
            self._lin(d, i, DS, "ADD_V", a.add_v_proj, n + "attn.add_v_proj")
            self._set(d, i, DS["NORM_Q"], a.norm_q.weight, n + "attn.norm_q.weight", a.norm_q, "weight")
            self._set(d, i, DS["NORM_K"], a.norm_k.weight, n + "attn.norm_k.weight", a.norm_k, "weight")
            self._set(d, i, DS["NORM_ADD_Q"], a.norm_added_q.weight, n + "attn.norm_added_q.weight", a.norm_added_q,
                      "weight")
            self._set(d, i, DS["NORM_ADD_K"], a.norm_added_k.weight, n + "attn.norm_added_k.weight", a.norm_added_k,
                      "weight")
            self._lin(d, i, DS, "OUT", a.to_out[0], n + "attn.to_out.0")
            self._lin(d, i, DS, "ADD_OUT", a.to_add_out, n + "attn.to_add_out")
            self._lin(d, i, DS, "FF_UP", b.ff.net[0].proj, n + "ff.net.0.proj")
            self._lin(d, i, DS, "FF_DOWN", b.ff.net[2], n + "ff.net.2")
            self._lin(d, i, DS, "FFC_UP", b.ff_context.net[0].proj, n + "ff_context.net.0.proj")
            self._lin(d, i, DS, "FFC_DOWN", b.ff_context.net[2], n + "ff_context.net.2")
        for i, b in enumerate(singles):
            n = f"single_transformer_blocks.{i}."
            a = b.attn
            self._lin(s, i, SS, "MOD", b.norm.linear, n + "norm.linear")
            self._lin(s, i, SS, "Q", a.to_q, n + "attn.to_q")
            self._lin(s, i, SS, "K", a.to_k, n + "attn.to_k")
            self._lin(s, i, SS, "V", a.to_v, n + "attn.to_v")
            self._set(s, i, SS["NORM_Q"], a.norm_q.weight, n + "attn.norm_q.weight", a.norm_q, "weight")
            self._set(s, i, SS["NORM_K"], a.norm_k.weight, n + "attn.norm_k.weight", a.norm_k, "weight")
            self._lin(s, i, SS, "MLP", b.proj_mlp, n + "proj_mlp")
            self._lin(s, i, SS, "OUT", b.proj_out, n + "proj_out")

    # ------------------------------------------------------------------ per image / per step
    def begin_image(self, txt_ids, img_ids, prompt_embeds, pooled, guidance_x1000: float, pass_id: int = 0):
        """txt_ids [T,3], img_ids [L+C,3] (any float dtype), prompt_embeds [T,ctx] bf16, pooled [pooled] bf16."""
        ti = txt_ids.to(torch.float32).contiguous()
        ii = img_ids.to(torch.float32).contiguous()
        pe = prompt_embeds.reshape(-1, prompt_embeds.shape[-1]).contiguous()
        po = pooled.reshape(-1).contiguous()
        if pe.dtype != torch.bfloat16 or po.dtype != torch.bfloat16:
            raise _lib.RegionEB200Error("prompt embeddings must be bf16")
        if ti.shape[0] != self.cfg.txt_len or ii.shape[0] != self.cfg.lat_len + self.cfg.cond_len:
            raise _lib.RegionEB200Error("id tensors do not match the engine's sequence lengths")
        check(self.lib.rge_begin_image(self._h, pass_id, ptr(ti), ptr(ii), ptr(pe), ptr(po), float(guidance_x1000),
                                       stream_ptr()), "rge_begin_image")

    @staticmethod
    def _rows(x, what):
        if x is None:
            return None, 0
        x = x.contiguous()
        if x.dtype != torch.bfloat16 or not x.is_cuda:
            raise _lib.RegionEB200Error(f"{what} must be bf16 CUDA tensors")
        return x, x.shape[0]

    def step(self, x_in, sel, timestep_x1000: float, n_out: int, pass_id: int = 0, out=None, x_cond=None):
        """x_in [n_x, C] bf16 (+ x_cond [n_cond, C]: the condition rows that follow, read in place instead of a
        concatenated copy); sel int32 [n_x + n_cond] or None (identity over L+C); returns velocity [n_out, C]."""
        x, n_x = self._rows(x_in, "latents")
        xc, n_c = self._rows(x_cond, "condition latents")
        if out is None:
            out = torch.empty(n_out, self.in_channels, dtype=torch.bfloat16, device=x.device)
        sel_ptr = ptr(sel)
        if sel is not None and sel.numel() == 0:   # empty edited set: NULL would mean "identity"
            sel_ptr = ptr(self._dummy_sel(x.device))
        check(self.lib.rge_dit_step(self._h, pass_id, ptr(x) if n_x else None, n_x, ptr(xc) if n_c else None, n_c,
                                    sel_ptr, float(timestep_x1000), ptr(out) if n_out else None, n_out, stream_ptr()),
              "rge_dit_step")
        return out

    def _dummy_sel(self, device):
        if getattr(self, "_dummy", None) is None:
            self._dummy = torch.zeros(1, dtype=torch.int32, device=device)
        return self._dummy

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            torch.cuda.synchronize()
            self.lib.rge_destroy(self._h)
            self._h = C.c_void_p()
        self._keep = []

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def cached_engine(transformer, key, factory):
    """One resident engine (KV cache) per transformer, keyed on the sequence shape. Rebuilt when the shape changes or
    when any borrowed weight pointer went stale (`weights_current`); the stale handle is destroyed first."""
    cache = transformer.__dict__.setdefault("_regione_b200_engines", {})
    eng = cache.get(key)
    if eng is not None and not eng.weights_current():
        eng = None
    if eng is None:
        for old in list(cache.values()):   # shapes rarely change: never hold two KV caches
            old.close()
        cache.clear()
        eng = factory()
        cache[key] = eng
    return eng
