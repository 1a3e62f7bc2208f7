"""Torch-tensor wrappers over the kernel-level C-ABI entry points (device memory and streams only; no compute here)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import AttnDesc, GemmDesc, check, ptr, stream_ptr


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise _lib.RegionEB200Error(f"{name}: expected a CUDA tensor (no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.RegionEB200Error(f"{name}: expected {dtype}, got {t.dtype}")
    if t.dim() >= 2 and t.stride(-1) != 1:
        raise _lib.RegionEB200Error(f"{name}: innermost dimension must be contiguous")


def _gemm_desc(d, a, w, bias=None, *, epilogue=_lib.EPI_STORE, out=None, row_map=None, row_off=0, col_off=0, gate=None,
               res=None, norm_w=None, rope_cs=None, rope_map=None, rope_off=0, rope_ld=0, fp16_roundtrip=False):
    _req(a, torch.bfloat16, "a"); _req(w, torch.bfloat16, "w")
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=torch.bfloat16, device=a.device)
    _req(out, torch.bfloat16, "out")
    d.A, d.lda = ptr(a), a.stride(0)
    d.W, d.ldw = ptr(w), w.stride(0)
    d.bias = ptr(bias)
    d.M, d.N, d.K = M, N, K
    d.epilogue = epilogue
    d.out, d.ldo = ptr(out), out.stride(0)
    d.row_map, d.row_off, d.col_off = ptr(row_map), row_off, col_off
    d.gate, d.res, d.ldr = ptr(gate), ptr(res), (res.stride(0) if res is not None else 0)
    d.norm_w, d.rope_cs = ptr(norm_w), ptr(rope_cs)
    d.rope_map, d.rope_off = ptr(rope_map), rope_off
    d.rope_ld = rope_ld              # 0: rope_cs is [S,64,2]; > 0: pair-major [64, rope_ld, 2]
    d.flags = _lib.GEMM_FP16_ROUNDTRIP if fp16_roundtrip else 0
    return out


def gemm(a, w, bias=None, **kw):
    """out[(row_map[m] or m)+row_off, col_off+n] = epilogue(a[M,K] @ w[N,K]^T + bias)."""
    lib = _lib.load()
    d = GemmDesc()
    out = _gemm_desc(d, a, w, bias, **kw)
    check(lib.rge_op_gemm(C.byref(d), stream_ptr()), "rge_op_gemm")
    return out


def gemm_group(members):
    """`members`: list of (a, w, bias, kwargs-of-gemm) run as ONE persistent launch; returns the list of outputs."""
    lib = _lib.load()
    descs = (GemmDesc * len(members))()
    outs = [_gemm_desc(descs[i], a, w, bias, **kw) for i, (a, w, bias, kw) in enumerate(members)]
    check(lib.rge_op_gemm_group(descs, len(members), stream_ptr()), "rge_op_gemm_group")
    return outs


def set_option(name: str, value: int) -> None:
    """Run-time tuning knob of the kernels (rge_set_option): "attn_kernel", "attn_poly", "attn_split", "gemm_bn", "gemm2_bn", "2cta_min_m",
    "raster", "wide_store", "trim_last", "nvtx"."""
    check(_lib.load().rge_set_option(name.encode(), int(value)), f"rge_set_option({name})")


def attention_workspace(heads: int, device="cuda"):
    """Scratch that lets `attention` cut the ragged last query tile of every head along K/V (rge_attn_desc.workspace)."""
    n = _lib.load().rge_attention_workspace_bytes(heads)
    return torch.empty(n, dtype=torch.uint8, device=device)


def attention(q, k, v, heads: int, out=None, scale: float | None = None, workspace=None):
    """q [Sq, H*128], k/v [Skv, H*128] -> out [Sq, H*128]."""
    lib = _lib.load()
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, torch.bfloat16, n)
    if out is None:
        out = torch.empty(q.shape[0], heads * 128, dtype=torch.bfloat16, device=q.device)
    d = AttnDesc()
    d.Q, d.ldq = ptr(q), q.stride(0)
    d.K, d.ldk = ptr(k), k.stride(0)
    d.V, d.ldv = ptr(v), v.stride(0)
    d.O, d.ldo = ptr(out), out.stride(0)
    d.Sq, d.Skv, d.H = q.shape[0], k.shape[0], heads
    d.scale = scale if scale is not None else 128 ** -0.5
    if workspace is not None:
        d.workspace, d.workspace_bytes = ptr(workspace), workspace.numel() * workspace.element_size()
    check(lib.rge_op_attention(C.byref(d), stream_ptr()), "rge_op_attention")
    return out


def ln_modulate(x, scale, shift, out=None):
    lib = _lib.load()
    _req(x, torch.bfloat16, "x")
    if out is None:
        out = torch.empty_like(x)
    M, Dm = x.shape
    check(lib.rge_op_ln_modulate(ptr(x), x.stride(0), ptr(scale), ptr(shift), ptr(out), out.stride(0), M, Dm,
                                 stream_ptr()), "rge_op_ln_modulate")
    return out


def rmsnorm(x, weight, eps: float = 1e-6, out=None):
    lib = _lib.load()
    _req(x, torch.bfloat16, "x"); _req(weight, torch.bfloat16, "weight")
    if out is None:
        out = torch.empty_like(x)
    M, Dm = x.shape
    check(lib.rge_op_rmsnorm(ptr(x), x.stride(0), ptr(weight), ptr(out), out.stride(0), M, Dm, eps, stream_ptr()),
          "rge_op_rmsnorm")
    return out


def cfg_rescale(pos, neg, scale: float, out=None):
    """Norm-rescaled CFG on [M, channels] velocities."""
    lib = _lib.load()
    _req(pos, torch.bfloat16, "pos"); _req(neg, torch.bfloat16, "neg")
    pos = pos.contiguous(); neg = neg.contiguous()
    if out is None:
        out = torch.empty_like(pos)
    check(lib.rge_cfg_rescale(ptr(pos), ptr(neg), float(scale), ptr(out), pos.shape[0], pos.shape[1], stream_ptr()),
          "rge_cfg_rescale")
    return out


def cfg_diff_norm(pos, neg):
    """bf16 [M]: per-token norm of (pos - neg)."""
    lib = _lib.load()
    _req(pos, torch.bfloat16, "pos"); _req(neg, torch.bfloat16, "neg")
    pos = pos.contiguous(); neg = neg.contiguous()
    out = torch.empty(pos.shape[0], dtype=torch.bfloat16, device=pos.device)
    check(lib.rge_cfg_diff_norm(ptr(pos), ptr(neg), ptr(out), pos.shape[0], pos.shape[1], stream_ptr()),
          "rge_cfg_diff_norm")
    return out


def cfg_combine(pos, neg, scale: float, denom=None, out=None):
    """neg + scale * (pos - neg) [/ denom[m]] on [M, channels]."""
    lib = _lib.load()
    _req(pos, torch.bfloat16, "pos"); _req(neg, torch.bfloat16, "neg")
    pos = pos.contiguous(); neg = neg.contiguous()
    if denom is not None:
        denom = denom.reshape(-1).contiguous()
        _req(denom, torch.bfloat16, "denom")
    if out is None:
        out = torch.empty_like(pos)
    check(lib.rge_cfg_combine(ptr(pos), ptr(neg), float(scale), ptr(denom), ptr(out), pos.shape[0], pos.shape[1],
                              stream_ptr()), "rge_cfg_combine")
    return out


def rope_table(ids: torch.Tensor) -> torch.Tensor:
    """ids fp32 [S,3] -> fp32 [S,64,2] (cos, sin)."""
    lib = _lib.load()
    _req(ids, torch.float32, "ids")
    ids = ids.contiguous()
    cs = torch.empty(ids.shape[0], 64, 2, dtype=torch.float32, device=ids.device)
    check(lib.rge_op_rope_table(ptr(ids), ptr(cs), ids.shape[0], stream_ptr()), "rge_op_rope_table")
    return cs


def gather_rows(src, ids, out=None):
    lib = _lib.load()
    _req(src, torch.bfloat16, "src"); _req(ids, torch.int32, "ids")
    n, width = ids.numel(), src.shape[1]
    if out is None:
        out = torch.empty(n, width, dtype=torch.bfloat16, device=src.device)
    check(lib.rge_gather_rows(ptr(src), src.stride(0), ptr(ids), n, width, ptr(out), out.stride(0), stream_ptr()),
          "rge_gather_rows")
    return out


def scatter_rows(src, ids, dst):
    lib = _lib.load()
    _req(src, torch.bfloat16, "src"); _req(ids, torch.int32, "ids"); _req(dst, torch.bfloat16, "dst")
    n, width = ids.numel(), src.shape[1]
    check(lib.rge_scatter_rows(ptr(src), src.stride(0) if n else width, ptr(ids), n, width, ptr(dst), dst.stride(0),
                               stream_ptr()), "rge_scatter_rows")
    return dst


def pack_latents(latents):
    """[B, C, H, W] bf16 -> [B, (H/2)(W/2), 4C] (FluxKontextPipeline._pack_latents)."""
    lib = _lib.load()
    _req(latents, torch.bfloat16, "latents")
    latents = latents.contiguous()
    B, Cc, H, W = latents.shape
    out = torch.empty(B, (H // 2) * (W // 2), 4 * Cc, dtype=torch.bfloat16, device=latents.device)
    check(lib.rge_pack_latents(ptr(latents), ptr(out), B, Cc, H, W, stream_ptr()), "rge_pack_latents")
    return out


def unpack_latents(packed, height: int, width: int, vae_scale_factor: int = 8):
    """[B, L, 4C] bf16 -> [B, C, H, W] with H = 2 * (height // (2 * vae_scale_factor)), likewise W
    (FluxKontextPipeline._unpack_latents; `height` / `width` are in pixels)."""
    lib = _lib.load()
    _req(packed, torch.bfloat16, "packed")
    packed = packed.contiguous()
    B, L, ch = packed.shape
    H = 2 * (int(height) // (vae_scale_factor * 2))
    W = 2 * (int(width) // (vae_scale_factor * 2))
    if L != (H // 2) * (W // 2) or ch % 4:
        raise _lib.RegionEB200Error(f"unpack_latents: {L} tokens x {ch} channels do not match {height} x {width}")
    out = torch.empty(B, ch // 4, H, W, dtype=torch.bfloat16, device=packed.device)
    check(lib.rge_unpack_latents(ptr(packed), ptr(out), B, ch // 4, H, W, stream_ptr()), "rge_unpack_latents")
    return out


def euler(x, v, dt: float, dt_direct: float = 0.0, edited_mask=None, reuse_ratio: float | None = None, out=None):
    lib = _lib.load()
    _req(x, torch.bfloat16, "x"); _req(v, torch.bfloat16, "v")
    x = x.contiguous(); v = v.contiguous()
    if out is None:
        out = torch.empty_like(x)
    M, ch = x.shape
    check(lib.rge_euler(ptr(x), ptr(v), ptr(out), M, ch, dt, dt_direct, ptr(edited_mask),
                        0 if reuse_ratio is None else 1, 0.0 if reuse_ratio is None else reuse_ratio, stream_ptr()),
          "rge_euler")
    return out


def partition(x, v, cond, dt_final: float, threshold: float, want_sim: bool = False):
    """Raw (pre-morphology) edited mask uint8 [L] (+ fp32 similarity if requested)."""
    lib = _lib.load()
    for t, n in ((x, "x"), (v, "v"), (cond, "cond")):
        _req(t, torch.bfloat16, n)
    x = x.contiguous(); v = v.contiguous(); cond = cond.contiguous()
    L, ch = x.shape
    mask = torch.empty(L, dtype=torch.uint8, device=x.device)
    sim = torch.empty(L, dtype=torch.float32, device=x.device) if want_sim else None
    check(lib.rge_partition(ptr(x), ptr(v), ptr(cond), dt_final, threshold, ptr(mask), ptr(sim), L, ch,
                            stream_ptr()), "rge_partition")
    return (mask, sim) if want_sim else mask


_PINNED_COUNTS = {}


def compact(mask, grid_h: int, grid_w: int, erosion_dilation: bool, before_sync=None):
    """-> (final mask uint8 [L], edited_ids int32 [n_e], unedited_ids int32 [L-n_e]).

    One host synchronisation for the two counts (the reference syncs at the same point through boolean indexing,
    utils.py:347). It is the ONLY sync of an image, and it is taken late: the counts travel through an asynchronous
    copy into pinned host memory followed by an event; `before_sync(final_mask)` - the caller's chance to enqueue
    everything that needs the mask but not the counts (two-speed Euler update, mask all-gather) - runs first, and only
    then does the host wait on the event."""
    lib = _lib.load()
    _req(mask, torch.uint8, "mask")
    L = grid_h * grid_w
    dev = mask.device
    out_mask = torch.empty(L, dtype=torch.uint8, device=dev)
    edited = torch.empty(L, dtype=torch.int32, device=dev)
    unedited = torch.empty(L, dtype=torch.int32, device=dev)
    counts = torch.empty(2, dtype=torch.int32, device=dev)
    check(lib.rge_compact(ptr(mask), ptr(out_mask), grid_h, grid_w, 1 if erosion_dilation else 0, ptr(edited),
                          ptr(unedited), ptr(counts), stream_ptr()), "rge_compact")
    host = _PINNED_COUNTS.get(dev)
    if host is None:
        host = _PINNED_COUNTS[dev] = torch.empty(2, dtype=torch.int32).pin_memory()
    host.copy_(counts, non_blocking=True)
    done = torch.cuda.Event()
    done.record()
    if before_sync is not None:
        before_sync(out_mask)
    done.synchronize()
    n_e = int(host[0])
    return out_mask, edited[:n_e], unedited[: L - n_e]
