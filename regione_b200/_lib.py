"""ctypes binding of the C ABI declared in include/regione_b200.h.

There is no CPU fallback: if the shared library is missing and cannot be built, importing the hot path raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

c_void_p, c_int32, c_int64, c_float = C.c_void_p, C.c_int32, C.c_int64, C.c_float

RGE_OK = 0
EPI_STORE, EPI_GELU, EPI_GATE_RES, EPI_NORM_ROPE = 0, 1, 2, 3
GEMM_FP16_ROUNDTRIP = 1
ABI_VERSION = 5
BLK_GLOBAL, BLK_DOUBLE, BLK_SINGLE = 0, 1, 2

GLOBAL_SLOTS = [
    "X_EMBED_W", "X_EMBED_B", "CTX_EMBED_W", "CTX_EMBED_B",
    "TIME1_W", "TIME1_B", "TIME2_W", "TIME2_B",
    "GUID1_W", "GUID1_B", "GUID2_W", "GUID2_B",
    "POOL1_W", "POOL1_B", "POOL2_W", "POOL2_B",
    "NORM_OUT_W", "NORM_OUT_B", "PROJ_OUT_W", "PROJ_OUT_B",
]
DOUBLE_SLOTS = [
    "MOD_W", "MOD_B", "MOD_CTX_W", "MOD_CTX_B",
    "Q_W", "Q_B", "K_W", "K_B", "V_W", "V_B",
    "ADD_Q_W", "ADD_Q_B", "ADD_K_W", "ADD_K_B", "ADD_V_W", "ADD_V_B",
    "NORM_Q", "NORM_K", "NORM_ADD_Q", "NORM_ADD_K",
    "OUT_W", "OUT_B", "ADD_OUT_W", "ADD_OUT_B",
    "FF_UP_W", "FF_UP_B", "FF_DOWN_W", "FF_DOWN_B",
    "FFC_UP_W", "FFC_UP_B", "FFC_DOWN_W", "FFC_DOWN_B",
]
SINGLE_SLOTS = [
    "MOD_W", "MOD_B",
    "Q_W", "Q_B", "K_W", "K_B", "V_W", "V_B",
    "NORM_Q", "NORM_K",
    "MLP_W", "MLP_B", "OUT_W", "OUT_B",
]
G = {n: i for i, n in enumerate(GLOBAL_SLOTS)}
D = {n: i for i, n in enumerate(DOUBLE_SLOTS)}
S = {n: i for i, n in enumerate(SINGLE_SLOTS)}


class GemmDesc(C.Structure):
    _fields_ = [
        ("A", c_void_p), ("lda", c_int64),
        ("W", c_void_p), ("ldw", c_int64),
        ("bias", c_void_p),
        ("M", c_int32), ("N", c_int32), ("K", c_int32),
        ("epilogue", c_int32),
        ("out", c_void_p), ("ldo", c_int64),
        ("row_map", c_void_p), ("row_off", c_int32), ("col_off", c_int32),
        ("gate", c_void_p), ("res", c_void_p), ("ldr", c_int64),
        ("norm_w", c_void_p), ("rope_cs", c_void_p),
        ("rope_map", c_void_p), ("rope_off", c_int32),
        ("rope_ld", c_int64), ("flags", c_int32),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("Q", c_void_p), ("ldq", c_int64),
        ("K", c_void_p), ("ldk", c_int64),
        ("V", c_void_p), ("ldv", c_int64),
        ("O", c_void_p), ("ldo", c_int64),
        ("Sq", c_int32), ("Skv", c_int32), ("H", c_int32),
        ("scale", c_float),
        ("workspace", c_void_p), ("workspace_bytes", c_int64),
    ]


class Config(C.Structure):
    _fields_ = [(n, c_int32) for n in (
        "dim", "heads", "n_double", "n_single", "mlp_ratio", "in_channels", "ctx_dim", "pooled_dim",
        "txt_len", "lat_len", "cond_len", "guidance_embeds", "n_pass", "device", "external_embed", "shared_cache")]


# name -> (restype, argtypes); every name here must be declared in include/regione_b200.h (tests check both ways)
PROTOTYPES = {
    "rge_abi_version": (c_int32, []),
    "rge_last_error": (C.c_char_p, []),
    "rge_launch_count": (c_int64, []),
    "rge_set_option": (c_int32, [C.c_char_p, c_int32]),
    "rge_profile_enable": (c_int32, [c_int32]),
    "rge_profile_collect": (c_int32, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                     C.POINTER(c_int64)]),
    "rge_op_gemm": (c_int32, [C.POINTER(GemmDesc), c_void_p]),
    "rge_op_gemm_group": (c_int32, [C.POINTER(GemmDesc), c_int32, c_void_p]),
    "rge_op_attention": (c_int32, [C.POINTER(AttnDesc), c_void_p]),
    "rge_attention_workspace_bytes": (c_int64, [c_int32]),
    "rge_op_ln_modulate": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32,
                                     c_void_p]),
    "rge_op_rmsnorm": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_float, c_void_p]),
    "rge_cfg_rescale": (c_int32, [c_void_p, c_void_p, c_float, c_void_p, c_int32, c_int32, c_void_p]),
    "rge_cfg_diff_norm": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "rge_cfg_combine": (c_int32, [c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "rge_op_rope_table": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p]),
    "rge_gather_rows": (c_int32, [c_void_p, c_int64, c_void_p, c_int32, c_int32, c_void_p, c_int64, c_void_p]),
    "rge_scatter_rows": (c_int32, [c_void_p, c_int64, c_void_p, c_int32, c_int32, c_void_p, c_int64, c_void_p]),
    "rge_pack_latents": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "rge_unpack_latents": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p]),
    "rge_euler": (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_float, c_float, c_void_p, c_int32,
                            c_float, c_void_p]),
    "rge_partition": (c_int32, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_int32,
                                c_int32, c_void_p]),
    "rge_compact": (c_int32, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
                              c_void_p]),
    "rge_create": (c_int32, [C.POINTER(Config), C.POINTER(c_void_p)]),
    "rge_destroy": (c_int32, [c_void_p]),
    "rge_set_weight": (c_int32, [c_void_p, c_int32, c_int32, c_int32, c_void_p]),
    "rge_finalize_weights": (c_int32, [c_void_p]),
    "rge_begin_image": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p]),
    "rge_set_pass_text_len": (c_int32, [c_void_p, c_int32, c_int32]),
    "rge_begin_image_ex": (c_int32, [c_void_p, c_int32, c_void_p, c_void_p, c_void_p]),
    "rge_dit_step_ex": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_int32, c_void_p]),
    "rge_dit_step": (c_int32, [c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_int32, c_void_p, c_float, c_void_p,
                               c_int32, c_void_p]),
}


class RegionEB200Error(RuntimeError):
    pass


def library_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if the .so is absent and nvcc exists). Raises if the CUDA library is unavailable."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    # build_library() is a no-op when the library is newer than every source (mtime check) and compiles under a file
    # lock, so concurrent ranks do not write the same objects; without nvcc a prebuilt library is used as is
    try:
        _build.build_library()
    except Exception as e:  # noqa: BLE001
        if not os.path.exists(path):
            raise RegionEB200Error(
                f"regione_b200: CUDA library {path} is missing and could not be built ({e}); "
                "run `python -m regione_b200.build` — there is no CPU fallback") from e
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.rge_abi_version() != ABI_VERSION:
        raise RegionEB200Error(f"regione_b200: ABI version mismatch (library {lib.rge_abi_version()}, binding "
                               f"{ABI_VERSION}): rebuild with `python -m regione_b200.build --force`")
    _LIB = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != RGE_OK:
        msg = load().rge_last_error().decode("utf-8", "replace")
        raise RegionEB200Error(f"regione_b200 {what} failed ({status}): {msg}")


def ptr(t) -> int | None:
    """Raw device pointer of a torch tensor (None stays None)."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
