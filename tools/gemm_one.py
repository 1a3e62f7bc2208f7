"""A few launches of ONE GEMM (for `ncu -k regex:gemm`): python tools/gemm_one.py M N K [store|gelu|gate_res|norm_rope]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regione_b200 import _lib, ops  # noqa: E402
M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (8704, 3072, 3072)
epi = sys.argv[4] if len(sys.argv) > 4 else "store"
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
b = torch.randn(N, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
kw = {}
if epi == "gelu":
    kw = dict(epilogue=_lib.EPI_GELU)
elif epi == "gate_res":
    kw = dict(epilogue=_lib.EPI_GATE_RES, gate=torch.randn(N, device="cuda").bfloat16(), res=out)
elif epi == "norm_rope":
    S = 8704
    kw = dict(epilogue=_lib.EPI_NORM_ROPE, norm_w=torch.ones(128, device="cuda").bfloat16(),
              rope_cs=torch.randn(64, S, 2, device="cuda"), rope_map=torch.arange(M, device="cuda", dtype=torch.int32),
              rope_ld=S)
for _ in range(3):
    ops.gemm(a, w, b, out=out, **kw)
torch.cuda.synchronize()
