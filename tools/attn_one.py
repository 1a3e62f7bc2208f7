"""One attention launch per shape (for `ncu -k regex:attention`): python tools/attn_one.py [Sq Skv] [kernel]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regione_b200 import ops  # noqa: E402
Sq = int(sys.argv[1]) if len(sys.argv) > 1 else 8704
Skv = int(sys.argv[2]) if len(sys.argv) > 2 else 8704
if len(sys.argv) > 3:
    ops.set_option("attn_kernel", int(sys.argv[3]))
H = 24
q, k, v = (torch.randn(n, H * 128, device="cuda").bfloat16() for n in (Sq, Skv, Skv))
o = torch.empty_like(q)
for _ in range(3):
    ops.attention(q, k, v, H, out=o)
torch.cuda.synchronize()
