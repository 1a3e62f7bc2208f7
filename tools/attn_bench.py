"""Isolated attention throughput (sustained ~1.5 s) at the FULL and REGION shapes of the hot path.
RGE_ATTN_POLY=0|1|2 selects how many of every four exponentials run on the FMA pipe."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regione_b200 import ops  # noqa: E402

H = 24
for Sq, Skv in [(8704, 8704), (1576, 8704), (3000, 8704)]:
    q = torch.randn(Sq, H * 128, device="cuda").bfloat16()
    k = torch.randn(Skv, H * 128, device="cuda").bfloat16()
    v = torch.randn(Skv, H * 128, device="cuda").bfloat16()
    o = torch.empty_like(q)
    for _ in range(3):
        ops.attention(q, k, v, H, out=o)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < 1.5:
        for _ in range(10):
            ops.attention(q, k, v, H, out=o)
        n += 10
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"poly={os.environ.get('RGE_ATTN_POLY','1')} Sq={Sq} Skv={Skv}: {ms*1e3:.0f} us, {4.0*Sq*Skv*128*H/ms/1e9:.0f} TF/s sustained")
