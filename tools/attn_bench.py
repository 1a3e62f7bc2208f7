"""Isolated attention throughput (sustained ~1.2 s per entry, back to back under the power cap) at the FULL and REGION
shapes of the hot path: this library's kernel for every exponential-offload variant (attn_poly = 0 / 2 / 3 / 4 of 8
pairs on the FMA pipe) next to the reference's own attention call `flash_attn_func` (inplace.py:796-801; flash-attn 2.8,
FA2 kernels compiled for sm_100) and cuDNN's fused SDPA through torch. Usage: python tools/attn_bench.py [--quick]"""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regione_b200 import ops  # noqa: E402

H = 24


def sustained(fn, seconds):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(5):
            fn()
        n += 5
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    secs = 0.5 if "--quick" in sys.argv else 1.2
    shapes = [(8704, 8704), (1576, 8704), (872, 8704), (4864, 4864)]
    for Sq, Skv in shapes:
        q = torch.randn(Sq, H * 128, device="cuda").bfloat16()
        k = torch.randn(Skv, H * 128, device="cuda").bfloat16()
        v = torch.randn(Skv, H * 128, device="cuda").bfloat16()
        o = torch.empty_like(q)
        fl = 4.0 * Sq * Skv * 128 * H
        res = {}
        for poly in (0, 2, 3, 4):
            ops.set_option("attn_poly", poly)
            res[f"ours poly={poly}"] = sustained(lambda: ops.attention(q, k, v, H, out=o), secs)
        ops.set_option("attn_poly", -1)
        ops.set_option("attn_kernel", 1)
        res["ours kernel=attention64"] = sustained(lambda: ops.attention(q, k, v, H, out=o), secs)
        ops.set_option("attn_kernel", -1)
        res["ours default"] = sustained(lambda: ops.attention(q, k, v, H, out=o), secs)
        ws = ops.attention_workspace(H)     # lets the launcher cut a ragged last query tile along K/V (REGION shapes)
        res["ours default + workspace"] = sustained(lambda: ops.attention(q, k, v, H, out=o, workspace=ws), secs)
        try:
            from flash_attn import flash_attn_func
            q4, k4, v4 = q.view(1, Sq, H, 128), k.view(1, Skv, H, 128), v.view(1, Skv, H, 128)
            res["flash_attn_func"] = sustained(lambda: flash_attn_func(q4, k4, v4, causal=False), secs)
        except Exception as e:  # noqa: BLE001
            print("flash_attn unavailable:", e)
        try:
            from torch.nn.attention import SDPBackend, sdpa_kernel
            qh, kh, vh = (t.view(1, -1, H, 128).transpose(1, 2) for t in (q, k, v))
            with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
                res["cudnn sdpa"] = sustained(lambda: F.scaled_dot_product_attention(qh, kh, vh), secs)
        except Exception as e:  # noqa: BLE001
            print("cuDNN SDPA unavailable:", str(e)[:200])
        print(f"Sq={Sq} Skv={Skv} H={H}: " + "  |  ".join(
            f"{name} {ms * 1e3:.0f} us {fl / ms / 1e9:.0f} TF/s" for name, ms in res.items()), flush=True)


if __name__ == "__main__":
    main()
