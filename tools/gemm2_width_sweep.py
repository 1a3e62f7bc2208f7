"""Tile-width sweep of the CTA-pair GEMM (gemm2.cu): sustained TFLOP/s per forced width (`gemm2_bn`) on the FULL- and
REGION-step shapes, next to the host's own choice (pick_bn2), the 1-CTA kernel and cuBLAS. The table is what the cost
model in pick_bn2 is fitted to. Usage: python tools/gemm2_width_sweep.py [--quick]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regione_b200 import _lib, ops  # noqa: E402

SHAPES = [(8704, 3072, 3072), (8192, 3072, 3072), (8704, 12288, 3072), (8192, 12288, 3072), (8704, 3072, 15360),
          (8192, 3072, 12288), (4096, 3072, 15360), (4096, 12288, 3072), (2560, 3072, 3072),
          (1576, 3072, 3072), (1576, 12288, 3072), (1576, 3072, 15360), (1064, 3072, 3072), (1064, 12288, 3072),
          (1064, 3072, 12288), (872, 3072, 3072), (872, 3072, 15360), (512, 3072, 3072), (512, 12288, 3072)]
WIDTHS = [256, 240, 224, 208, 192, 176, 160, 144, 128, 96, 64]


def sustained(fn, seconds):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(10):
            fn()
        n += 10
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    secs = 0.15 if "--quick" in sys.argv else 0.4
    for (M, N, K) in SHAPES:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        fl = 2.0 * M * N * K
        res = {}
        ops.set_option("2cta_min_m", 1)          # force the CTA-pair kernel at every M
        for bn in WIDTHS:
            ops.set_option("gemm2_bn", bn)
            res[f"{bn}"] = sustained(lambda: ops.gemm(a, w, b, out=out), secs)
        ops.set_option("gemm2_bn", 0)
        res["auto"] = sustained(lambda: ops.gemm(a, w, b, out=out), secs)
        ops.set_option("2cta_min_m", 0)          # 1-CTA kernel
        res["1cta"] = sustained(lambda: ops.gemm(a, w, b, out=out), secs)
        ops.set_option("2cta_min_m", -1)
        res["default"] = sustained(lambda: ops.gemm(a, w, b, out=out), secs)
        res["cublas"] = sustained(lambda: torch.matmul(a, w.t(), out=out), secs)
        print(f"M={M} N={N} K={K}: " + "  ".join(f"{k}:{fl / v / 1e9:.0f}" for k, v in res.items()), flush=True)


if __name__ == "__main__":
    main()
