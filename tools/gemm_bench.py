"""Isolated GEMM throughput: this library's kernel vs torch.matmul (cuBLAS) on the shapes of the hot path, burst (best
of a few launches) and sustained (back-to-back for ~1.5 s, under the power cap). Usage: python tools/gemm_bench.py
(set RGE_2CTA_MIN_M=0 to force the 1-CTA kernel)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regione_b200 import _lib, ops  # noqa: E402

SHAPES = [(8704, 3072, 3072), (8704, 12288, 3072), (8704, 3072, 15360), (8192, 3072, 12288), (1576, 3072, 3072),
          (1576, 12288, 3072), (512, 3072, 3072)]


def timeit(fn, seconds):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.perf_counter()
    e0.record()
    while True:
        for _ in range(10):
            fn()
        n += 10
        if time.perf_counter() - t0 > seconds:
            break
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    mode = os.environ.get("RGE_2CTA_MIN_M", "2048")
    for (M, N, K) in SHAPES:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        fl = 2.0 * M * N * K
        ours = lambda: ops.gemm(a, w, b, out=out)                     # noqa: E731
        gelu = lambda: ops.gemm(a, w, b, out=out, epilogue=_lib.EPI_GELU)   # noqa: E731
        cublas = lambda: torch.matmul(a, w.t(), out=out)              # noqa: E731
        for _ in range(3):
            ours(); cublas()
        res = {}
        for name, fn in (("ours", ours), ("ours_gelu", gelu), ("cublas", cublas)):
            burst = min(timeit(fn, 0.02) for _ in range(3))
            sustained = timeit(fn, 1.5)
            res[name] = (fl / burst / 1e9, fl / sustained / 1e9)
            time.sleep(0.5)
        print(f"2cta_min_m={mode} M={M} N={N} K={K}: " + "  ".join(
            f"{k} burst {v[0]:.0f} sustained {v[1]:.0f} TF/s" for k, v in res.items()))


if __name__ == "__main__":
    main()
