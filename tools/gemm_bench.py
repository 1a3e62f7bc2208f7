"""Isolated GEMM throughput: this library's kernels vs torch.matmul (cuBLAS) on the shapes of the hot path, burst (best
of a few short runs) and sustained (back to back for ~1.2 s, under the power cap); every epilogue on the two shapes
where it runs; and the scatter-GEMM against the reference's own Triton `_partially_linear` (SURVEY §2.2 K-a: "that JIT
output is the bar to beat on this box") at its call-site shapes. Usage: python tools/gemm_bench.py [--quick]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regione_b200 import _lib, ops  # noqa: E402

SHAPES = [(8704, 3072, 3072), (8192, 3072, 3072), (4096, 3072, 15360), (8704, 12288, 3072), (8704, 3072, 15360), (8704, 3072, 12288), (8192, 3072, 12288),
          (1576, 3072, 3072), (1576, 12288, 3072), (1576, 3072, 15360), (1064, 3072, 3072), (1064, 3072, 12288),
          (512, 3072, 3072), (512, 12288, 3072)]


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
    quick = "--quick" in sys.argv
    secs = 0.4 if quick else 1.2
    for (M, N, K) in SHAPES:
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        b = torch.randn(N, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        res_t = torch.randn(M, N, device="cuda").bfloat16()
        gate = torch.randn(N, device="cuda").bfloat16()
        fl = 2.0 * M * N * K
        fns = {"ours": lambda: ops.gemm(a, w, b, out=out),
               "ours_gelu": lambda: ops.gemm(a, w, b, out=out, epilogue=_lib.EPI_GELU),
               "ours_gate_res": lambda: ops.gemm(a, w, b, out=res_t, epilogue=_lib.EPI_GATE_RES, gate=gate, res=res_t),
               "cublas": lambda: torch.matmul(a, w.t(), out=out)}
        if N % 128 == 0 and N <= 3072:
            S = 8704
            nw = torch.ones(128, device="cuda").bfloat16()
            cs = torch.randn(S, 64, 2, device="cuda")
            cs_pm = cs.permute(1, 0, 2).contiguous()
            pos = torch.arange(M, device="cuda", dtype=torch.int32)
            fns["ours_norm_rope(row-major table)"] = lambda: ops.gemm(
                a, w, b, out=out, epilogue=_lib.EPI_NORM_ROPE, norm_w=nw, rope_cs=cs, rope_map=pos)
            fns["ours_norm_rope(pair-major table)"] = lambda: ops.gemm(
                a, w, b, out=out, epilogue=_lib.EPI_NORM_ROPE, norm_w=nw, rope_cs=cs_pm, rope_map=pos, rope_ld=S)
        if M >= 2048:                        # the CTA-pair kernel with its earlier fixed 256-wide tile
            def fixed256():
                ops.set_option("gemm2_bn", 256)
                ops.gemm(a, w, b, out=out)
                ops.set_option("gemm2_bn", 0)
            fns["ours(256-wide tiles)"] = fixed256
        if M < 2048 and N % 256 == 0:        # both kernels where the per-shape rule decides between them
            def pair():
                ops.set_option("2cta_min_m", 1)
                ops.gemm(a, w, b, out=out)
                ops.set_option("2cta_min_m", -1)
            fns["ours(cta-pair kernel)"] = pair

            def single():
                ops.set_option("2cta_min_m", 0)
                ops.gemm(a, w, b, out=out)
                ops.set_option("2cta_min_m", -1)
            fns["ours(1-cta kernel)"] = single
        res = {}
        for name, fn in fns.items():
            for _ in range(3):
                fn()
            burst = min(timeit(fn, 0.02) for _ in range(3))
            sustained = timeit(fn, secs)
            res[name] = (fl / burst / 1e9, fl / sustained / 1e9)
            time.sleep(0.2)
        print(f"M={M} N={N} K={K}: " + "  |  ".join(
            f"{k} {v[0]:.0f}/{v[1]:.0f}" for k, v in res.items()) + "   (burst/sustained TF/s)", flush=True)
    # ---- the q / k / v (+ text q / k / v) projections of one double block as ONE grouped launch vs six launches
    D, S = 3072, 8704
    nw = torch.ones(128, device="cuda").bfloat16()
    cs_pm = torch.randn(64, S, 2, device="cuda")
    for M_img, T in ((8192, 512), (1064, 512), (360, 512), (4608, 256)):
        xi = torch.randn(M_img, D, device="cuda").bfloat16()
        xt = torch.randn(T, D, device="cuda").bfloat16()
        ws = [(torch.randn(D, D, device="cuda") * 0.02).bfloat16() for _ in range(6)]
        bs = [torch.randn(D, device="cuda").bfloat16() for _ in range(6)]
        q = torch.empty(T + M_img, D, device="cuda", dtype=torch.bfloat16)
        kc = torch.empty(S, D, device="cuda", dtype=torch.bfloat16)
        vc = torch.empty(S, D, device="cuda", dtype=torch.bfloat16)
        pos = torch.arange(M_img, device="cuda", dtype=torch.int32)
        nr = dict(epilogue=_lib.EPI_NORM_ROPE, norm_w=nw, rope_cs=cs_pm, rope_ld=S)
        members = [(xi, ws[0], bs[0], dict(out=q, row_off=T, rope_map=pos, rope_off=T, **nr)),
                   (xi, ws[1], bs[1], dict(out=kc, row_map=pos, row_off=T, rope_map=pos, rope_off=T, **nr)),
                   (xi, ws[2], bs[2], dict(out=vc, row_map=pos, row_off=T)),
                   (xt, ws[3], bs[3], dict(out=q, **nr)), (xt, ws[4], bs[4], dict(out=kc, **nr)),
                   (xt, ws[5], bs[5], dict(out=vc))]
        fl = 2.0 * (M_img + T) * D * D * 3

        def six():
            for a_, w_, b_, kw in members:
                ops.gemm(a_, w_, b_, **kw)

        r = {}
        for name, fn in (("six launches (one stream)", six), ("grouped kernel", lambda: ops.gemm_group(members))):
            for _ in range(3):
                fn()
            r[name] = timeit(fn, secs)
        print(f"q/k/v + text q/k/v, image rows {M_img}, text rows {T}: " + "  |  ".join(
            f"{k} {v * 1e3:.1f} us {fl / v / 1e9:.0f} TF/s" for k, v in r.items()), flush=True)
    # ---- scatter-GEMM vs the reference's Triton kernel at its call sites (inplace.py:734-747)
    try:
        from oracle.build_ref import load_partially_linear
        pl = load_partially_linear()
    except Exception as e:  # noqa: BLE001
        pl, err = None, e
    if pl is None:
        print("reference Triton kernel not staged (oracle/_ref): skipped")
        return
    for M, S in ((1064, 8192), (1576, 8704), (360, 8192), (872, 8704)):
        D = 3072
        a = torch.randn(1, M, D, device="cuda").bfloat16()
        w = (0.02 * torch.randn(D, D, device="cuda")).bfloat16()
        b = (0.01 * torch.randn(D, device="cuda")).bfloat16()
        idx = torch.randperm(S, device="cuda")[:M].sort().values
        idx32 = idx.int()
        cache = torch.zeros(1, S, D, device="cuda", dtype=torch.bfloat16)
        fl = 2.0 * M * D * D
        r = {}
        for name, fn in (("triton _partially_linear", lambda: pl(a, w, b, idx, cache)),
                         ("ours (row_map epilogue)", lambda: ops.gemm(a[0], w, b, out=cache[0], row_map=idx32)),
                         ("cublas + index_put", lambda: cache[0].index_copy_(0, idx, torch.addmm(b, a[0], w.t())))):
            for _ in range(3):
                fn()
            r[name] = timeit(fn, secs)
        print(f"scatter-GEMM M={M} N=K=3072 into [{S}, 3072]: " + "  |  ".join(
            f"{k} {v * 1e3:.1f} us {fl / v / 1e9:.0f} TF/s" for k, v in r.items()), flush=True)


if __name__ == "__main__":
    main()
