"""HBM-bound kernels of the path (partition, Euler / AVDC reuse, two-speed Euler, row gather / scatter, LayerNorm +
modulation, CFG combine, adaLN GEMV batch) timed alone with CUDA events at the BASELINE size (L = 4096 tokens, one
image: launch-latency bound) and at a batch of `--images` images' worth of rows (bandwidth bound), against the measured
HBM peak. Algorithmic bytes per row as in DESIGN.md §3.
    python tools/hbm_ops_bench.py [--images 256] [--once]      # --once: one launch per case (for ncu)"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regione_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=256)
ap.add_argument("--once", action="store_true")
args = ap.parse_args()
peak = 6553.3
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p))["hbm_gbs"]
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2


def timeit(fn, reps):
    if args.once:
        fn(); torch.cuda.synchronize(); return float("nan")
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(reps):
        flush.max()                                              # evict L2 with a READ pass (clean lines: a memset
                                                                 # would leave 126 MB of dirty lines to write back
                                                                 # under the timed kernel)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


print(f"# peak {peak} GB/s (MEASURED_PEAKS.json hbm_gbs); L2 flushed (read pass over 256 MB) between launches")
print("kernel rows bytes_per_launch us GB/s frac_of_peak")
for rows in (4096, 4096 * args.images):
    ch = 64
    x = torch.randn(rows, ch, device=dev, generator=g).bfloat16()
    v = torch.randn(rows, ch, device=dev, generator=g).bfloat16()
    c = torch.randn(rows, ch, device=dev, generator=g).bfloat16()
    mask = (torch.rand(rows, device=dev, generator=g) < 0.25).to(torch.uint8)
    ids = torch.nonzero(mask).flatten().int()
    n = ids.numel()
    sub = torch.randn(n, ch, device=dev, generator=g).bfloat16()
    dst = torch.empty_like(x); out = torch.empty_like(x); gout = torch.empty(n, ch, device=dev, dtype=torch.bfloat16)
    cases = [
        ("arp_similarity_kernel", rows, rows * (ch * 2 * 3 + 1), lambda: ops.partition(x, v, c, -0.9, 0.88)),
        ("euler_kernel", rows, rows * ch * 2 * 3, lambda: ops.euler(x, v, -0.03, out=out)),
        ("euler_kernel(reuse)", rows, rows * ch * 2 * 3, lambda: ops.euler(x, v, -0.03, reuse_ratio=0.98, out=out)),
        ("euler_kernel(two-speed)", rows, rows * (ch * 2 * 3 + 1),
         lambda: ops.euler(x, v, -0.03, -0.4, edited_mask=mask, out=out)),
        ("move_rows_kernel<gather>", n, n * (ch * 2 * 2 + 4), lambda: ops.gather_rows(x, ids, out=gout)),
        ("move_rows_kernel<scatter>", n, n * (ch * 2 * 2 + 4), lambda: ops.scatter_rows(sub, ids, dst)),
        ("cfg_rescale_kernel", rows, rows * ch * 2 * 3, lambda: ops.cfg_rescale(x, v, 4.0, out=out)),
    ]
    for name, r, nbytes, fn in cases:
        us = timeit(fn, 20) * 1e3
        print(f"{name} {r} {nbytes} {us:.2f} {nbytes / us / 1e3:.1f} {nbytes / us / 1e3 / peak:.3f}", flush=True)
# LayerNorm + modulation at the FULL-step shape and the adaLN GEMV batch are timed inside the engine (bench.py); here
# the LN kernel alone at [8704, 3072] and [8 x 8704, 3072]
for rows in (8704, 8704 * 8):
    D = 3072
    x = torch.randn(rows, D, device=dev, generator=g).bfloat16()
    sc = torch.randn(D, device=dev, generator=g).bfloat16(); sh = torch.randn(D, device=dev, generator=g).bfloat16()
    out = torch.empty_like(x)
    us = timeit(lambda: ops.ln_modulate(x, sc, sh, out=out), 20) * 1e3
    nbytes = rows * D * 2 * 2
    print(f"ln_modulate_kernel {rows} {nbytes} {us:.2f} {nbytes / us / 1e3:.1f} {nbytes / us / 1e3 / peak:.3f}", flush=True)
