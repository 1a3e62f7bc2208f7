"""Short driver for ncu: a few transformer steps at BASELINE shapes (full depth by default) through the engine, so a
`ncu --set full` capture does not have to sit through whole images.

    python tools/profile_step.py [--blocks ND NS] [--full N] [--region N] [--edited 1100]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, nargs=2, default=[19, 38])
    ap.add_argument("--full", type=int, default=2)
    ap.add_argument("--region", type=int, default=2)
    ap.add_argument("--edited", type=int, default=1100)
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the LAST step of each kind (ncu --profile-from-start off)")
    args = ap.parse_args()
    from standins import synthetic as syn
    from regione_b200.engine import FluxEngine
    from regione_b200.schedule import latent_image_ids

    dev = "cuda"
    arch = dict(syn.FLUX_KONTEXT, n_double=args.blocks[0], n_single=args.blocks[1])
    G, T = 64, 512
    L = G * G
    pipe = syn.build_pipeline(arch, seed=110, device=dev)
    inp = syn.make_inputs(110, G, G, T, arch["ctx_dim"], arch["pooled_dim"], rho=0.25, device=dev)
    ids = torch.cat([latent_image_ids(G, G, 0.0, dev), latent_image_ids(G, G, 1.0, dev)])
    eng = FluxEngine(pipe.transformer, T, L, L)
    eng.begin_image(torch.zeros(T, 3, device=dev), ids, inp["prompt_embeds"][0], inp["pooled_prompt_embeds"][0], 2496.0)
    x_full = torch.cat([inp["latents"][0], inp["image_latents"][0]])
    edited = torch.randperm(L)[: args.edited].sort().values.to(dev).int()
    x_reg = inp["latents"][0][edited.long()]
    for name, n, fn in (("FULL", args.full, lambda: eng.step(x_full, None, 936.0, L)),
                        ("REGION", args.region, lambda: eng.step(x_reg, edited, 920.0, edited.numel()))):
        for i in range(n):
            torch.cuda.synchronize()
            ranged = args.profiler_range and i == n - 1
            if ranged:
                torch.cuda.profiler.start()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            if ranged:
                torch.cuda.profiler.stop()
            print(f"{name} step {i}: {(time.perf_counter() - t0) * 1e3:.2f} ms")
    eng.close()


if __name__ == "__main__":
    main()
