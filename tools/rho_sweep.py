"""Edited-fraction sweep of SURVEY §8d (rho in {5, 10, 25, 50, 100} %) at configs[1] shapes, for both cache thresholds
the reference ships (0.04 evaluation, 0.01 demo): images/s and mean FULL / REGION step times per setting.
    python tools/rho_sweep.py > profiles/rNN_rho_sweep.log"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regione_b200 import RegionEHelper
from standins import synthetic as syn

pipe = syn.build_pipeline(syn.FLUX_KONTEXT, seed=110, device="cuda")
print("rho_target cache_thr edited schedule img_per_s ms_image full_ms region_ms")
for delta in (0.04, 0.01):
    h = RegionEHelper(pipe)
    h.set_params(warmup_step=6, post_step=2, refresh_step="16", threshold=0.88, cache_threshold=delta,
                 erosion_dilation=True)
    h.enable()
    p = h.pipeline
    for rho in (0.05, 0.10, 0.25, 0.50, 1.0):
        inp = syn.make_inputs(110, 64, 64, 512, 4096, 768, rho=rho, device="cuda")
        kw = {k: v for k, v in inp.items() if k != "intended_mask"}
        best = 1e9
        for it in range(3):
            p.regione_time_steps = it == 2
            torch.cuda.synchronize(); t0 = time.perf_counter()
            p(guidance_scale=2.5, num_inference_steps=28, output_type="latent", return_dict=False, **kw)
            torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0) if it else best
        tr = p.regione_trace
        per = {}
        for m, ms in zip(tr["modes"], tr["step_ms"]):
            per.setdefault(m, []).append(ms)
        mean = lambda m: sum(per.get(m, [0])) / max(len(per.get(m, [])), 1)
        print(f"{rho:.2f} {delta:.2f} {tr['edited_ids'].numel()} {''.join(m[0] for m in tr['modes'])} "
              f"{1 / best:.4f} {best * 1e3:.1f} {mean('FULL'):.2f} {mean('REGION'):.2f}", flush=True)
    h.disable()
