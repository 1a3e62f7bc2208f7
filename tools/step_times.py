"""Per-step device times of one image (diagnostics): python tools/step_times.py"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from regione_b200 import RegionEHelper
from standins import synthetic as syn

pipe = syn.build_pipeline(syn.FLUX_KONTEXT, seed=110, device="cuda")
h = RegionEHelper(pipe)
h.set_params(warmup_step=6, post_step=2, refresh_step="16", threshold=0.88, cache_threshold=0.04, erosion_dilation=True)
h.enable()
inp = syn.make_inputs(110, 64, 64, 512, 4096, 768, rho=0.25, device="cuda")
kw = {k: v for k, v in inp.items() if k != "intended_mask"}
for it in range(3):
    pipe.regione_time_steps = it == 2
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pipe(guidance_scale=2.5, num_inference_steps=28, output_type="latent", return_dict=False, **kw)
    torch.cuda.synchronize(); print(f"image {it}: {(time.perf_counter()-t0)*1e3:.1f} ms wall")
tr = pipe.regione_trace
tot = {}
for m, ms in zip(tr["modes"], tr["step_ms"]):
    tot.setdefault(m, []).append(ms)
for m, v in tot.items():
    print(m, len(v), "steps, mean %.2f ms, sum %.1f ms" % (sum(v) / len(v), sum(v)), [round(x, 1) for x in v])
print("sum of steps %.1f ms" % sum(tr["step_ms"]))
