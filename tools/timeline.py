"""Per-launch timeline of one REGION and one FULL transformer step (GEMM / attention launches only) from the library's
own CUDA-event profiler (RGE_PROFILE_DUMP): stream-ready time, end time, shape. No nsys in this image.
    python tools/timeline.py [--blocks 2 3] [--edited 1064] > profiles/rNN_timeline.log"""
import argparse, ctypes as C, os, sys, tempfile
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--blocks", type=int, nargs=2, default=[2, 3])
ap.add_argument("--edited", type=int, default=1064)
args = ap.parse_args()
dump = tempfile.mktemp(suffix=".txt")
os.environ["RGE_PROFILE_DUMP"] = dump
from regione_b200 import _lib
from standins import synthetic as syn
from regione_b200.engine import FluxEngine
from regione_b200.schedule import latent_image_ids

dev = "cuda"
lib = _lib.load()
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
steps = {"FULL": lambda: eng.step(x_full, None, 936.0, L), "REGION": lambda: eng.step(x_reg, edited, 920.0, edited.numel())}
for name in ("FULL", "REGION", "FULL", "REGION"):
    steps[name]()
torch.cuda.synchronize()
buf = [(C.c_double * 2)(), (C.c_double * 2)(), (C.c_double * 2)(), (C.c_int64 * 2)()]
for name in ("REGION", "FULL"):
    open(dump, "w").close()
    lib.rge_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); steps[name](); e1.record(); torch.cuda.synchronize()
    lib.rge_profile_collect(*buf)
    lib.rge_profile_enable(0)
    rows = [l.split() for l in open(dump) if not l.startswith("#")]
    rows = sorted(((float(r[1]), float(r[2]), int(r[0]), r[3], r[4], r[5]) for r in rows))
    t_first = rows[0][0]
    print(f"== {name} step, blocks {args.blocks}: {e0.elapsed_time(e1):.3f} ms; gemm busy {buf[0][0]:.3f} ms, attention busy {buf[0][1]:.3f} ms")
    print("ready_us end_us dur_us kind M/Sq N/Skv K/H")
    for t0, t1, cls, m, n, k in rows:
        print(f"{(t0 - t_first) * 1e3:9.1f} {(t1 - t_first) * 1e3:9.1f} {(t1 - t0) * 1e3:8.1f} {'attn' if cls else 'gemm'} {m} {n} {k}")
eng.close()
os.unlink(dump)
