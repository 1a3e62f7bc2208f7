#!/usr/bin/env python
"""Benchmark of the RegionE hot path on B200: images/sec @1024^2, 28 steps, FLUX.1-Kontext shapes + RegionE.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (CUDA library through RegionEHelper)
    python bench.py --impl reference ...                          # the reference's algorithm on the host CPU cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...         # one rank per GPU, one image stream per rank

A "step" is one image: the full 28-step denoise (latents in -> latents out) with the default RegionE schedule
(9 FULL + 5 REGION + 14 SKIP transformer steps at cache_threshold 0.04). Synthetic weights / inputs (no checkpoints
or datasets offline); text encoders and VAE are outside the path. Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec @1024^2, 28 steps, FLUX.1-Kontext+RegionE"
PARAMS = dict(warmup_step=6, post_step=2, refresh_step="16", threshold=0.88, cache_threshold=0.04,
              erosion_dilation=True)
GRID, TXT = 64, 512            # 1024x1024 -> 64x64 latent tokens; max_sequence_length 512 (inplace.py:108)
FULL_STEPS = 9
# What the default workload (seed 110, rho 0.25) yields, bit-exact against the oracle at full size and depth in
# tests/test_flux_fullimage_gpu.py: the CPU reference arm, which cannot run the partition itself, times its REGION
# sample on this many edited tokens, so that both arms describe the same configuration.
EDITED_TOKENS = 1064
REGION_STEPS = 5
WORKLOAD = ("configs[1]: FLUX.1-Kontext-dev shapes 1024x1024 (T=512, L=C=4096, D=3072, 19+38 blocks), 28 steps, "
            "warmup_step=6 post_step=2 refresh=16 threshold=0.88 cache_threshold=0.04, bf16, 1 image stream per GPU")


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant GEMM shape, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json, written by hand from profiles/rNN_prof_gemm2_summary.csv).
    None if no capture is committed."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        return json.load(open(p))
    return None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                r = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                f = [x.strip() for x in r.stdout.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = sorted(int(float(s[0])) for s in self.samples if s[0].replace(".", "").isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": int(float(self.samples[0][1])) if self.samples else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU reference
def cpu_reference_sample(n_edited: int, region_steps: int, repeats: int = 1):
    """The reference's algorithm (oracle restatement, torch CPU, all host threads) on a bounded sample of the
    workload: ONE FULL step (S = 8704) and ONE REGION step (n_edited tokens) of a depth-reduced FLUX (1 double + 1
    single block, full width), scaled by 19 / 38 blocks and by the step schedule. Returns (images/sec, seconds
    measured, description)."""
    from oracle.flux import FluxOracle
    from oracle import region_ops as ro
    from standins import synthetic as syn
    from regione_b200.schedule import latent_image_ids

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    L = C_ = GRID * GRID
    arch = dict(syn.FLUX_KONTEXT, n_double=1, n_single=1)
    pipe = syn.build_pipeline(arch, seed=110, device="cpu")
    w = {k: v.detach() for k, v in pipe.transformer.state_dict().items()}
    inp = syn.make_inputs(110, GRID, GRID, TXT, arch["ctx_dim"], arch["pooled_dim"], rho=0.25)
    ids = torch.cat([latent_image_ids(GRID, GRID, 0.0), latent_image_ids(GRID, GRID, 1.0)])
    txt_ids = torch.zeros(TXT, 3)
    st = ro.RegionState()
    st.set_parameters(dict(num_inference_steps=28, **PARAMS))
    st.refresh(inp["latents"], inp["image_latents"], ids, txt_ids, 1024, 1024)
    model = FluxOracle(w, arch["heads"], 1, 1, True)
    guidance = torch.full([1], 2.5)
    t_bf16 = torch.tensor([936.0]).bfloat16() / 1000
    ed = torch.arange(n_edited).unsqueeze(0) * (L // max(n_edited, 1))
    st.edited_ids = ed.clamp_max(L - 1)

    def one_pair():
        st.current_step = st.warmup_step - 1          # FULL + cache write
        x_full = torch.cat([inp["latents"], inp["image_latents"]], dim=1)
        t0 = time.perf_counter()
        model.forward(st, x_full, inp["prompt_embeds"], inp["pooled_prompt_embeds"], t_bf16, ids, txt_ids, guidance)
        t1 = time.perf_counter()
        st.current_step = st.warmup_step              # REGION
        x_reg = ro.gather_rows(inp["latents"], st.edited_ids)
        rid = ro.gather_rows(ids.unsqueeze(0), st.edited_ids).squeeze(0)
        model.forward(st, x_reg, inp["prompt_embeds"], inp["pooled_prompt_embeds"], t_bf16, rid, txt_ids, guidance)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1

    best_full, best_reg, total = 1e30, 1e30, 0.0
    with torch.no_grad():
        for _ in range(repeats):
            f, r = one_pair()
            total += f + r
            best_full, best_reg = min(best_full, f), min(best_reg, r)
    # 1 double + 1 single measured together; blocks are 19 + 38 = 57 -> scale the pair by the mean block count
    # weighted by their FLOPs (both block types cost 12 D^2 MAC per token + the same attention): x 28.5
    scale = (19 + 38) / 2.0
    t_image = FULL_STEPS * best_full * scale + region_steps * best_reg * scale
    desc = (f"oracle (torch CPU, {threads} threads): 1 FULL step (S=8704) + 1 REGION step (N_e={n_edited}) of a "
            f"1-double+1-single-block FLUX at full width, x28.5 blocks, x({FULL_STEPS} FULL + {region_steps} REGION) "
            f"steps; extrapolated")
    return 1.0 / t_image, total, desc, threads


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist
    from regione_b200 import RegionEHelper, _lib
    from regione_b200 import flux_kontext as fk
    from standins import synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        fk.enable_mask_allgather(True)
    lib = _lib.load()
    arch = syn.TINY if args.arch == "tiny" else syn.FLUX_KONTEXT
    grid = 16 if args.arch == "tiny" else GRID
    txt = 32 if args.arch == "tiny" else TXT
    pipe = syn.build_pipeline(arch, seed=110, device=dev)
    helper = RegionEHelper(pipe)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):   # the helper prints its parameters (RegionE.py:51); keep stdout = JSON
        helper.set_params(**PARAMS)
    helper.enable()
    pipe = helper.pipeline
    host = syn.make_inputs(110 + rank, grid, grid, txt, arch["ctx_dim"], arch["pooled_dim"], rho=args.rho)
    keys = ("latents", "image_latents", "prompt_embeds", "pooled_prompt_embeds")
    pinned = {k: host[k].pin_memory() for k in keys}
    resident = {k: pinned[k].to(dev) for k in keys}
    hw = dict(height=host["height"], width=host["width"])
    out_host = torch.empty(1, grid * grid, 64, dtype=torch.bfloat16).pin_memory()

    def image_resident():
        return pipe(guidance_scale=2.5, num_inference_steps=28, output_type="latent", return_dict=False,
                    **resident, **hw)[0]

    def image_e2e():
        dev_in = {k: pinned[k].to(dev, non_blocking=True) for k in keys}
        out = pipe(guidance_scale=2.5, num_inference_steps=28, output_type="latent", return_dict=False,
                   **dev_in, **hw)[0]
        out_host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        per_rank = [float(ms)]
        if world > 1:
            every = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(every, ms)
            per_rank = [float(x) for x in every]
        return max(per_rank), per_rank

    for _ in range(args.warmup):
        image_resident()
    torch.cuda.synchronize()
    modes = list(pipe.regione_trace["modes"])
    n_edited = int(pipe.regione_trace["edited_ids"].numel())
    region_steps = modes.count("REGION")

    sampler = ClockSampler(local)
    if rank == 0:            # one nvidia-smi poller per box is enough (8 of them at 5 Hz perturb the host)
        sampler.start()
    lib.rge_profile_enable(1)
    launches0 = lib.rge_launch_count()
    if args.profiler_range:           # ncu --profile-from-start off: capture exactly the timed images
        torch.cuda.profiler.start()
    ms, ms_ranks = timed(image_resident, args.steps)
    if args.profiler_range:
        torch.cuda.profiler.stop()
    launches = lib.rge_launch_count() - launches0
    pms, psum, pwork, pcnt = (C.c_double * 2)(), (C.c_double * 2)(), (C.c_double * 2)(), (C.c_int64 * 2)()
    lib.rge_profile_collect(pms, psum, pwork, pcnt)
    lib.rge_profile_enable(0)
    ms_e2e, ms_e2e_ranks = timed(image_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- multi-GPU: every rank's schedule and edited-token count, and a replica check (untimed): all ranks denoise
    # rank 0's image (seed 110) once more and must produce bit-identical latents and the same region partition
    replica = None
    if world > 1:
        mine = {"rank": rank, "schedule": "".join(m[0] for m in modes), "edited_tokens": n_edited}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        # the NCCL mask all-gather of the last image (side stream, step warmup-1) must describe the same partitions
        from_masks = fk.MANAGER.batch_edited_counts()
        assert from_masks == [g_["edited_tokens"] for g_ in gathered], \
            f"gathered partition masks {from_masks} disagree with the ranks' own counts {gathered}"
        host0 = syn.make_inputs(110, grid, grid, txt, arch["ctx_dim"], arch["pooled_dim"], rho=args.rho)
        out0 = pipe(guidance_scale=2.5, num_inference_steps=28, output_type="latent", return_dict=False,
                    **{k: host0[k].to(dev) for k in keys}, **hw)[0]
        torch.cuda.synchronize()
        digest = torch.stack([out0.view(torch.int16).to(torch.int64).sum(),
                              (out0.view(torch.int16).to(torch.int64) * torch.arange(
                                  1, out0.numel() + 1, device=dev).view_as(out0)).sum(),
                              torch.tensor(int(pipe.regione_trace["edited_ids"].numel()), device=dev)])
        all_digests = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(all_digests, digest)
        identical = all(torch.equal(d, all_digests[0]) for d in all_digests)
        replica = {"per_rank": gathered, "edited_tokens_from_gathered_masks": from_masks,
                   "seed_110_edited_tokens": int(all_digests[0][2]),
                   "seed_110_bit_identical_across_ranks": bool(identical),
                   "same_schedule_on_all_ranks": len({g["schedule"] for g in gathered}) == 1}
        assert identical, "replicas disagree on the same input: " + str([d.tolist() for d in all_digests])
        assert replica["same_schedule_on_all_ranks"], "ranks ran different step schedules: " + str(gathered)

    value = world * args.steps / (ms / 1e3)
    e2e_value = world * args.steps / (ms_e2e / 1e3)
    peaks = load_peaks()
    gemm_tf = pwork[0] / (pms[0] * 1e-3) / 1e12 if pms[0] > 0 else None
    attn_tf = pwork[1] / (pms[1] * 1e-3) / 1e12 if pms[1] > 0 else None
    h2d = sum(pinned[k].numel() * pinned[k].element_size() for k in keys)
    line = {
        "metric": METRIC, "value": round(value, 4), "unit": "images/sec", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, "".join(m[0] for m in modes), n_edited, world),
        "e2e": {"value": round(e2e_value, 4), "unit": "images/sec", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": out_host.numel() * 2, "ms_per_step": round(ms_e2e / args.steps, 2)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "kernel": "gemm_kernel (tcgen05 bf16 GEMM, all epilogues)", "bound": "tensor",
            "achieved": round(gemm_tf, 1) if gemm_tf else None, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
            "frac": round(gemm_tf / peaks["tf_sustained"], 4) if gemm_tf else None,
            "traffic": (load_traffic() or {}).get("dram_bytes_per_launch"),
            "traffic_note": (load_traffic() or {}).get("note", "no ncu --set full capture committed"),
            "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
            "launches": int(pcnt[0]), "avg_launch_ms": round(psum[0] / max(pcnt[0], 1), 4),
            "share_of_step": round(pms[0] / ms, 4),
            "note": "achieved = sum(2MNK) / busy time, busy = union of the launches' CUDA-event intervals "
                    "(independent GEMMs of a block overlap on side streams)",
        },
        "roofline_attention": {
            "kernel": "attention_kernel (tcgen05/TMEM flash attention)", "bound": "tensor",
            "achieved": round(attn_tf, 1) if attn_tf else None, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
            "frac": round(attn_tf / peaks["tf_sustained"], 4) if attn_tf else None,
            "launches": int(pcnt[1]), "avg_launch_ms": round(psum[1] / max(pcnt[1], 1), 4),
            "share_of_step": round(pms[1] / ms, 4),
            "note": "busy time = union of the attention launches' intervals; in the single-stream blocks the MLP GEMM "
                    "shares the SMs with attention on purpose (tail fill), which this figure charges to attention",
        },
        "whole_step": {"tflop_per_image": None, "achieved_tflops": None},
    }
    # whole-image arithmetic rate: algorithmic FLOPs of every launched GEMM and attention / wall time of the images
    total_tf = (pwork[0] + pwork[1]) / 1e12 / args.steps
    line["whole_step"] = {"tflop_per_image": round(total_tf, 1),
                          "achieved_tflops": round(total_tf / (ms / args.steps / 1e3), 1),
                          "frac_of_sustained_peak": round(total_tf / (ms / args.steps / 1e3) / peaks["tf_sustained"], 4)}
    if world > 1:
        line["per_rank_ms_per_step"] = [round(m / args.steps, 2) for m in ms_ranks]
        line["per_rank_ms_per_step_e2e"] = [round(m / args.steps, 2) for m in ms_e2e_ranks]
        line["replica_check"] = replica
    helper.disable()
    if rank == 0 and world == 1 and args.arch != "tiny":
        if not args.no_reference_gpu:
            # the reference-equivalent eager PyTorch path on this very GPU (SURVEY §8d "how the reference path is
            # timed (1)"): same weights (shared tensors), inputs and schedule; the north star's >= 3x is this ratio
            ref = reference_gpu_images(args, pipe, max(args.steps, 5), 3)
            line["reference_gpu"] = ref
            line["speedup_vs_reference_gpu"] = {"device_resident": round(value / ref["value"], 3),
                                                "e2e": round(e2e_value / ref["value"], 3)}
        if not args.no_cpu_baseline:
            v, secs, desc, threads = cpu_reference_sample(n_edited, region_steps)
            line["cpu_baseline"] = {"value": round(v, 6), "unit": "images/sec (extrapolated from a bounded sample)",
                                    "cores": threads, "kind": "port", "sample": desc,
                                    "measured_seconds": round(secs, 1)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def workload_config(args, schedule: str, n_edited: int, world: int) -> dict:
    """The `config` object both arms print (the reference arm with the workload's known schedule / edited count)."""
    if args.arch == "tiny":
        return {"workload": "tiny smoke configuration"}
    return {"workload": WORKLOAD, "schedule": schedule, "edited_tokens": n_edited, "rho_target": args.rho,
            "l2": "inputs larger than L2 (23.7 GB of weights + 6 GB KV cache per image >> 126 MB)",
            "parallelism": f"replica x{world}; NCCL all-gather of the partition mask only" if world > 1 else "1 GPU"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_edited = args.ref_edited
    vals = []
    total = 0.0
    for i in range(args.warmup + args.steps):
        v, secs, desc, threads = cpu_reference_sample(n_edited, REGION_STEPS)
        total += secs
        if i >= args.warmup:
            vals.append(v)
    v = sum(vals) / len(vals)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 6), "unit": "images/sec",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 / v, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        # the same configuration object as our arm; schedule and edited-token count are those of this workload
        # (tests/test_flux_fullimage_gpu.py pins them against the oracle), the value is EXTRAPOLATED from the sample
        "config": workload_config(args, "FFFFFFRSRSSRSRSFSSSSSSSRSSFF", n_edited, world),
        "extrapolated": True,
        "cpu_baseline": {"value": round(v, 6), "unit": "images/sec (extrapolated from a bounded sample)",
                         "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": round(v, 6), "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def reference_gpu_images(args, pipe, n_timed: int, n_warm: int) -> dict:
    """The reference-equivalent PyTorch path of SURVEY §8d on the SAME GPU: the oracle restatement of the reference's
    loop / forward / processor run in eager torch on cuda:0, calling what the reference calls where it calls it -
    flash-attn's `flash_attn_func` (inplace.py:796-801) and the reference's OWN Triton `_partially_linear`
    (fused_kernels.py:81-101, staged unmodified under oracle/_ref by oracle/build_ref.py; a cuBLAS GEMM + fp16 round
    trip stands in, and says so, only if that file was never staged) - plus eager torch ops (cuBLAS nn.Linear, fp32
    RMSNorm / RoPE over the whole cache every step, torch.cat of K / V: K-d ... K-g of SURVEY §2.2). Same weights
    (the very tensors of `pipe`), inputs and schedule as our arm; `n_warm` warm-up images, then `cuda.synchronize()` +
    wall clock around each image exactly like src/FluxKontext/main.py:49-73."""
    import torch.nn.functional as F
    import oracle.flux as of
    from flash_attn import flash_attn_func
    from oracle.build_ref import load_partially_linear
    from oracle.loop import run_regione
    from oracle.schedule import GAMMA
    from standins import synthetic as syn
    from regione_b200.schedule import latent_image_ids

    dev = pipe.transformer.x_embedder.weight.device

    def fa(q, k, v):                       # [B,H,S,hd] like the processor after RoPE; flash-attn wants [B,S,H,hd]
        o = flash_attn_func(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), causal=False)
        return o.reshape(o.shape[0], o.shape[1], -1)

    triton_pl = None if args.ref_no_triton else load_partially_linear()

    def pl_cublas(x, w, b, index, cache):
        cache[:, index, :] = F.linear(x, w, b).to(torch.float16).to(cache.dtype)

    def pl_triton(x, w, b, index, cache):
        triton_pl(x.contiguous(), w, b, index, cache)

    saved = of.exact_attention, of.partially_linear
    of.exact_attention, of.partially_linear = fa, (pl_triton if triton_pl is not None else pl_cublas)
    try:
        arch = syn.FLUX_KONTEXT
        w = {k: v.detach() for k, v in pipe.transformer.state_dict().items()}
        model = of.FluxOracle(w, arch["heads"], arch["n_double"], arch["n_single"], True)
        inp = syn.make_inputs(110, GRID, GRID, TXT, arch["ctx_dim"], arch["pooled_dim"], rho=args.rho, device=dev)
        ids = torch.cat([latent_image_ids(GRID, GRID, 0.0, dev), latent_image_ids(GRID, GRID, 1.0, dev)])
        txt_ids = torch.zeros(TXT, 3, device=dev)

        def image():
            with torch.no_grad():
                return run_regione(model, dict(num_inference_steps=28, **PARAMS), GAMMA["FluxKontext"],
                                   inp["latents"], inp["image_latents"], ids, txt_ids, inp["prompt_embeds"],
                                   inp["pooled_prompt_embeds"], 2.5, inp["height"], inp["width"])

        for _ in range(n_warm):
            out, tr = image()
        times = []
        for _ in range(n_timed):
            torch.cuda.synchronize()
            t0 = time.time()
            out, tr = image()
            torch.cuda.synchronize()
            times.append(time.time() - t0)
    finally:
        of.exact_attention, of.partially_linear = saved
    sec = sum(times) / len(times)
    return {"value": round(1.0 / sec, 4), "unit": "images/sec", "ms": round(sec * 1e3, 1),
            "per_image_s": [round(t, 4) for t in times], "warmup_images": n_warm,
            "attention": "flash_attn_func (flash-attn %s)" % __import__("flash_attn").__version__,
            "partially_linear": ("reference Triton kernel (oracle/_ref/fused_kernels.py, triton %s)"
                                 % __import__("triton").__version__) if triton_pl is not None
                                else "cuBLAS GEMM + fp16 round trip (oracle/_ref not staged)",
            "schedule": "".join(m[0] for m in tr["modes"]), "edited_tokens": int(tr["edited_ids"].numel()),
            "timing": "cuda.synchronize() + wall clock per image, as src/FluxKontext/main.py:62-73"}


def run_reference_gpu(args):
    """`--impl reference_gpu`: only the reference-equivalent GPU arm (the default N = 1 run of our arm embeds it)."""
    from standins import synthetic as syn
    pipe = syn.build_pipeline(syn.FLUX_KONTEXT, seed=110, device=torch.device("cuda", 0))
    ref = reference_gpu_images(args, pipe, args.steps, args.warmup)
    print(json.dumps({
        "impl": "reference_gpu", "metric": METRIC, "value": ref["value"], "unit": "images/sec", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ref["ms"], "higher_is_better": True,
        "dtype": "bf16", "data": "synthetic",
        "config": workload_config(args, ref["schedule"], ref["edited_tokens"], 1), "reference_gpu": ref}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference_gpu"])
    ap.add_argument("--rho", type=float, default=0.25, help="target edited fraction of the synthetic image")
    ap.add_argument("--arch", default="flux", choices=["flux", "tiny"])
    ap.add_argument("--ref-edited", type=int, default=EDITED_TOKENS,
                    help="edited tokens of the CPU reference sample (default: what the default workload yields)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true",
                    help="skip the reference-equivalent PyTorch arm on the same GPU (N = 1 runs embed it by default)")
    ap.add_argument("--ref-no-triton", action="store_true",
                    help="reference GPU arm: cuBLAS + fp16 round trip instead of the staged Triton _partially_linear")
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed (device-resident) images, for ncu launch lists")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference_gpu":
        run_reference_gpu(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
