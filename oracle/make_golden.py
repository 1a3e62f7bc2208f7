"""Generates tests/golden/*.pt by EXECUTING THE REFERENCE'S OWN CODE in the build container. TEST INFRASTRUCTURE.

    python -m oracle.make_golden            # needs /root/reference; never runs on the GPU box

What runs from the reference (read where it lies, nothing is copied into this repository):
  * RegionE/FluxKontext/utils.py, imported unchanged behind a stub `diffusers` package (the only names it needs at
    import time are five symbols, utils.py:6-16): token_selector, morphology, ids_gather/ids_scatter and
    FluxKontextManager produce the fixtures for oracle/region_ops.py;
  * the AVDC decision block (inplace.py:295-313), the condition-concat test (:331), the scheduler's refresh
    bookkeeping (:630-641) and the gamma table (:47-50) are read from inplace.py by line number and exec'd in a
    harness (inplace.py itself cannot be imported: it subclasses diffusers classes). Together with the real
    FluxKontextManager.step this yields the golden 28-step schedule for oracle/schedule.py and the product's planner.
The sigma schedule fed to that harness comes from oracle.schedule.flow_match_sigmas (diffusers restated; unpinned).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import textwrap
import types

import torch

REF = "/root/reference/RegionE"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference_utils(family: str = "FluxKontext"):
    names = {
        "diffusers": dict(FluxKontextPipeline=object, Step1XEditPipeline=object, QwenImageEditPipeline=object),
        "diffusers.image_processor": dict(PipelineImageInput=object),
        "diffusers.pipelines": {},
        "diffusers.pipelines.flux": dict(FluxPipelineOutput=object),
        "diffusers.utils": dict(BaseOutput=object, is_torch_xla_available=lambda: False),
    }
    saved = {k: sys.modules.get(k) for k in names}
    for mod, attrs in names.items():
        m = types.ModuleType(mod)
        m.__dict__.update(attrs)
        sys.modules[mod] = m
    try:
        spec = importlib.util.spec_from_file_location(f"_ref_{family}_utils", f"{REF}/{family}/utils.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def ref_lines(path: str, first: int, last: int) -> str:
    with open(path) as f:
        lines = f.readlines()[first - 1:last]
    return textwrap.dedent("".join(lines))


def synthetic_partition_inputs(seed: int, gh: int, gw: int, frac: float):
    """Condition latent + an x0 estimate that equals it except inside blobs (edited) and salt noise, so the cosine
    map has structure for the morphology to clean up."""
    g = torch.Generator().manual_seed(seed)
    L = gh * gw
    cond = (torch.randn(1, L, 64, generator=g) * 0.6).bfloat16()
    est = cond.float() + 0.15 * torch.randn(1, L, 64, generator=g)
    yy, xx = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
    blob = torch.zeros(gh, gw, dtype=torch.bool)
    n_blobs = 3
    for b in range(n_blobs):
        cy = int(torch.randint(0, gh, (1,), generator=g))
        cx = int(torch.randint(0, gw, (1,), generator=g))
        r = max(1.5, (frac * L / n_blobs / 3.1416) ** 0.5)
        blob |= ((yy - cy) ** 2 + (xx - cx) ** 2) <= r * r
    salt = torch.rand(gh, gw, generator=g) < 0.02
    edit = (blob | salt).flatten()
    est[0, edit] = torch.randn(int(edit.sum()), 64, generator=g) * 0.6
    return est, cond


def make_region_ops(u):
    cases = []
    for seed, gh, gw, frac, thr, ed in [(1, 16, 16, 0.25, 0.88, True), (2, 32, 32, 0.10, 0.88, True),
                                        (3, 32, 32, 0.25, 0.93, False), (4, 24, 40, 0.5, 0.88, True),
                                        (5, 64, 64, 0.25, 0.88, True), (6, 16, 16, 0.0, 0.88, True)]:
        est, cond = synthetic_partition_inputs(seed, gh, gw, frac)
        if frac == 0.0:  # nothing edited -> empty edited set
            est = cond.float().clone()
        e, un = u.token_selector(est, cond, thr, similarity_type="cosine", height=gh * 16, width=gw * 16,
                                 erosion_dilation=ed, patch_size=2, vae_scale_factor=8)
        sim = torch.sum(torch.nn.functional.normalize(est, dim=-1) * torch.nn.functional.normalize(cond, dim=-1), -1)
        case = dict(seed=seed, gh=gh, gw=gw, frac=frac, threshold=thr, erosion_dilation=ed,
                    edited=e.to(torch.int32), unedited=un.to(torch.int32),
                    min_margin=float((sim - thr).abs().min()))
        if gh * gw <= 1024:
            case.update(estimate=est, condition=cond)
        cases.append(case)
    morph = []
    g = torch.Generator().manual_seed(7)
    grids = [torch.ones(6, 6), torch.zeros(5, 7), (torch.rand(16, 16, generator=g) < 0.6).float(),
             (torch.rand(64, 64, generator=g) < 0.8).float(), (torch.rand(9, 33, generator=g) < 0.7).float()]
    for m in grids:
        morph.append(dict(mask=m.to(torch.uint8), out=u.remove_scattered_points(m, 5, "square").to(torch.uint8)))
    lat = torch.randn(1, 40, 8, generator=g).bfloat16()
    ids = torch.randperm(40, generator=g)[:13].sort().values.unsqueeze(0)
    gathered = u.ids_gather(lat, ids)
    scattered = u.ids_scatter(gathered, ids, torch.zeros_like(lat))
    torch.save(dict(selector=cases, morphology=morph,
                    gather=dict(latent=lat, ids=ids.to(torch.int32), gathered=gathered, scattered=scattered)),
               os.path.join(OUT, "region_ops.pt"))
    return cases


def make_schedule(u):
    from oracle.schedule import flow_match_sigmas
    inplace = f"{REF}/FluxKontext/inplace.py"
    ns = {"torch": torch}
    exec(ref_lines(inplace, 47, 50), ns)                      # gamma
    gamma = ns["gamma"]
    avdc_src = ref_lines(inplace, 295, 313)
    concat_src = ref_lines(inplace, 331, 331).strip()
    concat_expr = concat_src[len("if image_latents is not None and "):-1]
    sched_src = ref_lines(inplace, 630, 641)
    out = []
    for params in [dict(warmup_step=6, post_step=2, refresh_step="16", cache_threshold=0.04),
                   dict(warmup_step=6, post_step=2, refresh_step="16", cache_threshold=0.01),
                   dict(warmup_step=6, post_step=2, refresh_step="16", cache_threshold=0.02),
                   dict(warmup_step=4, post_step=3, refresh_step="10,18", cache_threshold=0.04),
                   dict(warmup_step=8, post_step=1, refresh_step="12,20,24", cache_threshold=0.08)]:
        M = u.FluxKontextManager()
        M.set_parameters(dict(num_inference_steps=28, threshold=0.88, erosion_dilation=True, **params))
        L = 64
        sigmas, timesteps = flow_match_sigmas(28, 4096)
        lat = torch.zeros(1, L, 8)
        cond = torch.zeros(1, L, 8)
        ids = torch.zeros(2 * L, 3)
        M.refresh(lat, cond, ids, torch.zeros(5, 3), 2, 8, 128, 128)

        class _S:  # what the exec'd scheduler block reads through `self`
            pass
        S = _S()
        S.sigmas = sigmas
        env = dict(MANAGER=M, gamma=gamma, timesteps=timesteps, should_cache=False, accumulate=1, error=0,
                   self=S, torch=torch)
        steps = []
        for i, t in enumerate(timesteps):
            assert i == M.current_step
            env.update(i=i, t=t)
            exec(avdc_src, env)
            skip = bool(env["should_cache"])
            full = bool(eval(concat_expr, env))
            ratio = float(env["ratio"]) if "ratio" in env and i > M.warmup_step else None
            env.update(sigma=sigmas[i], sigma_next=sigmas[i + 1])
            exec(sched_src, env)
            if M.current_step == M.warmup_step - 1:
                M.edited_ids = torch.arange(0, L, 4).unsqueeze(0)
                M.unedited_ids = torch.tensor([j for j in range(L) if j % 4]).unsqueeze(0)
            write = M.current_step == M.warmup_step - 1 or M.current_step == M.prev_refresh_step
            rec = dict(mode="SKIP" if skip else ("FULL" if full else "REGION"), ratio=ratio,
                       write_cache=bool(write and not skip), dt=float(env["dt"]),
                       dt_direct=float(env["dt_direct"]) if "dt_direct" in env else None,
                       dt_final=float(env["dt_final"]) if "dt_final" in env else None)
            lat, ids = M.step(lat, ids)
            rec.update(rows_after=int(lat.shape[1]), prev_refresh_after=M.prev_refresh_step)
            steps.append(rec)
        out.append(dict(params=params, refresh_parsed=list(M.refresh_step), steps=steps,
                        timesteps=timesteps.clone(), sigmas=sigmas.clone()))
    torch.save(dict(gamma=gamma.clone(), plans=out), os.path.join(OUT, "schedule.pt"))
    return out


FAMILY_DIRS = {"FluxKontextPipeline": "FluxKontext", "Step1XEditPipeline": "Step1XEdit",
               "Step1XEditPipelineV1P2": "Step1XEditV1P2", "QwenImageEditPipeline": "QwenImageEdit",
               "QwenImageEditPlusPipeline": "QwenImageEditPlus"}
CLI_DIRS = ["FluxKontext", "Step1X-Edit", "Step1X-Edit-v1p2", "Qwen-Image", "Qwen-Image-Edit-2509"]


def make_front_end(u):
    """Constants and small host functions either side of the loop, produced by the reference's own code:
    calculate_shift (utils.py:38-48) on a sweep of token counts, every family's gamma table (inplace.py:47-50) and
    plugin defaults (tool/RegionE.py:1-7), and the argparse surface of src/<Family>/main.py:13-32 (the lines between
    `ArgumentParser()` and `parse_args()` exec'd against a real argparse)."""
    import argparse
    import json
    out = {"calculate_shift": {}, "gamma": {}, "defaults": {}, "cli": {}}
    for n in [256, 1024, 2304, 3600, 4050, 4096, 4104, 6400, 8192]:
        out["calculate_shift"][str(n)] = float(u.calculate_shift(n)).hex()
    out["calculate_shift_custom"] = float(u.calculate_shift(4050, 256, 8192, 0.5, 0.9)).hex()
    for name, d in FAMILY_DIRS.items():
        src = open(f"{REF}/{d}/inplace.py").read().split("\n")
        first = next(i for i, line in enumerate(src) if line.startswith("gamma = torch.tensor("))
        last = next(i for i in range(first, first + 8) if "dtype=torch.float16)" in src[i])
        ns = {"torch": torch}
        exec("\n".join(src[first:last + 1]), ns)
        out["gamma"][name] = [float(x) for x in ns["gamma"].tolist()]          # fp16 values, exactly representable
    ns = {}
    exec(ref_lines(f"{REF}/tool/RegionE.py", 1, 7), ns)
    out["defaults"] = ns["config"]
    for d in CLI_DIRS:
        src = open(f"{REF}/../src/{d}/main.py").read().split("\n")
        first = next(i for i, line in enumerate(src) if "argparse.ArgumentParser()" in line)
        last = next(i for i, line in enumerate(src) if "parser.parse_args()" in line)
        ns = {"argparse": argparse}
        exec(textwrap.dedent("\n".join(src[first:last])), ns)
        flags = {}
        for a in ns["parser"]._actions:
            if a.dest == "help":
                continue
            flags[a.dest] = dict(default=a.default, type=getattr(a.type, "__name__", None),
                                 flag=isinstance(a, argparse._StoreTrueAction))
        out["cli"][d] = flags
    with open(os.path.join(OUT, "front_end.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    return out


def make_family_schedules(u):
    """The 28-step AVDC plan of EVERY family, produced by that family's own loop code: its gamma table
    (inplace.py:47-50) and its AVDC block (`if MANAGER.current_step <= MANAGER.warmup_step ...` down to
    `should_cache = True`, found by pattern in RegionE/<Family>/inplace.py) are exec'd step by step, with the refresh
    bookkeeping of the scheduler (FluxKontext/inplace.py:630-641, identical in every family) and the real
    FluxKontextManager.step state machine in between. Parameters = the family's plugin defaults (tool/RegionE.py:1-7)
    plus the demo thresholds of script/*.sh."""
    import json
    from oracle.schedule import flow_match_sigmas
    ns = {}
    exec(ref_lines(f"{REF}/tool/RegionE.py", 1, 7), ns)
    defaults = ns["config"]
    sched_src = ref_lines(f"{REF}/FluxKontext/inplace.py", 630, 641)
    out = {}
    for name, d in FAMILY_DIRS.items():
        src = open(f"{REF}/{d}/inplace.py").read().split("\n")
        first = next(i for i, line in enumerate(src) if line.startswith("gamma = torch.tensor("))
        last = next(i for i in range(first, first + 8) if "dtype=torch.float16)" in src[i])
        gns = {"torch": torch}
        exec("\n".join(src[first:last + 1]), gns)
        a0 = next(i for i, line in enumerate(src) if line.strip().startswith("if MANAGER.current_step <= MANAGER.warmup_step"))
        a1 = next(i for i in range(a0, a0 + 40) if src[i].strip() == "should_cache = True")
        avdc_src = textwrap.dedent("\n".join(src[a0:a1 + 1]))
        plans = []
        for extra in ({}, {"cache_threshold": 0.01}, {"refresh_step": "12,20", "warmup_step": 5, "post_step": 3}):
            prm = dict(defaults[name], **extra)
            M = u.FluxKontextManager()
            M.set_parameters(prm)
            L = 64
            sigmas, timesteps = flow_match_sigmas(28, 4096)
            lat, ids = torch.zeros(1, L, 8), torch.zeros(2 * L, 3)
            M.refresh(lat, torch.zeros(1, L, 8), ids, torch.zeros(5, 3), 2, 8, 128, 128)

            class _S:
                pass
            S = _S()
            S.sigmas = sigmas
            env = dict(MANAGER=M, gamma=gns["gamma"], timesteps=timesteps, should_cache=False, accumulate=1, error=0,
                       self=S, torch=torch)
            steps = []
            for i, t in enumerate(timesteps):
                env.update(i=i, t=t)
                exec(avdc_src, env)
                skip = bool(env["should_cache"])
                steps.append(dict(skip=skip, ratio=float(env["ratio"]) if skip else None))
                env.update(sigma=sigmas[i], sigma_next=sigmas[i + 1])
                exec(sched_src, env)
                if M.current_step == M.warmup_step - 1:
                    M.edited_ids = torch.arange(0, L, 4).unsqueeze(0)
                    M.unedited_ids = torch.tensor([j for j in range(L) if j % 4]).unsqueeze(0)
                lat, ids = M.step(lat, ids)
            plans.append(dict(params=prm, steps=steps))
        out[name] = plans
    with open(os.path.join(OUT, "family_schedules.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    return out


def make_cfg():
    """Classifier-free-guidance arithmetic of the CFG families, produced by the reference's own lines on seeded bf16
    velocities: Step1XEdit/inplace.py:401-410 (norm-processed, above and below `timesteps_truncate`; `process_diff_norm`
    is the fork's function, absent here -> the stand-in the product tests use, passed in as `self`) and
    QwenImageEdit/inplace.py:401-405 (norm rescaling)."""
    from standins.step1x import Step1XEditPipeline
    g = torch.Generator().manual_seed(5)
    pos = torch.randn(1, 512, 64, generator=g).bfloat16()
    neg = (pos.float() + 0.5 * torch.randn(1, 512, 64, generator=g)).bfloat16()
    step1x_src = ref_lines(f"{REF}/Step1XEdit/inplace.py", 401, 410)
    qwen_src = ref_lines(f"{REF}/QwenImageEdit/inplace.py", 401, 405)
    out = dict(pos=pos, neg=neg, step1x=[], qwen=[])

    class _Self:
        process_diff_norm = staticmethod(Step1XEditPipeline.process_diff_norm)
    for t, scale in ((torch.tensor(976.2), 6.0), (torch.tensor(904.5), 6.0), (torch.tensor(935.6), 4.0)):
        env = dict(torch=torch, noise_pred=pos.clone(), neg_noise_pred=neg.clone(), t=t, timesteps_truncate=930.0,
                   true_cfg_scale=scale, process_norm_power=0.4, self=_Self())
        exec(step1x_src, env)
        out["step1x"].append(dict(t=float(t), truncate=930.0, scale=scale, k=0.4, out=env["noise_pred"].clone()))
    for scale in (4.0, 1.5):
        env = dict(torch=torch, noise_pred=pos.clone(), neg_noise_pred=neg.clone(), true_cfg_scale=scale)
        exec(qwen_src, env)
        out["qwen"].append(dict(scale=scale, out=env["noise_pred"].clone()))
    torch.save(out, os.path.join(OUT, "cfg.pt"))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    u = load_reference_utils("FluxKontext")
    cfg = make_cfg()
    print("cfg:", len(cfg["step1x"]), "step1x cases,", len(cfg["qwen"]), "qwen cases")
    fe = make_front_end(u)
    for name, plans in make_family_schedules(u).items():
        print(name, ["".join("S" if st["skip"] else "C" for st in p["steps"]) for p in plans])
    print("front end:", {k: len(v) if hasattr(v, "__len__") else v for k, v in fe.items()})
    cases = make_region_ops(u)
    for c in cases:
        print(f"selector seed {c['seed']} grid {c['gh']}x{c['gw']}: edited {c['edited'].shape[1]} "
              f"min|sim-thr| {c['min_margin']:.2e}")
    plans = make_schedule(u)
    for p in plans:
        print(p["params"], "".join(s["mode"][0] for s in p["steps"]))


if __name__ == "__main__":
    main()
