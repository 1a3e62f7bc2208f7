"""Command-line harness of the reference (src/<Family>/main.py:13-129, script/<Family>.sh) over the B200 hot path.

    python -m regione_b200.cli FluxKontext --use_regione --erosion_dilation --model_path synthetic \
        --image_path assets/data.jsonl --output_dir result/FluxKontext/Demo/RegionE

Same flags, defaults, warm-up count (3 images on the first demo item), per-image wall clock between two
`torch.cuda.synchronize()` calls and `time_consuming.json` / `metadata.json` layout as the reference, so the numbers
drop into evaluation/metric_merge.py:36-62 unchanged. Two model sources:

  * `--model_path <dir or hub id>`: the real diffusers pipeline (needs diffusers + weights; neither exists offline).
    The pipeline is loaded exactly like the reference's non-RegionE branch (main.py:38-39) and RegionE is switched on
    through the plugin surface, `RegionEHelper(pipe).set_params(...).enable()`; images are saved as PNG.
  * `--model_path synthetic[:tiny]`: the duck-typed stand-in pipeline of the family with seeded random weights at the
    real (or tiny) shapes. Text encoders and VAE do not exist, so each item's instruction / image are replaced by
    seeded synthetic embeddings and packed latents (seed = --seed + crc32(key)); the output latents are saved as
    `<key>.pt`. Timing and bookkeeping are the reference's.

Without `--use_regione` the reference runs the vanilla diffusers loop, which is not part of this repository: the
real-model path runs the un-patched pipeline, the synthetic path refuses.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
import zlib
from types import SimpleNamespace

import torch

# family -> (pipeline class name, guidance default, threshold, cache_threshold, model_path default, result dir,
#            name of the guidance keyword of the pipeline call, extra call keywords)        src/<Family>/main.py:15-32
FAMILIES = {
    "FluxKontext": ("FluxKontextPipeline", 2.5, 0.93, 0.04, "/mnt/jfs-test/lib/FLUX.1-Kontext-dev", "FluxKontext",
                    "guidance_scale", {}),
    "Step1X-Edit": ("Step1XEditPipeline", 6.0, 0.88, 0.02, "/mnt/jfs-test/lib/Step1X-Edit-v1p1-diffusers",
                    "Step1X-Edit", "true_cfg_scale", {}),
    "Step1X-Edit-v1p2": ("Step1XEditPipelineV1P2", 6.0, 0.88, 0.02, "/mnt/jfs-test/lib/Step1X-Edit-v1p2",
                         "Step1X-Edit-v1p2", "true_cfg_scale",
                         {"enable_thinking_mode": False, "enable_reflection_mode": False}),
    "Qwen-Image": ("QwenImageEditPipeline", 4.0, 0.80, 0.03, "/mnt/jfs-test/lib/Qwen-Image-Edit", "Qwen-Image",
                   "true_cfg_scale", {"negative_prompt": " "}),
    "Qwen-Image-Edit-2509": ("QwenImageEditPlusPipeline", 4.0, 0.80, 0.03, "/mnt/jfs-test/lib/Qwen-Image-Edit-2509",
                             "Qwen-Image-Edit-2509", "true_cfg_scale", {"negative_prompt": " ", "guidance_scale": 1.0}),
}
WARMUP_IMAGES = 3                      # main.py:50-58
WARMUP_KEY = "assets/demo_0"           # main.py:53


def build_parser(family: str) -> argparse.ArgumentParser:
    _, guidance, threshold, cache_threshold, model_path, result, _, _ = FAMILIES[family]
    p = argparse.ArgumentParser(prog=f"regione_b200.cli {family}")
    p.add_argument("--seed", type=int, default=110, help="Random seed for reproducibility")
    p.add_argument("--device", type=str, default="cuda", help="Device to run the model on (CUDA only here)")
    p.add_argument("--num_inference_steps", type=int, default=28, help="Number of inference steps for the model")
    p.add_argument("--guidance_scale", type=float, default=guidance, help="Guidance scale for the model")
    p.add_argument("--use_regione", action="store_true", help="Whether to use regione")
    p.add_argument("--warmup_step", type=int, default=6, help="Step of the stablization stage")
    p.add_argument("--post_step", type=int, default=2, help="Step of the smooth stage")
    p.add_argument("--refresh_step", type=str, default="16",
                   help="Steps are forcibly updated during the region-aware generation stage, format(str):16,22")
    p.add_argument("--threshold", type=float, default=threshold, help="Threshold for adaptive region partition")
    p.add_argument("--cache_threshold", type=float, default=cache_threshold,
                   help="Threshold for adaptive velocity decacy cache")
    p.add_argument("--erosion_dilation", action="store_true", help="Whether to use dilation and erosion")
    p.add_argument("--model_path", type=str, default=model_path,
                   help="Path to the pre-trained model, or synthetic / synthetic:tiny")
    p.add_argument("--evaluation", action="store_true", help="Whether to evaluate the model on the benchmark")
    p.add_argument("--image_path", type=str, default="assets/data.jsonl", help="Path to the input data")
    p.add_argument("--output_dir", type=str, default=f"result/{result}/Demo/RegionE",
                   help="Directory to save the output images")
    # synthetic source only
    p.add_argument("--grid", type=int, nargs=2, default=None, help="synthetic: latent token grid (rows cols)")
    p.add_argument("--txt_len", type=int, default=None, help="synthetic: prompt tokens")
    p.add_argument("--rho", type=str, default="0.25",
                   help="synthetic: edited fraction of each image, or `sweep` = per item one of 5/10/15/25/40/60/100 %% "
                        "(SURVEY §8d config 5)")
    p.add_argument("--no_warmup", action="store_true", help="skip the 3 warm-up images")
    return p


# ------------------------------------------------------------------------------------------------ model sources
def _load_real(family: str, args):
    try:
        import diffusers
    except ImportError as e:
        raise SystemExit(f"regione_b200.cli: --model_path {args.model_path} needs the diffusers package (fork "
                         f"Peyton-Chen/diffusers@step1xedit_v1p2, README.md:76-77), which is not installed; "
                         f"use --model_path synthetic") from e
    if args.use_regione and FAMILIES[family][0] != "FluxKontextPipeline":
        from .flux_kontext import LATENT_SPACE_ONLY
        raise SystemExit("regione_b200.cli: " + LATENT_SPACE_ONLY + " Run this family with --model_path synthetic, or "
                         "drive the patched pipeline from your own script with latents.")
    cls = getattr(diffusers, FAMILIES[family][0])
    pipe = cls.from_pretrained(args.model_path, torch_dtype=torch.bfloat16).to(args.device)
    return pipe


RHO_SWEEP = (0.05, 0.10, 0.15, 0.25, 0.40, 0.60, 1.00)


def _rho(args, seed: int) -> float:
    return RHO_SWEEP[seed % len(RHO_SWEEP)] if args.rho == "sweep" else float(args.rho)


def _shard(items):
    """Data-parallel runs (one process per GPU under torchrun; script/*.sh of the reference are run once per GPU on a
    slice of the benchmark): rank r of W takes items r, r + W, ... No collective is involved."""
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    return items[rank::world] if world > 1 else items


def _synthetic(family: str, args):
    """Stand-in pipeline of the family + a function item_key -> call keywords (seeded synthetic latents / embeds)."""
    try:   # test / benchmark scaffolding that lives next to the package, not in it
        from standins import diffusers_like as standin, step1x as sx
        from standins import synthetic as syn
    except ImportError as e:
        raise SystemExit("regione_b200.cli: --model_path synthetic needs the repository's `standins` package on "
                         "sys.path (run from the repository root)") from e

    tiny = args.model_path.endswith(":tiny")
    dev = torch.device(args.device)
    gh, gw = args.grid or ((16, 16) if tiny else (64, 64))
    T = args.txt_len or (32 if tiny else {"FluxKontext": 512, "Step1X-Edit": 640, "Step1X-Edit-v1p2": 640}.get(family, 256))
    base = dict(dim=256, heads=2, mlp_ratio=4, in_channels=64, ctx_dim=128) if tiny else \
        dict(dim=3072, heads=24, mlp_ratio=4, in_channels=64)
    scale = 0.3 / (0.02 * (base["dim"] ** 0.5))   # keeps |dt v| << |x| (synthetic.build_pipeline)

    def damp(tr):
        with torch.no_grad():
            tr.proj_out.weight.mul_(scale)
            tr.proj_out.bias.mul_(scale)
        return tr

    if family == "FluxKontext":
        arch = syn.TINY if tiny else syn.FLUX_KONTEXT
        pipe = syn.build_pipeline(arch, seed=110, device=dev)
        ctx, pooled = arch["ctx_dim"], arch["pooled_dim"]

        def inputs(seed):
            i = syn.make_inputs(seed, gh, gw, T, ctx, pooled, rho=_rho(args, seed), device=dev)
            i.pop("intended_mask")
            return i
    elif family in ("Step1X-Edit", "Step1X-Edit-v1p2"):
        v2 = family.endswith("v1p2")
        nd, ns = (2, 2) if tiny else (19, 38)
        extra = dict(vec_dim=64) if tiny else dict(ctx_dim=4096, vec_dim=768)
        if v2:
            extra["text_dim"] = 96 if tiny else 3584
        tr_cls = sx.Step1XEditV1P2Transformer2DModel if v2 else sx.Step1XEditTransformer2DModel
        tr = damp(tr_cls(**base, n_double=nd, n_single=ns, **extra).init_synthetic(110, dev))
        pipe = (sx.Step1XEditPipelineV1P2 if v2 else sx.Step1XEditPipeline)(tr)
        ctx = tr.context_embedder.in_features

        def inputs(seed):
            i = syn.make_inputs(seed, gh, gw, T, ctx, 64, rho=_rho(args, seed), device=dev)
            g = torch.Generator().manual_seed(seed + 1)
            neg = (0.1 * torch.randn(1, T, ctx, generator=g)).to(dev, torch.bfloat16)
            mask = torch.ones(1, T, dtype=torch.long, device=dev)
            kw = dict(latents=i["latents"], image_latents=i["image_latents"], height=i["height"], width=i["width"])
            if not v2:
                kw.update(prompt_embeds=i["prompt_embeds"], prompt_embeds_mask=mask, negative_prompt_embeds=neg,
                          negative_prompt_embeds_mask=mask)
                return kw
            td = tr.text_token_mapping.in_features

            def pack(emb):
                return SimpleNamespace(embedding=emb, mask=mask, txt_ids=torch.zeros(T, 3, device=dev),
                                       text_embeds=(0.1 * torch.randn(1, T, td, generator=g)).to(dev, torch.bfloat16),
                                       text_masks=torch.ones(1, T, device=dev, dtype=torch.bfloat16))
            kw.update(prompt_embeds=pack(i["prompt_embeds"]), negative_prompt_embeds=pack(neg))
            return kw
    else:
        nb = 3 if tiny else 60
        arch = dict(base, n_blocks=nb) if tiny else dict(base, n_blocks=nb, ctx_dim=3584)
        tr = damp(standin.QwenImageTransformer2DModel(**arch).init_synthetic(110, dev))
        cls = type(FAMILIES[family][0], (standin.QwenImageEditPipeline,), {})   # the helper dispatches on the NAME
        pipe = cls(tr)
        ctx = arch["ctx_dim"]

        def inputs(seed):
            i = syn.make_inputs(seed, gh, gw, T, ctx, 64, rho=_rho(args, seed), device=dev)
            g = torch.Generator().manual_seed(seed + 1)
            neg = (0.1 * torch.randn(1, T, ctx, generator=g)).to(dev, torch.bfloat16)
            return dict(latents=i["latents"], image_latents=i["image_latents"], prompt_embeds=i["prompt_embeds"],
                        negative_prompt_embeds=neg, height=i["height"], width=i["width"])
    return pipe, inputs


# ------------------------------------------------------------------------------------------------ the harness
def _read_jsonl(path):
    with open(path, "r") as f:
        return [json.loads(line) for line in f if line.strip()]


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] not in FAMILIES:
        print(f"usage: python -m regione_b200.cli {{{','.join(FAMILIES)}}} [options]", file=sys.stderr)
        return 2
    family = argv.pop(0)
    args = build_parser(family).parse_args(argv)
    if not torch.cuda.is_available() or not str(args.device).startswith("cuda"):
        raise SystemExit("regione_b200.cli: the hot path is CUDA-only (sm_100a); there is no CPU fallback")
    if "LOCAL_RANK" in os.environ and args.device == "cuda":      # torchrun: one process per GPU
        args.device = f"cuda:{os.environ['LOCAL_RANK']}"
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    from .helper import RegionEHelper

    synthetic = args.model_path.startswith("synthetic")
    guidance_kw, extra_kw = FAMILIES[family][6], dict(FAMILIES[family][7])
    if synthetic:
        if not args.use_regione:
            raise SystemExit("regione_b200.cli: the vanilla (non-RegionE) loop is diffusers code; pass --use_regione")
        pipe, make_inputs = _synthetic(family, args)
        for k in ("negative_prompt", "enable_thinking_mode", "enable_reflection_mode", "guidance_scale"):
            extra_kw.pop(k, None)        # handled by the real pipelines' text front ends
    else:
        pipe, make_inputs = _load_real(family, args), None
    if args.use_regione:                                               # main.py:34-36 through the plugin surface
        helper = RegionEHelper(pipe)
        helper.set_params(num_inference_steps=args.num_inference_steps, warmup_step=args.warmup_step,
                          post_step=args.post_step, refresh_step=args.refresh_step, threshold=args.threshold,
                          cache_threshold=args.cache_threshold, erosion_dilation=args.erosion_dilation)
        helper.enable()
        pipe = helper.pipeline

    def run(key, instruction, image_file):
        common = dict(num_inference_steps=args.num_inference_steps, **{guidance_kw: args.guidance_scale}, **extra_kw)
        if synthetic:
            seed = args.seed + (zlib.crc32(key.encode()) & 0xFFFF)
            out = pipe(output_type="latent", return_dict=False, **make_inputs(seed), **common)[0]
            return out
        from PIL import Image
        img = Image.open(image_file).convert("RGB")
        return pipe(image=img, prompt=instruction, generator=torch.Generator("cpu").manual_seed(args.seed),
                    **common).images[0]

    def save(result, path_no_ext):
        os.makedirs(os.path.dirname(path_no_ext) or ".", exist_ok=True)
        if synthetic:
            torch.save(result.cpu(), path_no_ext + ".pt")
            return path_no_ext + ".pt"
        result.save(path_no_ext + ".png")
        return path_no_ext + ".png"

    def warmup():
        if args.no_warmup:
            return
        print("Warmup...")
        for _ in range(WARMUP_IMAGES):
            run(WARMUP_KEY, "just warmup!", WARMUP_KEY + ".png")

    def timed(key, instruction, image_file):
        torch.cuda.synchronize()
        t0 = time.time()
        result = run(key, instruction, image_file)
        torch.cuda.synchronize()
        return result, time.time() - t0

    if not args.evaluation:                                            # main.py:41-76
        os.makedirs(args.output_dir, exist_ok=True)
        metadata = _shard(_read_jsonl(args.image_path))
        warmup()
        t_all = time.time()
        for index, data in enumerate(metadata):
            print(f"[{index + 1} / {len(metadata)}] Reference Image: {data['key']}.png, "
                  f"Instruction: {data['instruction']}")
            result, dt = timed(data["key"], data["instruction"], f"{data['key']}.png")
            print(f"Time consuming: {dt}s")
            save(result, os.path.join(args.output_dir, os.path.basename(data["key"])))
            print(f"Image has been saved to {args.output_dir}")
        if metadata:
            print(f"rank {os.environ.get('RANK', '0')}: {len(metadata)} images in {time.time() - t_all:.3f}s")
        return 0
    for task in sorted(os.listdir(args.image_path)):                   # main.py:78-129
        image_path = os.path.join(args.image_path, task)
        if not os.path.isfile(os.path.join(image_path, "metadata.jsonl")):
            continue
        output_dir = os.path.join(args.output_dir, task)
        os.makedirs(f"{output_dir}/generation", exist_ok=True)
        metadata = _shard(_read_jsonl(f"{image_path}/metadata.jsonl"))
        warmup()
        prefix_prompt, time_consuming = {}, []
        for idx, data in enumerate(metadata):
            prompt = data["instruction"]
            print(f"prompt:{prompt}")
            result, dt = timed(f"{task}/{data['key']}", prompt, f"{image_path}/img/{data['key']}.png")
            prefix_prompt[data["key"]] = prompt
            time_consuming.append(dt)
            where = save(result, f"{output_dir}/generation/{data['key']}")
            print(f"[task:{task} {idx + 1}/{len(metadata)}] {where}, save! cosuming:{dt}s")
        suffix = f".rank{os.environ['RANK']}" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else ""
        with open(f"{output_dir}/time_consuming{suffix}.json", "w") as f:
            json.dump({"num_item": len(time_consuming),
                       "ave_time_consuming": sum(time_consuming) / max(len(time_consuming), 1),
                       "time_consuming_list": time_consuming}, f, indent=4)
        with open(f"{output_dir}/metadata{suffix}.json", "w") as f:
            json.dump(prefix_prompt, f, indent=4)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
