"""Quality / latency bookkeeping of the reference's evaluation scripts (evaluation/metric_all_task.py:12-100,
evaluation/metric_merge.py:6-62) — SURVEY §8f row 4; host-side numpy, not on the hot path.

PSNR and SSIM restate what the reference gets from scikit-image (not installed here): `peak_signal_noise_ratio` on
uint8 images (data range 255) and `structural_similarity(multichannel, channel_axis=-1)` with skimage's defaults — 7x7
uniform window, K1 = 0.01, K2 = 0.03, sample covariance (N / (N - 1)), borders of (win - 1) / 2 pixels cropped before
the mean, channels averaged. LPIPS needs the `lpips` package and its AlexNet weights (neither exists offline): it is
computed when importable and reported as NaN otherwise.
"""
from __future__ import annotations

import json
import os

import numpy as np
from scipy.ndimage import uniform_filter

VALID_EXT = {".jpg", ".jpeg", ".png", ".bmp", ".tiff", ".tif"}


def psnr(reference: np.ndarray, test: np.ndarray, data_range: float = 255.0) -> float:
    err = np.mean((reference.astype(np.float64) - test.astype(np.float64)) ** 2)
    return float("inf") if err == 0 else float(10.0 * np.log10(data_range ** 2 / err))


def _ssim_plane(x: np.ndarray, y: np.ndarray, data_range: float, win: int = 7, k1: float = 0.01, k2: float = 0.03) -> float:
    x, y = x.astype(np.float64), y.astype(np.float64)
    n = win * win
    cov_norm = n / (n - 1.0)
    ux, uy = uniform_filter(x, size=win), uniform_filter(y, size=win)
    uxx, uyy, uxy = uniform_filter(x * x, size=win), uniform_filter(y * y, size=win), uniform_filter(x * y, size=win)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    pad = (win - 1) // 2
    return float(s[pad:-pad, pad:-pad].mean())


def ssim(reference: np.ndarray, test: np.ndarray, data_range: float = 255.0) -> float:
    """[H, W] or [H, W, C] arrays; channels are averaged like skimage's channel_axis=-1."""
    if reference.shape != test.shape:
        raise ValueError("Input images must have the same dimensions.")
    if min(reference.shape[:2]) < 7:
        raise ValueError("win_size exceeds image extent.")
    if reference.ndim == 2:
        return _ssim_plane(reference, test, data_range)
    return float(np.mean([_ssim_plane(reference[..., c], test[..., c], data_range) for c in range(reference.shape[-1])]))


def _lpips_fn():
    try:
        import lpips
        import torch
    except ImportError:
        return None
    model = lpips.LPIPS(net="alex")

    def fn(a: np.ndarray, b: np.ndarray) -> float:
        t = lambda x: (torch.from_numpy(x).permute(2, 0, 1)[None].float() / 255.0 - 0.5) / 0.5   # noqa: E731
        with torch.no_grad():
            return float(model(t(a), t(b)))
    return fn


def calculate_image_metrics(folder1_path: str, folder2_path: str) -> dict:
    """metric_all_task.py:12-146: PSNR / SSIM / LPIPS of equally named images of two folders (folder1 = reference)."""
    from PIL import Image
    if not os.path.exists(folder1_path) or not os.path.exists(folder2_path):
        raise ValueError("Specified folder path does not exist")
    names = lambda p: {f for f in os.listdir(p) if os.path.splitext(f.lower())[1] in VALID_EXT}   # noqa: E731
    common = names(folder1_path) & names(folder2_path)
    if not common:
        raise ValueError("No images with matching names found in both folders")
    lp = _lpips_fn()
    results = {"individual_metrics": {}, "average_metrics": {}}
    for filename in sorted(common):
        a = Image.open(os.path.join(folder1_path, filename)).convert("RGB")
        b = Image.open(os.path.join(folder2_path, filename)).convert("RGB")
        if a.size != b.size:
            b = b.resize(a.size, Image.LANCZOS)
        an, bn = np.array(a), np.array(b)
        results["individual_metrics"][filename] = {
            "PSNR": psnr(an, bn), "SSIM": ssim(an, bn), "LPIPS": lp(an, bn) if lp else float("nan")}
    vals = list(results["individual_metrics"].values())
    results["average_metrics"] = {k: float(np.mean([v[k] for v in vals])) for k in ("PSNR", "SSIM", "LPIPS")}
    return results


def save_results_to_csv(results: dict, output_path: str = "image_metrics_results.csv") -> None:
    """metric_all_task.py:148-181: one row per image and a final AVERAGE row (what metric_merge reads with tail(1))."""
    with open(output_path, "w") as f:
        f.write("Filename,PSNR,SSIM,LPIPS\n")
        for name, m in results["individual_metrics"].items():
            f.write(f"{name},{m['PSNR']},{m['SSIM']},{m['LPIPS']}\n")
        a = results["average_metrics"]
        f.write(f"AVERAGE,{a['PSNR']},{a['SSIM']},{a['LPIPS']}\n")


def merge_metrics(path: str, tasks=None) -> dict:
    """metric_merge.py:6-62: prompt-weighted mean of the per-task AVERAGE rows (metric.csv) and latencies
    (time_consuming.json, written by regione_b200.cli --evaluation); writes merged_metric.txt. A directory called
    `pretrain` (the vanilla run every other run is compared with) has no metric.csv: PSNR inf, SSIM 1, LPIPS 0."""
    import glob

    def latency_files(task):
        # a single-process run writes time_consuming.json; a data-parallel run (torchrun, regione_b200.cli) writes one
        # time_consuming.rank<r>.json shard per rank - the shards are summed, never mixed with a stale single file
        shards = sorted(glob.glob(os.path.join(path, task, "time_consuming.rank*.json")))
        single = os.path.join(path, task, "time_consuming.json")
        return shards or ([single] if os.path.isfile(single) else [])

    tasks = tasks or sorted(d for d in os.listdir(path) if os.path.isdir(os.path.join(path, d)) and latency_files(d))
    vanilla = os.path.basename(os.path.normpath(path)).lower() == "pretrain"
    sums = {"PSNR": 0.0, "SSIM": 0.0, "LPIPS": 0.0}
    items, latency = 0, 0.0
    for task in tasks:
        n, task_latency = 0, 0.0
        for name in latency_files(task):
            with open(name) as f:
                lat = json.load(f)
            n += lat["num_item"]
            task_latency += lat["ave_time_consuming"] * lat["num_item"]
        items += n
        latency += task_latency
        if not vanilla:
            with open(os.path.join(path, task, "metric.csv")) as f:
                last = f.read().strip().splitlines()[-1].split(",")
            for k, v in zip(("PSNR", "SSIM", "LPIPS"), last[1:4]):
                sums[k] += float(v) * n
    if items == 0:
        raise ValueError(f"merge_metrics: no time_consuming[.rank*].json with items under {path}")
    out = {"PSNR": float("inf"), "SSIM": 1.0, "LPIPS": 0.0} if vanilla else {k: v / items for k, v in sums.items()}
    out.update(Prompts=items, Latency=latency / items)
    with open(os.path.join(path, "merged_metric.txt"), "w") as f:
        for k in ("PSNR", "SSIM", "LPIPS", "Prompts", "Latency"):
            f.write(f"{k}: {out[k]} \n")
    return out
