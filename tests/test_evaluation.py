"""Evaluation bookkeeping: PSNR / SSIM restated from scikit-image's definitions (not installed
here, so the checker is a direct window-by-window evaluation of the published formula), CSV / merge formats of the
reference's evaluation scripts."""
import json
import os

import numpy as np
import pytest

from regione_b200 import evaluation as ev


def _direct_ssim_plane(x, y, L=255.0, win=7):
    x, y = x.astype(np.float64), y.astype(np.float64)
    c1, c2 = (0.01 * L) ** 2, (0.03 * L) ** 2
    vals = []
    for i in range(x.shape[0] - win + 1):
        for j in range(x.shape[1] - win + 1):
            a, b = x[i:i + win, j:j + win].ravel(), y[i:i + win, j:j + win].ravel()
            ma, mb = a.mean(), b.mean()
            va, vb = a.var(ddof=1), b.var(ddof=1)
            cab = ((a - ma) * (b - mb)).sum() / (a.size - 1)
            vals.append((2 * ma * mb + c1) * (2 * cab + c2) / ((ma * ma + mb * mb + c1) * (va + vb + c2)))
    return float(np.mean(vals))


def test_psnr_and_ssim_definitions():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (24, 31, 3), dtype=np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-20, 21, a.shape), 0, 255).astype(np.uint8)
    mse = np.mean((a.astype(float) - b.astype(float)) ** 2)
    assert ev.psnr(a, b) == pytest.approx(10 * np.log10(255.0 ** 2 / mse))
    assert ev.psnr(a, a) == float("inf")
    assert ev.ssim(a, a) == pytest.approx(1.0)
    want = np.mean([_direct_ssim_plane(a[..., c], b[..., c]) for c in range(3)])
    assert ev.ssim(a, b) == pytest.approx(want, rel=1e-9)
    assert 0 < ev.ssim(a, b) < 1
    with pytest.raises(ValueError):
        ev.ssim(a, b[:-1])


def test_folder_metrics_csv_and_merge(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(1)
    for run in ("pretrain", "RegionE"):
        for task, n in (("color_alter", 2), ("text_change", 3)):
            gen = tmp_path / run / task / "generation"
            os.makedirs(gen)
            for i in range(n):
                img = rng.integers(0, 256, (32, 40, 3), dtype=np.uint8) if run == "pretrain" else \
                    np.array(Image.open(tmp_path / "pretrain" / task / "generation" / f"{i}.png")) // 2 * 2
                Image.fromarray(img.astype(np.uint8)).save(gen / f"{i}.png")
            with open(tmp_path / run / task / "time_consuming.json", "w") as f:
                json.dump({"num_item": n, "ave_time_consuming": 2.0 if run == "pretrain" else 1.0,
                           "time_consuming_list": [1.0] * n}, f)
    for task in ("color_alter", "text_change"):
        res = ev.calculate_image_metrics(str(tmp_path / "pretrain" / task / "generation"),
                                         str(tmp_path / "RegionE" / task / "generation"))
        assert set(res["average_metrics"]) == {"PSNR", "SSIM", "LPIPS"} and res["average_metrics"]["PSNR"] > 40
        ev.save_results_to_csv(res, str(tmp_path / "RegionE" / task / "metric.csv"))
    lines = open(tmp_path / "RegionE" / "text_change" / "metric.csv").read().strip().splitlines()
    assert lines[0] == "Filename,PSNR,SSIM,LPIPS" and len(lines) == 5 and lines[-1].startswith("AVERAGE,")
    merged = ev.merge_metrics(str(tmp_path / "RegionE"))
    assert merged["Prompts"] == 5 and merged["Latency"] == pytest.approx(1.0) and merged["PSNR"] > 40
    base = ev.merge_metrics(str(tmp_path / "pretrain"))
    assert base["PSNR"] == float("inf") and base["SSIM"] == 1.0 and base["Latency"] == pytest.approx(2.0)
    txt = open(tmp_path / "RegionE" / "merged_metric.txt").read()
    assert [l.split(":")[0] for l in txt.strip().splitlines()] == ["PSNR", "SSIM", "LPIPS", "Prompts", "Latency"]


def test_merge_metrics_recombines_rank_shards(tmp_path):
    """A data-parallel evaluation (torchrun, regione_b200.cli) writes time_consuming.rank<r>.json per rank: the merge
    sums the shards (and ignores a stale single-process file next to them); an empty directory is an error, not a
    ZeroDivisionError."""
    task = tmp_path / "pretrain" / "color_alter"
    os.makedirs(task)
    for r, (n, t) in enumerate(((3, 2.0), (1, 4.0))):
        with open(task / f"time_consuming.rank{r}.json", "w") as f:
            json.dump({"num_item": n, "ave_time_consuming": t, "time_consuming_list": [t] * n}, f)
    with open(task / "time_consuming.json", "w") as f:          # stale file of an earlier single-process run
        json.dump({"num_item": 100, "ave_time_consuming": 9.0, "time_consuming_list": []}, f)
    merged = ev.merge_metrics(str(tmp_path / "pretrain"))
    assert merged["Prompts"] == 4 and merged["Latency"] == pytest.approx((3 * 2.0 + 1 * 4.0) / 4)
    os.makedirs(tmp_path / "empty" / "task")
    with pytest.raises(ValueError):
        ev.merge_metrics(str(tmp_path / "empty"))
