"""Kontext's trained (width, height) buckets (RegionE/FluxKontext/utils.py:18-36, from diffusers) — host-side data
for the pixel-space entry of the pipeline; not used by the latent-space hot path."""

KONTEXT_RESOLUTIONS = [(672, 1568), (688, 1504), (720, 1456), (752, 1392), (800, 1328), (832, 1248), (880, 1184),
                       (944, 1104), (1024, 1024), (1104, 944), (1184, 880), (1248, 832), (1328, 800), (1392, 752),
                       (1456, 720), (1504, 688), (1568, 672)]


def nearest_kontext_resolution(aspect_ratio: float):
    """(width, height) of the bucket whose aspect ratio is closest (inplace.py:120-124)."""
    _, w, h = min((abs(aspect_ratio - w / h), w, h) for w, h in KONTEXT_RESOLUTIONS)
    return w, h
