"""Reference-owned constants of the plugin surface: per-pipeline defaults and the fitted gamma tables.

Defaults: RegionE/tool/RegionE.py:1-7. Gamma (fp16 tensors in the reference): RegionE/<family>/inplace.py:47-50.
"""
from __future__ import annotations

DEFAULTS = {
    "FluxKontextPipeline": {"num_inference_steps": 28, "warmup_step": 6, "post_step": 2, "refresh_step": "16",
                            "threshold": 0.93, "cache_threshold": 0.04, "erosion_dilation": True},
    "Step1XEditPipeline": {"num_inference_steps": 28, "warmup_step": 6, "post_step": 2, "refresh_step": "16",
                           "threshold": 0.88, "cache_threshold": 0.02, "erosion_dilation": True},
    "Step1XEditPipelineV1P2": {"num_inference_steps": 28, "warmup_step": 6, "post_step": 2, "refresh_step": "16",
                               "threshold": 0.88, "cache_threshold": 0.02, "erosion_dilation": True},
    "QwenImageEditPipeline": {"num_inference_steps": 28, "warmup_step": 6, "post_step": 2, "refresh_step": "16",
                              "threshold": 0.80, "cache_threshold": 0.03, "erosion_dilation": True},
    "QwenImageEditPlusPipeline": {"num_inference_steps": 28, "warmup_step": 6, "post_step": 2, "refresh_step": "16",
                                  "threshold": 0.80, "cache_threshold": 0.03, "erosion_dilation": True},
}

GAMMA = {
    "FluxKontextPipeline": [
        0.8352, 0.9986, 1.0090, 1.0097, 1.0161, 1.0152, 1.0160, 1.0173, 1.0177, 1.0199, 1.0213, 1.0203, 1.0257, 1.0236,
        1.0235, 1.0278, 1.0302, 1.0311, 1.0352, 1.0371, 1.0391, 1.0459, 1.0498, 1.0581, 1.0693, 1.0866, 1.1090],
    "Step1XEditPipeline": [
        0.9746, 0.9593, 1.0036, 1.0084, 1.0106, 1.0114, 1.0138, 1.0163, 1.0152, 1.0163, 1.0197, 1.0186, 1.0219, 1.0218,
        1.0223, 1.0266, 1.0272, 1.0305, 1.0311, 1.0362, 1.0385, 1.0423, 1.0500, 1.0536, 1.0671, 1.0866, 1.1015],
    "Step1XEditPipelineV1P2": [
        0.7936, 0.9807, 1.0063, 1.0205, 0.9946, 1.0125, 1.0116, 1.0125, 1.0172, 1.0171, 1.0183, 1.0170, 1.0170, 1.0236,
        1.0263, 1.0264, 1.0277, 1.0321, 1.0338, 1.0361, 1.0396, 1.0454, 1.0492, 1.0566, 1.0696, 1.0879, 1.1179],
    "QwenImageEditPipeline": [
        1.0195, 1.0233, 1.0243, 1.0185, 1.0321, 1.0208, 1.0260, 1.0233, 1.0258, 1.0292, 1.0316, 1.0306, 1.0289, 1.0347,
        1.0329, 1.0402, 1.0378, 1.0384, 1.0413, 1.0444, 1.0526, 1.0400, 1.0555, 1.0439, 1.0357, 1.0118, 0.7603],
    "QwenImageEditPlusPipeline": [
        1.0186, 1.0241, 1.0236, 1.0205, 1.0298, 1.0221, 1.0248, 1.0246, 1.0269, 1.0275, 1.0323, 1.0311, 1.0298, 1.0353,
        1.0343, 1.0397, 1.0387, 1.0393, 1.0404, 1.0458, 1.0507, 1.0418, 1.0518, 1.0426, 1.0311, 1.0068, 0.7628],
}

# `0-dim fp32 CUDA tensor * bf16 tensor` rounds the scalar to bf16 first (TensorIterator casts every operand to the
# common dtype); the reference's dt, dt_final and AVDC ratio all pass through such a product (inplace.py:318, 650,
# 655-680). Measured on the B200 box with tools/scalar_semantics.py.
SCALAR_ROUNDS_TO_BF16 = True


def resample_gamma(table, num_inference_steps: int):
    """A fitted 27-entry (28-step) table linearly resampled to `num_inference_steps - 1` entries over the normalised
    step index (the stand-in for a 20-step run, for which the reference ships no table).
    NOT a fitted table: results with it are unpinned with respect to the reference."""
    n_src, n_dst = len(table), num_inference_steps - 1
    if n_dst == n_src:
        return list(table)
    out = []
    for j in range(n_dst):
        pos = j * (n_src - 1) / max(n_dst - 1, 1)
        lo = min(int(pos), n_src - 2)
        f = pos - lo
        out.append(round((1 - f) * table[lo] + f * table[lo + 1], 4))
    return out
