"""`RegionEHelper` — the plugin surface of the reference (RegionE/tool/RegionE.py:9-51), same constructor, same
`enable()` / `disable()` / `set_params()` semantics, dispatching on the pipeline's class NAME to the B200 hot path."""
from __future__ import annotations

from .params import DEFAULTS

config = DEFAULTS   # module-level and mutable, like the reference's `config` dict (RegionE.py:1-7, :13)

_FAMILY_MODULE = {
    "FluxKontextPipeline": "flux_kontext",
    "Step1XEditPipeline": "step1x_edit",
    "Step1XEditPipelineV1P2": "step1x_edit_v1p2",
    "QwenImageEditPipeline": "qwen_image_edit",
    "QwenImageEditPlusPipeline": "qwen_image_edit",
}


def _family(name: str):
    import importlib
    try:
        return importlib.import_module(f".{_FAMILY_MODULE[name]}", __package__)
    except ModuleNotFoundError as e:
        raise NotImplementedError(f"regione_b200: the {name} hot path is not built yet") from e


class RegionEHelper(object):
    def __init__(self, pipeline=None):
        if pipeline is not None:
            self.pipeline = pipeline
        self.name = self.pipeline.__class__.__name__
        self.config = config[self.name]          # KeyError for unknown pipelines, as in the reference (:13)

    def enable(self):
        assert self.pipeline is not None
        self.pipeline = _family(self.name).warp_modules(self.pipeline, **self.config)

    def disable(self):
        assert self.pipeline is not None
        self.pipeline = _family(self.name).unwarp_modules(self.pipeline)

    def set_params(self, num_inference_steps=28, warmup_step=None, post_step=None, refresh_step=None, threshold=None,
                   cache_threshold=None, erosion_dilation=None, gamma=None):
        """RegionE.py:43-51. One extension: the reference fixes 28 steps because its gamma tables have 27 fitted
        entries ("Changing the inference step requires fitting a new gamma", utils.py:391). Here the table is an
        input: `gamma` (num_inference_steps - 1 entries) lifts the 28-step restriction (BASELINE configs[2]: Qwen,
        20 steps). Without `gamma` the reference's assertion holds unchanged."""
        if gamma is None:
            assert num_inference_steps == 28, "num_inference_steps must be 28"
            self.config.pop("gamma", None)
            self.config["num_inference_steps"] = 28
        else:
            gamma = [float(g) for g in gamma]
            assert len(gamma) == num_inference_steps - 1, "gamma needs num_inference_steps - 1 entries"
            self.config["gamma"] = gamma
            self.config["num_inference_steps"] = int(num_inference_steps)
        for key, value in (("warmup_step", warmup_step), ("post_step", post_step), ("refresh_step", refresh_step),
                           ("threshold", threshold), ("cache_threshold", cache_threshold),
                           ("erosion_dilation", erosion_dilation)):
            if value is not None:
                self.config[key] = value
        print(f"RegionEHelper: set_params {self.config}")
