"""regione_b200 — B200-native implementation of RegionE's region-aware denoising hot path.

    from regione_b200 import RegionEHelper          # drop-in for `from RegionE import RegionEHelper`
    helper = RegionEHelper(pipeline); helper.set_params(...); helper.enable()
"""
from .helper import RegionEHelper  # noqa: F401

__version__ = "0.1.0"
__all__ = ["RegionEHelper"]
