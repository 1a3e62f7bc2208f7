"""Duck-typed stand-ins for the diffusers objects RegionE patches, plus seeded synthetic weights / inputs.

TEST AND BENCHMARK SCAFFOLDING, not product code: diffusers is not installed in this image and no model weights exist
offline, so tests, bench.py, the tools and the CLI's `--model_path synthetic` mode build pipelines from these classes
(whose class NAMES are what `RegionEHelper` dispatches on). The product package `regione_b200/` does not depend on it.
"""
from . import diffusers_like as standin  # noqa: F401
