"""regione_b200 — B200-native implementation of RegionE's region-aware denoising hot path."""
__version__ = "0.1.0"
