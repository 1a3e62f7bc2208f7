"""CPU oracle for the RegionE hot path — TEST INFRASTRUCTURE, not product code.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this
package, and only as the checker or as the timed CPU baseline. The product (`regione_b200/`) never imports it.

What is restated here and how it is pinned:

* region_ops.py  token_selector / morphology / ids_gather / ids_scatter / Manager state machine
                 (RegionE/FluxKontext/utils.py:124-465). PINNED: `oracle/make_golden.py` executed the reference's own
                 `utils.py` in the build container (through a 10-line `diffusers` import stub) and committed the
                 outputs under tests/golden/; tests/test_oracle_golden.py checks this restatement against them.
* schedule.py    sigma / timestep schedule and the adaptive velocity-decay-cache (AVDC) decision rule
                 (RegionE/FluxKontext/inplace.py:229-244, 295-313). The decision rule and the gamma tables are
                 reference-owned and pinned by the golden schedule; `set_timesteps` lives in
                 diffusers (absent here) and is restated from its published algorithm — PARITY UNPINNED for that part.
* flux.py        the FLUX.1-Kontext DiT (diffusers FluxTransformer2DModel, absent from /root/reference and from this
                 image) restated from its published architecture, plus the reference's forward / attention processor
                 (inplace.py:413-576, 694-824) with the pre-norm/pre-RoPE K/V cache semantics and the fp16 round trip
                 of `_partially_linear` (fused_kernels.py:80). PARITY UNPINNED w.r.t. the reference: it owns no test,
                 golden vector or fixture for this path and diffusers cannot be imported. Cross-pinned instead against
                 two independent implementations that do exist in this image: the original black-forest-labs/FLUX
                 blocks vendored by torchtitan (tests/test_oracle_vs_bfl_flux.py: double / single block, rotary
                 table, time embedding agree to 2e-4 in fp32 under diffusers' weight mapping) and, for the attention
                 op, the reference's own third-party kernel flash_attn_func (tests/test_kernels_gpu.py).
* loop.py        the denoising loop and the patched scheduler step (inplace.py:287-392, 594-691).
"""
