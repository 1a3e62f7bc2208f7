#!/bin/bash
# Round 2, GPU session 21: 256-bit stores in the GEMM epilogues (RGE_WIDE_STORE): tests, isolated throughput and step
# times on / off.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest -q -p no:cacheprovider -m gpu tests/test_kernels_gpu.py tests/test_trim_last_gpu.py \
  tests/test_flux_parity_gpu.py tests/test_flux_fullsize_gpu.py tests/test_partially_linear_vs_triton_gpu.py > gpurun_out/s21_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s21_tests.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/s21_tests.log | tail -6
cat > /tmp/ws_bench.py <<'PY'
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from regione_b200 import _lib, ops
def sustained(fn, secs=0.6):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.perf_counter(); e0.record()
    while time.perf_counter() - t0 < secs:
        for _ in range(10): fn()
        n += 10
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, K, epi) in [(8704, 12288, 3072, "gelu"), (8704, 3072, 3072, "store"), (8192, 3072, 3072, "gate_res"), (1576, 12288, 3072, "gelu")]:
    a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16(); out = torch.randn(M, N, device="cuda").bfloat16()
    gate = torch.randn(N, device="cuda").bfloat16(); fl = 2.0 * M * N * K
    kw = dict(epilogue=_lib.EPI_GELU) if epi == "gelu" else dict(epilogue=_lib.EPI_GATE_RES, gate=gate, res=out) if epi == "gate_res" else {}
    r = {}
    for ws in (0, 1, 0, 1):
        ops.set_option("wide_store", ws)
        r.setdefault(ws, []).append(fl / sustained(lambda: ops.gemm(a, w, b, out=out, **kw)) / 1e9)
    print(f"M={M} N={N} K={K} {epi}: " + "  ".join(f"wide_store={k}: " + "/".join(f"{x:.0f}" for x in v) for k, v in r.items()), flush=True)
PY
timeout 200 python /tmp/ws_bench.py > gpurun_out/s21_wide_store_bench.log 2>&1; cat gpurun_out/s21_wide_store_bench.log
run_steps() { echo "== $1"; env $1 timeout -k 10 150 python tools/step_times.py 2>&1 | tail -4; }
{ run_steps "RGE_WIDE_STORE=0"; run_steps "RGE_WIDE_STORE=1"; } > gpurun_out/s21_step_variants.log 2>&1
grep -v SKIP gpurun_out/s21_step_variants.log
