#!/bin/bash
# Round 2, GPU session 18: split of the trailing tile wave of K-heavy CTA-pair GEMMs along K (side-stream launch into
# fp32 slabs + finish kernel): tests, isolated throughput on / off, step times on / off.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -rA -p no:cacheprovider -m gpu tests/test_kernels_gpu.py tests/test_trim_last_gpu.py \
  tests/test_flux_parity_gpu.py tests/test_flux_fullsize_gpu.py > gpurun_out/s18_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s18_tests.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/s18_tests.log | tail -8
cat > /tmp/split_bench.py <<'PY'
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from regione_b200 import _lib, ops
def sustained(fn, secs=0.8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n, t0 = 0, time.perf_counter(); e0.record()
    while time.perf_counter() - t0 < secs:
        for _ in range(10): fn()
        n += 10
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for (M, N, K) in [(8192, 3072, 12288), (8704, 3072, 15360), (4096, 3072, 15360), (8704, 3072, 12288), (1576, 3072, 15360),
                  (512, 3072, 12288), (8192, 3072, 3072)]:
    a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16(); res = torch.randn(M, N, device="cuda").bfloat16()
    gate = torch.randn(N, device="cuda").bfloat16(); fl = 2.0 * M * N * K
    out = {}
    for sp in (0, 1, 0, 1):
        ops.set_option("split_tail", sp)
        t = sustained(lambda: ops.gemm(a, w, b, out=res, epilogue=_lib.EPI_GATE_RES, gate=gate, res=res))
        out.setdefault(sp, []).append(fl / t / 1e9)
    ops.set_option("split_tail", 1)
    t = sustained(lambda: torch.matmul(a, w.t(), out=res))
    print(f"M={M} N={N} K={K} gate_res: " + "  ".join(f"split_tail={k}: " + "/".join(f"{x:.0f}" for x in v) for k, v in out.items())
          + f"  cublas {fl / t / 1e9:.0f}", flush=True)
PY
timeout 300 python /tmp/split_bench.py > gpurun_out/s18_split_bench.log 2>&1; cat gpurun_out/s18_split_bench.log
run_steps() { echo "== $1"; env $1 timeout -k 10 200 python tools/step_times.py 2>&1 | tail -4; }
{ run_steps "RGE_SPLIT_TAIL=0"; run_steps "RGE_SPLIT_TAIL=1"; run_steps "RGE_SPLIT_TAIL=0 RGE_NOP=1"; run_steps "RGE_SPLIT_TAIL=1 RGE_NOP=1"; } > gpurun_out/s18_step_variants.log 2>&1
grep -v SKIP gpurun_out/s18_step_variants.log
