#!/bin/bash
# Round 2, GPU session 16: grouped N-fast tile order for the K-heavy GEMMs (W slice resident in L2): tests, isolated
# throughput with / without, ncu DRAM bytes of FF-down and the single-block proj_out, step times.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest -q -rA -p no:cacheprovider -m gpu tests/test_kernels_gpu.py tests/test_trim_last_gpu.py \
  > gpurun_out/s16_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s16_tests.log; grep -E "passed|failed|FAILED|rc=" gpurun_out/s16_tests.log | tail -6
cat > /tmp/ngroup.py <<'PY'
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
for (M, N, K) in [(8192, 3072, 12288), (8704, 3072, 15360), (4096, 3072, 15360), (8704, 3072, 12288)]:
    a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16(); res = torch.randn(M, N, device="cuda").bfloat16()
    gate = torch.randn(N, device="cuda").bfloat16(); fl = 2.0 * M * N * K
    out = {}
    for ng in (0, -1, 4, 0, -1):
        ops.set_option("n_group", ng)
        t = sustained(lambda: ops.gemm(a, w, b, out=res, epilogue=_lib.EPI_GATE_RES, gate=gate, res=res))
        out.setdefault(ng, []).append(fl / t / 1e9)
    ops.set_option("n_group", -1)
    t = sustained(lambda: torch.matmul(a, w.t(), out=res))
    print(f"M={M} N={N} K={K} gate_res: " + "  ".join(f"n_group={k}: " + "/".join(f"{x:.0f}" for x in v) for k, v in out.items())
          + f"  cublas {fl / t / 1e9:.0f}", flush=True)
PY
timeout 300 python /tmp/ngroup.py > gpurun_out/s16_ngroup_bench.log 2>&1; cat gpurun_out/s16_ngroup_bench.log
for ng in 0 -1; do
  RGE_N_GROUP=$ng timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:gemm2_kernel -c 1 -s 2 --csv python tools/gemm_one.py 8192 3072 12288 gate_res 2>/dev/null | tail -4 | cut -c1-300
  RGE_N_GROUP=$ng timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:gemm2_kernel -c 1 -s 2 --csv python tools/gemm_one.py 8704 3072 15360 gate_res 2>/dev/null | tail -4 | cut -c1-300
done > gpurun_out/s16_ncu_traffic.log 2>&1
cat gpurun_out/s16_ncu_traffic.log | awk -F'","' '{print $5, $(NF-2), $(NF)}' | cut -c1-200
run_steps() { echo "== $1"; env $1 timeout -k 10 200 python tools/step_times.py 2>&1 | tail -4; }
{ run_steps "RGE_N_GROUP=0"; run_steps "RGE_N_GROUP=-1"; } > gpurun_out/s16_step_variants.log 2>&1
grep -v SKIP gpurun_out/s16_step_variants.log
