"""Differential test of the scatter-GEMM against the REFERENCE'S OWN Triton kernel `_partially_linear`
(RegionE/FluxKontext/fused_kernels.py:9-101), staged unmodified under oracle/_ref/ by oracle/build_ref.py and JIT-compiled
by the image's triton on the box. This is the one device kernel the reference ships (SURVEY §2.2 K-a); everything else
on its path is diffusers / flash-attn / cuBLAS.

Shapes are the reference's call sites at FLUX 1024^2 (inplace.py:734-747): A [1, M, 3072] with M = N_e (double blocks)
or 512 + N_e (single blocks), W [3072, 3072], out = the persistent cache [1, 8192 | 8704, 3072], index = the selection.

Two epilogues:
  * RGE_GEMM_FP16_ROUNDTRIP: fp32 -> fp16 -> bf16 like the Triton store (`accumulator.to(tl.float16)` into a bf16
    buffer, :80). Both kernels accumulate in fp32 but in different K orders (tl.dot blocks of 64 vs tcgen05 K = 16 steps
    into TMEM), so a handful of sums land on the other side of a rounding boundary: >= 99.5 % of the elements must be
    bit-identical and no element may differ by more than one bf16 ulp.
  * direct fp32 -> bf16 (what the engine stores): differs from the Triton result only where the double rounding bites
    - an fp16 value whose three dropped mantissa bits are exactly 100b is a tie for the second rounding (1 in 8), which
    ties-to-even resolves against the sign of the first rounding's error half of the time: ~6 % of the elements, each
    by one bf16 ulp. Gate: >= 92 % bit-identical; at most one ulp wherever the value is a NORMAL fp16 number
    (|x| >= 2^-14); below that fp16 is subnormal (absolute spacing 2^-24), where the Triton result carries an absolute
    error of up to 2^-25 that spans several bf16 ulps of such tiny values - bounded absolutely there (2^-22: the fp16
    rounding plus the two bf16 roundings of values just under 2^-14).
Rows outside the index must stay untouched by both kernels (bit-exact)."""
import pytest
import torch

from oracle.build_ref import load_partially_linear

pytestmark = pytest.mark.gpu


def ulp_diff(a, b):
    """distance in bf16 ulps between two bf16 tensors (monotone integer mapping of the bit patterns)."""
    def key(x):
        i = x.view(torch.int16).to(torch.int32)
        return torch.where(i < 0, -(i & 0x7FFF), i)
    return (key(a) - key(b)).abs()


@pytest.mark.parametrize("M,S_cache,single", [(1064, 8192, False), (1576, 8704, True), (360, 8192, False)])
def test_scatter_gemm_matches_the_reference_triton_kernel(M, S_cache, single):
    from regione_b200 import ops
    pl = load_partially_linear()
    if pl is None:
        pytest.fail("oracle/_ref/fused_kernels.py is not staged: run `python -m oracle.build_ref` where /root/reference "
                    "exists (build() does) before snapshotting to the GPU box")
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(M)
    D = 3072
    a = torch.randn(1, M, D, device=dev, generator=g).bfloat16()
    w = (0.02 * torch.randn(D, D, device=dev, generator=g)).bfloat16()
    b = (0.01 * torch.randn(D, device=dev, generator=g)).bfloat16()
    n_txt = 512 if single else 0
    ed = torch.randperm(S_cache - n_txt, device=dev, generator=g)[: M - n_txt].sort().values
    index = torch.cat((torch.arange(n_txt, device=dev), ed + n_txt)) if single else ed           # inplace.py:729-732
    base = torch.randn(1, S_cache, D, device=dev, generator=g).bfloat16()

    ref = base.clone()
    pl(a, w, b, index, ref)                                   # the reference's kernel, in place, returns None
    torch.cuda.synchronize()

    got16 = base.clone()
    ops.gemm(a[0], w, b, out=got16[0], row_map=index.int(), fp16_roundtrip=True)
    got = base.clone()
    ops.gemm(a[0], w, b, out=got[0], row_map=index.int())
    torch.cuda.synchronize()

    untouched = torch.ones(S_cache, dtype=torch.bool, device=dev)
    untouched[index] = False
    for name, t in (("triton", ref), ("fp16-roundtrip", got16), ("direct", got)):
        assert torch.equal(t[0, untouched], base[0, untouched]), f"{name}: rows outside the index were modified"

    r, x16, x = ref[0, index], got16[0, index], got[0, index]
    assert torch.isfinite(r.float()).all()
    d16, d = ulp_diff(x16, r), ulp_diff(x, r)
    same16, same = float((d16 == 0).float().mean()), float((d == 0).float().mean())
    print(f"M={M} S={S_cache}: fp16-roundtrip epilogue bit-identical to Triton on {100 * same16:.3f} % (max {int(d16.max())} ulp); "
          f"direct fp32->bf16 epilogue on {100 * same:.3f} % (max {int(d.max())} ulp, fp16-subnormal values included)")
    assert int(d16.max()) <= 1 and same16 >= 0.995
    normal = r.float().abs() >= 2.0 ** -14
    assert same >= 0.92 and int(d[normal].max()) <= 1
    if bool((~normal).any()):
        assert float((x.float() - r.float())[~normal].abs().max()) <= 2.0 ** -22
    # and both are the same linear map as torch (fp32 reference of the op)
    want = (a[0].float() @ w.float().t() + b.float())
    assert float((x.float() - want).norm() / want.norm()) < 3e-3
