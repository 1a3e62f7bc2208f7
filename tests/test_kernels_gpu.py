"""Kernel-level parity through the C ABI: every CUDA kernel against the oracle / golden fixtures on seeded inputs.
Integer / index / byte results must be bit-exact; floating point within the stated tolerance."""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import flux as of
from oracle import region_ops as ro
from oracle.make_golden import synthetic_partition_inputs

pytestmark = pytest.mark.gpu

BF16_TOL = 4e-3   # one bf16 rounding of the output is 2^-9 relative; two rounding points allowed


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def _gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (77, 64, 3072), (300, 512, 256), (1, 128, 64), (1000, 3072, 3072),
                                   (8704, 768, 3072)])
def test_gemm_store_matches_linear(M, N, K):
    from regione_b200 import ops
    g = _gen(1)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    y = ops.gemm(a, w, b)
    ref = a.float() @ w.float().t() + b.float()
    assert rel_l2(y, ref) <= BF16_TOL
    y2 = ops.gemm(a, w, None)
    assert rel_l2(y2, a.float() @ w.float().t()) <= BF16_TOL


@pytest.mark.parametrize("M,N,K", [(2048, 256, 64), (2049, 512, 192), (4000, 3072, 256), (2304, 768, 3072),
                                   (4100, 512, 12288),    # A > 96 MB -> tiles walk along N (pick_n_fast)
                                   (1576, 3072, 192), (1200, 1536, 320)])   # REGION-sized rows that fill 256-row tiles
def test_gemm_cta_pair_kernel_all_epilogues(M, N, K):
    """M >= 2048, and REGION-sized M with <= 15 % padding to 256-row tiles, dispatch to the cta_group::2 kernel
    (gemm2.cu): ragged M (second CTA partly or fully out of range), every epilogue, scatter."""
    from regione_b200 import _lib, ops
    g = _gen(21)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    lin = F.linear(a, w, b)
    assert rel_l2(ops.gemm(a, w, b), a.float() @ w.float().t() + b.float()) <= BF16_TOL
    assert rel_l2(ops.gemm(a, w, b, epilogue=_lib.EPI_GELU), F.gelu(lin, approximate="tanh")) <= BF16_TOL
    gate = torch.randn(N, device="cuda", generator=g).bfloat16()
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    out = res.clone()
    ops.gemm(a, w, b, epilogue=_lib.EPI_GATE_RES, gate=gate, res=out, out=out)
    assert rel_l2(out, res + gate[None] * lin) <= BF16_TOL
    S = M + 100
    rows = torch.randperm(S, device="cuda", generator=g)[:M]
    cache = torch.zeros(S, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, b, out=cache, row_map=rows.int())
    assert rel_l2(cache[rows], lin) <= BF16_TOL
    if N % 128 == 0:
        H = N // 128
        nw = (1 + 0.1 * torch.randn(128, device="cuda", generator=g)).bfloat16()
        ids = torch.zeros(S, 3, device="cuda")
        ids[:, 1] = torch.arange(S, device="cuda") // 64
        ids[:, 2] = torch.arange(S, device="cuda") % 64
        cs = ops.rope_table(ids)
        cos, sin = of.rope_cos_sin(ids)
        q = ops.gemm(a, w, b, epilogue=_lib.EPI_NORM_ROPE, norm_w=nw, rope_cs=cs, rope_map=rows.int())
        x = of.apply_rope(of.rms_norm(lin.view(1, M, H, 128).transpose(1, 2), nw), (cos[rows], sin[rows]))
        assert rel_l2(q, x.transpose(1, 2).reshape(M, N)) <= BF16_TOL


@pytest.mark.parametrize("bn", [240, 208, 192, 176, 144, 128])
def test_gemm_cta_pair_kernel_tile_widths_bit_identical(bn):
    """The CTA-pair kernel's tile width is a launch parameter (any multiple of 16, chosen per shape to fill the last
    wave of the 74 pairs): ragged M, a ragged last N tile and the 16-column tail chunk of the epilogue. A narrower tile
    changes neither the k-order of any element's sum nor its epilogue: every width is BIT-identical to the 256-wide
    tile, for every epilogue that allows it."""
    from regione_b200 import _lib, ops
    g = _gen(22)
    M, N, K = 2304 + 77, 3072, 192
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    gate = torch.randn(N, device="cuda", generator=g).bfloat16()
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    rows = torch.randperm(M + 100, device="cuda", generator=g)[:M].int()

    def run_all():
        outs = [ops.gemm(a, w, b), ops.gemm(a, w, b, epilogue=_lib.EPI_GELU)]
        o = res.clone()
        ops.gemm(a, w, b, epilogue=_lib.EPI_GATE_RES, gate=gate, res=o, out=o)
        outs.append(o)
        cache = torch.zeros(M + 100, N + 64, device="cuda", dtype=torch.bfloat16)
        ops.gemm(a, w, b, out=cache, row_map=rows, col_off=32)
        outs.append(cache)
        return outs

    try:
        ops.set_option("gemm2_bn", 256)
        ref = run_all()
        ops.set_option("gemm2_bn", bn)
        got = run_all()
    finally:
        ops.set_option("gemm2_bn", 0)
    assert rel_l2(ref[0], a.float() @ w.float().t() + b.float()) <= BF16_TOL
    for x, y in zip(got, ref):
        assert torch.equal(x, y)
    auto = run_all()                                   # the width the host picks for this shape
    for x, y in zip(auto, ref):
        assert torch.equal(x, y)


@pytest.fixture
def one_cta_kernel():
    """Keeps every launch on the 1-CTA kernel (the per-shape rule would send well-filled REGION sizes to the CTA pair)."""
    from regione_b200 import ops
    ops.set_option("2cta_min_m", 0)
    yield
    ops.set_option("2cta_min_m", -1)


# (M, N) chosen so that the 1-CTA kernel's tile-width heuristic (gemm.cu pick_bn, 148 SMs) lands on every width it
# can choose: 1576x3072 -> 160, 512x3072 -> 96, 1064x3072 -> 192, 1576x12288 -> 224, 716x3072 -> 128, 300x64 -> 64,
# 100x3104 (ragged last tile) -> 64, 1900x1536 -> 96/128; K ragged too.
@pytest.mark.parametrize("M,N,K", [(1576, 3072, 320), (512, 3072, 256), (1064, 3072, 192), (1576, 12288, 128),
                                   (716, 3072, 456), (300, 64, 3072), (100, 3104, 64), (1900, 1536, 200)])
def test_gemm_region_step_tile_widths_all_epilogues(M, N, K, one_cta_kernel):
    from regione_b200 import _lib, ops
    g = _gen(31)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    lin = F.linear(a, w, b)
    assert rel_l2(ops.gemm(a, w, b), a.float() @ w.float().t() + b.float()) <= BF16_TOL
    assert rel_l2(ops.gemm(a, w, b, epilogue=_lib.EPI_GELU), F.gelu(lin, approximate="tanh")) <= BF16_TOL
    gate = torch.randn(N, device="cuda", generator=g).bfloat16()
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    out = res.clone()
    ops.gemm(a, w, b, epilogue=_lib.EPI_GATE_RES, gate=gate, res=out, out=out)
    assert rel_l2(out, res + gate[None] * lin) <= BF16_TOL
    # scatter into a wider buffer at a column offset: rows / columns outside the target stay untouched
    S = M + 50
    rows = torch.randperm(S, device="cuda", generator=g)[:M]
    wide = torch.full((S, N + 64), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, b, out=wide, row_map=rows.int(), col_off=32)
    assert rel_l2(wide[rows, 32:32 + N], lin) <= BF16_TOL
    assert bool((wide[:, :32] == 7).all()) and bool((wide[:, 32 + N:] == 7).all())
    keep = torch.ones(S, dtype=torch.bool, device="cuda")
    keep[rows] = False
    assert bool((wide[keep] == 7).all())
    if N % 128 == 0:
        H = N // 128
        nw = (1 + 0.1 * torch.randn(128, device="cuda", generator=g)).bfloat16()
        ids = torch.zeros(S, 3, device="cuda")
        ids[:, 1] = torch.arange(S, device="cuda") // 64
        ids[:, 2] = torch.arange(S, device="cuda") % 64
        cs = ops.rope_table(ids)
        cos, sin = of.rope_cos_sin(ids)
        q = ops.gemm(a, w, b, epilogue=_lib.EPI_NORM_ROPE, norm_w=nw, rope_cs=cs, rope_map=rows.int())
        x = of.apply_rope(of.rms_norm(lin.view(1, M, H, 128).transpose(1, 2), nw), (cos[rows], sin[rows]))
        assert rel_l2(q, x.transpose(1, 2).reshape(M, N)) <= BF16_TOL


def test_gemm_fused_epilogues():
    from regione_b200 import _lib, ops
    g = _gen(2)
    M, N, K = 777, 1024, 512
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    lin = F.linear(a, w, b)
    assert rel_l2(ops.gemm(a, w, b, epilogue=_lib.EPI_GELU), F.gelu(lin, approximate="tanh")) <= BF16_TOL
    gate = torch.randn(N, device="cuda", generator=g).bfloat16()
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    out = res.clone()
    ops.gemm(a, w, b, epilogue=_lib.EPI_GATE_RES, gate=gate, res=out, out=out)      # in place, as the engine does
    assert rel_l2(out, res + gate[None] * lin) <= BF16_TOL


def test_gemm_scatter_is_partially_linear():
    """row_map == the `index` of _partially_linear (fused_kernels.py:77-80): rows land at index[m], others untouched."""
    from regione_b200 import ops
    g = _gen(3)
    M, N, K, S = 333, 256, 512, 900
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    index = torch.randperm(S, device="cuda", generator=g)[:M].sort().values
    cache = torch.randn(1, S, N, device="cuda", generator=g).bfloat16()
    ref = cache.clone()
    of.partially_linear(a[None], w, b, index, ref)                                    # oracle (fp16 round trip)
    ops.gemm(a, w, b, out=cache[0], row_map=index.int())
    assert rel_l2(cache[0, index], ref[0, index]) <= BF16_TOL
    untouched = torch.ones(S, dtype=torch.bool, device="cuda")
    untouched[index] = False
    assert torch.equal(cache[0, untouched], ref[0, untouched])


def test_gemm_norm_rope_epilogue_matches_oracle():
    from regione_b200 import _lib, ops
    g = _gen(4)
    M, H, K, S = 333, 4, 512, 900
    N = H * 128
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    nw = (1 + 0.1 * torch.randn(128, device="cuda", generator=g)).bfloat16()
    ids = torch.zeros(S, 3, device="cuda")
    ids[:, 0] = (torch.arange(S, device="cuda") >= 450).float()
    ids[:, 1] = torch.arange(S, device="cuda") // 30
    ids[:, 2] = torch.arange(S, device="cuda") % 30
    cs = ops.rope_table(ids)
    cos, sin = of.rope_cos_sin(ids)
    assert torch.equal(cs[..., 0].repeat_interleave(2, -1), cos) and torch.equal(cs[..., 1].repeat_interleave(2, -1), sin)
    pos = torch.randperm(S - 7, device="cuda", generator=g)[:M]
    rows = torch.randperm(S, device="cuda", generator=g)[:M]
    cache = torch.zeros(S, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, b, epilogue=_lib.EPI_NORM_ROPE, out=cache, row_map=rows.int(), norm_w=nw, rope_cs=cs,
             rope_map=pos.int(), rope_off=7)
    x = F.linear(a, w, b).view(1, M, H, 128).transpose(1, 2)                          # [1,H,M,128]
    x = of.apply_rope(of.rms_norm(x, nw), (cos[pos + 7], sin[pos + 7]))
    ref = x.transpose(1, 2).reshape(M, N)
    assert rel_l2(cache[rows], ref) <= BF16_TOL
    # the pair-major table layout the engine keeps ([64][ld]: one rotary pair of consecutive rows is contiguous, so an
    # epilogue warp reads it coalesced) must give bit-identical results, on both the 1-CTA and the CTA-pair kernel
    ld = S + 12
    cs_pm = torch.zeros(64, ld, 2, device="cuda")
    cs_pm[:, :S] = cs.permute(1, 0, 2)
    cache_pm = torch.zeros(S, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, b, epilogue=_lib.EPI_NORM_ROPE, out=cache_pm, row_map=rows.int(), norm_w=nw, rope_cs=cs_pm,
             rope_map=pos.int(), rope_off=7, rope_ld=ld)
    assert torch.equal(cache_pm, cache)
    M2 = 2304                                                       # >= 2048 rows: CTA-pair kernel
    a2 = torch.randn(M2, K, device="cuda", generator=g).bfloat16()
    pos2 = torch.randint(0, S - 7, (M2,), device="cuda", generator=g)
    out_rm = ops.gemm(a2, w, b, epilogue=_lib.EPI_NORM_ROPE, norm_w=nw, rope_cs=cs, rope_map=pos2.int(), rope_off=7)
    out_pm = ops.gemm(a2, w, b, epilogue=_lib.EPI_NORM_ROPE, norm_w=nw, rope_cs=cs_pm, rope_map=pos2.int(), rope_off=7,
                      rope_ld=ld)
    assert torch.equal(out_rm, out_pm)
    x2 = F.linear(a2, w, b).view(1, M2, H, 128).transpose(1, 2)
    x2 = of.apply_rope(of.rms_norm(x2, nw), (cos[pos2 + 7], sin[pos2 + 7]))
    assert rel_l2(out_pm, x2.transpose(1, 2).reshape(M2, N)) <= BF16_TOL


@pytest.fixture(params=[0, 1], ids=["attention128", "attention64"])
def attn_kernel(request):
    """Both attention kernels must pass every attention test: attention.cu (128-row K/V tiles, P aliased onto S) and
    attention64.cu (64-row K/V tiles, P in its own TMEM columns, decoupled Q K^T / softmax pipeline)."""
    from regione_b200 import ops
    ops.set_option("attn_kernel", request.param)
    yield request.param
    ops.set_option("attn_kernel", -1)


@pytest.mark.parametrize("Sq,Skv,H", [(256, 256, 1), (128, 128, 2), (1, 130, 1), (200, 544, 2), (700, 1300, 3),
                                      (2048, 8704, 2), (300, 64, 1), (129, 65, 2), (513, 8704, 1)])
def test_attention_matches_exact_softmax(Sq, Skv, H, attn_kernel):
    from regione_b200 import ops
    g = _gen(5)
    q = torch.randn(Sq, H * 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
    k[Skv // 2:] *= 3.0                                            # late large keys: exercises the lazy O rescale
    o = ops.attention(q, k, v, H)
    hd = lambda t: t.view(1, -1, H, 128).transpose(1, 2)          # noqa: E731
    ref = of.exact_attention(hd(q), hd(k), hd(v))[0]
    assert rel_l2(o, ref) <= 6e-3                                  # bf16 P and bf16 output


def test_attention_lazy_rescale_paths(attn_kernel):
    """Late large keys (the running maximum jumps by far more than 2^8 in the middle of the key sequence, and again
    inside a tile), peaked query rows and a ragged KV tail: the lazy O-rescale path of both kernels."""
    from regione_b200 import ops
    g = _gen(23)
    Sq, Skv, H = 777, 2100, 3
    q = torch.randn(Sq, H * 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
    k[Skv // 2:] *= 3.0
    k[Skv // 2 + 70: Skv // 2 + 90] *= 4.0
    q[:64] *= 6.0
    hd = lambda t: t.view(1, -1, H, 128).transpose(1, 2)          # noqa: E731
    ref = of.exact_attention(hd(q), hd(k), hd(v))[0]
    o = ops.attention(q, k, v, H)
    torch.cuda.synchronize()
    assert torch.isfinite(o.float()).all()
    assert rel_l2(o, ref) <= 6e-3


@pytest.mark.parametrize("poly", [0, 2, 3, 4])
def test_attention_exponential_offload_variants(poly):
    """`attn_poly` of every 8 exponential pairs run as a Cody-Waite / degree-3 polynomial on the FMA pipe instead of
    MUFU.EX2 (relative error 7.5e-5, below the bf16 rounding of P): every variant stays within the same tolerance of
    the exact softmax, including rows with large late keys (lazy rescale) and a ragged KV tail (masked columns)."""
    from regione_b200 import ops
    g = _gen(21)
    Sq, Skv, H = 777, 2100, 3
    q = torch.randn(Sq, H * 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
    k[Skv // 2:] *= 3.0
    q[:64] *= 6.0                                                   # peaked rows: scores far below the row maximum
    hd = lambda t: t.view(1, -1, H, 128).transpose(1, 2)          # noqa: E731
    k[Skv // 2 + 70: Skv // 2 + 90] *= 4.0                          # a second jump inside a tile
    ref = of.exact_attention(hd(q), hd(k), hd(v))[0]
    ops.set_option("attn_poly", poly)
    try:
        o = ops.attention(q, k, v, H)
        torch.cuda.synchronize()
    finally:
        ops.set_option("attn_poly", -1)
    assert torch.isfinite(o.float()).all()
    assert rel_l2(o, ref) <= 6e-3


@pytest.mark.parametrize("Sq,Skv,H", [(1576, 8704, 4), (8704, 8704, 2), (333, 1000, 3)])
def test_attention_matches_the_reference_s_flash_attn_func(Sq, Skv, H, attn_kernel):
    """The reference calls flash-attn's `flash_attn_func(q, k, v, causal=False)` at inplace.py:796-801 (README pins
    flash-attn v2.8.2; this image has 2.8.3): the same call, on the same inputs, is the checker here — a pin of the
    attention op against the reference's own third-party kernel rather than against the restated oracle."""
    flash_attn = pytest.importorskip("flash_attn")
    from regione_b200 import ops
    g = _gen(9)
    q = torch.randn(Sq, H * 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
    o = ops.attention(q, k, v, H)
    ref = flash_attn.flash_attn_func(q.view(1, Sq, H, 128), k.view(1, Skv, H, 128), v.view(1, Skv, H, 128),
                                     causal=False).reshape(Sq, H * 128)
    assert rel_l2(o, ref) <= 6e-3


@pytest.mark.parametrize("Sq,Skv,H", [(1576, 8704, 24), (1600, 4000, 24), (1700, 2100, 24), (1537, 1100, 24),
                                      (1200, 1300, 30), (2048, 1300, 20), (4864, 4864, 24)])
def test_attention_kv_split_of_the_trailing_work_units(Sq, Skv, H):
    """With a workspace, the (256-row query tile, head) units that would form the last, mostly empty wave of CTAs on the
    148 SMs are cut along K/V (un-normalised fp32 partials, merge kernel) - REGION steps: 512 + 1064 rows x 24 heads =
    148 + 20 units, the ragged 40-row tiles last; (2048, ., 20) and (4864, ., 24): full tiles get split. Applies when the
    trailing units are at most a quarter of a wave (measured: wider splits do not pay).
    Same tolerance against the exact softmax as the plain kernel and close to it element-wise; late large keys
    exercise the merge of parts with different reference maxima, a ragged tile above 128 rows the two-group split
    CTAs; `attn_split` = 0 gives the plain kernel back bit for bit."""
    from regione_b200 import ops
    g = _gen(41)
    q = torch.randn(Sq, H * 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
    k[Skv // 2:] *= 3.0
    plain = ops.attention(q, k, v, H)
    ws = ops.attention_workspace(H)
    wide = torch.full((Sq, H * 128 + 256), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.attention(q, k, v, H, out=wide, workspace=ws)
    torch.cuda.synchronize()
    got = wide[:, : H * 128]
    assert bool((wide[:, H * 128:] == 7.0).all())
    assert torch.isfinite(got.float()).all()
    assert not torch.equal(got, plain)                              # the split path did run
    assert rel_l2(got, plain) <= 6e-3                               # two roundings of the same softmax
    heads = slice((H - 2) * 128, H * 128)                           # exact reference on the last two heads (split ones)
    hd = lambda t: t[:, heads].reshape(1, -1, 2, 128).transpose(1, 2)   # noqa: E731
    ref = of.exact_attention(hd(q), hd(k), hd(v))[0]
    assert rel_l2(got[:, heads], ref) <= 6e-3
    ops.set_option("attn_split", 0)
    try:
        off = ops.attention(q, k, v, H, workspace=ws)
    finally:
        ops.set_option("attn_split", 1)
    assert torch.equal(off, plain)


def test_attention_strided_output_and_row_independence(attn_kernel):
    """Output written into a wider buffer (the engine's [S, D + 4D] layout); untouched columns stay untouched and
    each query row depends only on its own query (permutation equivariance)."""
    from regione_b200 import ops
    g = _gen(6)
    Sq, Skv, H = 300, 700, 2
    q = torch.randn(Sq, H * 128, device="cuda", generator=g).bfloat16()
    k = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
    v = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
    wide = torch.full((Sq, H * 128 + 512), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.attention(q, k, v, H, out=wide)
    assert bool((wide[:, H * 128:] == 7.0).all())
    perm = torch.randperm(Sq, device="cuda", generator=g)
    o2 = ops.attention(q[perm].contiguous(), k, v, H)
    assert torch.equal(o2, wide[perm, : H * 128])


def test_ln_modulate_matches_oracle():
    from regione_b200 import ops
    g = _gen(7)
    for M, D in [(517, 3072), (3, 256), (64, 768), (5, 2048), (9, 4096), (1, 3072)]:
        x = torch.randn(M, D, device="cuda", generator=g).bfloat16()
        sc = (0.1 * torch.randn(D, device="cuda", generator=g)).bfloat16()
        sh = (0.1 * torch.randn(D, device="cuda", generator=g)).bfloat16()
        ref = F.layer_norm(x, (D,), eps=1e-6) * (1 + sc[None]) + sh[None]
        y = ops.ln_modulate(x, sc, sh)
        assert rel_l2(y, ref) <= 1e-3
        assert float((y == ref).float().mean()) > 0.99            # rounding points mirrored: almost all bits equal


def test_euler_and_reuse_bit_exact():
    """x' = bf16(float(x) + bf16(bf16(dt) * v)) — the CUDA semantics of the reference's `sample + dt * model_output`
    (inplace.py:680) and of `cache * ratio` (:318); bit-exact against torch on the same device."""
    from regione_b200 import ops
    g = _gen(8)
    x = torch.randn(4096, 64, device="cuda", generator=g).bfloat16()
    v = torch.randn(4096, 64, device="cuda", generator=g).bfloat16()
    dt = torch.tensor(-0.0371, device="cuda")
    dd = torch.tensor(-0.2113, device="cuda")
    ref = (x.float() + dt * v).bfloat16()
    assert torch.equal(ops.euler(x, v, float(dt.bfloat16())), ref)
    mask = (torch.rand(4096, device="cuda", generator=g) < 0.3).to(torch.uint8)
    ref2 = torch.where(mask.bool()[:, None], x.float() + dt * v, x.float() + dd * v).bfloat16()
    assert torch.equal(ops.euler(x, v, float(dt.bfloat16()), float(dd.bfloat16()), edited_mask=mask), ref2)
    ratio = torch.tensor(0.99263, device="cuda")
    ref3 = (x.float() + dt * (v * ratio)).bfloat16()
    assert torch.equal(ops.euler(x, v, float(dt.bfloat16()), reuse_ratio=float(ratio.bfloat16())), ref3)


def test_gather_scatter_bit_exact(golden_dir):
    from regione_b200 import ops
    g = torch.load(os.path.join(golden_dir, "region_ops.pt"), weights_only=False)["gather"]
    lat, ids = g["latent"][0].cuda(), g["ids"][0].cuda()
    got = ops.gather_rows(lat, ids)
    assert torch.equal(got.cpu(), g["gathered"][0])
    dst = torch.zeros_like(lat)
    ops.scatter_rows(got, ids, dst)
    assert torch.equal(dst.cpu(), g["scattered"][0])
    empty = ops.gather_rows(lat, ids[:0])
    assert empty.shape == (0, lat.shape[1])
    ops.scatter_rows(empty, ids[:0], dst)


def test_partition_masks_bit_exact_against_reference_fixtures(golden_dir):
    """rge_partition + rge_compact against ids produced by the reference's own token_selector."""
    from regione_b200 import ops
    cases = torch.load(os.path.join(golden_dir, "region_ops.pt"), weights_only=False)["selector"]
    for c in cases:
        if "estimate" in c:
            est, cond = c["estimate"], c["condition"]
        else:
            est, cond = synthetic_partition_inputs(c["seed"], c["gh"], c["gw"], c["frac"])
        # the kernel forms the estimate itself as x + bf16(dt_final * v): feed v = 0 and x = bf16(estimate);
        # the oracle is evaluated on the same bf16-rounded estimate so both see identical inputs
        x = est[0].bfloat16()
        e_ref, u_ref, raw_ref, _, sim_ref = ro.select_tokens(x.float()[None], cond, c["threshold"], c["gh"], c["gw"],
                                                             c["erosion_dilation"])
        raw, sim = ops.partition(x.cuda(), torch.zeros_like(x).cuda(), cond[0].cuda(), -1.0, c["threshold"],
                                 want_sim=True)
        margin = float((sim_ref - c["threshold"]).abs().min())
        assert float((sim.cpu() - sim_ref[0]).abs().max()) < 1e-5
        assert torch.equal(raw.cpu().bool(), raw_ref[0]), f"raw mask differs (min |sim-thr| = {margin:.2e})"
        _, ed, un = ops.compact(raw, c["gh"], c["gw"], c["erosion_dilation"])
        assert torch.equal(ed.cpu(), e_ref[0].to(torch.int32)) and torch.equal(un.cpu(), u_ref[0].to(torch.int32))
        if float((est - x.float()[None]).abs().max()) == 0 or margin > 1e-2:
            assert torch.equal(ed.cpu(), c["edited"][0]), "differs from the reference's token_selector output"


def test_morphology_bit_exact_against_reference_fixtures(golden_dir):
    from regione_b200 import ops
    for m in torch.load(os.path.join(golden_dir, "region_ops.pt"), weights_only=False)["morphology"]:
        gh, gw = m["mask"].shape
        out, ed, un = ops.compact(m["mask"].flatten().cuda(), gh, gw, True)
        assert torch.equal(out.cpu().view(gh, gw), m["out"])
        assert ed.numel() == int(m["out"].sum()) and ed.numel() + un.numel() == gh * gw


def test_full_size_properties():
    """BASELINE full size (L = 4096, 64 channels): idempotence / partition properties that need no oracle."""
    from regione_b200 import ops
    g = _gen(9)
    L = 4096
    x = torch.randn(L, 64, device="cuda", generator=g).bfloat16()
    raw = (torch.rand(L, device="cuda", generator=g) < 0.55).to(torch.uint8)
    out, ed, un = ops.compact(raw, 64, 64, True)
    allids = torch.cat([ed, un]).sort().values
    assert torch.equal(allids, torch.arange(L, device="cuda", dtype=torch.int32))       # a partition of the tokens
    assert bool((ed[1:] > ed[:-1]).all()) and bool((un[1:] > un[:-1]).all())              # ascending
    merged = torch.zeros_like(x)
    ops.scatter_rows(ops.gather_rows(x, ed), ed, merged)
    ops.scatter_rows(ops.gather_rows(x, un), un, merged)
    assert torch.equal(merged, x)                                                         # split + merge = identity
    out2, ed2, _ = ops.compact(out, 64, 64, False)
    assert torch.equal(out2, out) and torch.equal(ed2, ed)                                # compaction is idempotent
    assert torch.equal(ops.euler(x, x, 0.0), x)                                           # dt = 0 is the identity


@pytest.mark.parametrize("B,C,H,W", [(1, 16, 128, 128), (2, 16, 100, 162), (1, 4, 2, 2), (3, 16, 96, 96)])
def test_pack_unpack_latents_bit_exact(B, C, H, W):
    """diffusers _pack_latents / _unpack_latents (view / permute / reshape) restated in torch as the checker."""
    from regione_b200 import ops
    x = torch.randn(B, C, H, W, device="cuda", generator=_gen(41)).bfloat16()
    ref = x.view(B, C, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(B, (H // 2) * (W // 2), C * 4)
    got = ops.pack_latents(x)
    assert torch.equal(got, ref)
    back = ops.unpack_latents(got, H * 8, W * 8, 8)
    ref_back = ref.view(B, H // 2, W // 2, C, 2, 2).permute(0, 3, 1, 4, 2, 5).reshape(B, C, H, W)
    assert torch.equal(back, ref_back) and torch.equal(back, x)
    from standins.diffusers_like import FluxKontextPipeline
    assert torch.equal(FluxKontextPipeline._unpack_latents(FluxKontextPipeline._pack_latents(x), H * 8, W * 8, 8), x)


def test_gemm_group_equals_single_launches():
    """rge_op_gemm_group: the q / k / v (/ MLP-up) projections of a block as one persistent launch. Same tiles, same
    K order -> bit-identical to the members launched one by one; mixed epilogues, shapes, K, scatter, an empty member."""
    from regione_b200 import _lib, ops
    g = _gen(51)
    D, Dm, S, T, M = 512, 1024, 2600, 200, 1064

    def lin(n, k):
        return ((torch.randn(n, k, device="cuda", generator=g) * 0.05).bfloat16(),
                torch.randn(n, device="cuda", generator=g).bfloat16())

    x_img = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    x_txt = torch.randn(T, D, device="cuda", generator=g).bfloat16()
    sel = (torch.randperm(S - T, device="cuda", generator=g)[:M].sort().values + T).int()
    nw = (1 + 0.1 * torch.randn(128, device="cuda", generator=g)).bfloat16()
    ids = torch.zeros(S, 3, device="cuda")
    ids[:, 1] = torch.arange(S, device="cuda") // 50
    ids[:, 2] = torch.arange(S, device="cuda") % 50
    cs = ops.rope_table(ids)
    (wq, bq), (wk, bk), (wv, bv), (wm, bm) = lin(D, D), lin(D, D), lin(D, D), lin(Dm, D)
    (wtq, btq), (wdown, bdown) = lin(D, D), lin(D, Dm)
    gate = torch.randn(D, device="cuda", generator=g).bfloat16()
    hid = torch.randn(M, Dm, device="cuda", generator=g).bfloat16()
    res0 = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    empty = torch.empty(0, D, device="cuda", dtype=torch.bfloat16)

    def members(q, kc, vc, big, res, qe):
        return [
            (x_img, wq, bq, dict(epilogue=_lib.EPI_NORM_ROPE, out=q, row_off=T, norm_w=nw, rope_cs=cs, rope_map=sel)),
            (x_img, wk, bk, dict(epilogue=_lib.EPI_NORM_ROPE, out=kc, row_map=sel, norm_w=nw, rope_cs=cs, rope_map=sel)),
            (x_img, wv, bv, dict(out=vc, row_map=sel)),
            (x_txt, wtq, btq, dict(epilogue=_lib.EPI_NORM_ROPE, out=q, norm_w=nw, rope_cs=cs)),
            (x_img, wm, bm, dict(epilogue=_lib.EPI_GELU, out=big, col_off=D)),
            (hid, wdown, bdown, dict(epilogue=_lib.EPI_GATE_RES, out=res, gate=gate, res=res)),
        ], (empty, wq, bq, dict(out=qe))

    def buffers():
        return (torch.zeros(T + M, D, device="cuda", dtype=torch.bfloat16),
                torch.zeros(S, D, device="cuda", dtype=torch.bfloat16),
                torch.zeros(S, D, device="cuda", dtype=torch.bfloat16),
                torch.zeros(M, D + Dm, device="cuda", dtype=torch.bfloat16), res0.clone(),
                torch.empty(0, D, device="cuda", dtype=torch.bfloat16))

    ref = buffers()
    mem, emp = members(*ref)
    for a, w, b, kw in mem:
        ops.gemm(a, w, b, **kw)
    got = buffers()
    mem, emp = members(*got)
    ops.gemm_group(mem[:3] + [emp] + [mem[3]])          # 5 descriptors, one of them empty
    ops.gemm_group([mem[4], mem[5]])                    # different K (512 / 1024) and tile widths in one launch
    one = torch.zeros(S, D, device="cuda", dtype=torch.bfloat16)
    ops.gemm_group([(x_img, wv, bv, dict(out=one, row_map=sel))])   # a group of one falls through to the plain launch
    assert torch.equal(one, ref[2])
    torch.cuda.synchronize()
    for r, o in zip(ref[:5], got[:5]):
        assert torch.equal(r, o)
    assert rel_l2(got[2][sel.long()], F.linear(x_img, wv, bv)) <= BF16_TOL
    assert rel_l2(got[3][:, D:], F.gelu(F.linear(x_img, wm, bm), approximate="tanh")) <= BF16_TOL
    assert rel_l2(got[4], res0 + gate[None] * F.linear(hid, wdown, bdown)) <= BF16_TOL
