"""Bring-up probe for the sm_100a kernels: runs every case in its own subprocess (a trap or hang in one kernel must
not poison the others) and prints one line per case. Usage on the GPU box:

    python tools/gpu_probe.py            # all cases
    python tools/gpu_probe.py gemm_store # one case, in-process
"""
from __future__ import annotations

import json
import math
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rel_l2(a, b):
    a = a.float(); b = b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-20))


def case_gemm_store():
    import torch
    from regione_b200 import ops
    out = {}
    for (M, N, K) in [(128, 256, 64), (128, 128, 128), (300, 512, 256), (1000, 3072, 3072), (77, 64, 3072),
                      (8704, 3072, 3072)]:
        g = torch.Generator(device="cuda").manual_seed(1)
        a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
        w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
        b = torch.randn(N, device="cuda", generator=g).bfloat16()
        y = ops.gemm(a, w, b)
        torch.cuda.synchronize()
        ref = (a.float() @ w.float().t() + b.float())
        out[f"{M}x{N}x{K}"] = rel_l2(y, ref)
    return out


def case_gemm_epilogues():
    import torch
    import torch.nn.functional as F
    from regione_b200 import _lib, ops
    out = {}
    g = torch.Generator(device="cuda").manual_seed(2)
    M, N, K = 777, 1024, 512
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    lin = (a.float() @ w.float().t() + b.float()).bfloat16()
    y = ops.gemm(a, w, b, epilogue=_lib.EPI_GELU)
    out["gelu"] = rel_l2(y, F.gelu(lin.float(), approximate="tanh"))
    gate = torch.randn(N, device="cuda", generator=g).bfloat16()
    res = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    y = ops.gemm(a, w, b, epilogue=_lib.EPI_GATE_RES, gate=gate, res=res, out=res.clone())
    out["gate_res"] = rel_l2(y, res.float() + (gate.float() * lin.float()).bfloat16().float())
    # scatter + column offset into a wider buffer
    perm = torch.randperm(2000, device="cuda", generator=g)[:M].int()
    wide = torch.zeros(2000 + 5, N + 64, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, b, out=wide, row_map=perm, row_off=5, col_off=64)
    out["scatter"] = rel_l2(wide[perm.long() + 5, 64:], lin)
    out["scatter_untouched"] = float(wide[:, :64].abs().sum())
    torch.cuda.synchronize()
    return out


def _rope_ref(x, cs):  # x [M, H, 128] fp32-able, cs [M, 64, 2]
    import torch
    cos = cs[..., 0].repeat_interleave(2, dim=-1)[:, None, :]
    sin = cs[..., 1].repeat_interleave(2, dim=-1)[:, None, :]
    xr = x.float().reshape(*x.shape[:-1], -1, 2)
    rot = torch.stack([-xr[..., 1], xr[..., 0]], dim=-1).flatten(-2)
    return (x.float() * cos + rot * sin).to(x.dtype)


def case_gemm_norm_rope():
    import torch
    from regione_b200 import _lib, ops
    out = {}
    g = torch.Generator(device="cuda").manual_seed(3)
    M, H, K, S = 333, 4, 512, 900
    N = H * 128
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    b = torch.randn(N, device="cuda", generator=g).bfloat16()
    nw = (1 + 0.1 * torch.randn(128, device="cuda", generator=g)).bfloat16()
    ids = torch.zeros(S, 3, device="cuda")
    ids[:, 1] = torch.arange(S, device="cuda") // 30
    ids[:, 2] = torch.arange(S, device="cuda") % 30
    cs = ops.rope_table(ids)
    pos = torch.randperm(S - 7, device="cuda", generator=g)[:M].int()
    rows = torch.randperm(S, device="cuda", generator=g)[:M].int()
    cache = torch.zeros(S, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, b, epilogue=_lib.EPI_NORM_ROPE, out=cache, row_map=rows, norm_w=nw, rope_cs=cs, rope_map=pos,
             rope_off=7)
    torch.cuda.synchronize()
    lin = (a.float() @ w.float().t() + b.float()).bfloat16().view(M, H, 128)
    var = lin.float().pow(2).mean(-1, keepdim=True)
    nrm = (lin.float() * torch.rsqrt(var + 1e-6)).bfloat16() * nw
    ref = _rope_ref(nrm, cs[pos.long() + 7]).reshape(M, N)
    out["norm_rope"] = rel_l2(cache[rows.long()], ref)
    # rope table itself vs float64 reference
    pos64 = ids.double()
    parts = []
    for ax, dim in enumerate((16, 56, 56)):
        freqs = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float64, device="cuda") / dim))
        parts.append(pos64[:, ax:ax + 1] * freqs[None])
    ang = torch.cat(parts, dim=1)
    out["rope_table_maxabs"] = float((cs[..., 0] - ang.cos().float()).abs().max() +
                                     (cs[..., 1] - ang.sin().float()).abs().max())
    return out


def case_attention():
    import torch
    from regione_b200 import ops
    out = {}
    for (Sq, Skv, H) in [(256, 256, 1), (128, 128, 2), (200, 544, 2), (700, 1300, 3), (2048, 8704, 2)]:
        g = torch.Generator(device="cuda").manual_seed(4)
        q = torch.randn(Sq, H * 128, device="cuda", generator=g).bfloat16()
        k = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
        v = torch.randn(Skv, H * 128, device="cuda", generator=g).bfloat16()
        # a few large-magnitude keys late in the sequence force the lazy-rescale path
        k[Skv // 2:, :] *= 3.0
        o = ops.attention(q, k, v, H)
        torch.cuda.synchronize()
        qh = q.view(Sq, H, 128).transpose(0, 1).float()
        kh = k.view(Skv, H, 128).transpose(0, 1).float()
        vh = v.view(Skv, H, 128).transpose(0, 1).float()
        p = torch.softmax(qh @ kh.transpose(1, 2) * 128 ** -0.5, dim=-1)
        ref = (p @ vh).transpose(0, 1).reshape(Sq, H * 128)
        out[f"{Sq}x{Skv}x{H}"] = rel_l2(o, ref)
    return out


def case_elementwise():
    import torch
    import torch.nn.functional as F
    from regione_b200 import ops
    out = {}
    g = torch.Generator(device="cuda").manual_seed(5)
    M, Dm = 517, 3072
    x = torch.randn(M, Dm, device="cuda", generator=g).bfloat16()
    sc = (0.1 * torch.randn(Dm, device="cuda", generator=g)).bfloat16()
    sh = (0.1 * torch.randn(Dm, device="cuda", generator=g)).bfloat16()
    y = ops.ln_modulate(x, sc, sh)
    ref = F.layer_norm(x, (Dm,), eps=1e-6) * (1 + sc) + sh
    out["ln_modulate"] = rel_l2(y, ref)
    out["ln_modulate_exact_frac"] = float((y == ref).float().mean())
    L = 4096
    xs = torch.randn(L, 64, device="cuda", generator=g).bfloat16()
    vs = torch.randn(L, 64, device="cuda", generator=g).bfloat16()
    dt = torch.tensor(-0.0371, device="cuda")
    y = ops.euler(xs, vs, float(dt))
    ref = (xs.float() + dt * vs).bfloat16()
    out["euler_exact"] = bool(torch.equal(y, ref))
    ids = torch.randperm(L, device="cuda", generator=g)[:1000].int()
    gth = ops.gather_rows(xs, ids)
    out["gather_exact"] = bool(torch.equal(gth, xs[ids.long()]))
    dst = torch.zeros_like(xs)
    ops.scatter_rows(gth, ids, dst)
    out["scatter_exact"] = bool(torch.equal(dst[ids.long()], gth))
    cond = (xs.float() * 0.6 + 0.5 * torch.randn(L, 64, device="cuda", generator=g)).bfloat16()
    mask, sim = ops.partition(xs, vs, cond, -0.9, 0.5, want_sim=True)
    est = xs.float() + torch.tensor(-0.9, device="cuda") * vs
    simr = (F.normalize(est, dim=-1) * F.normalize(cond, dim=-1)).sum(-1)
    out["sim_maxabs"] = float((sim - simr).abs().max())
    out["mask_mismatch"] = int(((simr <= 0.5).to(torch.uint8) != mask).sum())
    fm, ed, un = ops.compact(mask, 64, 64, True)
    m2 = mask.float().view(1, 1, 64, 64)
    cross = torch.zeros(1, 1, 3, 3, device="cuda"); cross[0, 0, 1, :] = 1; cross[0, 0, :, 1] = 1
    er = (F.conv2d(m2, cross, padding=1) == 5).float()
    di = (F.conv2d(er, torch.ones(1, 1, 5, 5, device="cuda"), padding=2) > 0).flatten()
    out["morph_exact"] = bool(torch.equal(fm.bool(), di))
    ar = torch.arange(L, device="cuda", dtype=torch.int32)
    out["ids_exact"] = bool(torch.equal(ed, ar[di]) and torch.equal(un, ar[~di]))
    out["n_edited"] = int(ed.numel())
    return out


CASES = {k[5:]: v for k, v in globals().items() if k.startswith("case_")}


def main():
    if len(sys.argv) > 1:
        name = sys.argv[1]
        t0 = time.time()
        res = CASES[name]()
        print("PROBE " + json.dumps({"case": name, "ok": True, "sec": round(time.time() - t0, 2), "result": res}))
        return
    for name in CASES:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                               timeout=300)
            lines = [l for l in r.stdout.splitlines() if l.startswith("PROBE ")]
            if lines:
                print(lines[-1])
            else:
                print("PROBE " + json.dumps({"case": name, "ok": False, "rc": r.returncode,
                                             "stdout": r.stdout[-1500:], "stderr": r.stderr[-3000:]}))
        except subprocess.TimeoutExpired:
            print("PROBE " + json.dumps({"case": name, "ok": False, "error": "timeout"}))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
