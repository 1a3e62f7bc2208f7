"""How does torch round `0-dim fp32 tensor * bf16 tensor`? (decides the Euler / AVDC rounding points)"""
import sys, torch
dev = sys.argv[1] if len(sys.argv) > 1 else "cpu"
torch.manual_seed(0)
v = torch.randn(4096, 64).bfloat16().to(dev)
x = torch.randn(4096, 64).bfloat16().to(dev)
for val in (-0.0371, 0.9926, -0.9):
    s = torch.tensor(val, device=dev)
    for name, prod in (("s*v", s * v), ("v*s", v * s)):
        r_b = (s.bfloat16() * v)
        r_f = (v.float() * s).bfloat16()
        print(dev, val, name, prod.dtype, "eq_bf16_scalar", torch.equal(prod, r_b), "eq_fp32_scalar", torch.equal(prod, r_f))
    y = x.float() + s * v
    print(dev, val, "x.float()+s*v dtype", y.dtype)
