import numpy as np, torch
torch.manual_seed(0)
x = torch.rand(1_000_000, device="cuda") * 3
for s in (np.sqrt(1 - 0.999 ** 1), np.sqrt(1 - 0.999 ** 3), 0.7071, 3.3333333):
    y = x / s
    c1 = x * torch.tensor(np.float32(1.0) / np.float32(s), device="cuda")
    c2 = x * torch.tensor(np.float32(1.0 / s), device="cuda")
    c3 = x / torch.tensor(np.float32(s), device="cuda")
    print(s, "inv_f32", bool(torch.equal(y, c1)), "inv_f64", bool(torch.equal(y, c2)), "truediv", bool(torch.equal(y, c3)))
g = torch.randn(1_000_000, device="cuda")
v = torch.rand(1_000_000, device="cuda")
a = v.clone().addcmul_(g, g, value=1 - 0.999)
al = torch.tensor(np.float32(1 - 0.999), device="cuda")
print("addcmul a+alpha*(b*c)", bool(torch.equal(a, torch.addcmul(v, al * torch.ones_like(g), g * g))) )
m = torch.randn(1_000_000, device="cuda"); d = torch.rand(1_000_000, device="cuda") + 0.1
p = torch.randn(1_000_000, device="cuda")
r = p.clone().addcdiv_(m, d, value=-0.0123)
q = m / d
r2 = torch.addcmul(p, q, torch.full_like(q, np.float32(-0.0123)))
print("addcdiv a+alpha*(b/c) fma", bool(torch.equal(r, r2)), (r - r2).abs().max().item())
r3 = p + (q * np.float32(-0.0123))
print("addcdiv unfused", bool(torch.equal(r, r3)))
r4 = p + torch.tensor(np.float32(-0.0123), device="cuda") * m / d
print("addcdiv (alpha*b)/c", bool(torch.equal(r, r4)))
