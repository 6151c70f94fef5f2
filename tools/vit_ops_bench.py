import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagefolder_b200 import vit_ops


def timeit(fn, warm=3, it=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it


B, S, D = 256, 513, 768
x = torch.randn(B, S, D, device="cuda", requires_grad=True)
br = torch.randn(B, S, D, device="cuda").to(torch.bfloat16).requires_grad_(True)
g = torch.rand(D, device="cuda", requires_grad=True)
bb = torch.randn(D, device="cuda", requires_grad=True)
w = torch.rand(D, device="cuda", requires_grad=True)
b = torch.randn(D, device="cuda", requires_grad=True)
rs = torch.ones(B, device="cuda")
xo, y = vit_ops.residual_ln(x, br, bb, g, rs, w, b, 1e-6)
gx, gy = torch.randn_like(xo), torch.randn_like(y)
res = {}
res["ln_fwd_ms"] = timeit(lambda: vit_ops.residual_ln(x, br, bb, g, rs, w, b, 1e-6))
res["ln_bwd_ms"] = timeit(lambda: torch.autograd.grad((xo, y), (x, br, g, bb, w, b), (gx, gy), retain_graph=True))
h = torch.randn(B, S, 4 * D, device="cuda").to(torch.bfloat16).requires_grad_(True)
b1 = torch.randn(4 * D, device="cuda", requires_grad=True)
yy = vit_ops.gelu_bias(h, b1)
gg = torch.randn_like(yy)
res["gelu_fwd_ms"] = timeit(lambda: vit_ops.gelu_bias(h, b1))
res["gelu_bwd_ms"] = timeit(lambda: torch.autograd.grad(yy, (h, b1), gg, retain_graph=True))
n = B * S * D
res["ln_fwd_TBs"] = n * 12 / res["ln_fwd_ms"] / 1e9
res["ln_bwd_TBs"] = n * 18 / res["ln_bwd_ms"] / 1e9
res["gelu_fwd_TBs"] = n * 4 * 4 / res["gelu_fwd_ms"] / 1e9
res["gelu_bwd_TBs"] = n * 4 * 6 / res["gelu_bwd_ms"] / 1e9
print(json.dumps(res, indent=1))
