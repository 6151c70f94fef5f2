"""dev microbenchmark of the loss-stack kernels (row f-1) against the reference's op sequence on the same GPU."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagefolder_b200 import loss_ops
from imagefolder_b200.lpips import normalize_tensor, spatial_average


def timeit(fn, warm=2, it=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it


res = {}
B = 64
for name, C, HW in [("relu1_2", 64, 256), ("relu2_2", 128, 128), ("relu3_3", 256, 64), ("relu4_3", 512, 32), ("relu5_3", 512, 16)]:
    f0 = torch.relu(torch.randn(B, C, HW, HW, device="cuda")).to(torch.bfloat16)
    f1 = torch.relu(torch.randn(B, C, HW, HW, device="cuda")).to(torch.bfloat16).requires_grad_(True)
    w = torch.rand(1, C, 1, 1, device="cuda") * 0.1
    nbytes = f0.numel() * 2

    def fused():
        v = loss_ops.lpips_stage(f0, f1, w)
        torch.autograd.grad(v.sum(), f1)

    def eager():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            d = (normalize_tensor(f0) - normalize_tensor(f1)) ** 2
            v = spatial_average(torch.nn.functional.conv2d(d, w.to(d.dtype)), keepdim=True)
        torch.autograd.grad(v.float().sum(), f1)

    tf = timeit(lambda: loss_ops.lpips_stage(f0, f1, w))
    tfb = timeit(fused)
    te = timeit(eager)
    res[name] = {"fused_fwd_ms": tf, "fused_fwd_GBps": 2 * nbytes / tf / 1e6, "fused_fwdbwd_ms": tfb,
                 "fused_fwdbwd_GBps": 7 * nbytes / tfb / 1e6, "eager_fwdbwd_ms": te, "speedup": te / tfb}
    del f0, f1
    torch.cuda.empty_cache()

from imagefolder_b200.diffaug import DiffAug
x = (torch.rand(256, 3, 256, 256, device="cuda") * 2 - 1).requires_grad_(True)
aug = DiffAug(prob=1.0)


def ours():
    torch.manual_seed(0)
    y = aug.aug(x)
    torch.autograd.grad(y.sum(), x)


def ref_style():          # the reference's op sequence (diffaug.py:60-118), restated with library ops
    torch.manual_seed(0)
    torch.rand(3)
    Bn, _, H, W = x.shape
    r = torch.rand(7, Bn, 1, 1, device="cuda")
    dh = round(H * 0.125)
    th = r[0].mul(2 * dh + 1).floor().long() - dh
    tw = r[1].mul(2 * dh + 1).floor().long() - dh
    gb, gh, gw = torch.meshgrid(torch.arange(Bn, device="cuda"), torch.arange(H, device="cuda"), torch.arange(W, device="cuda"), indexing='ij')
    gh = (gh + th).add_(1).clamp_(0, H + 1)
    gw = (gw + tw).add_(1).clamp_(0, W + 1)
    y = torch.nn.functional.pad(x, [1, 1, 1, 1]).permute(0, 2, 3, 1).contiguous()[gb, gh, gw].permute(0, 3, 1, 2).contiguous()
    y = y.add(r[2].unsqueeze(-1).sub(0.5))
    m = y.mean(dim=1, keepdim=True)
    y = y.sub(m).mul(r[3].unsqueeze(-1).mul(2)).add_(m)
    m = y.mean(dim=(1, 2, 3), keepdim=True)
    y = y.sub(m).mul(r[4].unsqueeze(-1).add(0.5)).add_(m)
    ch = round(H * 0.2)
    oh = r[5].mul(H + (1 - ch % 2)).floor().long()
    ow = r[6].mul(W + (1 - ch % 2)).floor().long()
    gb, gh, gw = torch.meshgrid(torch.arange(Bn, device="cuda"), torch.arange(ch, device="cuda"), torch.arange(ch, device="cuda"), indexing='ij')
    gh = (gh + oh).sub_(ch // 2).clamp(min=0, max=H - 1)
    gw = (gw + ow).sub_(ch // 2).clamp(min=0, max=W - 1)
    mask = torch.ones(Bn, H, W, device="cuda")
    mask[gb, gh, gw] = 0
    y = y.mul(mask.unsqueeze(1))
    torch.autograd.grad(y.sum(), x)


t1, t2 = timeit(ours), timeit(ref_style)
res["diffaug_B256"] = {"fused_fwdbwd_ms": t1, "fused_GBps": 4 * x.numel() * 4 / t1 / 1e6, "reference_ops_fwdbwd_ms": t2, "speedup": t2 / t1}
print(json.dumps(res, indent=1))
