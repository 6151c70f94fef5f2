"""row f-1 on the GPU: the fused LPIPS-stage and DiffAug kernels (through the C ABI wrappers) against the oracle and the
reference goldens, and one full VQLoss generator / discriminator step with the DINO discriminator."""
import warnings

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import loss_oracle as lo

pytestmark = pytest.mark.gpu
warnings.filterwarnings("ignore", message=".*RANDOM.*")
warnings.filterwarnings("ignore", message=".*no vgg.pth.*")


def dev(a, dtype=torch.float32):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dtype)


def npy(t):
    return t.detach().float().cpu().numpy()


def close(a, b, rtol=1e-5, atol=1e-6):
    a = npy(a) if torch.is_tensor(a) else np.asarray(a)
    b = npy(b) if torch.is_tensor(b) else np.asarray(b)
    np.testing.assert_allclose(a.astype(np.float64), b.astype(np.float64), rtol=rtol, atol=atol)


def test_lpips_stage_golden_and_oracle():
    from imagefolder_b200.loss_ops import lpips_stage
    g = load_golden("loss_stack")
    for li in range(3):
        f0, f1, w = dev(g[f"lp{li}_f0"]), dev(g[f"lp{li}_f1"]).requires_grad_(True), dev(g[f"lp{li}_w"])
        val = lpips_stage(f0, f1, w)
        close(val, g[f"lp{li}_val"], rtol=2e-5, atol=1e-7)                       # the reference's own output
        close(val, lo.lpips_stage(g[f"lp{li}_f0"], g[f"lp{li}_f1"], g[f"lp{li}_w"]), rtol=1e-6, atol=1e-8)
        (gf1,) = torch.autograd.grad(val, f1, dev(g[f"lp{li}_g"]))
        close(gf1, g[f"lp{li}_gf1"], rtol=2e-4, atol=1e-7)
        close(gf1, lo.lpips_stage_backward(g[f"lp{li}_f0"], g[f"lp{li}_f1"], g[f"lp{li}_w"], g[f"lp{li}_g"]), rtol=2e-5, atol=1e-8)


@pytest.mark.parametrize("B,C,H,W,dt", [(2, 64, 37, 29, torch.float32), (1, 130, 9, 9, torch.float32),
                                        (3, 256, 16, 16, torch.bfloat16), (2, 512, 4, 4, torch.bfloat16)])
def test_lpips_stage_ragged_both_gradients(B, C, H, W, dt):
    from imagefolder_b200.loss_ops import lpips_stage
    torch.manual_seed(C + H)
    f0 = torch.relu(torch.randn(B, C, H, W, device="cuda")).to(dt).requires_grad_(True)
    f1 = torch.relu(torch.randn(B, C, H, W, device="cuda")).to(dt).requires_grad_(True)
    f1.data[0, :, 0, 0] = 0                                  # an all-zero pixel (the reference's autograd gives NaN here)
    w = torch.rand(C, device="cuda") * 0.1
    val = lpips_stage(f0, f1, w.view(1, C, 1, 1))
    go = torch.randn(B, device="cuda")
    g0, g1 = torch.autograd.grad(val, (f0, f1), go)
    assert g0.dtype == dt and g1.dtype == dt and torch.isfinite(g0).all() and torch.isfinite(g1).all()
    a, b = npy(f0), npy(f1)
    tol = 1e-5 if dt == torch.float32 else 1.5e-2             # bf16 gradients are rounded to bf16
    close(val, lo.lpips_stage(a, b, npy(w)), rtol=1e-5, atol=1e-8)
    want1 = lo.lpips_stage_backward(a, b, npy(w), npy(go))
    want0 = lo.lpips_stage_backward(b, a, npy(w), npy(go))
    close(g1, want1, rtol=tol, atol=tol * float(np.abs(want1).max()))
    close(g0, want0, rtol=tol, atol=tol * float(np.abs(want0).max()))
    # identical maps -> exactly zero distance is not guaranteed in floating point, but it must be tiny and non-negative-ish
    z = lpips_stage(f0.detach(), f0.detach(), w)
    assert float(z.abs().max()) < 1e-6


def test_diffaug_kernel_against_reference_goldens():
    from imagefolder_b200.loss_ops import diffaug_apply
    g = load_golden("loss_stack")
    for ci in g["aug_cases"]:
        flags3 = [int(f) for f in g[f"aug{ci}_flags"]]
        flags = flags3[0] | (flags3[1] << 1) | (flags3[2] << 2)
        if flags == 0:
            continue
        x = dev(g[f"aug{ci}_x"]).requires_grad_(True)
        _, _, H, W = x.shape
        y = diffaug_apply(x, dev(g[f"aug{ci}_rand01"]), flags, round(H * 0.2), round(W * 0.2))
        close(y, g[f"aug{ci}_y"], rtol=1e-5, atol=2e-6)
        (gx,) = torch.autograd.grad(y, x, dev(g[f"aug{ci}_g"]))
        close(gx, g[f"aug{ci}_gx"], rtol=1e-5, atol=2e-6)


def test_diffaug_module_draw_order_and_full_size():
    """DiffAug.aug consumes the generators exactly as the reference does: three CPU draws, then rand(7,B,1,1) on the device."""
    from imagefolder_b200.diffaug import DiffAug
    aug = DiffAug(prob=0.8, cutout=0.2)
    x = torch.rand(16, 3, 256, 256, device="cuda") * 2 - 1
    n_aug = 0
    for seed in range(6):
        torch.manual_seed(seed)
        y = aug.aug(x)
        torch.manual_seed(seed)
        flags3 = (torch.rand(3) <= 0.8).tolist()
        if not any(flags3):
            assert y is x
            continue
        rand01 = torch.rand(7, 16, 1, 1, device="cuda")
        want = lo.diffaug_forward(npy(x), flags3, npy(rand01).reshape(7, 16), cutout=0.2)
        close(y, want, rtol=1e-5, atol=3e-6)
        n_aug += 1
    assert n_aug >= 3
    assert DiffAug(prob=0.0).aug(x) is x
    # linearity / transpose property at full size: <A x, g> == <x, A^T g> for the affine part
    torch.manual_seed(1)
    xs = x.clone().requires_grad_(True)
    y = aug.aug(xs)
    torch.manual_seed(1)
    y0 = aug.aug(torch.zeros_like(x))
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, xs, gy)
    lhs = ((y - y0) * gy).double().sum()
    rhs = (x * gx).double().sum()
    assert abs(float(lhs - rhs)) <= 1e-4 * max(1.0, abs(float(lhs)))


def test_lpips_module_fused_equals_library_formula():
    from imagefolder_b200.lpips import LPIPS, normalize_tensor, spatial_average
    torch.manual_seed(0)
    m = LPIPS().eval().cuda()
    for i in range(5):                                        # non-trivial lin weights
        getattr(m, f"lin{i}").model[1].weight.data.uniform_(0, 0.1)
    x = torch.rand(3, 3, 64, 64, device="cuda") * 2 - 1
    y = (x + 0.2 * torch.randn_like(x)).clamp(-1, 1).requires_grad_(True)
    v = m(x, y)
    (gy,) = torch.autograd.grad(v.sum(), y)
    y2 = y.detach().clone().requires_grad_(True)
    o0, o1 = m.net(m.scaling_layer(x)), m.net(m.scaling_layer(y2))
    ref = 0
    for k in range(5):
        d = (normalize_tensor(o0[k]) - normalize_tensor(o1[k])) ** 2
        ref = ref + spatial_average(getattr(m, f"lin{k}").model(d), keepdim=True)
    (gy2,) = torch.autograd.grad(ref.sum(), y2)
    close(v, ref, rtol=2e-4, atol=1e-6)
    close(gy, gy2, rtol=2e-3, atol=2e-3 * float(gy2.abs().max()))
    with torch.autocast("cuda", dtype=torch.bfloat16):        # bf16 features through the bf16 kernel instantiation
        vb = m(x, y)
    close(vb, ref, rtol=5e-2, atol=1e-4)


def test_vqloss_dinodisc_generator_and_discriminator_steps():
    from imagefolder_b200.vq_loss import VQLoss, hinge_d_loss, lecam_reg
    torch.manual_seed(0)
    loss = VQLoss(disc_start=0, disc_type='dinodisc', disc_loss='hinge', gen_adv_loss='hinge', disc_adaptive_weight=True,
                  lecam_loss_weight=0.001, perceptual_weight=1.0, aug_prob=1.0).cuda().train()
    assert sorted(k.split(".")[0] for k in loss.state_dict().keys()).count("discriminator") > 0
    x = torch.rand(8, 3, 256, 256, device="cuda") * 2 - 1
    last = torch.nn.Parameter(torch.randn(3, 3, device="cuda") * 0.1)            # stands in for decoder.last_layer
    rec = torch.einsum("oc,bchw->bohw", torch.eye(3, device="cuda") + last, x) * 0.9
    cb = (torch.tensor(0.3, device="cuda"), torch.tensor(0.2, device="cuda"), torch.tensor(0.0, device="cuda"), [10.0])
    torch.manual_seed(5)
    g_loss = loss(cb, None, None, None, x, rec, 0, 1, last_layer=last)
    assert torch.isfinite(g_loss)
    g_loss.backward()
    assert torch.isfinite(last.grad).all() and float(last.grad.abs().sum()) > 0
    assert all(p.grad is None for p in loss.perceptual_loss.parameters())
    import random
    torch.manual_seed(6)
    random.seed(6)          # the frozen backbone picks "random 224-crop" vs "area resize" with Python's RNG (:331)
    d_loss = loss(cb, None, None, None, x, rec.detach(), 1, 1)
    # replay the same augmentation draws and rebuild the discriminator objective by hand
    torch.manual_seed(6)
    random.seed(6)
    lf = loss.discriminator(loss.daug.aug(rec.detach(), 0))
    lr = loss.discriminator(loss.daug.aug(x, 0))
    want = lecam_reg(lr, lf, loss.lecam_ema) * 0.001 + hinge_d_loss(lr, lf)
    close(d_loss, want, rtol=1e-4, atol=1e-6)
    d_loss.backward()
    head_grads = [p.grad for p in loss.discriminator.heads.parameters() if p.requires_grad]
    assert head_grads and all(g_ is not None and torch.isfinite(g_).all() for g_ in head_grads)
