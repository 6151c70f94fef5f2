"""row f-1: the loss-stack oracle against the reference's own functions (tests/golden/loss_stack.npz)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import loss_oracle as lo


def close(a, b, rtol=1e-5, atol=1e-6):
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


def test_diffaug_forward_backward_all_flag_combinations():
    g = load_golden("loss_stack")
    seen = set()
    for ci in g["aug_cases"]:
        flags, r = g[f"aug{ci}_flags"], g[f"aug{ci}_rand01"]
        seen.add(tuple(int(f) for f in flags))
        y = lo.diffaug_forward(g[f"aug{ci}_x"], flags, r)
        close(y, g[f"aug{ci}_y"], rtol=1e-5, atol=2e-6)
        gx = lo.diffaug_backward(g[f"aug{ci}_g"], flags, r)
        close(gx, g[f"aug{ci}_gx"], rtol=1e-5, atol=2e-6)
    assert len(seen) == 8                       # every combination of (translation, colour, cutout)


def test_lpips_stage_and_backward():
    g = load_golden("loss_stack")
    for li in range(3):
        v = lo.lpips_stage(g[f"lp{li}_f0"], g[f"lp{li}_f1"], g[f"lp{li}_w"])
        close(v, g[f"lp{li}_val"], rtol=2e-5, atol=1e-7)
        gf = lo.lpips_stage_backward(g[f"lp{li}_f0"], g[f"lp{li}_f1"], g[f"lp{li}_w"], g[f"lp{li}_g"])
        close(gf, g[f"lp{li}_gf1"], rtol=2e-4, atol=1e-7)


def test_gan_losses_and_schedules():
    g = load_golden("loss_stack")
    lr, lf = g["gan_lr"].astype(np.float64), g["gan_lf"].astype(np.float64)
    close(lo.hinge_d_loss(lr, lf), g["hinge_d"])
    close(lo.vanilla_d_loss(lr, lf), g["vanilla_d"])
    close(lo.non_saturating_d_loss(lr, lf), g["nonsat_d"])
    close(lo.hinge_gen_loss(lf), g["hinge_g"])
    close(lo.non_saturating_gen_loss(lf), g["nonsat_g"])
    er, ef = lo.lecam_update(0.0, 0.0, lr, lf)
    er, ef = lo.lecam_update(er, ef, lr * 0.5, lf + 1)
    close([er, ef], g["lecam_ema"], rtol=1e-5, atol=1e-9)
    close(lo.lecam_reg(lr, lf, er, ef), g["lecam_reg"])
    close([lo.adopt_weight(0.5, s, threshold=10, value=0.0) for s in (0, 9, 10, 11)], g["adopt"])
    close([lo.anneal_weight(1.0, s, threshold=10, initial_value=0.3, final_value=0.1, anneal_steps=20)
           for s in (0, 10, 15, 30, 31, 100)], g["anneal"])


def test_batchnorm_local():
    g = load_golden("loss_stack")
    close(lo.batchnorm_local(g["bnl_x"], g["bnl_w"], g["bnl_b"], virtual_bs=4), g["bnl_y"], rtol=1e-5, atol=1e-5)


def test_loss_oracle_properties():
    """size-independent properties of the restatements themselves: DiffAug backward is the transpose of its affine forward,
    the LPIPS stage distance is symmetric, non-negative, zero on identical maps, and its analytic gradient matches finite
    differences."""
    rng = np.random.default_rng(0)
    for flags in [(1, 1, 1), (1, 0, 1), (0, 1, 0), (0, 0, 1)]:
        B, C, H, W = 3, 3, 12, 10
        r = rng.random((7, B)).astype(np.float32)
        x, g = rng.standard_normal((B, C, H, W)).astype(np.float32), rng.standard_normal((B, C, H, W)).astype(np.float32)
        y = lo.diffaug_forward(x, flags, r).astype(np.float64)
        y0 = lo.diffaug_forward(np.zeros_like(x), flags, r).astype(np.float64)
        lhs = ((y - y0) * g).sum()
        rhs = (x.astype(np.float64) * lo.diffaug_backward(g, flags, r)).sum()
        assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))
    f0 = np.maximum(rng.standard_normal((2, 16, 4, 5)), 0)
    f1 = np.maximum(rng.standard_normal((2, 16, 4, 5)), 0)
    w = rng.random(16) * 0.2
    v01, v10 = lo.lpips_stage(f0, f1, w), lo.lpips_stage(f1, f0, w)
    close(v01, v10, rtol=1e-12, atol=0)
    assert np.all(v01 >= 0) and np.all(lo.lpips_stage(f0, f0, w) < 1e-20)
    gvec = rng.standard_normal(2)
    grad = lo.lpips_stage_backward(f0, f1, w, gvec)
    eps = 1e-6
    for (b, c, h, ww) in [(0, 3, 1, 2), (1, 7, 0, 4), (0, 15, 3, 0)]:
        if f1[b, c, h, ww] == 0:
            continue                              # post-ReLU zero: one-sided
        fp, fm = f1.copy(), f1.copy()
        fp[b, c, h, ww] += eps
        fm[b, c, h, ww] -= eps
        num = ((lo.lpips_stage(f0, fp, w) - lo.lpips_stage(f0, fm, w)) * gvec).sum() / (2 * eps)
        assert abs(num - grad[b, c, h, ww]) <= 1e-5 * max(1e-3, abs(num))
