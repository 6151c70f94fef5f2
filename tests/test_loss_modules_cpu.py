"""row f-1 host side on CPU: module surfaces / checkpoint keys / library-kernel parts against the reference goldens."""
import warnings

import numpy as np
import pytest
import torch

from conftest import load_golden

warnings.filterwarnings("ignore", message=".*RANDOM.*")
warnings.filterwarnings("ignore", message=".*no vgg.pth.*")


def _np(a):
    return a.detach().cpu().double().numpy() if torch.is_tensor(a) else np.asarray(a, np.float64)


def close(a, b, rtol=1e-5, atol=1e-6):
    np.testing.assert_allclose(_np(a), _np(b), rtol=rtol, atol=atol)


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_gan_objectives_match_reference_functions():
    from imagefolder_b200 import vq_loss as L
    g = load_golden("loss_stack")
    lr, lf = T(g["gan_lr"]), T(g["gan_lf"])
    close(L.hinge_d_loss(lr, lf), g["hinge_d"])
    close(L.vanilla_d_loss(lr, lf), g["vanilla_d"])
    close(L.non_saturating_d_loss(lr, lf), g["nonsat_d"])
    close(L.hinge_gen_loss(lf), g["hinge_g"])
    close(L.non_saturating_gen_loss(lf), g["nonsat_g"])
    ema = L.LeCAM_EMA()
    ema.update(lr, lf)
    ema.update(lr * 0.5, lf + 1)
    close([float(ema.logits_real_ema), float(ema.logits_fake_ema)], g["lecam_ema"], rtol=1e-5, atol=1e-9)
    close(L.lecam_reg(lr, lf, ema), g["lecam_reg"])
    close([L.adopt_weight(0.5, s, threshold=10, value=0.0) for s in (0, 9, 10, 11)], g["adopt"])
    close([L.anneal_weight(1.0, s, threshold=10, initial_value=0.3, final_value=0.1, anneal_steps=20)
           for s in (0, 10, 15, 30, 31, 100)], g["anneal"])


def test_lpips_checkpoint_layout_and_library_path():
    from imagefolder_b200.lpips import LPIPS, normalize_tensor, spatial_average, ScalingLayer
    g = load_golden("loss_stack")
    m = LPIPS().eval()
    assert sorted(m.state_dict().keys()) == list(g["lpips_keys"])
    assert [str(tuple(v.shape)) for _, v in sorted(m.state_dict().items())] == list(g["lpips_shapes"])
    assert not any(p.requires_grad for p in m.parameters())
    close(ScalingLayer()(T(g["scal_x"])), g["scal_y"])
    for li in range(3):                           # the stage formula on the reference's synthetic features
        f0, f1, w = T(g[f"lp{li}_f0"]), T(g[f"lp{li}_f1"]), T(g[f"lp{li}_w"]).view(1, -1, 1, 1)
        val = spatial_average(torch.nn.functional.conv2d((normalize_tensor(f0) - normalize_tensor(f1)) ** 2, w), keepdim=True)
        close(val.reshape(-1), g[f"lp{li}_val"])
    x, y = torch.rand(2, 3, 32, 32) * 2 - 1, (torch.rand(2, 3, 32, 32) * 2 - 1).requires_grad_(True)
    v = m(x, y)
    assert v.shape == (2, 1, 1, 1)
    v.sum().backward()
    assert torch.isfinite(y.grad).all() and float(y.grad.abs().sum()) > 0
    assert float(m(x, x).abs().max()) < 1e-6


def test_dino_discriminator_pieces_match_reference():
    from imagefolder_b200 import discriminator_dino as dd
    g = load_golden("loss_stack")
    vit = dd.FrozenDINOSmallNoDrop(depth=3, key_depths=(0, 2), embed_dim=48, num_heads=3)
    assert sorted(vit.state_dict().keys()) == list(g["dino_keys"])
    assert not any(p.requires_grad for p in vit.parameters()) and not vit.training
    vit.load_state_dict({k: T(g["dinow_" + k]) for k in vit.state_dict().keys()})
    img = T(g["dino_img"]).repeat_interleave(8, 2).repeat_interleave(8, 3)
    acts = vit(img)
    assert len(acts) == 3
    for i, a in enumerate(acts):
        assert a.shape == (2, 48, 196)
        close(a, g[f"dino_act{i}"], rtol=1e-4, atol=2e-6)
    head = dd._make_head(48, 9, 'bn', 1e-6, True).eval()
    assert sorted(head.state_dict().keys()) == list(g["head_keys"])
    head.load_state_dict({k: T(g["headw_" + k]) for k in head.state_dict().keys()})
    close(head(T(g["dino_act0"])), g["head_y"], rtol=1e-5, atol=1e-6)
    bn = dd.BatchNormLocal(6, virtual_bs=4)
    bn.weight.data, bn.bias.data = T(g["bnl_w"]), T(g["bnl_b"])
    close(bn(T(g["bnl_x"])), g["bnl_y"], rtol=1e-5, atol=1e-6)
    disc = dd.DinoDisc(device='cpu')
    assert all(k.startswith("heads.") for k in disc.state_dict().keys())        # frozen backbone is not checkpointed
    assert len(disc.heads) == 5
    out = disc(torch.rand(2, 3, 64, 64) * 2 - 1)
    assert out.shape == (2, 5 * 196)
    with pytest.raises(NotImplementedError):
        dd.make_block(8, 1, 'lbn', 1e-6, True)


def test_patchgan_layout_and_vqloss_patchgan_on_cpu():
    from imagefolder_b200.vq_loss import PatchGANDiscriminator, VQLoss, hinge_d_loss, hinge_gen_loss
    g = load_golden("loss_stack")
    pg = PatchGANDiscriminator(input_nc=3, n_layers=3, ndf=16)
    assert sorted(pg.state_dict().keys()) == list(g["patchgan_keys"])
    assert [str(tuple(v.shape)) for _, v in sorted(pg.state_dict().items())] == list(g["patchgan_shapes"])
    torch.manual_seed(0)
    loss = VQLoss(disc_start=5, disc_type='patchgan', disc_dim=8, disc_loss='hinge', gen_adv_loss='hinge',
                  perceptual_weight=0.7, reconstruction_weight=1.3, disc_weight=0.5)
    assert not loss.perceptual_loss.training
    loss.train()
    assert not loss.perceptual_loss.training                  # LPIPS stays in eval mode
    x = torch.rand(2, 3, 32, 32) * 2 - 1
    rec = (x + 0.1 * torch.randn_like(x)).requires_grad_(True)
    cb = (torch.tensor(0.3), torch.tensor(0.2), torch.tensor(0.1), [50.0])
    for step, dw in ((3, 0.0), (9, 0.5)):                     # before / after disc_start (adopt_weight)
        total = loss(cb, None, None, None, x, rec, 0, step)
        with torch.no_grad():
            p = loss.perceptual_loss(x, rec).mean()
            adv = hinge_gen_loss(loss.discriminator(rec))
        want = 1.3 * torch.nn.functional.mse_loss(x, rec) + 0.7 * p + dw * adv + 0.6
        close(total.detach(), want.detach(), rtol=1e-5)
        d = loss(cb, None, None, None, x, rec, 1, step)
        with torch.no_grad():
            wd = dw * hinge_d_loss(loss.discriminator(x), loss.discriminator(rec.detach()))
        close(d.detach(), wd, rtol=1e-5)
    with pytest.raises(ValueError):
        loss(cb, None, None, None, x, rec, 2, 0)
    with pytest.raises(NotImplementedError):
        VQLoss(disc_start=0, disc_type='samdisc')
    with pytest.raises(AssertionError):
        VQLoss(disc_start=0, disc_loss='wgan')


def test_stylegan_discriminator_layout_and_blur():
    from imagefolder_b200.vq_loss import StyleGANDiscriminator, VQLoss, _Blur3
    g = load_golden("loss_stack")
    m = StyleGANDiscriminator(input_nc=3, image_size=32)
    assert sorted(m.state_dict().keys()) == list(g["stylegan_keys"])
    assert [str(tuple(v.shape)) for _, v in sorted(m.state_dict().items())] == list(g["stylegan_shapes"])
    assert m(torch.randn(2, 3, 32, 32)).shape == (2, 1)
    # the blur is the normalised [1,2,1]x[1,2,1] filter with reflect borders (kornia filter2d semantics; kornia itself is
    # not installed here, so this is checked against a direct evaluation, not against the reference)
    x = torch.arange(2 * 2 * 5 * 5, dtype=torch.float32).reshape(2, 2, 5, 5)
    y = _Blur3()(x)
    xp = torch.nn.functional.pad(x, (1, 1, 1, 1), mode='reflect')
    k = torch.tensor([[1., 2., 1.], [2., 4., 2.], [1., 2., 1.]]) / 16
    want = sum(k[i, j] * xp[:, :, i:i + 5, j:j + 5] for i in range(3) for j in range(3))
    close(y, want, rtol=1e-6, atol=1e-6)
    close(_Blur3()(torch.ones(1, 3, 8, 8)), torch.ones(1, 3, 8, 8))         # normalised: constants are preserved
    loss = VQLoss(disc_start=0, disc_type='stylegan', image_size=32)
    x = torch.rand(2, 3, 32, 32) * 2 - 1
    cb = (torch.tensor(0.1), torch.tensor(0.1), torch.tensor(0.0), [1.0])
    assert torch.isfinite(loss(cb, None, None, None, x, x * 0.9, 0, 1)) and torch.isfinite(loss(cb, None, None, None, x, x * 0.9, 1, 1))


# ---- round 2: pretrained-weight plumbing is explicit and loud ------------------------------------
def test_lpips_reports_a_random_trunk_and_loads_torchvision_layout(tmp_path, monkeypatch):
    import warnings
    import torch
    from imagefolder_b200.lpips import LPIPS
    monkeypatch.delenv("XQ_VGG16_CKPT", raising=False)
    monkeypatch.delenv("XQ_LPIPS_CKPT", raising=False)
    with warnings.catch_warnings(record=True):
        warnings.simplefilter("always")
        ref = LPIPS().eval()
    lin_only = {k: v.clone() for k, v in ref.state_dict().items() if k.startswith("lin")}       # what vgg.pth carries
    torch.save(lin_only, tmp_path / "vgg.pth")
    monkeypatch.setenv("XQ_LPIPS_CKPT", str(tmp_path / "vgg.pth"))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m = LPIPS().eval()
    assert any("VGG16 trunk" in str(x.message) for x in w)                   # lin-only file: the trunk is random, said loudly
    assert m.unloaded_keys and all(k.startswith("net.") for k in m.unloaded_keys)
    assert torch.equal(m.lin0.model[-1].weight, ref.lin0.model[-1].weight)
    monkeypatch.setenv("XQ_REQUIRE_PRETRAINED", "1")
    with pytest.raises(RuntimeError):
        LPIPS()
    monkeypatch.delenv("XQ_REQUIRE_PRETRAINED")
    # torchvision layout: features.N.{weight,bias}
    tv = {}
    for k, v in ref.state_dict().items():
        if k.startswith("net.slice"):
            _, _, idx, leaf = k.split(".")
            tv[f"features.{idx}.{leaf}"] = v.clone()
    tv["classifier.0.weight"] = torch.zeros(4, 4)                            # ignored, as in the full torchvision file
    torch.save(tv, tmp_path / "vgg16.pth")
    monkeypatch.setenv("XQ_VGG16_CKPT", str(tmp_path / "vgg16.pth"))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m2 = LPIPS().eval()
    assert m2.unloaded_keys == [] and not w
    for k, v in ref.state_dict().items():
        assert torch.equal(m2.state_dict()[k], v), k


def test_create_model_pretrained_is_loud_and_resamples_local_checkpoints(tmp_path, monkeypatch):
    import warnings
    import torch
    from imagefolder_b200.dino_enc.vision_transformer import create_model
    name = "vit_small_patch14_dinov2.lvd142m"
    env = "XQ_TIMM_CKPT_VIT_SMALL_PATCH14_DINOV2_LVD142M"
    monkeypatch.delenv(env, raising=False)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m = create_model(name, pretrained=True, img_size=64, patch_size=16, depth=1)
    assert any("RANDOM" in str(x.message) for x in w) and not m.pretrained_loaded
    # a "timm" checkpoint at the native geometry (here 56px / patch 14 -> 4x4 grid) loads into 64px / patch 16
    torch.manual_seed(0)
    src = create_model(name, pretrained=False, img_size=56, patch_size=14, depth=1)
    torch.save(src.state_dict(), tmp_path / "vit.pth")
    monkeypatch.setenv(env, str(tmp_path / "vit.pth"))
    dst = create_model(name, pretrained=True, img_size=64, patch_size=16, depth=1)
    assert dst.pretrained_loaded
    assert dst.pos_embed.shape == (1, 17, 384) and dst.patch_embed.proj.weight.shape[-1] == 16
    assert torch.equal(dst.blocks[0].attn.qkv.weight, src.blocks[0].attn.qkv.weight)
    assert torch.allclose(dst.pos_embed[:, :1], src.pos_embed[:, :1])         # the class-token position is kept as is
    clip = create_model("vit_base_patch16_clip_224.openai", pretrained=False, depth=1)
    assert clip.norm.eps == 1e-5 and src.norm.eps == 1e-6                      # timm's CLIP variants use nn.LayerNorm defaults
