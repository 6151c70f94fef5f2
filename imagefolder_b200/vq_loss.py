"""Tokenizer training loss (reconstruction + LPIPS + adversarial + codebook terms) -- tokenizer/tokenizer_image/vq_loss.py.

`VQLoss` keeps the reference's constructor, its `forward(codebook_loss, sem_loss, detail_loss, dependency_loss, inputs,
reconstructions, optimizer_idx, global_step, last_layer, logger, log_every, fade_blur_schedule)` call (:150-152) and its
sub-module names (`discriminator`, `perceptual_loss`: checkpoint keys).  The HBM-bound pieces under it run as fused CUDA
kernels: the LPIPS stage distances (lpips.py) and DiffAug (diffaug.py); the VGG16 / DINO ViT-S trunks are library kernels.
Differences, all host-side: wandb and a process group are optional (the reference needs both even on one GPU, :143-144),
and the LeCAM running means stay on the device (the reference `.item()`s two scalars per discriminator step, :68-69).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .diffaug import DiffAug
from .discriminator_dino import DinoDisc as DINODiscriminator
from .lpips import LPIPS


# ---- GAN objectives (vq_loss.py:18-45) -------------------------------------------------------------------------------
def _bce_const_input(const: float, logits):
    """BCE-with-logits exactly as the reference CALLS it (:33-34, :44): the constant tensor sits in the *input* slot and the
    logits in the *target* slot."""
    return F.binary_cross_entropy_with_logits(torch.full_like(logits, const), logits).mean()


def hinge_d_loss(logits_real, logits_fake):
    margin_real = F.relu(1. - logits_real).mean()
    margin_fake = F.relu(1. + logits_fake).mean()
    return (margin_real + margin_fake) * 0.5


def vanilla_d_loss(logits_real, logits_fake):
    return (F.softplus(-logits_real).mean() + F.softplus(logits_fake).mean()) * 0.5


def non_saturating_d_loss(logits_real, logits_fake):
    return (_bce_const_input(1.0, logits_real) + _bce_const_input(0.0, logits_fake)) * 0.5


def hinge_gen_loss(logit_fake):
    return logit_fake.mean().neg()


def non_saturating_gen_loss(logit_fake):
    return _bce_const_input(1.0, logit_fake)


def adopt_weight(weight, global_step, threshold=0, value=0.):
    """`value` until the discriminator starts (:47-50)."""
    return weight if global_step >= threshold else value


def anneal_weight(weight, global_step, threshold=0, initial_value=0.3, final_value=0.1, anneal_steps=2000):
    """linear ramp initial -> final over `anneal_steps` after `threshold` (:52-62)."""
    done = (global_step - threshold) / anneal_steps
    if done < 0:
        return initial_value
    return final_value if done >= 1 else initial_value + done * (final_value - initial_value)


class LeCAM_EMA(object):
    """running means of the real / fake logits (:64-73); kept as 0-dim tensors on the logits' device (no host sync)."""

    def __init__(self, init=0., decay=0.999):
        self.decay = decay
        self.logits_real_ema = self.logits_fake_ema = init

    def update(self, logits_real, logits_fake):
        keep, new = self.decay, 1 - self.decay
        self.logits_real_ema = keep * self.logits_real_ema + new * logits_real.detach().float().mean()
        self.logits_fake_ema = keep * self.logits_fake_ema + new * logits_fake.detach().float().mean()


def lecam_reg(real_pred, fake_pred, lecam_ema):
    over = F.relu(real_pred - lecam_ema.logits_fake_ema).square().mean()
    under = F.relu(lecam_ema.logits_real_ema - fake_pred).square().mean()
    return over + under


class PatchGANDiscriminator(nn.Module):
    """pix2pix N-layer PatchGAN (discriminator_patchgan.py:6-63; BatchNorm variant), keys `main.{i}.*`."""

    def __init__(self, input_nc=3, ndf=64, n_layers=3, use_actnorm=False):
        super().__init__()
        if use_actnorm:
            raise NotImplementedError("ActNorm variant of the PatchGAN discriminator is not selected by any shipped config")
        seq = [nn.Conv2d(input_nc, ndf, 4, stride=2, padding=1), nn.LeakyReLU(0.2, True)]
        mult = 1
        for n in range(1, n_layers + 1):
            prev, mult = mult, min(2 ** n, 8)
            seq += [nn.Conv2d(ndf * prev, ndf * mult, 4, stride=2 if n < n_layers else 1, padding=1, bias=False),
                    nn.BatchNorm2d(ndf * mult), nn.LeakyReLU(0.2, True)]
        seq += [nn.Conv2d(ndf * mult, 1, 4, stride=1, padding=1)]
        self.main = nn.Sequential(*seq)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight.data, 0.0, 0.02)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.normal_(m.weight.data, 1.0, 0.02)
                nn.init.constant_(m.bias.data, 0)

    def forward(self, input):
        return self.main(input)


class _Blur3(nn.Module):
    """[1,2,1] x [1,2,1] / 16 binomial blur per channel with reflect padding (what kornia.filters.filter2d(x, f,
    normalized=True) computes in discriminator_stylegan.py:86-95); buffer name `f` as in the reference."""

    def __init__(self):
        super().__init__()
        self.register_buffer('f', torch.tensor([1., 2., 1.]))

    def forward(self, x):
        k = self.f[None, :] * self.f[:, None]
        k = (k / k.sum()).to(x.dtype)[None, None].expand(x.shape[1], 1, 3, 3)
        return F.conv2d(F.pad(x, (1, 1, 1, 1), mode='reflect'), k, groups=x.shape[1])


class _StyleGANBlock(nn.Module):
    """residual down-block: 1x1 stride-2 skip + (3x3, 3x3, blur, 3x3 stride-2), summed and scaled by 1/sqrt2 (:58-82)"""

    def __init__(self, input_channels, filters, downsample=True):
        super().__init__()
        act = lambda: nn.LeakyReLU(0.2, inplace=True)
        self.conv_res = nn.Conv2d(input_channels, filters, 1, stride=2 if downsample else 1)
        self.net = nn.Sequential(nn.Conv2d(input_channels, filters, 3, padding=1), act(),
                                 nn.Conv2d(filters, filters, 3, padding=1), act())
        self.downsample = nn.Sequential(_Blur3(), nn.Conv2d(filters, filters, 3, padding=1, stride=2)) if downsample else None

    def forward(self, x):
        y = self.net(x)
        if self.downsample is not None:
            y = self.downsample(y)
        return (y + self.conv_res(x)) * (1 / np.sqrt(2))


class StyleGANDiscriminator(nn.Module):
    """StyleGAN2-style residual discriminator down to 4x4 + 2-layer MLP (discriminator_stylegan.py:13-55); keys
    `blocks.{i}...`, `final_conv.0`, `final_linear.{0,2}`."""

    def __init__(self, input_nc=3, ndf=64, n_layers=3, channel_multiplier=1, image_size=256):
        super().__init__()
        width = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier, 128: 128 * channel_multiplier,
                 256: 64 * channel_multiplier, 512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}
        stages = int(np.log2(image_size))
        cin = width[image_size]
        blocks = [nn.Conv2d(input_nc, cin, 3, padding=1), nn.LeakyReLU(0.2, inplace=True)]
        for i in range(stages, 2, -1):
            blocks.append(_StyleGANBlock(cin, width[2 ** (i - 1)]))
            cin = width[2 ** (i - 1)]
        self.blocks = nn.ModuleList(blocks)
        self.final_conv = nn.Sequential(nn.Conv2d(cin, width[4], 3, padding=1), nn.LeakyReLU(0.2, inplace=True))
        self.final_linear = nn.Sequential(nn.Linear(width[4] * 16, width[4]), nn.LeakyReLU(0.2, inplace=True), nn.Linear(width[4], 1))

    def forward(self, x):
        for blk in self.blocks:
            x = blk(x)
        return self.final_linear(self.final_conv(x).flatten(1))


class VQLoss(nn.Module):
    def __init__(self, disc_start, disc_loss="hinge", disc_dim=64, disc_type='patchgan', image_size=256,
                 disc_num_layers=3, disc_in_channels=3, disc_weight=1.0, disc_adaptive_weight=False,
                 gen_adv_loss='hinge', reconstruction_loss='l2', reconstruction_weight=1.0,
                 codebook_weight=1.0, perceptual_weight=1.0, lecam_loss_weight=None, norm_type='bn', aug_prob=1,
                 ):
        super().__init__()
        assert disc_type in ["patchgan", "stylegan", 'dinodisc', 'samdisc']
        assert disc_loss in ["hinge", "vanilla", "non-saturating"]
        assert gen_adv_loss in ["hinge", "non-saturating"]
        self.disc_type = disc_type
        if disc_type == "patchgan":
            self.discriminator = PatchGANDiscriminator(input_nc=disc_in_channels, n_layers=disc_num_layers, ndf=disc_dim)
        elif disc_type == "stylegan":
            self.discriminator = StyleGANDiscriminator(input_nc=disc_in_channels, image_size=image_size)
        elif disc_type == "dinodisc":
            self.discriminator = DINODiscriminator(norm_type=norm_type)
            self.daug = DiffAug(prob=aug_prob, cutout=0.2)
        else:
            raise NotImplementedError("disc_type='samdisc' names a class the reference never defines (vq_loss.py:107)")
        self.disc_loss = {"hinge": hinge_d_loss, "vanilla": vanilla_d_loss, "non-saturating": non_saturating_d_loss}[disc_loss]
        self.discriminator_iter_start = disc_start
        self.disc_weight = disc_weight
        self.disc_adaptive_weight = disc_adaptive_weight
        self.gen_adv_loss = {"hinge": hinge_gen_loss, "non-saturating": non_saturating_gen_loss}[gen_adv_loss]
        self.perceptual_loss = LPIPS().eval()
        self.perceptual_weight = perceptual_weight
        if reconstruction_loss not in ("l1", "l2"):
            raise ValueError(f"Unknown rec loss '{reconstruction_loss}'.")
        self.rec_loss = F.l1_loss if reconstruction_loss == "l1" else F.mse_loss
        self.rec_weight = reconstruction_weight
        self.codebook_weight = codebook_weight
        self.lecam_loss_weight = lecam_loss_weight
        if self.lecam_loss_weight is not None:
            self.lecam_ema = LeCAM_EMA()
        self.wandb_tracker = None
        try:                                            # optional experiment tracking (the reference requires it)
            import torch.distributed as tdist
            if (not tdist.is_initialized()) or tdist.get_rank() == 0:
                import wandb
                self.wandb_tracker = wandb.init(project='MSVQ')
        except Exception:
            self.wandb_tracker = None

    def train(self, mode: bool = True):
        super().train(mode)
        self.perceptual_loss.eval()                     # LPIPS is frozen and always in eval mode (:128)
        return self

    def calculate_adaptive_weight(self, nll_loss, g_loss, last_layer):
        nll_grads = torch.autograd.grad(nll_loss, last_layer, retain_graph=True)[0]
        g_grads = torch.autograd.grad(g_loss, last_layer, retain_graph=True)[0]
        d_weight = torch.norm(nll_grads) / (torch.norm(g_grads) + 1e-4)
        return torch.clamp(d_weight, 0.0, 1e4).detach()

    def _disc_in(self, x, fade_blur_schedule):
        if self.disc_type != "dinodisc":
            return x
        return self.daug.aug(x, 0 if fade_blur_schedule < 1e-6 else fade_blur_schedule)

    def _log(self, tag, stats, logger, global_step, extra=""):
        stats = {k: float(v) for k, v in stats.items()}
        if logger is not None:
            logger.info(f"({tag}) " + ", ".join(f"{k}: {v:.4f}" for k, v in stats.items()) + extra)
        if self.wandb_tracker is not None:
            self.wandb_tracker.log(stats, step=global_step)

    def _generator_step(self, codebook_loss, extras, inputs, recon, global_step, last_layer, disc_weight, fade, log):
        """vq_loss.py:153-207"""
        rec = self.rec_loss(inputs, recon)
        perceptual = self.perceptual_loss(inputs, recon).mean()
        adv = self.gen_adv_loss(self.discriminator(self._disc_in(recon, fade)))
        if self.disc_adaptive_weight:
            nll = self.rec_weight * rec + self.perceptual_weight * perceptual
            balance = self.calculate_adaptive_weight(nll, adv, last_layer=last_layer)
        else:
            balance = 1
        sem, detail, dep = (0 if e is None else e for e in extras)
        total = self.rec_weight * rec + self.perceptual_weight * perceptual + balance * disc_weight * adv \
            + codebook_loss[0] + codebook_loss[1] + codebook_loss[2] + sem + detail + dep
        if log:
            self._log("Generator", {"rec_loss": self.rec_weight * rec, "perceptual_loss": self.perceptual_weight * perceptual,
                                    "sem_loss": sem, "detail_loss": detail, "dependency_loss": dep,
                                    "vq_loss": codebook_loss[0], "commit_loss": codebook_loss[1],
                                    "entropy_loss": codebook_loss[2], "generator_adv_loss": balance * disc_weight * adv,
                                    "disc_adaptive_weight": balance, "disc_weight": disc_weight},
                      log if log is not True else None, global_step, extra=f", codebook_usage: {codebook_loss[3]}")
        return total

    def _discriminator_step(self, inputs, recon, global_step, disc_weight, fade, log):
        """vq_loss.py:210-247"""
        fake = self.discriminator(self._disc_in(recon.detach(), fade))
        real = self.discriminator(self._disc_in(inputs.detach(), fade))
        objective = self.disc_loss(real, fake)
        if self.lecam_loss_weight is not None:
            self.lecam_ema.update(real, fake)
            objective = lecam_reg(real, fake, self.lecam_ema) * self.lecam_loss_weight + objective
        d_loss = disc_weight * objective
        if log:
            self._log("Discriminator", {"discriminator_adv_loss": d_loss, "disc_weight": disc_weight,
                                        "logits_real": real.detach().mean(), "logits_fake": fake.detach().mean()},
                      log if log is not True else None, global_step)
        return d_loss

    def forward(self, codebook_loss, sem_loss, detail_loss, dependency_loss, inputs, reconstructions, optimizer_idx,
                global_step, last_layer=None, logger=None, log_every=100, fade_blur_schedule=0):
        if optimizer_idx not in (0, 1):
            raise ValueError(f"optimizer_idx must be 0 (generator) or 1 (discriminator), got {optimizer_idx}")
        disc_weight = adopt_weight(self.disc_weight, global_step, threshold=self.discriminator_iter_start)
        due = global_step % log_every == 0
        log = (logger if logger is not None else True) if (due and (logger is not None or self.wandb_tracker is not None)) else False
        x, y = inputs.contiguous(), reconstructions.contiguous()
        if optimizer_idx == 0:
            return self._generator_step(codebook_loss, (sem_loss, detail_loss, dependency_loss), x, y, global_step, last_layer,
                                        disc_weight, fade_blur_schedule, log)
        return self._discriminator_step(x, y, global_step, disc_weight, fade_blur_schedule, log)
