"""DiffAug (differentiable augmentation in front of the discriminator) -- tokenizer/tokenizer_image/diffaug.py:23-118.

Same constructor, same `aug(BCHW, warmup_blur_schedule)` call and -- for parity under a fixed seed -- the SAME random
draws in the same order from the same generators (three Bernoulli draws from the CPU generator, one `rand(7,B,1,1)` from
the device generator).  The translation / colour / cutout chain runs as one fused CUDA pass (csrc/loss_kernels.cu) whose
backward is the exact transpose; the reference gathers through a padded NHWC copy, takes three mean reductions and
scatters a mask (~15 full passes over the batch).  The warm-up blur (only active before `disc_start`, xqgan_train.py:443)
is a depthwise convolution and stays a library call.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import loss_ops


class DiffAug(object):
    def __init__(self, prob=1.0, cutout=0.2):
        self.prob = abs(prob)
        self.using_cutout = prob > 0
        self.cutout = cutout
        self._blur_radius = -1
        self._blur_kh = self._blur_kw = None

    def _blur(self, x, schedule: float):
        """Gaussian warm-up blur, sigma = sqrt(H/2) * schedule, reflect padding (diffaug.py:49-66)."""
        ch = x.shape[1]
        sigma = (x.shape[-2] * 0.5) ** 0.5 * schedule
        radius = math.floor(sigma * 3)
        if radius < 1:
            return x
        if radius != self._blur_radius or self._blur_kh is None or self._blur_kh.device != x.device \
                or self._blur_kh.shape[0] != ch:
            k = torch.arange(-radius, radius + 1, dtype=torch.float32, device=x.device)
            k = k.mul_(1 / sigma).square_().neg_().exp2_()
            k.div_(k.sum())
            self._blur_radius = radius
            self._blur_kh = k.view(1, 1, -1, 1).repeat(ch, 1, 1, 1).contiguous()
            self._blur_kw = k.view(1, 1, 1, -1).repeat(ch, 1, 1, 1).contiguous()
        x = F.pad(x, [radius] * 4, mode='reflect')
        x = F.conv2d(x, self._blur_kh, groups=ch)
        return F.conv2d(x, self._blur_kw, groups=ch)

    def aug(self, BCHW: torch.Tensor, warmup_blur_schedule: float = 0) -> torch.Tensor:
        if BCHW.dtype != torch.float32:
            BCHW = BCHW.float()
        if warmup_blur_schedule > 0:
            BCHW = self._blur(BCHW, warmup_blur_schedule)
        if self.prob < 1e-6:
            return BCHW
        trans, color, cut = (torch.rand(3) <= self.prob).tolist()          # CPU generator, as the reference (:61-62)
        if not (trans or color or cut):
            return BCHW
        B, _, H, W = BCHW.shape
        rand01 = torch.rand(7, B, 1, 1, device=BCHW.device)                 # device generator (:64)
        cut_h, cut_w = round(H * self.cutout), round(W * self.cutout)
        # a zero-sized cutout is an empty index grid in the reference (diffaug.py:100-112: nothing is zeroed)
        flags = int(trans) | (int(color) << 1) | (int(bool(self.using_cutout and cut and cut_h > 0 and cut_w > 0)) << 2)
        if flags == 0:
            return BCHW
        return loss_ops.diffaug_apply(BCHW, rand01.view(7, B), flags, cut_h, cut_w)
