"""Convolutional encoder / decoder for `enc_type='cnn'` / `dec_type='cnn'`.

Same architecture, hyper-parameters and parameter names (= checkpoint keys: conv_in, conv_blocks.{i}.res.{j}.*,
conv_blocks.{i}.attn.{j}.*, conv_blocks.{i}.downsample|upsample.conv, mid.{0,1,2}.*, norm_out, conv_out) as the
reference's taming-style autoencoder (tokenizer/tokenizer_image/xqgan_model.py:454-704).  No shipped YAML selects
it, so it runs on library kernels (cuDNN convolutions, GroupNorm, SDPA) -- it is API coverage, not a hot path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _norm(channels: int, norm_type: str = 'group') -> nn.Module:
    """xqgan_model.py:672-677"""
    assert norm_type in ['group', 'batch']
    if norm_type == 'group':
        return nn.GroupNorm(num_groups=32, num_channels=channels, eps=1e-6, affine=True)
    return nn.SyncBatchNorm(channels)


class ResnetBlock(nn.Module):
    """norm -> swish -> conv3x3, twice, plus a (1x1 | 3x3) shortcut when the width changes (:587-622)."""

    def __init__(self, in_channels, out_channels=None, conv_shortcut=False, dropout=0.0, norm_type='group'):
        super().__init__()
        out_channels = out_channels or in_channels
        self.in_channels, self.out_channels, self.use_conv_shortcut = in_channels, out_channels, conv_shortcut
        self.norm1 = _norm(in_channels, norm_type)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.norm2 = _norm(out_channels, norm_type)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            if conv_shortcut:
                self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
            else:
                self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(self.dropout(F.silu(self.norm2(h))))
        if self.in_channels != self.out_channels:
            x = self.conv_shortcut(x) if self.use_conv_shortcut else self.nin_shortcut(x)
        return x + h


class AttnBlock(nn.Module):
    """single-head spatial self-attention over the h*w positions with 1x1-conv projections (:625-660)."""

    def __init__(self, in_channels, norm_type='group'):
        super().__init__()
        self.norm = _norm(in_channels, norm_type)
        self.q = nn.Conv2d(in_channels, in_channels, 1)
        self.k = nn.Conv2d(in_channels, in_channels, 1)
        self.v = nn.Conv2d(in_channels, in_channels, 1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, 1)

    def forward(self, x):
        b, c, h, w = x.shape
        y = self.norm(x)
        # tokens = positions, one head of width c; softmax(q k^T / sqrt(c)) v
        q, k, v = (t(y).flatten(2).transpose(1, 2).unsqueeze(1) for t in (self.q, self.k, self.v))
        o = F.scaled_dot_product_attention(q, k, v)          # default scale = c ** -0.5
        o = o.squeeze(1).transpose(1, 2).reshape(b, c, h, w)
        return x + self.proj_out(o)


class Upsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, 3, 1, 1)

    def forward(self, x):
        x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return self.conv(x) if self.with_conv else x


class Downsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, 3, 2, 0)   # asymmetric (0,1,0,1) padding done by hand

    def forward(self, x):
        if self.with_conv:
            return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))
        return F.avg_pool2d(x, kernel_size=2, stride=2)


def _level(block_in, block_out, n_blocks, with_attn, dropout, norm_type):
    lvl = nn.Module()
    lvl.res, lvl.attn = nn.ModuleList(), nn.ModuleList()
    for _ in range(n_blocks):
        lvl.res.append(ResnetBlock(block_in, block_out, dropout=dropout, norm_type=norm_type))
        block_in = block_out
        if with_attn:
            lvl.attn.append(AttnBlock(block_in, norm_type))
    return lvl, block_in


def _mid(ch, dropout, norm_type):
    return nn.ModuleList([ResnetBlock(ch, ch, dropout=dropout, norm_type=norm_type), AttnBlock(ch, norm_type=norm_type),
                          ResnetBlock(ch, ch, dropout=dropout, norm_type=norm_type)])


def _run_level(lvl, h):
    for i, res in enumerate(lvl.res):
        h = res(h)
        if len(lvl.attn) > 0:
            h = lvl.attn[i](h)
    return h


class Encoder(nn.Module):
    """:454-515 -- image [B,3,H,W] -> [B, z_channels, H / 2^(L-1), W / 2^(L-1)]"""

    def __init__(self, in_channels=3, ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, norm_type='group', dropout=0.0,
                 resamp_with_conv=True, z_channels=256):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.conv_in = nn.Conv2d(in_channels, ch, 3, 1, 1)
        widths = (1,) + tuple(ch_mult)
        self.conv_blocks = nn.ModuleList()
        block_in = ch
        for lvl_i in range(self.num_resolutions):
            last = lvl_i == self.num_resolutions - 1
            lvl, block_in = _level(ch * widths[lvl_i], ch * ch_mult[lvl_i], num_res_blocks, last, dropout, norm_type)
            if not last:
                lvl.downsample = Downsample(block_in, resamp_with_conv)
            self.conv_blocks.append(lvl)
        self.mid = _mid(block_in, dropout, norm_type)
        self.norm_out = _norm(block_in, norm_type)
        self.conv_out = nn.Conv2d(block_in, z_channels, 3, 1, 1)

    def forward(self, x):
        h = self.conv_in(x)
        for lvl_i, lvl in enumerate(self.conv_blocks):
            h = _run_level(lvl, h)
            if lvl_i != self.num_resolutions - 1:
                h = lvl.downsample(h)
        for blk in self.mid:
            h = blk(h)
        return self.conv_out(F.silu(self.norm_out(h)))


class Decoder(nn.Module):
    """:518-584 -- latent [B, z_channels, h, w] -> image"""

    def __init__(self, z_channels=256, ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, norm_type="group", dropout=0.0,
                 resamp_with_conv=True, out_channels=3):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        block_in = ch * ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, 1, 1)
        self.mid = _mid(block_in, dropout, norm_type)
        self.conv_blocks = nn.ModuleList()
        for lvl_i in reversed(range(self.num_resolutions)):
            lvl, block_in = _level(block_in, ch * ch_mult[lvl_i], num_res_blocks + 1, lvl_i == self.num_resolutions - 1,
                                   dropout, norm_type)
            if lvl_i != 0:
                lvl.upsample = Upsample(block_in, resamp_with_conv)
            self.conv_blocks.append(lvl)
        self.norm_out = _norm(block_in, norm_type)
        self.conv_out = nn.Conv2d(block_in, out_channels, 3, 1, 1)

    @property
    def last_layer(self):
        return self.conv_out.weight

    def forward(self, z):
        h = self.conv_in(z)
        for blk in self.mid:
            h = blk(h)
        for i, lvl in enumerate(self.conv_blocks):
            h = _run_level(lvl, h)
            if i != self.num_resolutions - 1:
                h = lvl.upsample(h)
        return self.conv_out(F.silu(self.norm_out(h)))
