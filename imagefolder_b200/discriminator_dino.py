"""DINO-feature discriminator (StyleGAN-T style) -- tokenizer/tokenizer_image/discriminator_dino.py:152-362.

A frozen DINO ViT-S/16 (eval mode, no dropout) produces token activations at the input and after blocks 2 / 5 / 8 / 11;
each activation [B, 384, 196] feeds a small trainable head (spectral-norm Conv1d k=1 -> residual spectral-norm Conv1d
k=9, circular -> spectral-norm Conv1d to 1 channel, with BatchNormLocal + LeakyReLU), and the five head outputs are
concatenated into the logits [B, 5*196].

Parameter / buffer names equal the reference's (checkpoint keys `heads.{i}.{0|1.fn|2}....`; the frozen backbone lives in a
tuple and is deliberately NOT part of the state dict, as in the reference :190).  No network: the DINO checkpoint is read
from `$XQ_DINO_CKPT` or the torch-hub cache when it is there, otherwise the backbone keeps its random initialisation (with
a warning).  All layers here are library kernels; the hand-written parts of the loss stack are LPIPS and DiffAug.
"""
from __future__ import annotations

import math
import os
import random
import warnings
from typing import List

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.spectral_norm import SpectralNorm

_DINO_URL = 'https://dl.fbaipublicfiles.com/dino/dino_deitsmall16_pretrain/dino_deitsmall16_pretrain.pth'


class MLPNoDrop(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, fused_if_available=True):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.act = nn.GELU(approximate='tanh')
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class SelfAttentionNoDrop(nn.Module):
    def __init__(self, block_idx, embed_dim=768, num_heads=12, flash_if_available=True):
        super().__init__()
        assert embed_dim % num_heads == 0
        self.block_idx, self.num_heads, self.head_dim = block_idx, num_heads, embed_dim // num_heads
        self.scale = 1 / math.sqrt(self.head_dim)
        self.qkv = nn.Linear(embed_dim, embed_dim * 3, bias=True)
        self.proj = nn.Linear(embed_dim, embed_dim, bias=True)

    def forward(self, x):
        B, L, C = x.shape
        q, k, v = self.qkv(x).view(B, L, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4).unbind(0)
        o = F.scaled_dot_product_attention(q, k, v, scale=self.scale)
        return self.proj(o.transpose(1, 2).reshape(B, L, C))


class SABlockNoDrop(nn.Module):
    def __init__(self, block_idx, embed_dim, num_heads, mlp_ratio, norm_eps):
        super().__init__()
        self.norm1 = nn.LayerNorm(embed_dim, eps=norm_eps)
        self.attn = SelfAttentionNoDrop(block_idx, embed_dim, num_heads)
        self.norm2 = nn.LayerNorm(embed_dim, eps=norm_eps)
        self.mlp = MLPNoDrop(embed_dim, round(embed_dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class ResidualBlock(nn.Module):
    def __init__(self, fn):
        super().__init__()
        self.fn = fn
        self.ratio = 1 / np.sqrt(2)

    def forward(self, x):
        return (self.fn(x).add(x)).mul_(self.ratio)


class SpectralConv1d(nn.Conv1d):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        SpectralNorm.apply(self, name='weight', n_power_iterations=1, dim=0, eps=1e-12)


class BatchNormLocal(nn.Module):
    """batch statistics over groups of `virtual_bs` samples (no cross-rank sync; discriminator_dino.py:122-149)."""

    def __init__(self, num_features: int, affine: bool = True, virtual_bs: int = 8, eps: float = 1e-6):
        super().__init__()
        self.virtual_bs, self.eps, self.affine = virtual_bs, eps, affine
        if affine:
            self.weight = nn.Parameter(torch.ones(num_features))
            self.bias = nn.Parameter(torch.zeros(num_features))

    def forward(self, x):
        shape = x.size()
        x = x.float()
        groups = int(np.ceil(x.size(0) / self.virtual_bs))
        x = x.view(groups, -1, x.size(-2), x.size(-1))
        mean = x.mean([1, 3], keepdim=True)
        var = x.var([1, 3], keepdim=True, unbiased=False)
        x = (x - mean) / torch.sqrt(var + self.eps)
        if self.affine:
            x = x * self.weight[None, :, None] + self.bias[None, :, None]
        return x.view(shape)


def make_block(channels: int, kernel_size: int, norm_type: str, norm_eps: float, using_spec_norm: bool) -> nn.Module:
    if norm_type == 'bn':
        norm = BatchNormLocal(channels, eps=norm_eps)
    elif norm_type == 'sbn':
        norm = nn.SyncBatchNorm(channels, eps=norm_eps, process_group=None)
    elif norm_type == 'gn':
        norm = nn.GroupNorm(num_groups=32, num_channels=channels, eps=norm_eps, affine=True)
    else:
        raise NotImplementedError(f"norm_type {norm_type!r} (the reference's 'lbn'/'hbn' need its node-local process groups)")
    conv = SpectralConv1d if using_spec_norm else nn.Conv1d
    return nn.Sequential(conv(channels, channels, kernel_size=kernel_size, padding=kernel_size // 2, padding_mode='circular'),
                         norm, nn.LeakyReLU(negative_slope=0.2, inplace=True))


def _make_head(C, ks, norm_type, norm_eps, using_spec_norm):
    conv = SpectralConv1d if using_spec_norm else nn.Conv1d
    return nn.Sequential(make_block(C, 1, norm_type, norm_eps, using_spec_norm),
                         ResidualBlock(make_block(C, ks, norm_type, norm_eps, using_spec_norm)),
                         conv(C, 1, kernel_size=1, padding=0))


class PatchEmbed(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None, flatten=True):
        super().__init__()
        self.img_size, self.patch_size, self.flatten = img_size, patch_size, flatten
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

    def forward(self, x):
        return self.norm(self.proj(x).flatten(2).transpose(1, 2))


class FrozenDINOSmallNoDrop(nn.Module):
    """frozen ViT-S/16 @ 224 returning the 'readout' activations (patch tokens + cls token) [B, C, L] at the input and
    after every block in `key_depths` (discriminator_dino.py:269-343)."""

    def __init__(self, depth=12, key_depths=(2, 5, 8, 11), norm_eps=1e-6, patch_size=16, in_chans=3, num_classes=0,
                 embed_dim=384, num_heads=6, mlp_ratio=4.):
        super().__init__()
        self.num_classes, self.num_features, self.embed_dim = num_classes, embed_dim, embed_dim
        self.img_size, self.patch_size = 224, patch_size
        self.patch_nums = self.img_size // patch_size
        self.patch_embed = PatchEmbed(self.img_size, patch_size, in_chans, embed_dim)
        mean, std = torch.tensor((0.485, 0.456, 0.406)), torch.tensor((0.229, 0.224, 0.225))
        self.register_buffer('x_scale', (0.5 / std).reshape(1, 3, 1, 1))          # [-1,1] -> ImageNet-normalised
        self.register_buffer('x_shift', ((0.5 - mean) / std).reshape(1, 3, 1, 1))
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.dist_token = None
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_nums ** 2 + 1, embed_dim))
        self.key_depths = set(d for d in key_depths if d < depth)
        self.blocks = nn.Sequential(*[SABlockNoDrop(i, embed_dim, num_heads, mlp_ratio, norm_eps)
                                      for i in range(max(depth, 1 + max(self.key_depths)))])
        self.norm = nn.LayerNorm(embed_dim, eps=norm_eps)
        self.eval()
        for p in self.parameters():
            p.requires_grad_(False)

    def _to_224(self, x):
        H, W = x.shape[-2:]
        if H > self.img_size and W > self.img_size and random.random() <= 0.5:
            from torchvision.transforms import RandomCrop
            return RandomCrop(self.img_size)(x)
        return F.interpolate(x, size=(self.img_size, self.img_size), mode='area' if H > self.img_size else 'bicubic')

    def forward(self, x, grad_ckpt=False) -> List[torch.Tensor]:
        with torch.autocast(x.device.type, enabled=False):
            x = self._to_224((self.x_scale * x.float()).add_(self.x_shift))
        x = self.patch_embed(x)
        with torch.autocast(x.device.type, enabled=False):
            x = torch.cat((self.cls_token.expand(x.shape[0], -1, -1), x.float()), dim=1) + self.pos_embed
            acts = [(x[:, 1:] + x[:, :1]).transpose_(1, 2)]
        for i, blk in enumerate(self.blocks):
            x = torch.utils.checkpoint.checkpoint(blk, x, use_reentrant=False) if grad_ckpt else blk(x)
            if i in self.key_depths:
                acts.append((x[:, 1:].float() + x[:, :1].float()).transpose_(1, 2))
        return acts


def _find_dino_ckpt(path_or_url: str):
    cands = [os.environ.get("XQ_DINO_CKPT", "")]
    if path_or_url and not path_or_url.startswith("http"):
        cands.append(path_or_url)
    cands.append(os.path.join(torch.hub.get_dir(), "checkpoints", os.path.basename(path_or_url or _DINO_URL)))
    return next((c for c in cands if c and os.path.exists(c)), None)


class DinoDisc(nn.Module):
    def __init__(self, dino_ckpt_path=_DINO_URL, device='cuda', ks=9, depth=12, key_depths=(2, 5, 8, 11), norm_type='bn',
                 using_spec_norm=True, norm_eps=1e-6):
        super().__init__()
        key_depths = tuple(d for d in key_depths if d < depth)
        backbone = FrozenDINOSmallNoDrop(depth=depth, key_depths=key_depths, norm_eps=norm_eps)
        ckpt = _find_dino_ckpt(dino_ckpt_path)
        if ckpt is not None:
            state = torch.load(ckpt, map_location='cpu')
            for k in sorted(state.keys()):
                if '.attn.qkv.bias' in k:                      # the reference zeroes the key bias (:166-170)
                    c = state[k].numel() // 3
                    state[k][c:2 * c].zero_()
            res = backbone.load_state_dict(state, strict=False)
            # the reference's own check (discriminator_dino.py:171-175): only the input-normalisation buffers may be
            # absent from the file, and nothing may be left over -- a wrapped / prefixed checkpoint must not pass silently
            missing = [k for k in res.missing_keys if not k.startswith(('x_scale', 'x_shift'))]
            if missing or res.unexpected_keys:
                raise RuntimeError(f"DinoDisc: {ckpt} does not match the frozen DINO ViT-S/16 backbone "
                                   f"(missing {missing[:4]}, unexpected {list(res.unexpected_keys)[:4]})")
        else:
            warnings.warn("DinoDisc: no DINO ViT-S/16 checkpoint found (set XQ_DINO_CKPT); the frozen backbone is RANDOM")
        if device == 'cuda' and not torch.cuda.is_available():
            device = 'cpu'
        self.dino_proxy = (backbone.to(device=device),)       # a tuple: not registered, not in the state dict
        self._head_cfg = dict(ks=ks, norm_type=norm_type, norm_eps=norm_eps, using_spec_norm=using_spec_norm)
        C = backbone.embed_dim
        self.heads = nn.ModuleList([_make_head(C, ks, norm_type, norm_eps, using_spec_norm) for _ in range(len(key_depths) + 1)])

    def reinit(self, *_, **__):
        cfg, C = self._head_cfg, self.dino_proxy[0].embed_dim
        fresh = nn.ModuleList([_make_head(C, cfg['ks'], cfg['norm_type'], cfg['norm_eps'], cfg['using_spec_norm'])
                               for _ in range(len(self.heads))])
        self.heads.load_state_dict(fresh.state_dict())

    def forward(self, x_in_pm1, grad_ckpt=False):
        backbone = self.dino_proxy[0]
        if next(backbone.parameters()).device != x_in_pm1.device:
            self.dino_proxy = (backbone.to(x_in_pm1.device),)
            backbone = self.dino_proxy[0]
        acts = backbone(x_in_pm1.float(), grad_ckpt=grad_ckpt and x_in_pm1.requires_grad)
        B = x_in_pm1.shape[0]
        outs = [(torch.utils.checkpoint.checkpoint(h, a, use_reentrant=False) if grad_ckpt else h(a)).view(B, -1)
                for h, a in zip(self.heads, acts)]
        return torch.cat(outs, dim=1)
