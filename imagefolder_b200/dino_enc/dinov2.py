"""DINOv2Encoder / DINOv2Decoder -- drop-in for tokenizer/tokenizer_image/dino_enc/dinov2.py
(:18 and :201): ViT backbone + learnable latent tokens, level embedding, mask tokens, ToPixel.

Only tuning_method 'full' / 'frozen' are built (peft LoRA variants need the un-vendored peft
package and are not selected by any shipped config, xqgan_model.py:96,114).
Sub-module / parameter names match the reference so released checkpoints load unchanged.

Both classes split their forward into (1) the token ASSEMBLY -- everything between the batch-dependent rows (patch
tokens / quantised latents) and the first transformer block: prefix tokens, positional embeddings, latent or mask
tokens, level embedding -- and (2) the blocks.  (1) is affine in the batch-dependent rows, so under bf16 autocast it
runs as `table[t] + src[b, t - t0]` in one fused pass (vit_ops.assemble_tokens, which checks that property once per
module); (2) runs on the fused ViT glue (vit_ops.run_blocks).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ..vit_ops import assemble_tokens, patch_embed, run_blocks
from .to_pixel import ToPixel
from .vision_transformer import Attention, create_model, trunc_normal_

_NAMES = ['vit_small_patch14_dinov2.lvd142m', 'vit_base_patch14_dinov2.lvd142m', 'vit_large_patch14_dinov2.lvd142m']
_LORA_MSG = "tuning_method={!r} needs peft (LoRA); not built"


def _autocast_off(x):
    return torch.autocast(device_type=x.device.type, enabled=False)


def _main_dtype(x):
    temp = x.new_ones(8, 8)
    return torch.matmul(temp, temp).dtype


def _freeze(module):
    for param in module.parameters():
        param.requires_grad = False


def _adopt_backbone(owner, model, tuning_method):
    """`self.model = model` for 'full', the same with frozen parameters for 'frozen' (dinov2.py:41-62 / 228-249)."""
    if tuning_method not in ('full', 'frozen'):
        raise NotImplementedError(_LORA_MSG.format(tuning_method))
    if tuning_method == 'frozen':
        _freeze(model)
    owner.model = model
    owner.embed_dim = model.embed_dim
    owner.num_img_tokens = model.patch_embed.num_patches
    owner.num_prefix_tokens = model.num_prefix_tokens


def _level_embedding(owner, n_levels, dim, segment_lengths):
    """`lvl_embed` (trunc-normal, std sqrt(1/3D)) + the `lvl1LC` index row: segment i of the sequence gets level i."""
    owner.lvl_embed = nn.Embedding(n_levels, dim)
    nn.init.trunc_normal_(owner.lvl_embed.weight.data, mean=0, std=math.sqrt(1 / dim / 3))
    idx = torch.cat([torch.full((n,), lvl) for lvl, n in enumerate(segment_lengths)])
    owner.register_buffer('lvl1LC', idx.view(1, -1))


def _square_side(n):
    side = int(math.sqrt(n))
    assert side * side == n
    return side


def _static_sequence(vit, training):
    """nothing stochastic between the batch-dependent rows and the first block (patch_drop / pos_drop inactive)"""
    return (isinstance(vit.patch_drop, nn.Identity) and not vit.no_embed_class
            and (not training or getattr(vit.pos_drop, "p", 0.0) == 0.0))


class _Tunable:
    def finetine(self, tuning_method, tuning_kwargs={'r': 8}):            # (sic) the reference's spelling
        if tuning_method == 'frozen':
            _freeze(self.model)
        elif tuning_method != 'full':
            raise NotImplementedError(_LORA_MSG.format(tuning_method))


class DINOv2Encoder(_Tunable, nn.Module):
    def __init__(self, in_channels=3, num_latent_tokens=32, use_attn_mask=False,
                 model_name='vit_small_patch14_dinov2.lvd142m',
                 model_kwargs={'img_size': 224, 'patch_size': 14, 'drop_path_rate': 0.0, },
                 pretrained=True, tuning_method='lora', tuning_kwargs={'r': 8}, abs_pos_embed=False, product_quant=1):
        super().__init__()
        assert model_name in _NAMES, f"{model_name} not found"
        self.num_latent_tokens, self.use_attn_mask = num_latent_tokens, use_attn_mask
        self.product_quant, self.abs_pos_embed = product_quant, abs_pos_embed
        _adopt_backbone(self, create_model(model_name, pretrained=pretrained, **model_kwargs), tuning_method)
        if not num_latent_tokens:
            return
        D, L = self.embed_dim, num_latent_tokens
        self.latent_tokens = nn.Parameter(torch.zeros(1, L, D))
        nn.init.normal_(self.latent_tokens, std=1e-6)
        if abs_pos_embed:
            # level 0 = [cls | image tokens] -- the reference sizes it as patch_size^2 + 1 (dinov2.py:72,80), i.e. it assumes
            # a 16 x 16 token grid for patch 16 -- then one level per product-quantisation branch
            n_img = model_kwargs['patch_size'] ** 2 + 1
            _level_embedding(self, 1 + product_quant, D, [n_img] + [L // product_quant] * product_quant)
        else:
            self.latent_pos_embed = nn.Parameter(torch.zeros(1, L, D))
            trunc_normal_(self.latent_pos_embed, std=.02)
        if use_attn_mask:                        # image tokens must not look at the latent tokens (dinov2.py:95-101)
            n_front = self.num_prefix_tokens + self.num_img_tokens
            mask = torch.zeros(n_front + L, n_front + L)
            mask[:n_front, -L:] = -torch.inf
            self.register_buffer('attn_mask', mask[None, None])

    def no_weight_decay(self):
        return ['model.pos_embed', 'model.cls_token', 'model.dist_token', 'latent_tokens', 'latent_pos_embed']

    def _latent_rows(self, batch):
        """the latent tokens with their positional term: 2-D pos-embed resampled to each branch's grid when abs_pos_embed
        (the cls row that _pos_embed prepends is dropped, dinov2.py:160-166), else the learned latent_pos_embed"""
        z = self.latent_tokens.expand(batch, -1, -1)
        if not self.abs_pos_embed:
            return [z + self.latent_pos_embed]
        side = _square_side(self.num_latent_tokens // self.product_quant)
        grids = z.view(batch, self.product_quant * side, side, -1).chunk(chunks=self.product_quant, dim=1)
        return [self.model._pos_embed(g)[:, 1:] for g in grids]

    def _assemble(self, x):
        """dinov2.py:151-170.  x: patch tokens [B, N, D] -> fp32 [B, prefix + N + L, D]."""
        with _autocast_off(x):
            x = self.model.patch_drop(self.model._pos_embed(x))
            if self.num_latent_tokens:
                x = torch.cat([x] + self._latent_rows(x.size(0)), dim=1)
                if self.abs_pos_embed:
                    x += self.lvl_embed(self.lvl1LC.expand(x.size(0), -1))
        return x

    def forward(self, x, masks=None):
        """dinov2.py:146-198 -> [B, num_latent_tokens, D]"""
        x = patch_embed(self.model.patch_embed, x)
        if _static_sequence(self.model, self.training):
            x = assemble_tokens(self, self._assemble, x, self.num_prefix_tokens)
        else:
            x = self._assemble(x)
        # norm_pre -> blocks -> norm (dinov2.py:176-190); fused CUDA glue under bf16 autocast
        x = run_blocks(self.model, x, self.attn_mask if self.use_attn_mask else None)
        return x[:, -self.num_latent_tokens:] if self.num_latent_tokens else x[:, self.num_prefix_tokens:]


class DINOv2Decoder(_Tunable, nn.Module):
    def __init__(self, in_channels=3, model_name='vit_small_patch14_dinov2.lvd142m',
                 model_kwargs={'img_size': 224, 'patch_size': 14, 'drop_path_rate': 0.0}, pretrained=True,
                 tuning_method='lora', tuning_kwargs={'r': 8}, num_latent_tokens=32, to_pixel='linear', use_rope=False,
                 cond_latent=False, abs_pos_embed=False):
        super().__init__()
        assert model_name in _NAMES
        for flag, name in ((use_rope, "use_rope=True (RoPEAttention)"), (cond_latent, "cond_latent=True")):
            if flag:
                raise NotImplementedError(f"{name} is not selected by any shipped config; not built")
        self.use_rope, self.cond_latent = use_rope, cond_latent
        self.num_latent_tokens, self.abs_pos_embed = num_latent_tokens, abs_pos_embed
        vit_kwargs = dict(model_kwargs, num_latent_tokens=num_latent_tokens, attn_layer=Attention)
        _adopt_backbone(self, create_model(model_name, pretrained=pretrained, **vit_kwargs), tuning_method)
        D = self.embed_dim
        self.mask_token = nn.Parameter(torch.zeros(1, 1, D))
        nn.init.normal_(self.mask_token, std=1e-6)
        if abs_pos_embed:
            # level 0 = [cls | mask tokens], level 1 = the latents WITH the cls slot _pos_embed gives them (dinov2.py:266-272)
            _level_embedding(self, 2, D, [model_kwargs['patch_size'] ** 2 + 1, num_latent_tokens + 1])
        else:
            self.latent_pos_embed = nn.Parameter(torch.zeros(1, num_latent_tokens, D))
            trunc_normal_(self.latent_pos_embed, std=.02)
        self.to_pixel = ToPixel(to_pixel=to_pixel, img_size=model_kwargs['img_size'], in_channels=in_channels, in_dim=D,
                                patch_size=model_kwargs['patch_size'])
        # the decoder never embeds pixels: drop the unused projection so that it is neither trained nor checkpointed
        del self.model.patch_embed.proj.bias
        del self.model.patch_embed.proj.weight

    def no_weight_decay(self):
        return ['model.pos_embed', 'model.cls_token', 'model.dist_token', 'mask_token', 'latent_pos_embed']

    @property
    def last_layer(self):
        return self.to_pixel.model.weight

    def _assemble(self, z):
        """dinov2.py:318-336.  z: latents [B, L, D] -> fp32 [B, prefix + N_img + (prefix if abs_pos_embed) + L, D]."""
        masks = self.mask_token.expand(z.size(0), self.num_img_tokens, -1)
        with _autocast_off(masks):
            front = self.model._pos_embed(masks)
            if self.abs_pos_embed:
                side = _square_side(self.num_latent_tokens)
                back = self.model._pos_embed(z.view(z.size(0), side, side, -1))   # keeps its cls slot (L + 1 rows), :330
            else:
                back = z + self.latent_pos_embed
            x = torch.cat([self.model.patch_drop(front), back], dim=1)
            if self.abs_pos_embed:
                x += self.lvl_embed(self.lvl1LC.expand(x.size(0), -1))
        return x

    def forward(self, z):
        """dinov2.py:313-365: z [B, L, D] -> image [B, 3, H, W]"""
        if _static_sequence(self.model, self.training):
            n_front = self.num_prefix_tokens + self.num_img_tokens
            t0 = n_front + (self.num_prefix_tokens if self.abs_pos_embed else 0)      # where the latent rows start
            x = assemble_tokens(self, self._assemble, z, t0)
        else:
            x = self._assemble(z)
        x = run_blocks(self.model, x)
        return self.to_pixel(x[:, self.num_prefix_tokens:self.num_prefix_tokens + self.num_img_tokens])
