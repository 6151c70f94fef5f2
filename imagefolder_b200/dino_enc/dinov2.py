"""DINOv2Encoder / DINOv2Decoder -- drop-in for tokenizer/tokenizer_image/dino_enc/dinov2.py
(:18 and :201): ViT backbone + learnable latent tokens, level embedding, mask tokens, ToPixel.

Only tuning_method 'full' / 'frozen' are built (peft LoRA variants need the un-vendored peft
package and are not selected by any shipped config, xqgan_model.py:96,114).
Sub-module / parameter names match the reference so released checkpoints load unchanged.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ..vit_ops import assemble_tokens, patch_embed, run_blocks
from .to_pixel import ToPixel
from .vision_transformer import Attention, create_model, trunc_normal_

_NAMES = ['vit_small_patch14_dinov2.lvd142m', 'vit_base_patch14_dinov2.lvd142m', 'vit_large_patch14_dinov2.lvd142m']


def _autocast_off(x):
    return torch.autocast(device_type=x.device.type, enabled=False)


def _main_dtype(x):
    temp = x.new_ones(8, 8)
    return torch.matmul(temp, temp).dtype


class DINOv2Encoder(nn.Module):
    def __init__(self, in_channels=3, num_latent_tokens=32, use_attn_mask=False,
                 model_name='vit_small_patch14_dinov2.lvd142m',
                 model_kwargs={'img_size': 224, 'patch_size': 14, 'drop_path_rate': 0.0, },
                 pretrained=True, tuning_method='lora', tuning_kwargs={'r': 8}, abs_pos_embed=False, product_quant=1):
        super().__init__()
        assert model_name in _NAMES, f"{model_name} not found"
        self.num_latent_tokens = num_latent_tokens
        self.use_attn_mask = use_attn_mask
        self.product_quant = product_quant
        model = create_model(model_name, pretrained=pretrained, **model_kwargs)
        self.embed_dim = model.embed_dim
        self.num_img_tokens = model.patch_embed.num_patches
        self.num_prefix_tokens = model.num_prefix_tokens
        self.abs_pos_embed = abs_pos_embed
        if tuning_method == 'full':
            self.model = model
        elif tuning_method == 'frozen':
            for param in model.parameters():
                param.requires_grad = False
            self.model = model
        else:
            raise NotImplementedError(f"tuning_method={tuning_method!r} needs peft (LoRA); not built")

        if self.num_latent_tokens:
            self.latent_tokens = nn.Parameter(torch.zeros(1, self.num_latent_tokens, model.embed_dim))
            nn.init.normal_(self.latent_tokens, std=1e-6)
            if self.abs_pos_embed:
                patch_size = model_kwargs['patch_size']
                if self.product_quant > 1:
                    self.lvl_embed = nn.Embedding(1 + self.product_quant, model.embed_dim)
                    nn.init.trunc_normal_(self.lvl_embed.weight.data, mean=0, std=math.sqrt(1 / model.embed_dim / 3))
                    lvl1LC = torch.cat([torch.full((patch_size * patch_size + 1,), 0), ] +
                                       [torch.full((self.num_latent_tokens // self.product_quant,), i + 1)
                                        for i in range(self.product_quant)]).view(1, -1)
                else:
                    self.lvl_embed = nn.Embedding(2, model.embed_dim)
                    nn.init.trunc_normal_(self.lvl_embed.weight.data, mean=0, std=math.sqrt(1 / model.embed_dim / 3))
                    lvl1LC = torch.cat([torch.full((patch_size * patch_size + 1,), 0),
                                        torch.full((self.num_latent_tokens,), 1)]).view(1, -1)
                self.register_buffer('lvl1LC', lvl1LC)
            else:
                self.latent_pos_embed = nn.Parameter(torch.zeros(1, self.num_latent_tokens, model.embed_dim))
                trunc_normal_(self.latent_pos_embed, std=.02)
            if self.use_attn_mask:
                total_length = self.num_img_tokens + self.num_latent_tokens + self.num_prefix_tokens
                attn_mask = torch.zeros((total_length, total_length))
                attn_mask[:self.num_prefix_tokens + self.num_img_tokens, -self.num_latent_tokens:] = -torch.inf
                self.register_buffer('attn_mask', attn_mask.view(1, 1, total_length, total_length))

    def finetine(self, tuning_method, tuning_kwargs={'r': 8}):
        if tuning_method == 'full':
            return
        if tuning_method == 'frozen':
            for param in self.model.parameters():
                param.requires_grad = False
            return
        raise NotImplementedError(f"tuning_method={tuning_method!r} needs peft (LoRA); not built")

    def no_weight_decay(self):
        return ['model.pos_embed', 'model.cls_token', 'model.dist_token', 'latent_tokens', 'latent_pos_embed']

    def _assembly_is_static(self):
        """no stochastic op between the patch tokens and the blocks (pos_drop / patch_drop inactive)"""
        m = self.model
        return (isinstance(m.patch_drop, nn.Identity)
                and (not self.training or getattr(m.pos_drop, "p", 0.0) == 0.0) and not m.no_embed_class)

    def _assemble(self, x):
        """dinov2.py:151-170: [cls | patch tokens] + pos-embed, then the latent tokens (+ their resampled pos-embed) and the
        level embedding.  x: patch tokens [B, N, D] -> fp32 [B, 1 + N + L, D]."""
        with _autocast_off(x):
            x = self.model._pos_embed(x)
            x = self.model.patch_drop(x)
            if self.num_latent_tokens:
                z = self.latent_tokens.expand(x.size(0), -1, -1)
                if self.abs_pos_embed:
                    if self.product_quant > 1:
                        H = W = int(math.sqrt(self.num_latent_tokens // self.product_quant))
                        assert H * W == self.num_latent_tokens // self.product_quant
                        z = z.view(x.size(0), self.product_quant * H, W, -1)
                        z_list = z.chunk(chunks=self.product_quant, dim=1)
                        z_list = [self.model._pos_embed(z)[:, 1:, ] for z in z_list]  # remove cls token
                        x = torch.cat([x, ] + z_list, dim=1)
                        x += self.lvl_embed(self.lvl1LC.expand(x.size(0), -1))
                    else:
                        H = W = int(math.sqrt(self.num_latent_tokens))
                        assert H * W == self.num_latent_tokens
                        z = z.view(x.size(0), H, W, -1)
                        z = self.model._pos_embed(z)[:, 1:, ]
                        x = torch.cat([x, z], dim=1)
                        x += self.lvl_embed(self.lvl1LC.expand(x.size(0), -1))
                else:
                    x = torch.cat([x, z + self.latent_pos_embed], dim=1)
        return x

    def forward(self, x, masks=None):
        """dinov2.py:146-198 -> [B, num_latent_tokens, D]"""
        x = patch_embed(self.model.patch_embed, x)
        if self._assembly_is_static():
            x = assemble_tokens(self, self._assemble, x, self.num_prefix_tokens)
        else:
            x = self._assemble(x)
        # norm_pre -> blocks -> norm (dinov2.py:176-190); fused CUDA glue under bf16 autocast
        x = run_blocks(self.model, x, self.attn_mask if self.use_attn_mask else None)
        if self.num_latent_tokens:
            return x[:, -self.num_latent_tokens:]
        return x[:, self.num_prefix_tokens:]


class DINOv2Decoder(nn.Module):
    def __init__(self, in_channels=3, model_name='vit_small_patch14_dinov2.lvd142m',
                 model_kwargs={'img_size': 224, 'patch_size': 14, 'drop_path_rate': 0.0}, pretrained=True,
                 tuning_method='lora', tuning_kwargs={'r': 8}, num_latent_tokens=32, to_pixel='linear', use_rope=False,
                 cond_latent=False, abs_pos_embed=False):
        super().__init__()
        assert model_name in _NAMES
        if use_rope:
            raise NotImplementedError("use_rope=True (RoPEAttention) is not selected by any shipped config; not built")
        if cond_latent:
            raise NotImplementedError("cond_latent=True is not selected by any shipped config; not built")
        model_kwargs = dict(model_kwargs)
        model_kwargs['num_latent_tokens'] = num_latent_tokens
        model_kwargs['attn_layer'] = Attention
        model = create_model(model_name, pretrained=pretrained, **model_kwargs)
        self.use_rope = use_rope
        self.embed_dim = model.embed_dim
        self.num_img_tokens = model.patch_embed.num_patches
        self.num_prefix_tokens = model.num_prefix_tokens
        self.num_latent_tokens = num_latent_tokens
        self.abs_pos_embed = abs_pos_embed
        if tuning_method == 'full':
            self.model = model
        elif tuning_method == 'frozen':
            for param in model.parameters():
                param.requires_grad = False
            self.model = model
        else:
            raise NotImplementedError(f"tuning_method={tuning_method!r} needs peft (LoRA); not built")
        self.mask_token = nn.Parameter(torch.zeros(1, 1, model.embed_dim))
        nn.init.normal_(self.mask_token, std=1e-6)
        if self.abs_pos_embed:
            self.lvl_embed = nn.Embedding(2, model.embed_dim)
            patch_size = model_kwargs['patch_size']
            nn.init.trunc_normal_(self.lvl_embed.weight.data, mean=0, std=math.sqrt(1 / model.embed_dim / 3))
            lvl1LC = torch.cat([torch.full((patch_size * patch_size + 1,), 0),
                                torch.full((self.num_latent_tokens + 1,), 1)]).view(1, -1)
            self.register_buffer('lvl1LC', lvl1LC)
        else:
            self.latent_pos_embed = nn.Parameter(torch.zeros(1, self.num_latent_tokens, model.embed_dim))
            trunc_normal_(self.latent_pos_embed, std=.02)
        self.to_pixel = ToPixel(to_pixel=to_pixel, img_size=model_kwargs['img_size'], in_channels=in_channels,
                                in_dim=model.embed_dim, patch_size=model_kwargs['patch_size'])
        self.cond_latent = cond_latent
        del self.model.patch_embed.proj.bias
        del self.model.patch_embed.proj.weight

    def finetine(self, tuning_method, tuning_kwargs={'r': 8}):
        if tuning_method == 'full':
            return
        if tuning_method == 'frozen':
            for param in self.model.parameters():
                param.requires_grad = False
            return
        raise NotImplementedError(f"tuning_method={tuning_method!r} needs peft (LoRA); not built")

    def no_weight_decay(self):
        return ['model.pos_embed', 'model.cls_token', 'model.dist_token', 'mask_token', 'latent_pos_embed']

    @property
    def last_layer(self):
        return self.to_pixel.model.weight

    def _assemble(self, z):
        """dinov2.py:318-336: [cls | mask tokens] + pos-embed, then the latents (with their own cls slot and resampled
        pos-embed when abs_pos_embed) and the level embedding.  z [B, L, D] -> fp32 [B, T, D]."""
        x = self.mask_token.expand(z.size(0), self.num_img_tokens, -1)
        with _autocast_off(x):
            x = self.model._pos_embed(x)
            if self.abs_pos_embed:
                H = W = int(math.sqrt(self.num_latent_tokens))
                assert H * W == self.num_latent_tokens
                z = z.view(x.size(0), H, W, -1)
                z = self.model._pos_embed(z)  # NB: keeps the cls slot (L+1 tokens), dinov2.py:330
            else:
                z = z + self.latent_pos_embed
            x = self.model.patch_drop(x)
            x = torch.cat([x, z], dim=1)
            if self.abs_pos_embed:
                x += self.lvl_embed(self.lvl1LC.expand(x.size(0), -1))
        return x

    def forward(self, z):
        """dinov2.py:313-365: z [B, L, D] -> image [B, 3, H, W]"""
        m = self.model
        static = isinstance(m.patch_drop, nn.Identity) and (not self.training or getattr(m.pos_drop, "p", 0.0) == 0.0) \
            and not m.no_embed_class
        if static:
            t0 = self.num_img_tokens + self.num_prefix_tokens + (self.num_prefix_tokens if self.abs_pos_embed else 0)
            x = assemble_tokens(self, self._assemble, z, t0)
        else:
            x = self._assemble(z)
        x = run_blocks(self.model, x)
        x = x[:, self.num_prefix_tokens:self.num_img_tokens + self.num_prefix_tokens]
        return self.to_pixel(x)
