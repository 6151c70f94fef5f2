"""VQModel -- drop-in for tokenizer/tokenizer_image/xqgan_model.py (ModelArgs :31, VQModel :75,
VQ_models :851): encode -> [PQ split] -> quantizer(s) -> [latent perturbation] -> decode, plus the
CLIP-style semantic / detail regularisers.

Same constructor (a ModelArgs dataclass with the same fields and defaults), same sub-module names
(= checkpoint keys: encoder, decoder, quant_conv, post_quant_conv, quantize | quantizes.{i},
semantic_model, sem_linear, detail_model) and the same forward return structure
`(dec, (vq, commit, entropy, usages), sem_loss, detail_loss, dependency_loss)` (:365).

What differs, all host-side and documented in DESIGN.md:
  * quantizers / perturbation are the CUDA implementations of this package;
  * no `print(alpha, beta, delta)` per step (:296);
  * the frozen semantic / detail teachers are built with random weights when no checkpoint is
    available (the reference downloads them, :175,209); `semantic_guide='none'` skips them;
  * enc_type / dec_type 'cnn' (never selected by a shipped YAML) runs on library conv kernels (cnn.py);
  * `img_to_idxBl`, `encode_to_codes`, `decode_tokens` exist (callers in trainer.py:69,122,
    scripts/pretokenization.py:233, demo_util.py:109 expect them; the reference class lacks them).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from math import sqrt
from typing import List

import torch
import torch.distributed as tdist
import torch.distributed.nn  # noqa: F401  (differentiable all_gather)
import torch.nn as nn
import torch.nn.functional as F

from .cnn import Decoder, Encoder
from .dino_enc import DINOv2Decoder, DINOv2Encoder, create_model
from .latent_perturbation import add_perturbation
from .lookup_free_quantize import LFQ
from .quant import VectorQuantizer2
from .vq import VectorQuantizer

__all__ = ["ModelArgs", "VQModel", "VQ_models", "VQ_8", "VQ_16", "VectorQuantizer", "orthogonal_cosine_loss",
           "ClipLoss", "Normalize", "Denormalize"]


@dataclass
class ModelArgs:
    codebook_size: int = 16384
    codebook_embed_dim: int = 8
    codebook_l2_norm: bool = True
    codebook_show_usage: bool = True
    commit_loss_beta: float = 0.25
    entropy_loss_ratio: float = 0.0

    encoder_ch_mult: List[int] = field(default_factory=lambda: [1, 1, 2, 2, 4])
    decoder_ch_mult: List[int] = field(default_factory=lambda: [1, 1, 2, 2, 4])
    z_channels: int = 256
    dropout_p: float = 0.0

    v_patch_nums: List[int] = field(default_factory=lambda: [1, 2, 3, 4, 5, 6, 8, 10, 13, 16])
    enc_type: str = 'cnn'
    dec_type: str = 'cnn'
    semantic_guide: str = 'dinov2'
    detail_guide: str = 'clip'
    num_latent_tokens: int = 256
    encoder_model: str = 'vit_small_patch14_dinov2.lvd142m'
    decoder_model: str = 'vit_small_patch14_dinov2.lvd142m'
    abs_pos_embed: bool = False
    share_quant_resi: int = 4
    product_quant: int = 1
    codebook_drop: float = 0.0
    half_sem: bool = False
    start_drop: int = 1
    sem_loss_weight: float = 0.1
    detail_loss_weight: float = 0.1
    clip_norm: bool = False
    sem_loss_scale: float = 1.0
    detail_loss_scale: float = 1.0
    guide_type_1: str = "class"
    guide_type_2: str = "class"

    lfq: bool = False
    scale: float = 1.0
    soft_entropy: bool = True

    dependency_loss_weight: float = 0.0

    test_model: bool = False


class Normalize(nn.Module):
    """datasets/normalize.py:7 (mean/std follow the module's device)."""

    def __init__(self, mean, std, device=None):
        super().__init__()
        self.register_buffer('mean', torch.tensor(mean).view(1, -1, 1, 1), persistent=False)
        self.register_buffer('std', torch.tensor(std).view(1, -1, 1, 1), persistent=False)

    def forward(self, x):
        return (x - self.mean) / self.std


class Denormalize(nn.Module):
    """datasets/normalize.py:18."""

    def __init__(self, mean, std, device=None):
        super().__init__()
        self.register_buffer('mean', torch.tensor(mean).view(1, -1, 1, 1), persistent=False)
        self.register_buffer('std', torch.tensor(std).view(1, -1, 1, 1), persistent=False)

    def forward(self, x):
        return x * self.std + self.mean


class ClipLoss(nn.Module):
    """Symmetric contrastive loss with a differentiable all-gather -- same contract as
    tokenizer/vqgan/cliploss.py:66-130 (local_loss=False, gather_with_grad=True as VQModel uses it)."""

    def __init__(self, local_loss=False, gather_with_grad=False, cache_labels=False, rank=0, world_size=1,
                 use_horovod=False):
        super().__init__()
        self.local_loss, self.gather_with_grad = local_loss, gather_with_grad
        self.rank, self.world_size = rank, world_size

    def forward(self, image_features, text_features, logit_scale, output_dict=False):
        if self.world_size > 1:
            if self.gather_with_grad:
                all_i = torch.cat(torch.distributed.nn.all_gather(image_features), dim=0)
                all_t = torch.cat(torch.distributed.nn.all_gather(text_features), dim=0)
            else:
                gi = [torch.zeros_like(image_features) for _ in range(self.world_size)]
                gt = [torch.zeros_like(text_features) for _ in range(self.world_size)]
                tdist.all_gather(gi, image_features)
                tdist.all_gather(gt, text_features)
                gi[self.rank], gt[self.rank] = image_features, text_features
                all_i, all_t = torch.cat(gi, dim=0), torch.cat(gt, dim=0)
            logits_per_image = logit_scale * all_i @ all_t.T
            logits_per_text = logits_per_image.T
        else:
            logits_per_image = logit_scale * image_features @ text_features.T
            logits_per_text = logit_scale * text_features @ image_features.T
        labels = torch.arange(logits_per_image.shape[0], device=image_features.device, dtype=torch.long)
        total_loss = (F.cross_entropy(logits_per_image, labels) + F.cross_entropy(logits_per_text, labels)) / 2
        return {"contrastive_loss": total_loss} if output_dict else total_loss


def _dist_rank_world():
    if tdist.is_available() and tdist.is_initialized():
        return tdist.get_rank(), tdist.get_world_size()
    return 0, 1


def orthogonal_cosine_loss(A, B):
    """xqgan_model.py:836-840."""
    A_norm = A / A.norm(dim=1, keepdim=True)
    B_norm = B / B.norm(dim=1, keepdim=True)
    return (A_norm * B_norm).sum(dim=1).mean()


class VQModel(nn.Module):
    def __init__(self, config: ModelArgs, ):
        super().__init__()
        self.config = config
        self.enc_type = config.enc_type
        self.dec_type = config.dec_type
        self.product_quant = config.product_quant
        self.half_sem = config.half_sem
        self.start_drop = config.start_drop
        self.clip_norm = config.clip_norm
        config.num_latent_tokens = config.num_latent_tokens * config.product_quant  # scale num_latent_tokens for PQ

        if config.enc_type == 'dinov2':
            self.encoder = DINOv2Encoder(
                in_channels=3, num_latent_tokens=config.num_latent_tokens, model_name=config.encoder_model,
                model_kwargs={'img_size': 256, 'patch_size': 16, 'drop_path_rate': 0.1}, pretrained=True,
                tuning_method='full', tuning_kwargs={'r': 8}, abs_pos_embed=config.abs_pos_embed,
                product_quant=config.product_quant)
            self.quant_conv = nn.Conv2d(self.encoder.embed_dim, config.codebook_embed_dim, 1)
        elif config.enc_type == 'cnn':
            self.encoder = Encoder(ch_mult=config.encoder_ch_mult, z_channels=config.z_channels, dropout=config.dropout_p)
            self.quant_conv = nn.Conv2d(config.z_channels, config.codebook_embed_dim, 1)
        else:
            raise NotImplementedError

        if config.dec_type == 'dinov2':
            self.decoder = DINOv2Decoder(
                in_channels=3, num_latent_tokens=config.num_latent_tokens // self.product_quant,
                model_name=config.decoder_model,
                model_kwargs={'img_size': 256, 'patch_size': 16, 'drop_path_rate': 0.1}, pretrained=True,
                tuning_method='full', tuning_kwargs={'r': 8}, to_pixel='linear', use_rope=False, cond_latent=False,
                abs_pos_embed=config.abs_pos_embed)
            self.post_quant_conv = nn.Conv2d(config.codebook_embed_dim, self.decoder.embed_dim, 1)
        elif config.dec_type == 'cnn':
            self.decoder = Decoder(ch_mult=config.decoder_ch_mult, z_channels=config.z_channels, dropout=config.dropout_p)
            self.post_quant_conv = nn.Conv2d(config.codebook_embed_dim, config.z_channels, 1)
        else:
            raise NotImplementedError

        self.V = self.vocab_size = config.codebook_size * self.product_quant
        self.Cvae = config.codebook_embed_dim * self.product_quant
        if self.product_quant > 1:
            if len(config.v_patch_nums) == 1:
                self.quantizes = nn.ModuleList([
                    VectorQuantizer(config.codebook_size, config.codebook_embed_dim, config.commit_loss_beta,
                                    config.codebook_l2_norm) for _ in range(self.product_quant)])
            elif not config.lfq:
                self.quantizes = nn.ModuleList([
                    VectorQuantizer2(config.codebook_size, config.codebook_embed_dim, v_patch_nums=config.v_patch_nums,
                                     num_latent_tokens=config.num_latent_tokens // self.product_quant,
                                     share_quant_resi=config.share_quant_resi, codebook_drop=config.codebook_drop, )
                    for _ in range(self.product_quant)])
            else:
                self.quantizes = nn.ModuleList([
                    LFQ(config.codebook_size, config.codebook_embed_dim, v_patch_nums=config.v_patch_nums,
                        num_latent_tokens=config.num_latent_tokens // self.product_quant,
                        share_quant_resi=config.share_quant_resi, codebook_drop=config.codebook_drop,
                        using_znorm=config.codebook_l2_norm, scale=config.scale,
                        entropy_weight=config.entropy_loss_ratio, soft_entropy=config.soft_entropy, )
                    for _ in range(self.product_quant)])
            self.post_quant_conv = nn.Conv2d(config.codebook_embed_dim * self.product_quant,
                                             self.decoder.embed_dim if config.dec_type == 'dinov2' else config.z_channels, 1)
        else:
            if len(config.v_patch_nums) == 1:
                self.quantize = VectorQuantizer(config.codebook_size, config.codebook_embed_dim,
                                                config.commit_loss_beta, config.codebook_l2_norm)
            elif not config.lfq:
                self.quantize = VectorQuantizer2(config.codebook_size, config.codebook_embed_dim,
                                                 v_patch_nums=config.v_patch_nums,
                                                 num_latent_tokens=config.num_latent_tokens,
                                                 share_quant_resi=config.share_quant_resi, )
            else:
                self.quantize = LFQ(config.codebook_size, config.codebook_embed_dim, v_patch_nums=config.v_patch_nums,
                                    num_latent_tokens=config.num_latent_tokens,
                                    share_quant_resi=config.share_quant_resi, codebook_drop=config.codebook_drop,
                                    using_znorm=config.codebook_l2_norm, scale=config.scale,
                                    entropy_weight=config.entropy_loss_ratio, soft_entropy=config.soft_entropy)

        self.codebook_embed_dim = config.codebook_embed_dim
        self.v_patch_nums = config.v_patch_nums
        self.codebook_drop = config.codebook_drop
        self.semantic_guide = config.semantic_guide
        self.denormalize = Denormalize(mean=[0.5, 0.5, 0.5], std=[0.5, 0.5, 0.5])
        self.normalize = Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
        rank, world_size = _dist_rank_world()
        if self.semantic_guide == 'dinov2':
            semantic_model = create_model(config.encoder_model, pretrained=True, img_size=256, patch_size=16,
                                          drop_path_rate=0.0)
            semantic_model.eval()
            for param in semantic_model.parameters():
                param.requires_grad = False
            self.semantic_model = semantic_model
            self.sem_loss_scale = config.sem_loss_scale
            self.semantic_loss = ClipLoss(local_loss=False, gather_with_grad=True, cache_labels=True, rank=rank,
                                          world_size=world_size, use_horovod=False)
            if not self.half_sem and self.product_quant > 1:
                self.sem_linear = nn.Conv2d(self.product_quant * config.codebook_embed_dim, config.codebook_embed_dim, 1)
            elif self.half_sem and self.product_quant == 1:
                self.sem_linear = nn.Conv2d(768, config.codebook_embed_dim // 2, 1)
            if self.enc_type == 'cnn':
                self.sem_linear = torch.nn.Linear(384, config.codebook_embed_dim)
            self.sem_loss_weight = config.sem_loss_weight

        self.detail_guide = config.detail_guide
        if self.detail_guide != 'none':
            detail_model = create_model("vit_base_patch16_clip_224.openai", pretrained=True, img_size=256,
                                        patch_size=16, drop_path_rate=0.0)
            detail_model.eval()
            for param in detail_model.parameters():
                param.requires_grad = False
            self.detail_model = detail_model
            self.detail_loss_scale = config.detail_loss_scale
            self.detail_loss = ClipLoss(local_loss=False, gather_with_grad=True, cache_labels=True, rank=rank,
                                        world_size=world_size, use_horovod=False)
            self.detail_loss_weight = config.detail_loss_weight

        self.guide_type_1 = config.guide_type_1
        self.guide_type_2 = config.guide_type_2
        self.dependency_loss_weight = config.dependency_loss_weight
        self.test_mode = config.test_model
        if self.test_mode:
            self.eval()
            [p.requires_grad_(False) for p in self.parameters()]

    def train(self, mode: bool = True):
        super().train(mode)
        for name in ("semantic_model", "detail_model"):  # frozen teachers stay in eval (:176, :211)
            m = getattr(self, name, None)
            if m is not None:
                m.eval()
        return self

    def finetune(self, enc_tuning_method, dec_tuning_method):
        self.encoder.finetine(enc_tuning_method)
        self.decoder.finetine(dec_tuning_method)

    # ------------------------------------------------------------------ encode / decode
    def _tokens_to_map(self, h):
        b, l, c = h.shape
        if self.product_quant > 1:
            assert int(sqrt(l // self.product_quant)) ** 2 * self.product_quant == l
            h = h.view(b, l, 1, c)
        else:
            assert int(sqrt(l)) ** 2 == l
            h = h.view(b, int(sqrt(l)), int(sqrt(l)), c)
        return h.permute(0, 3, 1, 2)

    def encode(self, x):
        """:241-254 -> continuous latent B x C x (sqrt L) x (sqrt L)  (PQ>1: B x C x (PQ*L) x 1)"""
        h = self.encoder(x)
        if self.enc_type == 'dinov2':
            h = self._tokens_to_map(h)
        return self.quant_conv(h)

    def decode(self, quant, return_quant=False):
        """:256-261"""
        quant = self.post_quant_conv(quant)
        if self.dec_type == 'dinov2':
            quant = quant.flatten(2).permute(0, 2, 1)
        return self.decoder(quant)

    def _split_branches(self, h):
        b, c, l, _ = h.shape
        s = int(sqrt(l // self.product_quant))
        return [t.reshape(b, -1, s, s) for t in h.chunk(chunks=self.product_quant, dim=2)]

    # ------------------------------------------------------------------ training forward
    def forward(self, input, epoch, alpha, beta, delta):
        """:268-365"""
        h = self.encode(input)
        b, c, l, _ = h.shape
        if len(self.v_patch_nums) == 1:
            dropout_rand = None
        else:
            # CPU generator, shared by all PQ branches (:274)
            dropout_rand = torch.randint(self.start_drop, len(self.v_patch_nums) + 1, (b,))

        if self.product_quant > 1:
            quant_list, usages_list, vq_list, commit_list, entropy_list = [], [], [], [], []
            for i, hi in enumerate(self._split_branches(h)):
                quant, usages, vq_loss, commit_loss, entropy_loss = self.quantizes[i].forward(
                    hi, ret_usages=True, dropout=dropout_rand)
                quant_list.append(quant)
                usages_list.append(usages)
                vq_list.append(vq_loss)
                commit_list.append(commit_loss)
                entropy_list.append(entropy_loss)
            dependency_loss = self.dependency_loss_weight * orthogonal_cosine_loss(
                torch.mean(quant_list[0], dim=(2, 3)).contiguous(), torch.mean(quant_list[-1], dim=(2, 3)).contiguous())
            usages = [sum(us) / self.product_quant for us in zip(*usages_list)]
            mean_vq_loss = sum(vq_list) / self.product_quant
            mean_commit_loss = sum(commit_list) / self.product_quant
            mean_entropy = sum(entropy_list) / self.product_quant
            quant = torch.cat(quant_list, dim=1)
        else:
            dependency_loss = 0.0
            quant, usages, mean_vq_loss, mean_commit_loss, mean_entropy = self.quantize.forward(
                h, ret_usages=True, dropout=dropout_rand)
            # the reference calls add_perturbation unconditionally here and would fail for the
            # multi-scale quantizers (they have no z_channels / codebook_norm); same restriction:
            quant = add_perturbation(h, quant, self.quantize.z_channels, self.quantize.codebook_norm,
                                     self.quantize.embedding, alpha, beta, delta)
            quant_list = [quant]

        dec = self.decode(quant)

        sem_loss = None
        detail_loss = None
        if self.semantic_guide != 'none' or self.detail_guide != 'none':
            input = self.normalize(self.denormalize(input))
        if self.semantic_guide != 'none':
            if self.guide_type_1 == 'class':
                z_s = self.semantic_model(input)
                z_s = z_s[..., None, None]
            else:
                z_s = self.semantic_model.forward_features(input)[:, 1:, :]
                z_s = z_s.reshape(b, 768, 16, 16)
            if self.enc_type == 'dinov2':
                z_s = self.quant_conv(z_s).contiguous()
                semantic_quant = quant_list[-1]
                z_s = torch.mean(z_s, dim=(2, 3)).contiguous()
                z_q_ = torch.mean(semantic_quant, dim=(2, 3)).contiguous()
            else:  # cnn (:318-320)
                z_q_ = torch.mean(h, dim=(2, 3)).contiguous()
                z_s = self.sem_linear(z_s.flatten(1)).contiguous()
            n_drop = int(b * self.codebook_drop)
            with torch.autocast(device_type=input.device.type, enabled=False):
                sem_loss_scale = self.sem_loss_scale
                feat1 = z_s[n_drop:].float()
                feat2 = z_q_[n_drop:].float()
                if self.clip_norm:
                    feat1 = feat1 / feat1.norm(dim=1, keepdim=True)
                    feat2 = feat2 / feat2.norm(dim=1, keepdim=True)
                    sem_loss_scale = (epoch % 200) / 200 * (100 - sem_loss_scale) + sem_loss_scale if epoch < 200 else 100
                sem_loss = self.semantic_loss.forward(feat1, feat2, logit_scale=sem_loss_scale)
                sem_loss = sem_loss * self.sem_loss_weight

        if self.detail_guide != 'none':
            assert self.guide_type_2 == 'patch', "current only accept patch for detail guide"
            z_d = self.detail_model.forward_features(input)[:, 1:, :]
            z_d = z_d.reshape(b, 768, 16, 16)
            z_d = self.quant_conv(z_d).contiguous()
            detail_quant = quant_list[0]
            z_d = torch.mean(z_d, dim=(2, 3)).contiguous()
            z_q_ = torch.mean(detail_quant, dim=(2, 3)).contiguous()
            n_drop = int(b * self.codebook_drop)
            with torch.autocast(device_type=input.device.type, enabled=False):
                detail_loss_scale = self.detail_loss_scale
                feat1 = z_d[n_drop:].float()
                feat2 = z_q_[n_drop:].float()
                if self.clip_norm:
                    feat1 = feat1 / feat1.norm(dim=1, keepdim=True)
                    feat2 = feat2 / feat2.norm(dim=1, keepdim=True)
                    detail_loss_scale = (epoch % 200) / 200 * (100 - detail_loss_scale) + detail_loss_scale if epoch < 200 else 100
                detail_loss = self.detail_loss.forward(feat1, feat2, logit_scale=detail_loss_scale)
                detail_loss = detail_loss * self.detail_loss_weight

        return dec, (mean_vq_loss, mean_commit_loss, mean_entropy, usages), sem_loss, detail_loss, dependency_loss

    # ------------------------------------------------------------------ inference
    def _quantizers(self):
        return list(self.quantizes) if self.product_quant > 1 else [self.quantize]

    def _latent_branches(self, x):
        h = self.encoder(x)
        f = self.quant_conv(self._tokens_to_map(h) if self.enc_type == 'dinov2' else h)
        if self.product_quant > 1:
            return self._split_branches(f)
        return [f]

    def img_to_reconstructed_img(self, x, last_one=True, ) -> List[torch.Tensor]:
        """:367-403"""
        f_list = self._latent_branches(x)
        vpn = None if len(self.v_patch_nums) == 1 else self.v_patch_nums
        f_hats_list = [q.f_to_idxBl_or_fhat(f, to_fhat=True, v_patch_nums=vpn) for q, f in zip(self._quantizers(), f_list)]
        f_hats = [self.post_quant_conv(torch.cat(f_hats, dim=1)) for f_hats in zip(*f_hats_list)]
        if self.dec_type == 'dinov2':
            f_hats = [f_hat.flatten(2).permute(0, 2, 1) for f_hat in f_hats]
        if last_one:
            return self.decoder(f_hats[-1]).clamp_(-1, 1)
        return [self.decoder(f_hat).clamp_(-1, 1) for f_hat in f_hats]

    def img_to_idxBl(self, x, v_patch_nums=None):
        """tokens of an image batch: PQ=1 -> list over scales of [B, pn^2] int64 (single scale: [B, L]);
        PQ>1 -> list over branches of such lists.  (expected by trainer.py:69,122; the original
        lives in models/vqvae.py:65)"""
        f_list = self._latent_branches(x)
        vpn = None if len(self.v_patch_nums) == 1 else (v_patch_nums or self.v_patch_nums)
        out = []
        for q, f in zip(self._quantizers(), f_list):
            ls = q.f_to_idxBl_or_fhat(f, to_fhat=False, v_patch_nums=vpn)
            if len(self.v_patch_nums) == 1:
                ls = [ls[0].view(f.shape[0], -1)]
            out.append(ls)
        return out if self.product_quant > 1 else out[0]

    encode_to_codes = img_to_idxBl  # scripts/pretokenization.py:233 treats the tokenizer output as codes

    def decode_tokens(self, idxBl):
        """inverse of img_to_idxBl (demo_util.py:109): token lists -> reconstructed image in [-1, 1]."""
        per_branch = idxBl if self.product_quant > 1 else [idxBl]
        f_hats = [q.idx_to_fhat(ls) for q, ls in zip(self._quantizers(), per_branch)]
        return self.fhat_to_img(torch.cat(f_hats, dim=1))

    def img_to_sem_feat(self, x, ) -> List[torch.Tensor]:
        """:405-427"""
        f_list = self._latent_branches(x)
        f_hats_list = [q.f_to_idxBl_or_fhat(f, to_fhat=True, v_patch_nums=self.v_patch_nums)
                       for q, f in zip(self._quantizers(), f_list)]
        return f_hats_list[-1][-1]

    def fhat_to_img(self, f_hat: torch.Tensor):
        f_hat = self.post_quant_conv(f_hat)
        if self.dec_type == 'dinov2':
            f_hat = f_hat.flatten(2).permute(0, 2, 1)
        return self.decoder(f_hat).clamp_(-1, 1)

    def idxBl_to_var_input(self, gt_idx_Bl):
        if self.product_quant > 1:
            return torch.cat([self.quantizes[i].idxBl_to_var_input(gt_idx_Bl[i]) for i in range(self.product_quant)], dim=-1)
        return self.quantize.idxBl_to_var_input(gt_idx_Bl)

    def get_next_autoregressive_input(self, si, SN, f_hat, h_BChw):
        f_hat_list = f_hat.chunk(self.product_quant, dim=1)
        h_BChw_list = h_BChw.chunk(self.product_quant, dim=1)
        out_fhat_list, out_next_token_map_list = [], []
        for i, (f_hat, h_BChw) in enumerate(zip(f_hat_list, h_BChw_list)):
            out_fhat, out_next_token_map = self._quantizers()[i].get_next_autoregressive_input(si, SN, f_hat, h_BChw)
            out_fhat_list.append(out_fhat)
            out_next_token_map_list.append(out_next_token_map)
        return torch.cat(out_fhat_list, dim=1), torch.cat(out_next_token_map_list, dim=1)


def VQ_8(**kwargs):
    return VQModel(ModelArgs(encoder_ch_mult=[1, 2, 2, 4], decoder_ch_mult=[1, 2, 2, 4], **kwargs))


def VQ_16(**kwargs):
    return VQModel(ModelArgs(encoder_ch_mult=[1, 1, 2, 2, 4], decoder_ch_mult=[1, 1, 2, 2, 4], **kwargs))


VQ_models = {'VQ-16': VQ_16, 'VQ-8': VQ_8}
