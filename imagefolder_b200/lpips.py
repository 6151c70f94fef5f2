"""LPIPS perceptual loss (VGG16) -- tokenizer/tokenizer_image/lpips.py:52-159.

Checkpoint-compatible with the reference (`scaling_layer.{shift,scale}`, `net.slice{1..5}.{i}.{weight,bias}`,
`lin{0..4}.model.1.weight`); `vgg.pth` carries the lin layers only (see load_from_pretrained).  The VGG16 trunk is a stack of library convolutions; what the
reference does AFTER the trunk -- per stage: channel-normalise both maps, subtract, square, 1x1 `lin` conv, spatial mean,
i.e. ~10 full passes over feature maps of up to 2 GB -- is one fused CUDA pass per stage (xq_lpips_layer_forward /
_backward, csrc/loss_kernels.cu).

On CUDA tensors the stage distance ALWAYS goes through libxqb200 (and fails loudly if the library is missing); there is no
silent substitute.  Two situations keep the reference's op sequence on library kernels by design: train-mode dropout in
front of the `lin` conv (a semantic the fused kernel does not have; `VQLoss` keeps LPIPS in eval mode), and CPU tensors,
where the whole module -- trunk included -- is a plain library network.

There is no network here.  The reference takes the trunk from torchvision (`models.vgg16(pretrained=True)`) and only the
`lin` layers from vgg.pth; `load_from_pretrained` mirrors that with two local files ($XQ_VGG16_CKPT for the trunk,
$XQ_LPIPS_CKPT or `<this dir>/cache/vgg.pth` for the lin layers), reports every tensor that is still at its random
initialisation (`self.unloaded_keys`, a warning -- or an exception with XQ_REQUIRE_PRETRAINED=1).
"""
from __future__ import annotations

import os
import warnings
from collections import namedtuple

import torch
import torch.nn as nn

from . import loss_ops

_VGG16_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512]   # features[:30]
_SLICE_ENDS = (4, 9, 16, 23, 30)        # relu1_2, relu2_2, relu3_3, relu4_3, relu5_3  (lpips.py:122-131)
VggOutputs = namedtuple("VggOutputs", ['relu1_2', 'relu2_2', 'relu3_3', 'relu4_3', 'relu5_3'])


def _vgg16_feature_layers():
    layers, cin = [], 3
    for v in _VGG16_CFG:
        if v == 'M':
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            layers += [nn.Conv2d(cin, v, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
            cin = v
    return layers


class vgg16(nn.Module):
    """torchvision `vgg16().features[:30]` cut into the five LPIPS slices; sub-module names = torchvision indices."""

    def __init__(self, requires_grad=False, pretrained=True):
        super().__init__()
        feats = _vgg16_feature_layers()
        start = 0
        for si, end in enumerate(_SLICE_ENDS, 1):
            seq = nn.Sequential()
            for i in range(start, end):
                seq.add_module(str(i), feats[i])
            setattr(self, f"slice{si}", seq)
            start = end
        self.N_slices = 5
        if not requires_grad:
            for p in self.parameters():
                p.requires_grad = False

    def forward(self, X):
        outs, h = [], X
        for si in range(1, 6):
            h = getattr(self, f"slice{si}")(h)
            outs.append(h)
        return VggOutputs(*outs)


class ScalingLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer('shift', torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer('scale', torch.tensor([.458, .448, .450])[None, :, None, None])

    def forward(self, inp):
        return (inp - self.shift) / self.scale


class NetLinLayer(nn.Module):
    """the 1x1 conv of one stage (no bias); `model.0` is the Dropout of the released checkpoint layout."""

    def __init__(self, chn_in, chn_out=1, use_dropout=False):
        super().__init__()
        layers = [nn.Dropout()] if use_dropout else []
        layers += [nn.Conv2d(chn_in, chn_out, 1, stride=1, padding=0, bias=False)]
        self.model = nn.Sequential(*layers)


def normalize_tensor(x, eps=1e-10):
    return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + eps)


def spatial_average(x, keepdim=True):
    return x.mean([2, 3], keepdim=keepdim)


class LPIPS(nn.Module):
    def __init__(self, use_dropout=True):
        super().__init__()
        self.scaling_layer = ScalingLayer()
        self.chns = [64, 128, 256, 512, 512]
        self.net = vgg16(pretrained=True, requires_grad=False)
        for i, c in enumerate(self.chns):
            setattr(self, f"lin{i}", NetLinLayer(c, use_dropout=use_dropout))
        self.load_from_pretrained()
        for p in self.parameters():
            p.requires_grad = False

    def load_from_pretrained(self, name="vgg_lpips"):
        """Two files, as in the reference: the torchvision VGG16 trunk (`models.vgg16(pretrained=True)`, lpips.py:119)
        and vgg.pth, which carries ONLY the `lin{k}.model.1.weight` 1x1 convs (which is why the reference loads it with
        strict=False, lpips.py:72).  There is no network here, so both come from local files:
            $XQ_VGG16_CKPT   torchvision vgg16 state_dict (`features.N.{weight,bias}`), remapped to `net.slice{k}.N.*`
            $XQ_LPIPS_CKPT   vgg.pth (or <this dir>/cache/vgg.pth, the reference's cache location)
        A checkpoint that already holds `net.slice*` keys (a full LPIPS state_dict) serves both.  Whatever is still
        unset afterwards is reported loudly: a perceptual loss on a random trunk is not a perceptual loss."""
        loaded = set()
        own = self.state_dict()
        trunk = os.environ.get("XQ_VGG16_CKPT", "")
        if trunk and os.path.exists(trunk):
            sd = torch.load(trunk, map_location="cpu")
            sd = sd.get("state_dict", sd)
            remap = {}
            for k, v in sd.items():
                if k.startswith("features."):
                    idx = int(k.split(".")[1])
                    if idx < _SLICE_ENDS[-1]:
                        si = next(i for i, e in enumerate(_SLICE_ENDS, 1) if idx < e)
                        remap[f"net.slice{si}.{idx}.{k.split('.')[2]}"] = v
            bad = [k for k, v in remap.items() if k not in own or own[k].shape != v.shape]
            if bad:
                raise RuntimeError(f"LPIPS: {trunk} is not a torchvision vgg16 state_dict (unexpected {bad[:3]})")
            self.load_state_dict(remap, strict=False)
            loaded |= set(remap)
        used = None
        for path in (os.environ.get("XQ_LPIPS_CKPT", ""), os.path.join(os.path.dirname(os.path.abspath(__file__)), "cache", "vgg.pth")):
            if path and os.path.exists(path):
                sd = torch.load(path, map_location="cpu")
                res = self.load_state_dict(sd, strict=False)
                if res.unexpected_keys:
                    raise RuntimeError(f"LPIPS: unexpected keys in {path}: {res.unexpected_keys[:4]}")
                loaded |= set(sd)
                used = path
                break
        missing = [k for k in own if k not in loaded and not k.startswith("scaling_layer.")]
        self.unloaded_keys = missing
        if missing:
            trunk_missing = [k for k in missing if k.startswith("net.")]
            lin_missing = [k for k in missing if k.startswith("lin")]
            msg = "LPIPS runs on RANDOM weights for: "
            if trunk_missing:
                msg += f"the VGG16 trunk ({len(trunk_missing)} tensors; set XQ_VGG16_CKPT to a torchvision vgg16 state_dict) "
            if lin_missing:
                msg += f"the lin layers ({len(lin_missing)} tensors; set XQ_LPIPS_CKPT to vgg.pth)"
            if os.environ.get("XQ_REQUIRE_PRETRAINED", "0") == "1":
                raise RuntimeError(msg)
            warnings.warn(msg)
        return used

    def _stage(self, i, f0, f1):
        lin = getattr(self, f"lin{i}").model[-1]
        if f0.is_cuda and lin.weight.shape[0] == 1 and not (self.training and len(getattr(self, f"lin{i}").model) > 1):
            return loss_ops.lpips_stage(f0, f1, lin.weight).view(-1, 1, 1, 1)
        # dropout active (train mode) or a multi-output lin layer: the reference's op sequence on library kernels
        diff = (normalize_tensor(f0) - normalize_tensor(f1)) ** 2
        return spatial_average(getattr(self, f"lin{i}").model(diff), keepdim=True)

    def forward(self, input, target):
        outs0 = self.net(self.scaling_layer(input))
        outs1 = self.net(self.scaling_layer(target))
        val = self._stage(0, outs0[0], outs1[0])
        for i in range(1, len(self.chns)):
            val = val + self._stage(i, outs0[i], outs1[i])
        return val
