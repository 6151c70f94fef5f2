"""Deterministic, name-keyed parameter values shared by make_vit_golden.py (which applies them to the REFERENCE's
modules) and the tests (which apply them to the product / oracle).  Test infrastructure only.

The checkpoint key names are the contract between the two sides, so seeding each tensor from
its own name gives both sides identical weights without committing ~90 MB of ViT parameters."""
import math
import zlib

import torch


def det_tensor(name: str, shape, like: torch.Tensor = None) -> torch.Tensor:
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7fffffff)
    r = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "gamma":                                   # LayerScale: large enough that every block matters
        return 0.3 + 0.1 * r
    if leaf == "bias":
        return 0.02 * r
    if leaf == "weight" and len(shape) == 1:              # LayerNorm scale
        return 1.0 + 0.1 * r
    if leaf == "weight" and len(shape) >= 2:              # Linear / Conv / Embedding: unit-variance outputs
        fan_in = 1
        for s in shape[1:]:
            fan_in *= int(s)
        return r / math.sqrt(max(fan_in, 1))
    return 0.2 * r                                        # tokens, positional embeddings


def apply_det_init(module: torch.nn.Module) -> None:
    """Overwrite every floating-point PARAMETER (buffers keep their constructed values)."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.is_floating_point():
                p.copy_(det_tensor(name, p.shape))


def golden_inputs(codebook_channels: int, side: int):
    """The image and the decoder-side latent of the vit_*.npz cases (CPU generator: identical everywhere)."""
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(1, 3, 256, 256, generator=g) * 2 - 1
    q = torch.randn(1, codebook_channels, side, side, generator=g)
    return x, q
