"""Latent perturbation (RobustTok) -- drop-in for the reference's
tokenizer/tokenizer_image/latent_perturbation.py:4 `add_perturbation`, plus the `add_perturb`
spelling used in the reference README (README.md:79).

The reference recomputes the full N x V distance matrix and a top-delta over V for EVERY row, then
keeps the result only for the first int(B*beta) samples.  Here only those samples' rows are
touched, and per row only the selected rank is extracted (radix select in shared memory,
csrc/vq_kernels.cu: rank_select_kernel).  The random draws are made with the same two torch calls
in the same order as the reference (:21-22), so the device RNG stream is identical.
"""
from __future__ import annotations

import torch

from . import ops

__all__ = ["add_perturbation", "add_perturb"]


def add_perturbation(z, z_q, z_channels, codebook_norm, codebook, alpha, beta, delta, rand_u=None, rand_j=None):
    """z: pre-quant latent [B,C,H,W]; z_q: quantizer output; codebook: nn.Embedding.
    rand_u / rand_j let a caller inject the two random tensors (tests); by default they are drawn
    exactly like the reference does."""
    assert z.shape[1] == z_channels
    B = z.shape[0]
    N = z.numel() // z_channels
    delta = int(delta)
    if rand_u is None:
        rand_u = torch.rand(N, device=z.device)                         # latent_perturbation.py:21
    if rand_j is None:
        rand_j = torch.randint(0, delta, (N,), device=z.device)         # :22
    n_perturb = int(B * beta)                                           # :32
    return ops.perturb(z, z_q, codebook.weight, rand_u, rand_j, bool(codebook_norm), float(alpha), n_perturb, delta)


def add_perturb(x, z_q=None, *, z_channels, codebook_norm, codebook, alpha, beta, delta):
    """README.md:79 pseudo-code spelling.  Without z_q the un-perturbed samples get the plain
    nearest-code output."""
    if z_q is None:
        z_q, _ = ops.vq_lookup(x, codebook.weight, codebook_norm)
    return add_perturbation(x, z_q, z_channels, codebook_norm, codebook, alpha, beta, delta)
