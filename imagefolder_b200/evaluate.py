"""Reconstruction-evaluation loop of the tokenizer (the rFID data path) -- tokenizer/tokenizer_image/xqgan_train.py:517-535.

The reference, every `ckpt_every` steps: puts the model in eval mode, reconstructs the whole validation set with
`img_to_reconstructed_img`, converts both the reconstruction and the ground truth to NHWC uint8
(`clamp(127.5 * x + 128, 0, 255)`), ALL-GATHERs them across ranks (`dist.nn.all_gather`), and accumulates them on the host
for the (TensorFlow) Inception-statistics evaluator.  This module is that loop up to the evaluator boundary: it returns the
two uint8 arrays the reference hands to `Evaluator.read_activations` (the evaluator itself -- a frozen TF Inception graph
with downloaded weights -- is out of scope).

Data-path notes: the gather moves uint8 (one byte per value, as the reference), is issued once per batch for each of the
two tensors, and every rank returns the same arrays (rank 0 is the one that evaluates)."""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def to_uint8_nhwc(x: torch.Tensor) -> torch.Tensor:
    """[-1, 1] float NCHW -> uint8 NHWC exactly as xqgan_train.py:527-528."""
    return torch.clamp(127.5 * x + 128.0, 0, 255).permute(0, 2, 3, 1).to(torch.uint8).contiguous()


def _all_gather_cat(t: torch.Tensor) -> torch.Tensor:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)          # rank-major concatenation, like torch.cat(dist.nn.all_gather(t), dim=0)
    return torch.cat(parts, dim=0)


@torch.no_grad()
def reconstruct_for_fid(model, loader: Iterable, device=None, max_batches: Optional[int] = None,
                        autocast_dtype: Optional[torch.dtype] = None) -> Tuple[np.ndarray, np.ndarray, int]:
    """-> (samples uint8 [T,H,W,3], ground_truth uint8 [T,H,W,3], T)  with T = all ranks' images in rank-major batch order.

    `model` is a VQModel (or a DDP wrapper of one); its training flag is restored on exit.  `loader` yields (images, _)
    with images in [-1, 1]; every rank must yield the same number of equally sized batches (the reference's sampler
    guarantees that)."""
    core = getattr(model, "module", model)
    was_training = core.training
    core.eval()
    device = device if device is not None else next(core.parameters()).device
    samples, gt, total = [], [], 0
    try:
        for bi, (x, _) in enumerate(loader):
            if max_batches is not None and bi >= max_batches:
                break
            x = x.to(device, non_blocking=True)
            if autocast_dtype is not None and x.is_cuda:
                with torch.autocast("cuda", dtype=autocast_dtype):
                    rec = core.img_to_reconstructed_img(x)
            else:
                rec = core.img_to_reconstructed_img(x)
            s8 = _all_gather_cat(to_uint8_nhwc(rec.float()))
            x8 = _all_gather_cat(to_uint8_nhwc(x.float()))
            samples.append(s8.cpu().numpy())
            gt.append(x8.cpu().numpy())
            total += s8.shape[0]
    finally:
        core.train(was_training)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if not samples:
        return np.zeros((0, 0, 0, 3), np.uint8), np.zeros((0, 0, 0, 3), np.uint8), 0
    return np.concatenate(samples, axis=0), np.concatenate(gt, axis=0), total
