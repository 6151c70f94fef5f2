"""Pre-tokenisation of an image stream into the on-disk token format the generators train on
(SURVEY.md 8f-2).  Mirrors scripts/pretokenization.py:221-251: every sample is tokenised together with
its horizontal flip, one record per image, written as JSON lines

    {"class_id": <int>, "tokens": [<int>, ...]}

`tokens` is the flat code list of `VQModel.encode_to_codes` (single scale: L codes; multi-scale: the
scales concatenated small -> large; product quantisation: the branches concatenated).  The encoder ->
quantizer work runs through libxqb200; only the file writing is host code.
"""
from __future__ import annotations

import json
from typing import Iterable, List, Tuple

import torch


def flatten_codes(codes) -> torch.Tensor:
    """img_to_idxBl output (list over scales, or list over PQ branches of lists) -> [B, n_tokens] int64."""
    if isinstance(codes[0], (list, tuple)):
        return torch.cat([flatten_codes(c) for c in codes], dim=1)
    return torch.cat([c.reshape(c.shape[0], -1) for c in codes], dim=1)


def unflatten_codes(tokens: torch.Tensor, model) -> List:
    """inverse of flatten_codes for `model` (a VQModel): [B, n_tokens] -> the img_to_idxBl structure."""
    pns = list(model.v_patch_nums)
    per_scale = [model.config.num_latent_tokens // model.product_quant] if len(pns) == 1 else [p * p for p in pns]
    out, off = [], 0
    for _ in range(model.product_quant):
        branch = []
        for n in per_scale:
            branch.append(tokens[:, off:off + n].contiguous())
            off += n
        out.append(branch)
    return out if model.product_quant > 1 else out[0]


@torch.no_grad()
def pretokenize(model, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], out_path: str, flip: bool = True,
                autocast_dtype=torch.bfloat16) -> int:
    """batches yields (images [B,3,H,W] in [-1,1], class ids [B]).  Returns the number of records written."""
    model.eval()
    n = 0
    dev = next(model.parameters()).device
    with open(out_path, "w", encoding="utf-8") as f:
        for images, targets in batches:
            images = images.to(dev, non_blocking=True)
            targets = torch.as_tensor(targets)
            if flip:
                images = torch.cat([images, torch.flip(images, dims=[-1])])
                targets = torch.cat([targets, targets])
            with torch.autocast(dev.type, dtype=autocast_dtype, enabled=autocast_dtype is not None and dev.type == "cuda"):
                codes = flatten_codes(model.encode_to_codes(images))
            codes = codes.cpu()
            for b in range(codes.shape[0]):
                f.write(json.dumps({"class_id": int(targets[b]), "tokens": codes[b].tolist()}) + "\n")
                n += 1
    return n


def read_tokens(path: str):
    with open(path, "r", encoding="utf-8") as f:
        for line in f:
            rec = json.loads(line)
            yield rec["class_id"], torch.tensor(rec["tokens"], dtype=torch.int64)
