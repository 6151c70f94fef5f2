"""autograd.Function wrappers over the loss-stack entry points of the C ABI (include/xqb200.h, row f-1)."""
from __future__ import annotations

import torch

from . import _capi as C


class _LpipsStage(torch.autograd.Function):
    """(f0, f1 [B,C,H,W] fp32 / bf16, lin_w [C]) -> [B]:  spatial mean of  sum_c w_c (f0^ - f1^)^2   (lpips.py:83-86).
    The function is symmetric in (f0, f1): the backward kernel differentiates its second map, so the gradient of the
    first is the same call with the maps swapped.  `lin_w` is a frozen LPIPS parameter (no gradient)."""

    @staticmethod
    def forward(ctx, f0, f1, lin_w, eps: float):
        if f0.shape != f1.shape or f0.dim() != 4:
            raise ValueError(f"feature maps must have the same [B,C,H,W] shape, got {tuple(f0.shape)} / {tuple(f1.shape)}")
        dt = torch.bfloat16 if (f0.dtype == torch.bfloat16 and f1.dtype == torch.bfloat16
                                and (f0.shape[2] * f0.shape[3]) % 2 == 0) else torch.float32   # bf16 kernel reads pixel pairs
        f0, f1 = f0.to(dt).contiguous(), f1.to(dt).contiguous()
        w = lin_w.detach().reshape(-1).float().contiguous()
        B, Cc, H, W = f0.shape
        if w.numel() != Cc:
            raise ValueError(f"lin weight has {w.numel()} channels, features have {Cc}")
        out = torch.empty(B, dtype=torch.float32, device=f0.device)
        L = C.lib()
        ws = C.workspace(L.xq_lpips_workspace_bytes(B, H * W), f0.device)
        C.call("xq_lpips_layer_forward", 2, L.xq_lpips_layer_forward, C.ptr(f0), C.ptr(f1), int(dt == torch.bfloat16), C.ptr(w),
               B, Cc, H * W, float(eps), 0, C.ptr(out), C.ptr(ws), ws.numel(), C.stream_ptr(f0.device),
               nbytes=2 * f0.numel() * f0.element_size())
        ctx.save_for_backward(f0, f1, w)
        ctx.eps = float(eps)
        return out

    @staticmethod
    def backward(ctx, g):
        f0, f1, w = ctx.saved_tensors
        B, Cc, H, W = f0.shape
        g = g.float().contiguous()
        L = C.lib()
        grads = [None, None]
        for which in (0, 1):
            if not ctx.needs_input_grad[which]:
                continue
            a, b = (f0, f1) if which == 1 else (f1, f0)
            gb = torch.empty_like(b)
            C.call("xq_lpips_layer_backward", 1, L.xq_lpips_layer_backward, C.ptr(a), C.ptr(b),
                   int(b.dtype == torch.bfloat16), C.ptr(w), B, Cc, H * W, ctx.eps, C.ptr(g), C.ptr(gb),
                   C.stream_ptr(b.device), nbytes=5 * b.numel() * b.element_size())
            grads[which] = gb
        return grads[0], grads[1], None, None


def lpips_stage(f0, f1, lin_w, eps: float = 1e-10):
    return _LpipsStage.apply(f0, f1, lin_w, eps)


class _DiffAug(torch.autograd.Function):
    """translation + colour + cutout of DiffAug.aug (diffaug.py:60-118) as one fused pass; backward = its transpose."""

    @staticmethod
    def forward(ctx, x, rand01, flags: int, cut_h: int, cut_w: int):
        x = x.contiguous()
        B, Cc, H, W = x.shape
        y = torch.empty_like(x)
        sums = torch.empty(B, dtype=torch.float32, device=x.device)
        L = C.lib()
        C.call("xq_diffaug_forward", 2 if flags & 2 else 1, L.xq_diffaug_forward, C.ptr(x), C.ptr(rand01), B, Cc, H, W,
               int(flags), int(cut_h), int(cut_w), C.ptr(y), C.ptr(sums), C.stream_ptr(x.device), nbytes=8 * x.numel())
        ctx.save_for_backward(rand01)
        ctx.cfg = (int(flags), int(cut_h), int(cut_w))
        return y

    @staticmethod
    def backward(ctx, g):
        (rand01,) = ctx.saved_tensors
        flags, cut_h, cut_w = ctx.cfg
        g = g.float().contiguous()
        B, Cc, H, W = g.shape
        gx = torch.empty_like(g)
        sums = torch.empty(B, dtype=torch.float32, device=g.device)
        L = C.lib()
        C.call("xq_diffaug_backward", 2 if flags & 2 else 1, L.xq_diffaug_backward, C.ptr(g), C.ptr(rand01), B, Cc, H, W,
               flags, cut_h, cut_w, C.ptr(gx), C.ptr(sums), C.stream_ptr(g.device), nbytes=8 * g.numel())
        return gx, None, None, None, None


def diffaug_apply(x, rand01, flags: int, cut_h: int, cut_w: int):
    return _DiffAug.apply(x, rand01, flags, cut_h, cut_w)
