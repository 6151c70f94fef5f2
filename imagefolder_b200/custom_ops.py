"""`torch.library` registration of the single-scale quantizer entry points ("loaded as PyTorch custom
ops through a thin C-ABI").

    torch.ops.xqb200.vq_forward(z, E, beta, codebook_norm)        -> (out, loss[2], idx, hist)
    torch.ops.xqb200.vq_backward(z, E, idx, g_out, g_loss, beta, codebook_norm) -> (gz, gE)
    torch.ops.xqb200.usage_ema_(ema, hit, counter, margin)        -> usage[rows]      (mutates ema and counter)

Each op is the same ctypes call into libxqb200 that `ops.py` makes (the C ABI stays the boundary, ctypes the loader); what the
registration adds is a schema, fake (meta) implementations and an autograd formula, so that `torch.compile(fullgraph=True)`
traces `VectorQuantizer.forward` without graph breaks and the ops can be captured in CUDA graphs.  The step counter of the
usage EMA is a device tensor here (`xq_usage_ema_dev`): nothing in the traced forward depends on a Python int that changes
every call.  Reference: VectorQuantizer.forward, tokenizer/tokenizer_image/xqgan_model.py:745-801."""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from . import _capi as C

__all__ = ["vq_forward", "vq_backward", "usage_ema_"]


@torch.library.custom_op("xqb200::vq_forward", mutates_args=())
def vq_forward(z: Tensor, E: Tensor, beta: float, codebook_norm: bool) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    z = z.float().contiguous()
    E = E.float().contiguous()
    B, Cc = z.shape[0], z.shape[1]
    HW = z[0, 0].numel()
    V = E.shape[0]
    dev = z.device
    idx = torch.empty(B * HW, dtype=torch.int64, device=dev)
    out = torch.empty_like(z)
    loss = torch.empty(2, dtype=torch.float32, device=dev)
    hist = torch.zeros(V, dtype=torch.float32, device=dev)
    L = C.lib()
    ws = C.workspace(L.xq_vq_workspace_bytes(B, Cc, HW, V), dev)
    C.call("xq_vq_forward", 3, L.xq_vq_forward, C.ptr(z), C.ptr(E), B, Cc, HW, V, int(codebook_norm), 1, float(beta),
           C.ptr(idx), C.ptr(out), C.ptr(loss), C.ptr(hist), C.ptr(ws), ws.numel(), C.stream_ptr(dev))
    return out, loss, idx, hist


@vq_forward.register_fake
def _(z, E, beta, codebook_norm):
    n = z.shape[0] * z[0, 0].numel()
    return (torch.empty_like(z, dtype=torch.float32, memory_format=torch.contiguous_format), z.new_empty(2, dtype=torch.float32),
            z.new_empty(n, dtype=torch.int64), z.new_empty(E.shape[0], dtype=torch.float32))


@torch.library.custom_op("xqb200::vq_backward", mutates_args=())
def vq_backward(z: Tensor, E: Tensor, idx: Tensor, g_out: Tensor, g_loss: Tensor, beta: float,
                codebook_norm: bool) -> Tuple[Tensor, Tensor]:
    z = z.float().contiguous()
    E = E.float().contiguous()
    g_out = g_out.float().contiguous()
    g_loss = g_loss.float().contiguous()
    B, Cc = z.shape[0], z.shape[1]
    HW = z[0, 0].numel()
    V = E.shape[0]
    gz = torch.empty_like(z)
    gE = torch.empty_like(E)
    L = C.lib()
    C.call("xq_vq_backward", 1, L.xq_vq_backward, C.ptr(z), C.ptr(E), C.ptr(idx), C.ptr(g_out), g_loss.data_ptr(),
           g_loss.data_ptr() + 4, B, Cc, HW, V, int(codebook_norm), float(beta), C.ptr(gz), C.ptr(gE), C.stream_ptr(z.device))
    return gz, gE


@vq_backward.register_fake
def _(z, E, idx, g_out, g_loss, beta, codebook_norm):
    return (torch.empty_like(z, dtype=torch.float32, memory_format=torch.contiguous_format),
            torch.empty_like(E, dtype=torch.float32, memory_format=torch.contiguous_format))


def _vq_setup(ctx, inputs, output):
    z, E, beta, codebook_norm = inputs
    ctx.save_for_backward(z, E, output[2])
    ctx.beta, ctx.codebook_norm = beta, codebook_norm


def _vq_bwd(ctx, g_out, g_loss, _g_idx, _g_hist):
    z, E, idx = ctx.saved_tensors
    if g_out is None:
        g_out = torch.zeros_like(z, dtype=torch.float32)
    if g_loss is None:
        g_loss = z.new_zeros(2, dtype=torch.float32)
    gz, gE = torch.ops.xqb200.vq_backward(z, E, idx, g_out, g_loss, ctx.beta, ctx.codebook_norm)
    return gz, gE, None, None


torch.library.register_autograd("xqb200::vq_forward", _vq_bwd, setup_context=_vq_setup)


@torch.library.custom_op("xqb200::usage_ema_", mutates_args=("ema", "counter"))
def usage_ema_(ema: Tensor, hit: Tensor, counter: Tensor, margin: float) -> Tensor:
    """ema [rows, V] or [V] (updated in place), hit same shape, counter int64 [2] on the device (see xq_usage_ema_dev)."""
    rows = 1 if ema.dim() == 1 else ema.shape[0]
    V = ema.shape[-1]
    usage = torch.empty(rows, dtype=torch.float32, device=ema.device)
    L = C.lib()
    C.call("xq_usage_ema", 1, L.xq_usage_ema_dev, C.ptr(ema), C.ptr(hit.contiguous()), rows, V, C.ptr(counter), float(margin),
           C.ptr(usage), C.stream_ptr(ema.device))
    return usage


@usage_ema_.register_fake
def _(ema, hit, counter, margin):
    return ema.new_empty(1 if ema.dim() == 1 else ema.shape[0], dtype=torch.float32)
