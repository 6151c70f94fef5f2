"""cuDNN SDPA time vs sequence length (dev tool): how much of the S = 513 / 514 / 769 attention cost is tile padding?"""
import torch, torch.nn.functional as F
from torch.nn.attention import sdpa_kernel, SDPBackend


def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it


H, D = 12, 64
for B, Sq, Sk in [(256, 513, 513), (256, 512, 512), (256, 512, 513), (256, 514, 514), (256, 384, 384), (256, 640, 640),
                  (128, 769, 769), (128, 768, 768), (128, 499, 499), (128, 379, 379)]:
    S = max(Sq, Sk)
    qkv = torch.randn(B, S, 3, H, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    q, k, v = qkv.permute(2, 0, 3, 1, 4).unbind(0)
    q, k, v = q[:, :, S - Sq:], k[:, :, S - Sk:], v[:, :, S - Sk:]
    with sdpa_kernel(SDPBackend.CUDNN_ATTENTION):
        o = F.scaled_dot_product_attention(q, k, v)
        g = torch.randn_like(o)
        f = t(lambda: F.scaled_dot_product_attention(q, k, v))
        b = t(lambda: torch.autograd.grad(o, (q, k, v), g, retain_graph=True))
    fl = 4.0 * B * H * Sq * Sk * D
    print(f"B={B} Sq={Sq} Sk={Sk}: fwd {f:.3f} ms ({fl/f/1e9:.0f} TF/s)  bwd {b:.3f} ms ({2.5*fl/b/1e9:.0f} TF/s)", flush=True)
