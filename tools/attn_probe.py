"""Attention library probe on the GPU box: cuDNN/flash SDPA vs flash_attn packed (fwd+bwd)."""
import time, torch, torch.nn.functional as F


def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it


B, N, H, D = 256, 513, 12, 64
qkv = torch.randn(B, N, 3, H, D, device="cuda", dtype=torch.bfloat16, requires_grad=True)


def sdpa():
    q, k, v = qkv.permute(2, 0, 3, 1, 4).unbind(0)
    o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, H * D)
    o.backward(torch.ones_like(o))


print("sdpa fwd+bwd ms", t(sdpa))
try:
    from flash_attn import flash_attn_qkvpacked_func

    def fa():
        o = flash_attn_qkvpacked_func(qkv).reshape(B, N, H * D)
        o.backward(torch.ones_like(o))
    print("flash_attn packed fwd+bwd ms", t(fa))
except Exception as e:
    print("flash_attn unavailable:", repr(e)[:300])
for be in ["CUDNN_ATTENTION", "FLASH_ATTENTION", "EFFICIENT_ATTENTION"]:
    try:
        from torch.nn.attention import sdpa_kernel, SDPBackend
        with sdpa_kernel(getattr(SDPBackend, be)):
            print(be, "fwd+bwd ms", t(sdpa))
    except Exception as e:
        print(be, "failed", repr(e)[:200])
