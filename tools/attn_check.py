"""dev tool: xq_vit_attn_fwd / xq_vit_attn_bwd against an explicit fp32 softmax reference, plus timings vs the SDPA
library kernel at the bench shapes.   python tools/attn_check.py [--bwd] [--time]"""
import argparse
import math
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagefolder_b200 import _capi as C  # noqa: E402


def ref_attn(qkv, H):
    B, N, _ = qkv.shape
    x = qkv.float().view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    q, k, v = x[0], x[1], x[2]
    s = (q @ k.transpose(-1, -2)) * 0.125
    p = torch.softmax(s, -1)
    o = p @ v
    lse2 = torch.logsumexp(s, -1) * math.log2(math.e)
    return o.transpose(1, 2).reshape(B, N, H * 64), lse2, (q, k, v, p)


def ours_fwd(qkv, H):
    B, N, _ = qkv.shape
    out = torch.empty(B, N, H * 64, dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty(B, H, N, dtype=torch.float32, device=qkv.device)
    L = C.lib()
    C.call("xq_vit_attn_fwd", 1, L.xq_vit_attn_fwd, C.ptr(qkv), C.ptr(out), C.ptr(lse), B, N, H, 64, 0.125,
           C.stream_ptr(qkv.device))
    return out, lse


def ours_bwd(qkv, out, lse, do, H):
    B, N, _ = qkv.shape
    dqkv = torch.empty_like(qkv)
    L = C.lib()
    ws = torch.empty(int(L.xq_vit_attn_bwd_workspace_bytes(B, N, H)), dtype=torch.uint8, device=qkv.device)
    C.call("xq_vit_attn_bwd", 3, L.xq_vit_attn_bwd, C.ptr(qkv), C.ptr(out), C.ptr(do), C.ptr(lse), C.ptr(dqkv), None, B, N, H, 64,
           0.125, C.ptr(ws), ws.numel(), C.stream_ptr(qkv.device))
    return dqkv


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--lib", type=str, default="", help="alternative libxqb200 build (experiments)")
    ap.add_argument("--skip-check", action="store_true")
    ap.add_argument("--prof", action="store_true", help="per-kernel device times of the backward call")
    a = ap.parse_args()
    if a.lib:
        C.LIB_PATH = os.path.abspath(a.lib)
    dev = torch.device("cuda")
    torch.manual_seed(0)
    shapes = [(2, 128, 2), (1, 16, 1), (2, 1, 1), (2, 129, 3), (3, 200, 2), (2, 513, 4), (2, 514, 2), (1, 769, 2), (2, 499, 3),
              (2, 379, 2), (1, 1024, 1), (1, 333, 12)]
    for (B, N, H) in ([] if a.skip_check else shapes):
        for amp in (1.0, 6.0):
            qkv = (torch.randn(B, N, 3 * H * 64, device=dev) * amp).to(torch.bfloat16)
            o_ref, l_ref, _ = ref_attn(qkv, H)
            o, l = ours_fwd(qkv, H)
            torch.cuda.synchronize()
            eo = (o.float() - o_ref).abs().max().item()
            el = (l - l_ref).abs().max().item()
            ok = eo < 2e-2 * max(1.0, amp) and el < 2e-2 and torch.isfinite(o.float()).all().item()
            print(f"fwd B={B} N={N} H={H} amp={amp}: max|dO|={eo:.3e} (ref max {o_ref.abs().max().item():.2f}) max|dLSE2|={el:.3e} {'OK' if ok else 'FAIL'}")
            if a.bwd:
                do = torch.randn(B, N, H * 64, device=dev).to(torch.bfloat16)
                q3 = qkv.float().requires_grad_(True)
                o2, _, _ = ref_attn_grad(q3, H)
                (o2 * do.float()).sum().backward()
                dqkv = ours_bwd(qkv, o, l, do, H)
                torch.cuda.synchronize()
                g = q3.grad.view(B, N, 3, H * 64)
                d = dqkv.float().view(B, N, 3, H * 64)
                errs = [((d[:, :, i] - g[:, :, i]).abs().max().item(), g[:, :, i].abs().max().item()) for i in range(3)]
                ok = all(e < 3e-2 * max(1.0, m) for e, m in errs) and torch.isfinite(d).all().item()
                print("   bwd  " + "  ".join(f"d{n}: err {e:.3e} / max {m:.2f}" for n, (e, m) in zip("qkv", errs)), "OK" if ok else "FAIL")
    if a.time:
        for (B, N, H) in [(256, 513, 12), (256, 514, 12), (128, 769, 12), (128, 499, 12), (128, 379, 12)]:
            qkv = torch.randn(B, N, 3 * H * 64, device=dev).to(torch.bfloat16)
            t = timeit(lambda: ours_fwd(qkv, H))
            fl = 4.0 * B * H * N * N * 64
            x = qkv.view(B, N, 3, H, 64)
            q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
            t_lib = timeit(lambda: F.scaled_dot_product_attention(q, k, v))
            print(f"time fwd B={B} N={N} H={H}: ours {t:.3f} ms ({fl / t / 1e9:.0f} TFLOP/s)   sdpa {t_lib:.3f} ms ({fl / t_lib / 1e9:.0f} TFLOP/s)")
            if a.bwd:
                o, l = ours_fwd(qkv, H)
                do = torch.randn(B, N, H * 64, device=dev).to(torch.bfloat16)
                tb = timeit(lambda: ours_bwd(qkv, o, l, do, H))
                if a.prof:
                    from torch.profiler import profile, ProfilerActivity
                    with profile(activities=[ProfilerActivity.CUDA]) as pr:
                        for _ in range(5):
                            ours_bwd(qkv, o, l, do, H)
                        torch.cuda.synchronize()
                    for ev in pr.key_averages():
                        print(f"      {ev.key[:60]:60s} {ev.device_time_total / 5:9.1f} us")
                qq, kk, vv = (t_.detach().requires_grad_(True) for t_ in (q, k, v))
                oo = F.scaled_dot_product_attention(qq, kk, vv)
                gg = do.view(B, N, H, 64).transpose(1, 2)
                tb_lib = timeit(lambda: torch.autograd.grad(oo, (qq, kk, vv), gg, retain_graph=True))
                print(f"time bwd B={B} N={N} H={H}: ours {tb:.3f} ms ({2.5 * fl / tb / 1e9:.0f} TFLOP/s)   sdpa {tb_lib:.3f} ms ({2.5 * fl / tb_lib / 1e9:.0f} TFLOP/s)")


def ref_attn_grad(qkv, H):
    return ref_attn_nograd_free(qkv, H)


def ref_attn_nograd_free(qkv, H):
    B, N, _ = qkv.shape
    x = qkv.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
    q, k, v = x[0], x[1], x[2]
    s = (q @ k.transpose(-1, -2)) * 0.125
    p = torch.softmax(s, -1)
    o = p @ v
    return o.transpose(1, 2).reshape(B, N, H * 64), None, None


if __name__ == "__main__":
    main()
