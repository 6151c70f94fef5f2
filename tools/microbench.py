"""Kernel micro-benchmarks (CUDA events, after warm-up).  Not the bench.py contract -- a dev tool."""
import json
import sys
import os
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagefolder_b200 import LFQ, VectorQuantizer, VectorQuantizer2, add_perturbation  # noqa: E402


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    res = {}
    torch.manual_seed(0)
    for (V, C, B) in [(8192, 32, 256), (4096, 64, 128), (16384, 32, 128)]:
        q = VectorQuantizer(V, C).cuda().train()
        z = torch.randn(B, C, 16, 16, device="cuda", requires_grad=True)
        res[f"vq_fwd_V{V}_C{C}_B{B}_ms"] = timeit(lambda: q(z))
        res[f"vq_lookup_V{V}_C{C}_B{B}_ms"] = timeit(lambda: q.f_to_idxBl_or_fhat(z.detach(), False))

        def fb():
            out, _, vq, cm, _ = q(z)
            (out.sum() + vq + cm).backward()
        res[f"vq_fwdbwd_V{V}_C{C}_B{B}_ms"] = timeit(fb)
    q = VectorQuantizer(4096, 64).cuda().train()
    z = torch.randn(128, 64, 16, 16, device="cuda")
    zq = q(z)[0].detach()
    res["perturb_V4096_C64_B128_beta0.1_ms"] = timeit(lambda: add_perturbation(z, zq, 64, True, q.embedding, 1.0, 0.1, 100))
    pn = [1, 1, 2, 3, 3, 4, 5, 6, 8, 11]
    for V in (4096, 16384):
        m = VectorQuantizer2(V, 32, v_patch_nums=pn, num_latent_tokens=121, codebook_drop=0.1).cuda().train()
        f = torch.randn(128, 32, 11, 11, device="cuda", requires_grad=True)
        dr = torch.randint(3, 11, (128,))
        res[f"msvr_fwd_V{V}_B128_ms"] = timeit(lambda: m(f, True, dr))

        def fb2():
            out, _, vq, cm, _ = m(f, True, dr)
            (out.sum() + vq + cm).backward()
        res[f"msvr_fwdbwd_V{V}_B128_ms"] = timeit(fb2)
    l = LFQ(16384, 14, using_znorm=True, v_patch_nums=pn, num_latent_tokens=121, codebook_drop=0.1).cuda().train()
    f = torch.randn(128, 14, 11, 11, device="cuda", requires_grad=True)
    dr = torch.randint(3, 11, (128,))
    res["msbr_fwd_B128_ms"] = timeit(lambda: l(f, True, dr))

    def fb3():
        out, _, vq, cm, en = l(f, True, dr)
        (out.sum() + vq + cm + en).backward()
    res["msbr_fwdbwd_B128_ms"] = timeit(fb3)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
