"""dev tool: the fused MLP GEMMs (xq_vit_fc1_gelu_fwd / xq_vit_fc2_dgelu_bwd, csrc/gemm_kernel.cu) against the two-call sequences they
replace (library GEMM + stand-alone bias / GELU kernel) at the bench shape, with a bit-level comparison.
   python tools/mlp_gemm_bench.py [M N K]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagefolder_b200 import _capi as C, vit_ops  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (131328, 3072, 768)
dev = torch.device("cuda")
torch.manual_seed(0)
y = torch.randn(M, K, device=dev).to(torch.bfloat16)
W1 = (torch.randn(N, K, device=dev) * 0.03).to(torch.bfloat16)
b1 = torch.randn(N, device=dev) * 0.1
W2t = (torch.randn(N, K, device=dev) * 0.03).to(torch.bfloat16)     # fc2.weight^T
g = torch.randn(M, K, device=dev).to(torch.bfloat16)
L = C.lib()
st = C.stream_ptr(dev)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


pre = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
act = torch.empty_like(pre)
dpre = torch.empty_like(pre)
db = torch.empty(N, device=dev)


def fused_fwd():
    C.call("xq_vit_fc1_gelu_fwd", 1, L.xq_vit_fc1_gelu_fwd, C.ptr(y), C.ptr(W1), C.ptr(b1), C.ptr(pre), C.ptr(act), M, N, K, st)


def fused_bwd():
    C.call("xq_vit_fc2_dgelu_bwd", 1, L.xq_vit_fc2_dgelu_bwd, C.ptr(g), C.ptr(W2t), C.ptr(pre), C.ptr(b1), C.ptr(dpre), C.ptr(db), M, N, K, st)


def lib_fwd():
    p = y @ W1.t()
    return p, vit_ops.gelu_bias(p, b1)


gy_holder = {}


def lib_bwd():
    da = g @ W2t.t()
    gx = torch.empty_like(da)
    gb = torch.empty_like(b1)
    C.call("xq_vit_gelu_bwd", 1, L.xq_vit_gelu_bwd, C.ptr(pre), C.ptr(b1), C.ptr(da), C.ptr(gx), C.ptr(gb), M, N, st)
    return gx, gb


fused_fwd()
p_ref, a_ref = lib_fwd()
fused_bwd()
gx_ref, gb_ref = lib_bwd()
torch.cuda.synchronize()
print(f"M={M} N={N} K={K}")
print("pre  identical:", torch.equal(pre, p_ref), "  act identical:", torch.equal(act, a_ref))
exact = dpre.double().sum(0)
print("dpre identical:", torch.equal(dpre, gx_ref),
      f"  d_bias vs fp64 column sums of d_pre (max err / max |sum|): fused {((db.double() - exact).abs().max() / exact.abs().max()).item():.2e}"
      f", stand-alone kernel {((gb_ref.double() - exact).abs().max() / exact.abs().max()).item():.2e}")
fl = 2.0 * M * N * K
t = timeit(fused_fwd); tl = timeit(lib_fwd)
print(f"forward : fused {t:.3f} ms ({fl / t / 1e9:.0f} TFLOP/s)   library GEMM + gelu_fwd kernel {tl:.3f} ms")
t = timeit(fused_bwd); tl = timeit(lib_bwd)
print(f"backward: fused {t:.3f} ms ({fl / t / 1e9:.0f} TFLOP/s)   library GEMM + gelu_bwd kernel {tl:.3f} ms")
