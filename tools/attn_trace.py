"""dev tool: in-kernel clock trace of CTA 0 of attn_bwd_kernel (needs tools/mb/libxq_exp_trace.so, built with
-DXQ_ATTN_TRACE).  Prints, per query block i, when the MMA thread saw p_full / ds_full and issued dQ, and when the two
compute warpgroups saw S, finished P, finished the dQ drain, saw dP and finished dS -- all relative to the kernel start."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagefolder_b200 import _capi as C  # noqa: E402

C.LIB_PATH = os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else "tools/mb/libxq_exp_trace.so")
from imagefolder_b200 import vit_ops  # noqa: E402

B, N, H = 256, 513, 12
torch.manual_seed(0)
qkv = torch.randn(B, N, 3 * H * 64, device="cuda").to(torch.bfloat16)
g = torch.randn(B, N, H * 64, device="cuda").to(torch.bfloat16)
out, lse = vit_ops.attn_tc_forward(qkv, H)
L = C.lib()
trace = torch.zeros(16 * 32, dtype=torch.int64, device="cuda")
L.xq_dev_set_attn_trace.argtypes = [ctypes.c_void_p]
for _ in range(2):
    dqkv = vit_ops.attn_tc_backward(qkv, out, lse, g, H)
torch.cuda.synchronize()
assert L.xq_dev_set_attn_trace(trace.data_ptr()) == 0
dqkv = vit_ops.attn_tc_backward(qkv, out, lse, g, H)
torch.cuda.synchronize()
t = trace.cpu().tolist()
t0 = t[16 * 30 + 2]
names = ["mma:p_full", "mma:ds_full", "mma:dQ issued", "-", "wgE:S seen", "wgE:P done", "-", "-", "wgD:P seen", "wgD:dP seen", "wgD:dS done",
         "wgD:chunks done", "wgQ:drained", "wgD:ch0", "wgD:ch1", "wgD:ch2"]
print("CTA 0 start -> end:", t[16 * 30 + 3] - t0, "clocks")
print("  kernel entry -> setup done:", t0 - t[16 * 30 + 4], " wgE0 reaches epilogue:", t[16 * 30 + 5] - t0, " dkv_done seen:", t[16 * 30 + 6] - t0,
      " dV store issued:", t[16 * 30 + 7] - t0, " dV store + bias sums done:", t[16 * 30 + 8] - t0)
for i in range(16):
    row = [(names[k], t[16 * i + k] - t0) for k in range(16) if t[16 * i + k]]
    print(f"i={i}: " + "  ".join(f"{n}={v}" for n, v in row))
