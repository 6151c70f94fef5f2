"""dev tool: one forward + one backward of the tcgen05 attention at the VQ-8192 bench shape, for ncu captures:
   ncu --set full --clock-control none --import-source on -k regex:attn_ -o gpurun_out/attn_prof python tools/attn_ncu_target.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagefolder_b200 import vit_ops  # noqa: E402

B, N, H = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (256, 513, 12)))
torch.manual_seed(0)
qkv = torch.randn(B, N, 3 * H * 64, device="cuda").to(torch.bfloat16)
g = torch.randn(B, N, H * 64, device="cuda").to(torch.bfloat16)
for _ in range(2):
    out, lse = vit_ops.attn_tc_forward(qkv, H)
    dqkv = vit_ops.attn_tc_backward(qkv, out, lse, g, H)
torch.cuda.synchronize()
print("done", float(out.float().abs().mean()), float(dqkv.float().abs().mean()))
