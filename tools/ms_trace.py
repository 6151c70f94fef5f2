import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagefolder_b200 import VectorQuantizer2, _capi as C
V = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
pn = [1, 1, 2, 3, 3, 4, 5, 6, 8, 11]
q = VectorQuantizer2(V, 32, v_patch_nums=pn, num_latent_tokens=121, codebook_drop=0.1).cuda().train()
f = torch.randn(128, 32, 11, 11, device="cuda")
dr = torch.randint(3, 11, (128,))
for _ in range(3): q(f, True, dr)
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
# needs a development build of the library: add -DXQ_MS_TRACE to the nvcc flags of csrc/build.sh
import ctypes
_L = C.lib()
_L.xq_dev_set_ms_trace.argtypes = [ctypes.c_void_p]
_L.xq_dev_set_ms_trace(dbg.data_ptr())
q(f, True, dr)
torch.cuda.synchronize()
d = dbg.cpu().tolist()
prev = None
for si, p in enumerate(pn):
    a = d[4 * si:4 * si + 4]
    start = prev if prev is not None else a[0]
    print(f"scale {si} pn={p:2d}: pool {a[0]-start:7d}  search {a[1]-a[0]:7d}  gather+bicubic {a[2]-a[1]:7d}  phi+update {a[3]-a[2]:7d}")
    prev = a[3]
print("total", d[4 * 9 + 3] - d[0])
