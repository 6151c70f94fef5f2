import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagefolder_b200 import _capi as C, ops
V, C, B = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (8192, 32, 256)
z = torch.randn(B, C, 16, 16, device="cuda")
E = torch.nn.functional.normalize(torch.randn(V, C, device="cuda"), dim=-1)
os.environ["XQ_VQ_ALGO"] = "tc"
for _ in range(3): ops.vq_lookup(z, E, True)
dbg = torch.zeros(512, dtype=torch.int64, device="cuda")
# needs a development build of the library: add -DXQ_VQ_TC_TRACE to the nvcc flags of csrc/build.sh
import ctypes
_L = C.lib()
_L.xq_dev_set_vq_trace.argtypes = [ctypes.c_void_p]
_L.xq_dev_set_vq_trace(dbg.data_ptr())
ops.vq_lookup(z, E, True)
torch.cuda.synchronize()
d = dbg.cpu().tolist()
t0 = d[0]
print("prologue", d[1] - t0, "epilogue-loop-end", d[2] - t0, "rescored", d[4] - t0, "kernel-end", d[3] - t0)
for t in range(12):
    a = [d[8 + 4 * t + i] - t0 for i in range(4)]
    print(f"tile {t:2d}: mma tempty-ok {a[0]:8d}  full-ok {a[1]:8d}  issued {a[2]:8d} | epi done {a[3]:8d}")
T = V // 128
for t in range(T - 4, T):
    if t < 60:
        a = [d[8 + 4 * t + i] - t0 for i in range(4)]
        print(f"tile {t:2d}: mma tempty-ok {a[0]:8d}  full-ok {a[1]:8d}  issued {a[2]:8d} | epi done {a[3]:8d}")
