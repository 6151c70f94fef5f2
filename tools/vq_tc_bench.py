import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagefolder_b200 import ops


def timeit(fn, warm=3, it=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / it


res = {}
for (V, C, B) in [(8192, 32, 256), (4096, 64, 128), (16384, 32, 128)]:
    torch.manual_seed(0)
    z = torch.randn(B, C, 16, 16, device="cuda")
    E = torch.nn.functional.normalize(torch.randn(V, C, device="cuda"), dim=-1)
    for algo in ("exact", "tc"):
        os.environ["XQ_VQ_ALGO"] = algo
        res[f"{algo}_V{V}_C{C}_B{B}_ms"] = timeit(lambda: ops.vq_lookup(z, E, True))
print(json.dumps(res, indent=1))
