"""torch.profiler kernel table of one training step (dev tool; numbers under a profiler are not bench values)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import bench
from imagefolder_b200 import config as xcfg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
wl = sys.argv[2] if len(sys.argv) > 2 else "VQ-8192"
dev = torch.device("cuda", 0)
model, margs = bench.build_model(wl, dev)
model.train()
opt = torch.optim.AdamW(model.parameters(), lr=3e-5, betas=(0.9, 0.95), fused=True)
x = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
al, be, de = xcfg.perturbation_schedule(margs, 0)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        dec, (vq, cm, en, us), _, _, _ = model(x, 0, al, be, de)
        loss = F.mse_loss(dec.float(), x) + vq + cm + en
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in ev)
print(f"total device ms {tot/1e3:.1f}")
for e in ev[:70]:
    print(f"{e.self_device_time_total/1e3:9.2f} ms {100*e.self_device_time_total/tot:5.1f}% n={e.count:5d} {e.key[:110]}")

# GPU idle analysis: gaps between consecutive kernels on the device timeline of the profiled step
ks = sorted(((e.time_range.start, e.time_range.end) for e in prof.events()
             if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda t: t[0])
if ks:
    busy_end, idle, gaps = ks[0][1], 0.0, []
    for st, en in ks[1:]:
        if st > busy_end:
            idle += st - busy_end
            gaps.append(st - busy_end)
        busy_end = max(busy_end, en)
    span = busy_end - ks[0][0]
    gaps.sort(reverse=True)
    print(f"device span {span/1e3:.1f} ms, idle {idle/1e3:.2f} ms in {len(gaps)} gaps; top gaps us: {[round(g) for g in gaps[:12]]}; "
          f"median gap us: {gaps[len(gaps)//2] if gaps else 0:.1f}")
