"""dev tool: why is bench.py's host-buffer (e2e) region slower than the resident region?  Alternates both and prints
per-step device times + allocator statistics."""
import os, sys, argparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import bench
from imagefolder_b200 import config as xcfg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
model, margs = bench.build_model("VQ-8192", dev)
model.train()
opt = torch.optim.AdamW(model.parameters(), lr=3e-5, betas=(0.9, 0.95), weight_decay=0.0, fused=True)
imgs_dev = torch.rand(B, 3, 256, 256, device=dev) * 2 - 1
imgs_host = imgs_dev.cpu().pin_memory()
al, be, de = xcfg.perturbation_schedule(margs, 0)


def step(x):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        dec, (vq, cm, en, us), _, _, _ = model(x, 0, al, be, de)
        loss = F.mse_loss(dec.float(), x) + vq + cm + en
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


def stats(tag):
    s = torch.cuda.memory_stats()
    print(f"  [{tag}] reserved {torch.cuda.memory_reserved()/2**30:.1f} GiB allocated {torch.cuda.memory_allocated()/2**30:.1f} "
          f"retries {s.get('num_alloc_retries', 0)} cudaMalloc calls {s.get('segment.all.allocated', 0)} "
          f"frees {s.get('segment.all.freed', 0)}", flush=True)


def resident(n):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    evs[0].record()
    for i in range(n):
        step(imgs_dev)
        evs[i + 1].record()
    torch.cuda.synchronize()
    print("resident ms/step:", [round(evs[i].elapsed_time(evs[i + 1]), 1) for i in range(n)], flush=True)


def e2e(n, mode):
    copy_stream = torch.cuda.Stream(device=dev)
    loss_pinned = torch.empty(n, dtype=torch.float32).pin_memory()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]

    def prefetch():
        with torch.cuda.stream(copy_stream):
            xb = imgs_host.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return xb, ev

    evs[0].record()
    if mode == "same_stream":
        for i in range(n):
            x = imgs_host.to(dev, non_blocking=True)
            loss = step(x)
            loss_pinned[i:i + 1].copy_(loss.detach().float().reshape(1), non_blocking=True)
            evs[i + 1].record()
    else:
        nxt = prefetch()
        for i in range(n):
            x, ev = nxt
            torch.cuda.current_stream().wait_event(ev)
            if mode == "record_stream":
                x.record_stream(torch.cuda.current_stream())
            if i + 1 < n:
                nxt = prefetch()
            loss = step(x)
            loss_pinned[i:i + 1].copy_(loss.detach().float().reshape(1), non_blocking=True)
            evs[i + 1].record()
    torch.cuda.synchronize()
    print(f"e2e[{mode}] ms/step:", [round(evs[i].elapsed_time(evs[i + 1]), 1) for i in range(n)], flush=True)


for _ in range(3):
    step(imgs_dev)
torch.cuda.synchronize()
stats("after warmup")
resident(5); stats("resident")
e2e(5, "record_stream"); stats("e2e record_stream")
resident(4)
e2e(5, "same_stream"); stats("e2e same_stream")
e2e(5, "no_record"); stats("e2e no_record")
resident(4)
