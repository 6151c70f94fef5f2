"""How many host threads should the CPU baseline use?  Times the oracle training step at several
torch thread counts (bounded by an alarm) on the box it runs on."""
import os, signal, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import vit_ref, xq_oracle as xo


class TO(Exception):
    pass


def alarm(*_):
    raise TO()


def main():
    model, _ = bench.build_model("VQ-8192", "cpu")
    cfg = vit_ref.cfg_from_model_args(model.config)
    ref = vit_ref.RefTokenizer(model.state_dict(), cfg, requires_grad=True)
    opt = torch.optim.AdamW(ref.parameters(), lr=3e-5)
    signal.signal(signal.SIGALRM, alarm)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    x = torch.rand(n, 3, 256, 256) * 2 - 1
    for t in [16, 32, 64, os.cpu_count()]:
        torch.set_num_threads(t)
        xo.set_num_threads(min(t, 16))
        try:
            signal.alarm(90)
            ref.train_step(x, opt)
            t0 = time.time()
            ref.train_step(x, opt)
            dt = time.time() - t0
            signal.alarm(0)
            print(f"threads={t} n={n} step={dt:.2f}s  {n/dt:.3f} img/s", flush=True)
        except TO:
            print(f"threads={t} n={n} TIMEOUT >90s", flush=True)


if __name__ == "__main__":
    main()
