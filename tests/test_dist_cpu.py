"""N > 1 host logic on CPU: world_size-2 gloo.  Covers the only collectives the path has besides DDP's
gradient all-reduce: the usage-histogram all-reduce (quant.py:104 collapsed to one [SN,V] op) and the
differentiable feature all-gather of the semantic ClipLoss (cliploss.py:49-50)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from imagefolder_b200.quant import _allreduce_hist_, _world_size
        from imagefolder_b200.xqgan_model import ClipLoss
        assert _world_size() == world
        # histogram: every rank contributes its local bincount, all ranks see the global one
        hist = torch.zeros(3, 16)
        hist[rank, rank::2] = 1.0 + rank
        _allreduce_hist_(hist)
        expect = torch.zeros(3, 16)
        for r in range(world):
            expect[r, r::2] = 1.0 + r
        assert torch.equal(hist, expect)
        # usage margin scales with world size (quant.py:137)
        margin = _world_size() * 128 / 64 * 0.08
        assert abs(margin - 2 * 128 / 64 * 0.08) < 1e-12
        # ClipLoss with gather_with_grad == single-process loss over the concatenated batch
        g = torch.Generator().manual_seed(0)
        a_all = torch.randn(8, 6, generator=g)
        b_all = torch.randn(8, 6, generator=g)
        a = a_all[rank * 4:(rank + 1) * 4].clone().requires_grad_(True)
        b = b_all[rank * 4:(rank + 1) * 4].clone().requires_grad_(True)
        loss = ClipLoss(local_loss=False, gather_with_grad=True, rank=rank, world_size=world)(a, b, 3.0)
        loss.backward()
        a1, b1 = a_all.clone().requires_grad_(True), b_all.clone().requires_grad_(True)
        ref = ClipLoss(world_size=1)(a1, b1, 3.0)
        ref.backward()
        assert abs(float(loss) - float(ref)) < 1e-6
        # torch.distributed.nn.all_gather's backward sums the gradient over ranks: each rank's loss is the same
        # global loss, so the local slice of d(loss)/d(features) is world x the single-process gradient
        np.testing.assert_allclose(a.grad.numpy(), world * a1.grad[rank * 4:(rank + 1) * 4].numpy(), rtol=1e-5, atol=1e-7)
        # rFID data path (xqgan_train.py:517-535): uint8 NHWC conversion + rank-major all-gather of reconstructions and ground truth
        from imagefolder_b200.evaluate import reconstruct_for_fid, to_uint8_nhwc

        class _Identity(torch.nn.Module):                # stands in for VQModel: the loop only needs img_to_reconstructed_img
            def __init__(self):
                super().__init__()
                self.w = torch.nn.Parameter(torch.zeros(1))

            def img_to_reconstructed_img(self, x):
                return x * 0.5

        gi = torch.Generator().manual_seed(100 + rank)
        batches = [(torch.rand(3, 3, 4, 4, generator=gi) * 2 - 1, None) for _ in range(2)]
        m = _Identity().train()
        smp, gt_, tot = reconstruct_for_fid(m, batches, device="cpu")
        assert m.training and tot == 2 * 3 * world and smp.dtype == np.uint8 and smp.shape == (tot, 4, 4, 3)
        ref_batches = []
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            ref_batches.append([torch.rand(3, 3, 4, 4, generator=gr) * 2 - 1 for _ in range(2)])
        want_gt = np.concatenate([np.concatenate([to_uint8_nhwc(ref_batches[r][bi]).numpy() for r in range(world)]) for bi in range(2)])
        want_s = np.concatenate([np.concatenate([to_uint8_nhwc(ref_batches[r][bi] * 0.5).numpy() for r in range(world)]) for bi in range(2)])
        assert np.array_equal(gt_, want_gt) and np.array_equal(smp, want_s)
        # rank-dependent synthetic data seeds (bench.py) differ
        g2 = torch.Generator().manual_seed(1234 * world + rank)
        q.put((rank, float(torch.rand(1, generator=g2))))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_host_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=180)
        assert p.exitcode == 0
    vals = dict(q.get(timeout=5) for _ in range(2))
    assert vals[0] != vals[1]
