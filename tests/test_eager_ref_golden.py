"""The plain-PyTorch eager restatement (oracle/eager_ref.py, the same-GPU baseline of bench.py --impl eager)
against the golden vectors from the reference's own modules: it must BE the reference's computation."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import eager_ref as er

RTOL = 2e-4


def close(a, b, rtol=RTOL):
    a = np.asarray(a.detach().numpy() if torch.is_tensor(a) else a, np.float64)
    b = np.asarray(b, np.float64)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=rtol * max(1e-30, float(np.abs(b).max())))


@pytest.mark.parametrize("name", ["vq4096_b1", "vq512_randn", "vq300_nonorm"])
def test_eager_vq(name):
    from imagefolder_b200 import VectorQuantizer
    g = load_golden(name)
    q = VectorQuantizer(*g["E"].shape, float(g["beta"]), bool(g["codebook_norm"])).train()
    q.embedding.weight.data.copy_(torch.tensor(g["E"]))
    z = torch.tensor(g["z"], requires_grad=True)
    for _ in range(int(g["steps"])):
        out, usage, vq, commit, idx = er.vq_forward(q, z)
    np.testing.assert_array_equal(idx.numpy(), g["idx"].reshape(-1))
    close(out, g["out"]), close(vq, g["vq"]), close(commit, g["commit"])
    assert abs(usage[0] - float(g["usage"])) < 1e-4
    (out * torch.tensor(g["g_out"])).sum().add(float(g["w_vq"]) * vq).add(float(g["w_commit"]) * commit).backward()
    close(z.grad, g["gz"])


@pytest.mark.parametrize("name", ["msvr_small", "msvr_l2", "msvr_shared1"])
def test_eager_vq2(name):
    from imagefolder_b200 import VectorQuantizer2
    g = load_golden(name)
    pn = [int(p) for p in g["patch_nums"]]
    V, C = g["E"].shape
    q = VectorQuantizer2(V, C, using_znorm=bool(g["using_znorm"]), v_patch_nums=pn, num_latent_tokens=pn[-1] ** 2,
                         share_quant_resi=int(g["share"]), codebook_drop=float(g["codebook_drop"])).train()
    q.embedding.weight.data.copy_(torch.tensor(g["E"]))
    for i, m in enumerate(q.quant_resi.modules_list()):
        m.weight.data.copy_(torch.tensor(g["phi_w"][i]))
        m.bias.data.copy_(torch.tensor(g["phi_b"][i]))
    f = torch.tensor(g["f"], requires_grad=True)
    for _ in range(int(g["steps"])):
        out, usages, vq, commit, _ = er.vq2_forward(q, f, torch.tensor(g["dropout"]))
    close(out, g["out"]), close(vq, g["vq"]), close(commit, g["commit"])
    close(np.array(usages), g["usages"], rtol=1e-5)
    (out * torch.tensor(g["g_out"])).sum().add(float(g["w_vq"]) * vq).add(float(g["w_commit"]) * commit).backward()
    close(f.grad, g["gf"]), close(q.embedding.weight.grad, g["gE"])


@pytest.mark.parametrize("name", ["msbr_small", "lfq_nonorm"])
def test_eager_lfq(name):
    from imagefolder_b200 import LFQ
    g = load_golden(name)
    pn = [int(p) for p in g["patch_nums"]]
    C = g["f"].shape[1]
    q = LFQ(2 ** C, C, using_znorm=bool(g["using_znorm"]), v_patch_nums=pn, num_latent_tokens=pn[-1] ** 2,
            codebook_drop=float(g["codebook_drop"]), scale=float(g["scale"]), entropy_weight=float(g["entropy_weight"])).train()
    for i, m in enumerate(q.quant_resi.modules_list()):
        m.weight.data.copy_(torch.tensor(g["phi_w"][i]))
        m.bias.data.copy_(torch.tensor(g["phi_b"][i]))
    f = torch.tensor(g["f"], requires_grad=True)
    out, usages, vq, commit, ent = er.lfq_forward(q, f, torch.tensor(g["dropout"]))
    close(out, g["out"]), close(vq, g["vq"]), close(commit, g["commit"]), close(ent, g["entropy"])
    (out * torch.tensor(g["g_out"])).sum().add(float(g["w_vq"]) * vq).add(float(g["w_commit"]) * commit).add(
        float(g["w_ent"]) * ent).backward()
    close(f.grad, g["gf"])


def test_eager_perturb():
    g = load_golden("perturb_a07")
    from imagefolder_b200 import VectorQuantizer
    q = VectorQuantizer(*g["E"].shape, 0.25, True)
    q.embedding.weight.data.copy_(torch.tensor(g["E"]))
    torch.manual_seed(101)            # make_golden.py: seed + 100 before the call
    out = er.perturb(torch.tensor(g["z"]), torch.tensor(g["zq"]), q, float(g["alpha"]), float(g["beta"]), int(g["delta"]))
    close(out, g["out"])
