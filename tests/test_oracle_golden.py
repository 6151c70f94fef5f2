"""Pins the CPU oracle to the reference: oracle outputs vs golden vectors produced by running
the reference's own modules (tests/golden/make_golden.py).  Indices must be equal except on
provable near-ties (top-2 margin below 1e-5), floats within 1e-4 relative."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import xq_oracle as xo

RTOL = 2e-4


def close(a, b, rtol=RTOL, atol=None):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if atol is None:
        atol = rtol * max(1e-30, float(np.abs(b).max()))
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


def check_idx(mine, ref, margin, tie=1e-5):
    mine, ref = np.asarray(mine).reshape(-1), np.asarray(ref).reshape(-1)
    bad = mine != ref
    if bad.any():
        assert np.all(np.asarray(margin).reshape(-1)[bad] < tie), "index mismatch that is not a near-tie"
    return int(bad.sum())


@pytest.mark.parametrize("name", ["vq4096_b1", "vq512_randn", "vq300_nonorm"])
def test_vq_forward_backward(name):
    g = load_golden(name)
    cn = bool(g["codebook_norm"])
    fwd = xo.vq_forward(g["z"], g["E"], beta=float(g["beta"]), codebook_norm=cn)
    assert check_idx(fwd["idx"], g["idx"], fwd["margin"]) == 0
    close(fwd["out"], g["out"])
    close(fwd["q_nchw"], g["fhat"])
    close(fwd["vq"], g["vq"])
    close(fwd["commit"], g["commit"])
    gz, gE = xo.vq_backward(fwd, g["E"], g["g_out"], float(g["w_vq"]), float(g["w_commit"]), float(g["beta"]), cn)
    close(gz, g["gz"])
    gE_ref = np.zeros_like(gE)
    gE_ref[g["gE_rows"]] = g["gE_vals"]
    close(gE, gE_ref)
    # EMA / usage after `steps` identical forwards (xqgan_model.py:773-788)
    ema = np.zeros(g["E"].shape[0], np.float32)
    for s in range(int(g["steps"])):
        ema = xo.ema_update(ema, fwd["hist"], s)
    close(ema, g["ema"], rtol=1e-6)
    N = g["z"].size // g["z"].shape[1]
    margin = 1 * N / g["E"].shape[0] * 0.08
    assert abs(float((ema >= margin).mean() * 100) - float(g["usage"])) < 1e-4


@pytest.mark.parametrize("name", ["perturb_a07", "perturb_a0"])
def test_add_perturbation(name):
    g = load_golden(name)
    cn = bool(g["codebook_norm"])
    fwd = xo.add_perturbation(g["z"], g["zq"], g["E"], cn, float(g["alpha"]), float(g["beta"]), int(g["delta"]),
                              g["rand_u"], g["rand_j"])
    close(fwd["out"], g["out"])
    gz, gzq = xo.add_perturbation_backward(fwd, g["g"], cn)
    close(gz, g["gz"], atol=1e-6)
    close(gzq, g["gzq"])


@pytest.mark.parametrize("name", ["msvr_small", "msvr_4096", "msvr_l2", "msvr_shared1"])
def test_vq2(name):
    g = load_golden(name)
    pn = [int(p) for p in g["patch_nums"]]
    zn = bool(g["using_znorm"])
    fwd = xo.vq2_forward(g["f"], g["E"], g["phi_w"], g["phi_b"], pn, using_znorm=zn, codebook_drop=float(g["codebook_drop"]),
                         dropout=g["dropout"])
    for si in range(len(pn)):
        assert check_idx(fwd["idx"][si], g[f"idx{si}"], fwd["margins"][si]) == 0
    close(fwd["out"], g["out"])
    close(fwd["vq"], g["vq"])
    close(fwd["commit"], g["commit"])
    gf, gE, gw, gb = xo.vq2_backward(fwd, g["f"], g["E"], g["phi_w"], g["phi_b"], pn, g["g_out"], float(g["w_vq"]),
                                     float(g["w_commit"]))
    close(gf, g["gf"])
    close(gE, g["gE"])
    close(gw, g["gphi_w"])
    close(gb, g["gphi_b"])
    fh = xo.vq2_f_to_idxBl_or_fhat(g["f"], g["E"], g["phi_w"], g["phi_b"], pn, using_znorm=zn, to_fhat=True)
    close(fh[-1], g["fhat_last"])
    close(fh[len(pn) // 2], g["fhat_mid"])
    # EMA rows: record_hit increments once per scale (quant.py:121-127)
    SN, V = len(pn), g["E"].shape[0]
    ema = np.zeros((SN, V), np.float32)
    rec = 0
    for _ in range(int(g["steps"])):
        for si in range(SN):
            ema[si] = xo.ema_update(ema[si], fwd["hist"][si], rec)
            rec += 1
    close(ema, g["ema"], rtol=1e-6)


@pytest.mark.parametrize("name", ["msbr_small", "msbr_14", "lfq_nonorm"])
def test_lfq(name):
    g = load_golden(name)
    pn = [int(p) for p in g["patch_nums"]]
    zn = bool(g["using_znorm"])
    kw = dict(using_znorm=zn, codebook_drop=float(g["codebook_drop"]), dropout=g["dropout"], scale=float(g["scale"]),
              entropy_weight=float(g["entropy_weight"]))
    fwd = xo.lfq_forward(g["f"], g["phi_w"], g["phi_b"], pn, **kw)
    close(fwd["scaler"], g["scaler"], rtol=1e-6)
    for si in range(len(pn)):
        np.testing.assert_array_equal(fwd["idx"][si], g[f"idx{si}"])
    close(fwd["out"], g["out"])
    close(fwd["vq"], g["vq"])
    close(fwd["commit"], g["commit"])
    close(fwd["entropy"], g["entropy"])
    gf, gw, gb = xo.lfq_backward(fwd, g["f"], g["phi_w"], g["phi_b"], pn, g["g_out"], float(g["w_vq"]),
                                 float(g["w_commit"]), float(g["w_ent"]), using_znorm=zn,
                                 entropy_weight=float(g["entropy_weight"]))
    close(gf, g["gf"])
    close(gw, g["gphi_w"])
    close(gb, g["gphi_b"])


@pytest.mark.parametrize("name", ["varhelp_msvr", "varhelp_shared1", "varhelp_lfq"])
def test_var_feature_map_helpers(name):
    """oracle embed_to_fhat / get_next_autoregressive_input vs the reference modules' outputs
    (quant.py:148-166, 247-258; lookup_free_quantize.py:311-328, 404-415)."""
    g = load_golden(name)
    pn = [int(p) for p in g["patch_nums"]]
    SN = len(pn)
    hs = [g[f"h{si}"] for si in range(SN)]
    fl = xo.embed_to_fhat(hs, g["phi_w"], g["phi_b"], pn)
    for si in range(SN):
        close(fl[si], g[f"fh{si}"])
    close(xo.embed_to_fhat(hs, g["phi_w"], g["phi_b"], pn, last_one=True), g["fh_last"])
    F = np.zeros_like(g["fh_last"])
    for si in range(SN):
        F, nxt = xo.get_next_autoregressive_input(si, F, hs[si], g["phi_w"], g["phi_b"], pn)
        assert nxt.shape == g[f"next{si}"].shape
        close(nxt, g[f"next{si}"])
    close(F, g["ar_f_hat"])
    np.testing.assert_array_equal(F, fl[-1])          # the AR chain and embed_to_fhat are the same arithmetic


def test_oracle_edge_cases_ties_and_degenerate_sizes():
    """first-index tie-break, single row / single code / single channel, zero rows (the normalisation clamp) --
    the conventions the CUDA kernels are held to."""
    rng = np.random.default_rng(0)
    # duplicated codes: the FIRST of the equal codes wins, for both metrics
    codes = rng.standard_normal((6, 5)).astype(np.float32)
    codes[4] = codes[1]
    rows = np.stack([codes[1], codes[4] * 1.0, codes[3]]).astype(np.float32)
    for metric in (0, 1):
        idx, best, second = xo.search(rows, codes if metric == 0 else xo.l2norm_rows(codes)[0], metric)
        assert list(idx) == [1, 1, 3]
    # V = 1, N = 1, C = 1
    idx, _, _ = xo.search(np.array([[2.0]], np.float32), np.array([[-1.0]], np.float32), 0)
    assert list(idx) == [0]
    # an all-zero row normalises to zero (den clamped to 1e-12) and still gets a valid index
    y, den = xo.l2norm_rows(np.zeros((2, 4), np.float32))
    assert np.all(y == 0) and np.allclose(den, 1e-12)
    fwd = xo.vq_forward(np.zeros((1, 4, 2, 2), np.float32), rng.standard_normal((7, 4)).astype(np.float32))
    assert fwd["idx"].shape == (4,) and np.all((fwd["idx"] >= 0) & (fwd["idx"] < 7)) and np.isfinite(fwd["out"]).all()
    # rank_select with delta > number of distinct distances: ranks follow (distance, index) order
    codes = np.array([[1.0, 0.0], [1.0, 0.0], [0.0, 1.0]], np.float32)
    r = np.array([[1.0, 0.0]], np.float32)
    out, topk = xo.rank_select(r, codes, np.array([1]), 3, want_topk=True)
    assert list(topk[0]) == [0, 1, 2] and list(out) == [1]
    # area pool P = H is the identity, P = 1 the global mean; bicubic P = H is the identity
    f = rng.standard_normal((2, 3, 5, 5)).astype(np.float32)
    np.testing.assert_array_equal(xo.rows_to_nchw(xo.area_pool_rows(f, 5), f.shape), f)
    np.testing.assert_allclose(xo.area_pool_rows(f, 1).reshape(2, 3), f.reshape(2, 3, -1).mean(-1), rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(xo.bicubic_up(xo.nchw_to_rows(f), 2, 3, 5, 5, 5), f)


# ---- round 2: BASELINE-shaped codebooks (configs #2 / #3) and an unscreened multi-scale case ------------------------
def _det_inputs(shape, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).numpy()


def _det_param(name, shape):
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from vit_det_init import det_tensor
    return det_tensor(name, shape).numpy()


def big_vq_inputs(g):
    V, C, B, hw, seed = (int(g[k]) for k in ("V", "C", "B", "hw", "seed"))
    E = _det_param("embedding.weight", (V, C))
    z = _det_inputs((B, C, hw, hw), seed)
    g_out = _det_inputs((B, C, hw, hw), seed + 1)
    return E, z, g_out


@pytest.mark.parametrize("name", ["vq8192_c32", "vq16384_c32"])
def test_vq_baseline_shaped_codebooks(name):
    """reference goldens at V = 8192 / 16384, C = 32 (BASELINE configs #2 / #3), 1024 rows."""
    g = load_golden(name)
    E, z, g_out = big_vq_inputs(g)
    fwd = xo.vq_forward(z, E, beta=0.25, codebook_norm=True)
    assert check_idx(fwd["idx"], g["idx"], fwd["margin"]) == 0
    close(fwd["out"][:, :, ::2, ::2], g["out_sub"])
    close(fwd["vq"], g["vq"])
    close(fwd["commit"], g["commit"])
    gz, gE = xo.vq_backward(fwd, E, g_out, float(g["w_vq"]), float(g["w_commit"]), 0.25, True)
    close(gz[:, :, ::2, ::2], g["gz_sub"])
    gE_ref = np.zeros_like(gE)
    gE_ref[g["gE_rows"]] = g["gE_vals"]
    close(gE, gE_ref)


def msvr_unscreened_inputs(g):
    V, C, B, seed = (int(g[k]) for k in ("V", "C", "B", "seed"))
    pn = [int(p) for p in g["patch_nums"]]
    E = _det_param("embedding.weight", (V, C))
    phi_w = np.stack([_det_param(f"quant_resi.qresi_ls.{k}.weight", (C, C, 3, 3)) for k in range(4)])
    phi_b = np.stack([_det_param(f"quant_resi.qresi_ls.{k}.bias", (C,)) for k in range(4)])
    f = _det_inputs((B, C, pn[-1], pn[-1]), seed)
    return E, phi_w, phi_b, f, pn


def count_first_divergences(idx_mine, idx_ref, margins, B, tie=1e-5):
    """Per sample: the first scale where the index lists differ must differ only at near-tie positions (a flipped index
    changes the residual of every later scale, so later scales of that sample are not comparable).  Returns the number of
    samples that diverged and the number of differing tokens at their first divergent scale."""
    diverged, tokens = 0, 0
    for b in range(B):
        for si in range(len(idx_ref)):
            bad = np.asarray(idx_mine[si][b]) != np.asarray(idx_ref[si][b])
            if bad.any():
                m = np.asarray(margins[si]).reshape(B, -1)[b]
                assert np.all(m[bad] < tie), f"sample {b} scale {si}: index mismatch with top-2 margin {m[bad].max():.3e}"
                diverged += 1
                tokens += int(bad.sum())
                break
    return diverged, tokens


def test_msvr_unscreened_seed_mismatches_are_near_ties():
    """make_golden.case_vq2 screens seeds for near-ties; this case does not: mismatches against the reference's indices
    are COUNTED and each must be a provable near-tie."""
    g = load_golden("msvr_unscreened")
    E, phi_w, phi_b, f, pn = msvr_unscreened_inputs(g)
    fw = xo.vq2_forward(f, E, phi_w, phi_b, pn, using_znorm=True)
    ref = [g[f"idx{si}"] for si in range(len(pn))]
    diverged, tokens = count_first_divergences(fw["idx"], ref, fw["margins"], f.shape[0])
    print("msvr_unscreened: samples diverged", diverged, "tokens at first divergence", tokens)
    assert diverged <= 1          # observed 0 on this seed; a single near-tie flip would still be legitimate
