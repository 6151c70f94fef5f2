"""numpy front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; the product (imagefolder_b200/) never does.

Index-bearing arithmetic (normalise, search, area pool, bicubic, Phi, residual chain) runs in
oracle/xq_oracle.c in canonical fp32; losses and gradients are closed forms evaluated here in
float64 (re-derived from the reference code cited per function).
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libxq_oracle.so")
        if not os.path.exists(path):
            import importlib.util
            spec = importlib.util.spec_from_file_location("_xq_oracle_build", os.path.join(_HERE, "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build()
        _LIB = ctypes.CDLL(path)
        _LIB.xqo_num_threads.restype = ctypes.c_int
    return _LIB


def _f(a):
    return None if a is None else a.ctypes.data_as(_f32p)


def _l(a):
    return None if a is None else a.ctypes.data_as(_i64p)


def _i(a):
    return None if a is None else a.ctypes.data_as(_i32p)


def _c32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads() -> int:
    return int(lib().xqo_num_threads())


def set_num_threads(n: int) -> None:
    lib().xqo_set_num_threads(int(n))


# ----------------------------------------------------------------------------------------
# primitives
# ----------------------------------------------------------------------------------------
def l2norm_rows(x: np.ndarray):
    x = _c32(x)
    n, C = x.shape
    y = np.empty_like(x)
    den = np.empty(n, np.float32)
    lib().xqo_l2norm_rows(_f(x), ctypes.c_int64(n), C, _f(y), _f(den))
    return y, den


def nchw_to_rows(x: np.ndarray) -> np.ndarray:
    x = _c32(x)
    B, C = x.shape[:2]
    HW = int(np.prod(x.shape[2:]))
    rows = np.empty((B * HW, C), np.float32)
    lib().xqo_nchw_to_rows(_f(x), B, C, HW, _f(rows))
    return rows


def rows_to_nchw(rows: np.ndarray, shape) -> np.ndarray:
    rows = _c32(rows)
    B, C = shape[:2]
    HW = int(np.prod(shape[2:]))
    x = np.empty(shape, np.float32)
    lib().xqo_rows_to_nchw(_f(rows), B, C, HW, _f(x))
    return x


def search(rows: np.ndarray, codes: np.ndarray, metric: int):
    """metric 0: argmin (zz+ee)-2dot ; metric 1: argmax dot.  Returns idx, best, second."""
    rows, codes = _c32(rows), _c32(codes)
    n, C = rows.shape
    V = codes.shape[0]
    idx = np.empty(n, np.int64)
    best = np.empty(n, np.float32)
    second = np.empty(n, np.float32)
    lib().xqo_search(_f(rows), ctypes.c_int64(n), _f(codes), V, C, metric, _l(idx), _f(best), _f(second))
    return idx, best, second


def rank_select(rows, codes, rank, delta, want_topk=False):
    rows, codes = _c32(rows), _c32(codes)
    n, C = rows.shape
    V = codes.shape[0]
    rank = np.ascontiguousarray(rank, np.int64)
    out = np.empty(n, np.int64)
    topk = np.empty((n, delta), np.int64) if want_topk else None
    lib().xqo_rank_select(_f(rows), ctypes.c_int64(n), _f(codes), V, C, _l(rank), int(delta), _l(out), _l(topk))
    return (out, topk) if want_topk else out


def cubic_table(in_size: int, out_size: int):
    idx = np.empty((out_size, 4), np.int32)
    w = np.empty((out_size, 4), np.float32)
    lib().xqo_cubic_table(in_size, out_size, _i(idx), _f(w))
    return idx, w


def cubic_matrix(in_size: int, out_size: int) -> np.ndarray:
    """Dense [out,in] float64 matrix of the bicubic map (for the closed-form backward)."""
    idx, w = cubic_table(in_size, out_size)
    M = np.zeros((out_size, in_size), np.float64)
    if in_size == out_size:
        return np.eye(out_size)
    for d in range(out_size):
        for k in range(4):
            M[d, idx[d, k]] += float(w[d, k])
    return M


def area_pool_rows(f: np.ndarray, P: int) -> np.ndarray:
    f = _c32(f)
    B, C, H, W = f.shape
    rows = np.empty((B * P * P, C), np.float32)
    lib().xqo_area_pool_rows(_f(f), B, C, H, W, P, _f(rows))
    return rows


def bicubic_up(src_rows: np.ndarray, B: int, C: int, P: int, H: int, W: int) -> np.ndarray:
    src_rows = _c32(src_rows)
    u = np.empty((B, C, H, W), np.float32)
    lib().xqo_bicubic_up(_f(src_rows), B, C, P, H, W, _f(u))
    return u


def phi(u: np.ndarray, w: np.ndarray, b: np.ndarray, r: float) -> np.ndarray:
    u, w, b = _c32(u), _c32(w), _c32(b)
    B, C, H, W = u.shape
    h = np.empty_like(u)
    lib().xqo_phi(_f(u), B, C, H, W, _f(w), _f(b), ctypes.c_float(r), _f(h))
    return h


def phi_map(SN: int, K: int) -> List[int]:
    """scale -> Phi index (PhiPartiallyShared.__getitem__, quant.py:279-288; PhiShared :271;
    PhiNonShared :294-302).  K = number of Phi modules."""
    if K == 1:
        return [0] * SN
    ticks = np.linspace(1 / 3 / K, 1 - 1 / 3 / K, K) if K == 4 else np.linspace(1 / 2 / K, 1 - 1 / 2 / K, K)
    if SN == 1:
        return [0]
    return [int(np.argmin(np.abs(ticks - si / (SN - 1)))) for si in range(SN)]


def _norm_jvp_T(g, y, den):
    """Transpose-Jacobian of y = x / max(|x|, eps) applied to g (rows).  Where the clamp is
    active (|x| < eps) the map is x/eps and the Jacobian is I/eps."""
    g = g.astype(np.float64)
    y = y.astype(np.float64)
    den = den.astype(np.float64)[:, None]
    proj = g - y * np.sum(y * g, axis=1, keepdims=True)
    return np.where(den > 1e-12, proj, g) / den


# ----------------------------------------------------------------------------------------
# VectorQuantizer (single scale)            xqgan_model.py:745-833
# ----------------------------------------------------------------------------------------
def vq_forward(z: np.ndarray, E: np.ndarray, beta: float = 0.25, codebook_norm: bool = True) -> Dict:
    z, E = _c32(z), _c32(E)
    B, C = z.shape[:2]
    rows = nchw_to_rows(z)
    if codebook_norm:
        zn, zden = l2norm_rows(rows)
        En, _ = l2norm_rows(E)
    else:
        zn, zden, En = rows, np.ones(rows.shape[0], np.float32), E
    idx, best, second = search(zn, En, 0)
    y = E[idx]
    if codebook_norm:
        q, yden = l2norm_rows(y)
    else:
        q, yden = y, np.ones(rows.shape[0], np.float32)
    diff = q.astype(np.float64) - zn.astype(np.float64)
    mse = float(np.mean(diff ** 2))
    out_rows = zn + (q - zn)  # fp32, xqgan_model.py:796
    return dict(idx=idx, zn=zn, zden=zden, q=q, yden=yden, out=rows_to_nchw(out_rows, z.shape),
                q_nchw=rows_to_nchw(q, z.shape), vq=mse, commit=beta * mse,
                hist=np.bincount(idx, minlength=E.shape[0]).astype(np.float32),
                margin=np.abs(second - best))


def vq_backward(fwd: Dict, E: np.ndarray, g_out: np.ndarray, g_vq: float, g_commit: float,
                beta: float = 0.25, codebook_norm: bool = True):
    """Gradients of (out, vq, commit) wrt z (NCHW) and E -- SURVEY.md Appendix A.3."""
    zn, q, idx = fwd["zn"], fwd["q"], fwd["idx"]
    n = zn.size
    g_rows = nchw_to_rows(g_out).astype(np.float64)
    diff = q.astype(np.float64) - zn.astype(np.float64)
    g_q = g_vq * 2.0 * diff / n
    g_zn = g_rows - g_commit * beta * 2.0 * diff / n
    if codebook_norm:
        g_y = _norm_jvp_T(g_q, q, fwd["yden"])
        g_zrows = _norm_jvp_T(g_zn, zn, fwd["zden"])
    else:
        g_y, g_zrows = g_q, g_zn
    gE = np.zeros(E.shape, np.float64)
    np.add.at(gE, idx, g_y)
    shape = fwd["out"].shape
    gz = rows_to_nchw(g_zrows.astype(np.float32), shape).astype(np.float64)
    # keep fp64 precision for gz
    B, C = shape[:2]
    gz = g_zrows.reshape(B, -1, C).transpose(0, 2, 1).reshape(shape)
    return gz, gE


# ----------------------------------------------------------------------------------------
# add_perturbation                          latent_perturbation.py:4-35
# ----------------------------------------------------------------------------------------
def add_perturbation(z, z_q, E, codebook_norm, alpha, beta, delta, rand_u, rand_j) -> Dict:
    """rand_u ~ torch.rand(N), rand_j ~ torch.randint(0, delta, (N,)) are INPUTS (the reference
    draws them at latent_perturbation.py:21-22)."""
    z, z_q, E = _c32(z), _c32(z_q), _c32(E)
    B, C = z.shape[:2]
    rows = nchw_to_rows(z)
    if codebook_norm:
        zn, zden = l2norm_rows(rows)
        En, _ = l2norm_rows(E)
    else:
        zn, zden, En = rows, np.ones(rows.shape[0], np.float32), E
    rank = np.where(np.asarray(rand_u) > alpha, 0, np.asarray(rand_j)).astype(np.int64)
    sel = rank_select(zn, En, rank, delta)
    y = E[sel]
    q = l2norm_rows(y)[0] if codebook_norm else y
    p_rows = zn + (q - zn)
    p = rows_to_nchw(p_rows, z.shape)
    nb = int(B * beta)
    out = z_q.copy()
    out[:nb] = p[:nb]
    return dict(out=out, sel=sel, rank=rank, nb=nb, zn=zn, zden=zden)


def add_perturbation_backward(fwd: Dict, g: np.ndarray, codebook_norm: bool):
    """Returns (g_z, g_zq): the perturbed samples pass the gradient to z through the
    normalisation Jacobian (straight-through), the rest to z_q untouched."""
    nb = fwd["nb"]
    g = np.asarray(g, np.float64)
    B, C = g.shape[:2]
    g_zq = g.copy()
    g_zq[:nb] = 0.0
    gm = g.copy()
    gm[nb:] = 0.0
    g_rows = gm.reshape(B, C, -1).transpose(0, 2, 1).reshape(-1, C)
    if codebook_norm:
        g_rows = _norm_jvp_T(g_rows, fwd["zn"], fwd["zden"])
    g_z = g_rows.reshape(B, -1, C).transpose(0, 2, 1).reshape(g.shape)
    return g_z, g_zq


# ----------------------------------------------------------------------------------------
# multi-scale residual quantizers           quant.py:64-223, lookup_free_quantize.py:149-380
# ----------------------------------------------------------------------------------------
def multiscale(f, E, mode, patch_nums, scaler, phi_w, phi_b, pmap, r):
    f = _c32(f)
    B, C, H, W = f.shape
    SN = len(patch_nums)
    pn = np.asarray(patch_nums, np.int32)
    tot = int(sum(B * p * p for p in patch_nums))
    idx_all = np.empty(tot, np.int64)
    margin = np.empty(tot, np.float32)
    u_all = np.empty((SN, B, C, H, W), np.float32)
    h_all = np.empty((SN, B, C, H, W), np.float32)
    f_rest = np.empty_like(f)
    Ec = _c32(E) if E is not None else None
    V = 0 if Ec is None else Ec.shape[0]
    sc = _c32(scaler) if scaler is not None else None
    pw = _c32(phi_w) if phi_w is not None else None
    pb = _c32(phi_b) if phi_b is not None else None
    pm = np.asarray(pmap, np.int32)
    lib().xqo_multiscale(_f(f), B, C, H, W, _f(Ec), V, int(mode), _i(pn), SN, _f(sc), _f(pw), _f(pb), _i(pm),
                         ctypes.c_float(r), _l(idx_all), _f(u_all), _f(h_all), _f(f_rest), _f(margin))
    idx_list, mar_list, off = [], [], 0
    for p in patch_nums:
        n = B * p * p
        idx_list.append(idx_all[off:off + n].reshape(B, p * p))
        mar_list.append(margin[off:off + n].reshape(B, p * p))
        off += n
    return idx_list, u_all, h_all, f_rest, mar_list


def n_quantizers_from_dropout(B: int, SN: int, codebook_drop: float, dropout: Optional[np.ndarray]) -> np.ndarray:
    """quant.py:79-86 / lookup_free_quantize.py:167-174."""
    nq = np.full(B, SN + 1, np.float64)
    if dropout is not None:
        nd = int(B * codebook_drop)
        nq[:nd] = np.asarray(dropout)[:nd]
    return nq


def _ms_losses(f_ref64, h_all, nq, beta):
    """masked cumulative F_si, vq/commit partial sums (Appendix A.2)."""
    SN, B = h_all.shape[:2]
    n = f_ref64.size
    F = np.zeros(h_all.shape[1:], np.float32)
    F_list, masks, ratios = [], [], []
    vq = 0.0
    commit = 0.0
    for si in range(SN):
        m = (si < nq).astype(np.float32)
        F = F + h_all[si] * m[:, None, None, None]
        ratio = float(m.sum()) / B
        d2 = (F.astype(np.float64) - f_ref64) ** 2 * m[:, None, None, None]
        vq += d2.sum() / n / ratio
        commit += beta / ratio * d2.sum() / n
        F_list.append(F.copy())
        masks.append(m)
        ratios.append(ratio)
    return F_list, masks, ratios, vq, commit


def vq2_forward(f, E, phi_w, phi_b, patch_nums, using_znorm=True, beta=0.25, resi_ratio=0.5,
                codebook_drop=0.0, dropout=None, training=True) -> Dict:
    """VectorQuantizer2.forward (quant.py:64-144)."""
    f = _c32(f)
    B, C, H, W = f.shape
    SN = len(patch_nums)
    K = 0 if phi_w is None else phi_w.shape[0]
    pmap = phi_map(SN, K) if K else [-1] * SN
    idx_list, u_all, h_all, f_rest, margins = multiscale(f, E, 0 if using_znorm else 1, patch_nums, None,
                                                        phi_w, phi_b, pmap, resi_ratio)
    nq = n_quantizers_from_dropout(B, SN, codebook_drop, dropout if training else None)
    f64 = f.astype(np.float64)
    F_list, masks, ratios, vq, commit = _ms_losses(f64, h_all, nq, beta)
    vq *= 1.0 / SN  # quant.py:134 (commit is NOT divided)
    out = (F_list[-1] - f) + f  # quant.py:135
    hist = np.stack([np.bincount(ix.reshape(-1), minlength=E.shape[0]).astype(np.float32) for ix in idx_list])
    return dict(out=out, idx=idx_list, vq=vq, commit=commit, hist=hist, F=F_list, masks=masks, ratios=ratios,
                u=u_all, h=h_all, pmap=pmap, margins=margins, f_rest=f_rest)


def _phi_backward(dh, u, w, r):
    """h = (1-r) u + r (conv3x3(u; w) + b).  Returns du, dw, db (float64)."""
    Bn, C, H, W = dh.shape
    dh64 = dh.astype(np.float64)
    u64 = u.astype(np.float64)
    w64 = w.astype(np.float64)
    du = (1.0 - r) * dh64
    dw = np.zeros_like(w64)
    up = np.pad(u64, ((0, 0), (0, 0), (1, 1), (1, 1)))
    dhp = np.pad(dh64, ((0, 0), (0, 0), (1, 1), (1, 1)))
    for ky in range(3):
        for kx in range(3):
            # forward: out[co,y,x] += w[co,ci,ky,kx] * u[ci, y+ky-1, x+kx-1]
            us = up[:, :, ky:ky + H, kx:kx + W]
            dw[:, :, ky, kx] = r * np.einsum("bohw,bihw->oi", dh64, us)
            # du[ci, y', x'] += w[co,ci,ky,kx] * dh[co, y'-ky+1, x'-kx+1]
            dhs = dhp[:, :, 2 - ky:2 - ky + H, 2 - kx:2 - kx + W]
            du += r * np.einsum("oi,bohw->bihw", w64[:, :, ky, kx], dhs)
    db = r * dh64.sum(axis=(0, 2, 3))
    return du, dw, db


def vq2_backward(fwd: Dict, f, E, phi_w, phi_b, patch_nums, g_out, g_vq, g_commit, beta=0.25, resi_ratio=0.5):
    """Closed-form gradients (Appendix A.2) wrt f, E, phi_w, phi_b."""
    f64 = np.asarray(f, np.float64)
    B, C, H, W = f64.shape
    SN = len(patch_nums)
    n = f64.size
    gf = np.asarray(g_out, np.float64).copy()
    gE = np.zeros(np.asarray(E).shape, np.float64)
    gw = np.zeros(np.asarray(phi_w).shape, np.float64) if phi_w is not None else None
    gb = np.zeros(np.asarray(phi_b).shape, np.float64) if phi_b is not None else None
    S = np.zeros_like(f64)
    for k in range(SN - 1, -1, -1):
        m = fwd["masks"][k].astype(np.float64)[:, None, None, None]
        D = (fwd["F"][k].astype(np.float64) - f64) * m
        S = S + (2.0 / (SN * n * fwd["ratios"][k])) * D * g_vq
        gf += g_commit * (-2.0 * beta / (n * fwd["ratios"][k])) * D
        dh = S * m
        kk = fwd["pmap"][k]
        if kk >= 0:
            du, dw, db = _phi_backward(dh, fwd["u"][k], np.asarray(phi_w)[kk], resi_ratio)
            gw[kk] += dw
            gb[kk] += db
        else:
            du = dh
        P = patch_nums[k]
        My, Mx = cubic_matrix(P, H), cubic_matrix(P, W)
        dg = np.einsum("yp,bcyx,xq->bpqc", My, du, Mx)  # [B,P,P,C]
        np.add.at(gE, fwd["idx"][k].reshape(-1), dg.reshape(-1, C))
    return gf, gE, gw, gb


def _h2(p):
    return -p * np.log(p + 1e-8) - (1.0 - p) * np.log(1.0 - p + 1e-8)


def _dh2(p):
    return -np.log(p + 1e-8) - p / (p + 1e-8) + np.log(1.0 - p + 1e-8) + (1.0 - p) / (1.0 - p + 1e-8)


def lfq_scaler(SN: int, C: int, scale: float, using_znorm: bool) -> np.ndarray:
    """lookup_free_quantize.py:124-128, in float32 like torch (scale ** arange, / sqrt(C))."""
    scaler = np.power(np.float32(scale), np.arange(SN, dtype=np.float32)).astype(np.float32)
    if using_znorm:
        scaler = (scaler / np.float32(np.sqrt(C))).astype(np.float32)
    return scaler


def lfq_forward(f, phi_w, phi_b, patch_nums, using_znorm=False, beta=0.25, resi_ratio=0.5, codebook_drop=0.0,
                dropout=None, scale=1.0, entropy_weight=0.1, w_sample=1.0, w_batch=1.0, scaler=None) -> Dict:
    """LFQ.forward (lookup_free_quantize.py:149-250), training mode, soft_entropy=True.
    `scaler` (the module's registered buffer) overrides the value derived from `scale`."""
    f = _c32(f)
    B, C, H, W = f.shape
    SN = len(patch_nums)
    scaler = lfq_scaler(SN, C, scale, using_znorm) if scaler is None else np.asarray(scaler, np.float32)
    if using_znorm:  # F.normalize(f, dim=1)  :153
        rows = nchw_to_rows(f)
        fn_rows, fden = l2norm_rows(rows)
        fn = rows_to_nchw(fn_rows, f.shape)
    else:
        fn, fn_rows, fden = f, None, None
    K = 0 if phi_w is None else phi_w.shape[0]
    pmap = phi_map(SN, K) if K else [-1] * SN
    idx_list, u_all, h_all, f_rest, _ = multiscale(fn, None, 2, patch_nums, scaler, phi_w, phi_b, pmap, resi_ratio)
    nq = n_quantizers_from_dropout(B, SN, codebook_drop, dropout)
    f64 = fn.astype(np.float64)
    F_list, masks, ratios, vq, commit = _ms_losses(f64, h_all, nq, beta)
    # entropy term (:197, :218-235, :283-300): x = fn - F_{si-1}; int-mask gather of batch rows 0/1
    ent = 0.0
    ent_parts = []
    HW = H * W
    for si in range(SN):
        Fprev = np.zeros_like(f64) if si == 0 else F_list[si - 1].astype(np.float64)
        x = f64 - Fprev  # [B,C,H,W]
        m = masks[si]
        n1 = float(m.sum())
        n0 = B - n1
        s = float(scaler[si])
        p0 = 1.0 / (1.0 + np.exp(4.0 * x[0] * s))  # sigmoid(-4 x s)   [C,H,W]
        p1 = 1.0 / (1.0 + np.exp(4.0 * x[1] * s)) if B > 1 else p0
        Hs = (n0 * _h2(p0).sum() + n1 * _h2(p1).sum()) / (B * HW)
        pbar = (n0 * p0.sum(axis=(1, 2)) + n1 * p1.sum(axis=(1, 2))) / (B * HW)  # [C]
        qbar = (n0 * (1.0 - p0).sum(axis=(1, 2)) + n1 * (1.0 - p1).sum(axis=(1, 2))) / (B * HW)
        Hc = float((-pbar * np.log(pbar + 1e-8) - qbar * np.log(qbar + 1e-8)).sum())
        aux = w_sample * Hs - w_batch * Hc
        ent += aux * entropy_weight / ratios[si]
        ent_parts.append(dict(p0=p0, p1=p1, pbar=pbar, qbar=qbar, n0=n0, n1=n1, s=s))
    vq *= 1.0 / SN
    commit *= 1.0 / SN
    ent *= 1.0 / SN
    out = (F_list[-1] - fn) + fn
    V = 2 ** C
    hist = np.stack([np.bincount(ix.reshape(-1), minlength=V).astype(np.float32) for ix in idx_list])
    return dict(out=out, idx=idx_list, vq=vq, commit=commit, entropy=ent, hist=hist, F=F_list, masks=masks,
                ratios=ratios, u=u_all, h=h_all, pmap=pmap, fn=fn, fn_rows=fn_rows, fden=fden, scaler=scaler,
                ent_parts=ent_parts)


def lfq_backward(fwd: Dict, f, phi_w, phi_b, patch_nums, g_out, g_vq, g_commit, g_ent, using_znorm=False,
                 beta=0.25, resi_ratio=0.5, entropy_weight=0.1, w_sample=1.0, w_batch=1.0):
    """Closed-form gradients (Appendix A.5) wrt f, phi_w, phi_b."""
    fn64 = fwd["fn"].astype(np.float64)
    B, C, H, W = fn64.shape
    SN = len(patch_nums)
    n = fn64.size
    HW = H * W
    gfn = np.asarray(g_out, np.float64).copy()
    gw = np.zeros(np.asarray(phi_w).shape, np.float64) if phi_w is not None else None
    gb = np.zeros(np.asarray(phi_b).shape, np.float64) if phi_b is not None else None
    S = np.zeros_like(fn64)
    for k in range(SN - 1, -1, -1):
        m = fwd["masks"][k].astype(np.float64)[:, None, None, None]
        D = (fwd["F"][k].astype(np.float64) - fn64) * m
        S = S + (2.0 / (SN * n * fwd["ratios"][k])) * D * g_vq
        gfn += g_commit * (-2.0 * beta / (SN * n * fwd["ratios"][k])) * D
        dh = S * m
        kk = fwd["pmap"][k]
        if kk >= 0:
            _, dw, db = _phi_backward(dh, fwd["u"][k], np.asarray(phi_w)[kk], resi_ratio)
            gw[kk] += dw
            gb[kk] += db
        # entropy gradient to fn[0], fn[1]
        ep = fwd["ent_parts"][k]
        coef = g_ent * entropy_weight / fwd["ratios"][k] / SN
        dHc_dp = -np.log(ep["pbar"] + 1e-8) - ep["pbar"] / (ep["pbar"] + 1e-8)      # d/dpbar of -pbar ln(pbar+eps)
        dHc_dq = -np.log(ep["qbar"] + 1e-8) - ep["qbar"] / (ep["qbar"] + 1e-8)
        for img, cnt, p in ((0, ep["n0"], ep["p0"]), (1, ep["n1"], ep["p1"])):
            if img >= B or cnt == 0:
                continue
            wgt = cnt / (B * HW)
            # sample entropy: d/dp [h2(p)]
            dp = w_sample * wgt * _dh2(p)
            # codebook entropy: pbar depends on p with weight wgt ; qbar on (1-p)
            dp -= w_batch * wgt * (dHc_dp[:, None, None] - dHc_dq[:, None, None])
            dx = dp * p * (1.0 - p) * (-4.0 * ep["s"])
            gfn[img] += coef * dx
    if using_znorm:
        g_rows = gfn.reshape(B, C, -1).transpose(0, 2, 1).reshape(-1, C)
        g_rows = _norm_jvp_T(g_rows, fwd["fn_rows"], fwd["fden"])
        gf = g_rows.reshape(B, -1, C).transpose(0, 2, 1).reshape(fn64.shape)
    else:
        gf = gfn
    return gf, gw, gb


# ----------------------------------------------------------------------------------------
# inference helpers
# ----------------------------------------------------------------------------------------
def vq2_f_to_idxBl_or_fhat(f, E, phi_w, phi_b, patch_nums, using_znorm=True, resi_ratio=0.5, to_fhat=False):
    """quant.py:182-223 (no masks, no losses)."""
    f = _c32(f)
    SN = len(patch_nums)
    K = 0 if phi_w is None else phi_w.shape[0]
    pmap = phi_map(SN, K) if K else [-1] * SN
    idx_list, u_all, h_all, _, _ = multiscale(f, E, 0 if using_znorm else 1, patch_nums, None, phi_w, phi_b, pmap,
                                              resi_ratio)
    if not to_fhat:
        return idx_list
    out, F = [], np.zeros_like(f)
    for si in range(SN):
        F = F + h_all[si]
        out.append(F.copy())
    return out


def _phi_or_id(u, phi_w, phi_b, k, r):
    return u if (phi_w is None or k < 0) else phi(u, phi_w[k], phi_b[k], r)


def embed_to_fhat(ms_h, phi_w, phi_b, patch_nums, resi_ratio=0.5, last_one=False):
    """VectorQuantizer2.embed_to_fhat(all_to_max_scale=True) quant.py:148-166 (LFQ: lookup_free_quantize.py:311-328):
    per-scale feature maps [B,C,pn,pn] -> cumulative f_hat after every scale (or only the last)."""
    SN = len(patch_nums)
    H = W = patch_nums[-1]
    K = 0 if phi_w is None else phi_w.shape[0]
    pmap = phi_map(SN, K) if K else [-1] * SN
    B, C = ms_h[0].shape[:2]
    F = np.zeros((B, C, H, W), np.float32)
    out = []
    for si, pn in enumerate(patch_nums):
        h = _c32(ms_h[si])
        u = bicubic_up(nchw_to_rows(h), B, C, pn, H, W) if si < SN - 1 else h      # :157-158
        F = (F + _phi_or_id(u, phi_w, phi_b, pmap[si], resi_ratio)).astype(np.float32)   # :159-160
        out.append(F.copy())
    return out[-1] if last_one else out


def get_next_autoregressive_input(si, f_hat, h, phi_w, phi_b, patch_nums, resi_ratio=0.5):
    """VectorQuantizer2.get_next_autoregressive_input quant.py:247-258 (LFQ: lookup_free_quantize.py:404-415).
    -> (new f_hat, area-pooled f_hat at scale si+1 as [B,C,pn,pn])  /  (f_hat, f_hat) at the last scale."""
    SN = len(patch_nums)
    H = W = patch_nums[-1]
    K = 0 if phi_w is None else phi_w.shape[0]
    pmap = phi_map(SN, K) if K else [-1] * SN
    h, f_hat = _c32(h), _c32(f_hat)
    B, C = h.shape[:2]
    if si != SN - 1:
        u = bicubic_up(nchw_to_rows(h), B, C, h.shape[2], H, W)                    # :251-252
        F = (f_hat + _phi_or_id(u, phi_w, phi_b, pmap[si], resi_ratio)).astype(np.float32)
        pn = patch_nums[si + 1]
        return F, rows_to_nchw(area_pool_rows(F, pn), (B, C, pn, pn))             # :254
    F = (f_hat + _phi_or_id(h, phi_w, phi_b, pmap[si], resi_ratio)).astype(np.float32)   # :256-258
    return F, F


def ema_update(ema_row: np.ndarray, hit: np.ndarray, record_hit: int) -> np.ndarray:
    """quant.py:121-126 / xqgan_model.py:777-782."""
    if record_hit == 0:
        return hit.astype(np.float32).copy()
    if record_hit < 100:
        return (ema_row * np.float32(0.9) + hit * np.float32(0.1)).astype(np.float32)
    return (ema_row * np.float32(0.99) + hit * np.float32(0.01)).astype(np.float32)
