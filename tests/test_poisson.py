"""Poisson blending: oracle vs the reference goldens and the kernels' host-executable scalar
recurrence on the CPU; CUDA batched LSQR vs oracle and goldens on the GPU.

Tolerances: the reference runs scipy's LSQR on float32 operands (float32 bidiagonalisation), the oracle and the
CUDA path run the same recurrences and stopping rule in float64: oracle vs reference 5e-6 absolute on [0,1]
images (measured 3e-7..6e-7), CUDA vs oracle 1e-6 (summation order only; a stop one iteration apart would show
up as ~1e-5 and is rejected), UnfilledMask bit-exact."""
import ctypes

import numpy as np
import pytest
import torch

from fgt_b200 import lib, synth
from oracle import poisson_oracle as PO
from tests.util import load_golden

CASES = {"poisson_small": (3, 64, 96, 7, False), "poisson_edge": (2, 48, 64, 8, True),
         "poisson_mid": (2, 120, 216, 9, False)}


def _inputs(name):
    F, H, W, seed, with_edge = CASES[name]
    g = load_golden(name)
    assert (g["meta"]["F"], g["meta"]["H"], g["meta"]["W"], g["meta"]["seed"]) == (F, H, W, seed)
    inp = synth.poisson_inputs(seed=seed, F=F, H=H, W=W, with_edge=with_edge)
    edge = inp[5] if with_edge else None
    hole = inp[3]
    unf = np.unpackbits(g["unfilled"])[:hole.size].reshape(hole.shape).astype(bool)
    return inp[0], inp[1], inp[2], hole, inp[4], edge, g["blend_hole"], unf


def _oracle_clip(trg, gx, gy, hole, gm, edge):
    blends, unfs, infos = [], [], []
    for f in range(trg.shape[0]):
        b, u, info = PO.poisson_blend(trg[f], gx[f], gy[f], hole[f], gm[f], None if edge is None else edge[f],
                                      return_info=True)
        blends.append(b); unfs.append(u); infos.append(info)
    return np.stack(blends), np.stack(unfs), infos


@pytest.mark.parametrize("name", ["poisson_small", "poisson_edge"])
def test_oracle_matches_reference_golden(name):
    trg, gx, gy, hole, gm, edge, gold, gunf = _inputs(name)
    blend, unf, infos = _oracle_clip(trg, gx, gy, hole, gm, edge)
    assert np.array_equal(unf, gunf)
    assert np.abs(blend[hole] - gold).max() < 5e-6
    assert np.array_equal(blend[~hole], trg.astype(np.float64)[~hole])
    assert all(istop in (1, 2) and itn > 10 for info in infos[:-1] for istop, itn in info)
    if name == "poisson_small":                          # last frame: empty hole -> "x = 0" exit, frame unchanged
        assert infos[-1] == [(0, 0)] * 3 and not unf[-1].any()


def test_oracle_operator_adjoint_and_duplicate_equations():
    """<A v, u> == <v, A^T u>, and every in-hole edge carries the same equation from both ends with opposite sign
    (what lets the kernels form A^T u from a pixel's own code bits)."""
    trg, gx, gy, hole, gm, edge = synth.poisson_inputs(seed=3, F=1, H=40, W=56, with_edge=True)
    code = PO.equation_codes(hole[0], gm[0], edge[0])
    rng = np.random.default_rng(0)
    v = rng.standard_normal(hole[0].shape) * hole[0]
    u = rng.standard_normal((4,) + hole[0].shape) * np.stack([(code >> n) & 1 for n in range(4)])
    assert abs((PO.matvec(code, v) * u).sum() - (v * PO.rmatvec(code, u)).sum()) < 1e-9
    b = PO.rhs(code, trg[0], gx[0].astype(np.float64), gy[0].astype(np.float64))
    for n, m in ((0, 2), (1, 3)):
        inh = ((code >> (4 + n)) & 1).astype(bool)
        ys, xs = np.nonzero(inh)
        qy, qx = ys + PO.DY[n], xs + PO.DX[n]
        assert ((code[qy, qx] >> (4 + m)) & 1).all() and hole[0][qy, qx].all()
        assert np.array_equal(b[n, ys, xs], -b[m, qy, qx])


def test_oracle_lsqr_equals_scipy_lsqr_on_the_assembled_system():
    """Independent check of the restated solver: assemble the sparse system explicitly from the equation codes and run
    the real scipy.sparse.linalg.lsqr (float64, default tolerances) on it — same stopping iteration, same istop, same
    iterate as the matrix-free restatement (random ragged holes, gradient masks and edges)."""
    from scipy import sparse
    from scipy.sparse.linalg import lsqr as scipy_lsqr
    rng = np.random.default_rng(7)
    H, W = 26, 34
    for trial in range(4):
        hole = rng.random((H, W)) < (0.25, 0.5, 0.7, 0.4)[trial]
        hole[8:18, 10:24] = True
        gm = (rng.random((H, W)) < 0.15) & hole
        edge = (rng.random((H, W)) < 0.05).astype(np.float32) if trial % 2 else np.zeros((H, W), np.float32)
        trg = rng.random((H, W, 3)).astype(np.float32)
        trg[hole] = 0
        gx = rng.standard_normal((H, W - 1, 3)).astype(np.float32) * 0.1
        gy = rng.standard_normal((H - 1, W, 3)).astype(np.float32) * 0.1
        code = PO.equation_codes(hole, gm, edge)
        b = PO.rhs(code, trg, gx.astype(np.float64), gy.astype(np.float64))
        rows, cols, vals, rhs_v = [], [], [], []
        for n in range(4):
            ys, xs = np.nonzero((code >> n) & 1)
            for y, x in zip(ys, xs):
                r = len(rhs_v)
                rows.append(r); cols.append(y * W + x); vals.append(1.0)
                if (code[y, x] >> (4 + n)) & 1:
                    rows.append(r); cols.append((y + PO.DY[n]) * W + x + PO.DX[n]); vals.append(-1.0)
                rhs_v.append(b[n, y, x, 0])
        A = sparse.csr_matrix((vals, (rows, cols)), shape=(len(rhs_v), H * W))
        ref = scipy_lsqr(A, np.asarray(rhs_v))
        x, istop, itn = PO.lsqr(code, b[..., 0])
        assert (istop, itn) == (ref[1], ref[2]), (trial, istop, itn, ref[1], ref[2])
        assert np.abs(x.reshape(-1) - ref[0]).max() < 1e-9


def _advance(L, prev, k, bbk, aak, wwk, iter_lim):
    cur = np.zeros(16)
    step = np.zeros(6)
    p = prev.ctypes.data_as(ctypes.c_void_p) if prev is not None else None
    rc = L.fgt_poisson_advance_host(p, cur.ctypes.data_as(ctypes.c_void_p), step.ctypes.data_as(ctypes.c_void_p), k,
                                    bbk, aak, wwk, 1e-6, 1e-6, 1e8, iter_lim)
    assert rc == 0
    return cur, step


def _emulate_kernels(L, code, b, max_k=5000):
    """numpy statement of psn_v_kernel / psn_ux_kernel for one system: unnormalised u^, v^ with scales applied on
    read, per-iteration sums, and the library's own scalar recurrence (psn_advance run on the host)."""
    H, W = code.shape
    u, v = b.copy(), np.zeros((H, W))
    w, x = np.zeros((H, W)), np.zeros((H, W))
    bb, aa, ww = {0: float((b * b).sum())}, {}, {0: 0.0}
    state = None
    for k in range(max_k):
        # psn_v_kernel(k)
        done = state is not None and state[12] != 0
        beta = np.sqrt(bb[k])
        aa[k] = 0.0
        if not done and beta > 0:
            coef = beta * state[9] if k >= 1 else 0.0
            v = PO.rmatvec(code, u) * (1.0 / beta) - coef * v
            aa[k] = float((v * v).sum())
        # psn_ux_kernel(k)
        state, step = _advance(L, state, k, bb[k], aa[k], ww[k], 2 * H * W)
        t1, t2, vscale, au, skip_all, skip_u = step
        ww[k + 1] = bb[k + 1] = 0.0
        if not skip_all:
            x = x + t1 * w
            w = v * vscale + t2 * w
            ww[k + 1] = float((w * w).sum())
        if not skip_u:
            u = PO.matvec(code, v * vscale) - au * u
            bb[k + 1] = float((u * u).sum())
        if state[12] != 0:
            return x, int(state[13]), int(state[14])
    raise AssertionError("emulation did not stop")


def test_kernel_scalar_recurrence_matches_oracle_lsqr():
    """The kernels' formulation (same code path for the scalars: fgt_poisson_advance_host) stops at the same
    iteration with the same istop as the oracle's LSQR and gives the same iterate."""
    L = lib.load()
    trg, gx, gy, hole, gm, edge = synth.poisson_inputs(seed=11, F=3, H=40, W=56, with_edge=True)
    for f, ed in ((0, None), (1, edge[1]), (2, None)):       # frame 2: empty hole -> immediate exit
        code = PO.equation_codes(hole[f], gm[f], np.zeros_like(hole[f]) if ed is None else ed)
        b = PO.rhs(code, trg[f], gx[f].astype(np.float64), gy[f].astype(np.float64))
        for c in (0, 2):
            xo, istop, itn = PO.lsqr(code, b[..., c])
            xe, istop_e, itn_e = _emulate_kernels(L, code, b[..., c])
            assert (istop_e, itn_e) == (istop, itn)
            assert np.abs(xe - xo).max() < 1e-11


def _sweep_words(hole, gm, back):
    """numpy statement of psn_sweep_kernel: 32-column ballot words, Kogge-Stone prefix, carry between words."""
    H, W = hole.shape
    out = np.zeros((H, W), dtype=bool)
    M = 0xFFFFFFFF
    for r in range(H):
        y = H - 1 - r if back else r
        yp = y + 1 if back else y - 1
        carry = 0
        for g0 in range(0, W, 32):
            g = p = 0
            for lane in range(32):
                jj = g0 + lane
                if jj >= W:
                    continue
                xc = W - 1 - jj if back else jj
                G = not hole[y, xc]
                if not G and r > 0:
                    G = bool(out[yp, xc]) and not gm[(y, xc) if back else (yp, xc)]
                P = jj > 0 and not gm[(y, xc) if back else (y, xc - 1)]
                g |= int(G) << lane
                p |= int(P) << lane
            g |= p & carry
            d = 1
            while d < 32:
                g |= p & ((g << d) & M)
                p &= (p << d) & M
                d <<= 1
            for lane in range(32):
                jj = g0 + lane
                if jj < W:
                    out[y, W - 1 - jj if back else jj] = (g >> lane) & 1
            carry = g >> 31
    return out


def test_sweep_prefix_formulation_matches_oracle():
    _, _, _, hole, gm = synth.poisson_inputs(seed=5, F=2, H=37, W=70)
    rng = np.random.default_rng(1)
    for f in range(2):
        h = hole[f] | (rng.random(hole[f].shape) < 0.3)          # ragged holes, W not a multiple of 32
        m = (gm[f] | (rng.random(hole[f].shape) < 0.2)) & h
        unf = h & ~_sweep_words(h, m, False) & ~_sweep_words(h, m, True)
        assert np.array_equal(unf, PO.unfilled_mask(h, m))


def _pipeline_case():
    """Driver-faithful inputs: the seven Poisson_blend_img calls of one run of the unmodified reference driver
    (tests/golden/make_pipeline_golden.py) — gradients after real flow-guided propagation, gradient masks after
    binary_fill_holes."""
    g = load_golden("pipeline_poisson")
    trg = g["trg"]
    F, H, W = trg.shape[:3]
    bits = lambda k: np.unpackbits(g[k])[:F * H * W].reshape(F, H, W).astype(bool)
    return trg, g["gx"], g["gy"], bits("hole"), bits("gmask"), g["blend_hole"], bits("unfilled")


def test_oracle_matches_reference_driver_poisson_calls():
    trg, gx, gy, hole, gm, gold, gunf = _pipeline_case()
    blend, unf, _ = _oracle_clip(trg, gx, gy, hole, gm, None)
    assert np.array_equal(unf, gunf) and gunf.sum() > 0
    assert np.abs(blend[hole] - gold).max() < 5e-6


@pytest.mark.gpu
def test_gpu_poisson_vs_reference_driver_calls():
    from fgt_b200 import poisson as P
    trg, gx, gy, hole, gm, gold, gunf = _pipeline_case()
    out, unf = P.poisson_blend_batch(trg, gx, gy, hole, gm)
    out, unf = out.cpu().numpy(), unf.cpu().numpy()
    assert np.array_equal(unf, gunf)
    assert np.abs(out[hole] - gold).max() < 5e-6
    assert np.array_equal(out[~hole], trg.astype(np.float64)[~hole])


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_poisson_vs_golden_and_oracle(name):
    from fgt_b200 import poisson as P
    trg, gx, gy, hole, gm, edge, gold, gunf = _inputs(name)
    out, unf, istop, itn = P.poisson_blend_batch(trg, gx, gy, hole, gm, edge, return_info=True)
    out, unf = out.cpu().numpy(), unf.cpu().numpy()
    assert out.dtype == np.float64 and unf.dtype == np.bool_
    assert np.array_equal(unf, gunf), "UnfilledMask vs reference golden"
    assert np.array_equal(out[~hole], trg.astype(np.float64)[~hole])          # untouched outside the holes
    assert np.abs(out[hole] - gold).max() < 5e-6, "vs reference golden"
    blend, ounf, infos = _oracle_clip(trg, gx, gy, hole, gm, edge)
    assert np.array_equal(unf, ounf)
    assert np.abs(out - blend).max() < 1e-6, "vs oracle"
    assert istop.cpu().tolist() == [[i for i, _ in info] for info in infos]
    assert itn.cpu().tolist() == [[n for _, n in info] for info in infos]


@pytest.mark.gpu
def test_gpu_poisson_api_forms():
    from fgt_b200 import poisson as P
    trg, gx, gy, hole, gm = synth.poisson_inputs(seed=12, F=3, H=40, W=56)
    one, unf = P.Poisson_blend_img(trg[1], gx[1], gy[1], hole[1], gm[1])
    ob, ou = PO.poisson_blend(trg[1], gx[1], gy[1], hole[1], gm[1])
    assert one.dtype == np.float64 and unf.dtype == np.bool_ and one.shape == (40, 56, 3)
    assert np.abs(one - ob).max() < 1e-6 and np.array_equal(unf, ou)
    nog, unf0 = P.Poisson_blend_img(trg[0], gx[0], gy[0], hole[0])            # gradientMask=None: nothing unfilled
    ob0, ou0 = PO.poisson_blend(trg[0], gx[0], gy[0], hole[0])
    assert np.abs(nog - ob0).max() < 1e-6 and not unf0.any() and not ou0.any()
    assert np.array_equal(P.getUnfilledMask(hole[1], gm[1]), PO.unfilled_mask(hole[1], gm[1]))
    # the driver's array layout, whole clip in one batch (frame 2 has no hole and comes back unchanged)
    H, W = hole.shape[1:]
    pad_x = np.concatenate([gx, np.zeros((3, H, 1, 3), np.float32)], 2)
    pad_y = np.concatenate([gy, np.zeros((3, 1, W, 3), np.float32)], 1)
    last = lambda a: np.ascontiguousarray(np.moveaxis(a, 0, -1))
    blend, unfs = P.poisson_blend_clip(last(trg), last(pad_x), last(pad_y), last(hole), last(gm))
    assert np.abs(blend[1] - ob).max() < 1e-6 and np.array_equal(unfs[1], ou)
    assert np.array_equal(blend[2], trg[2].astype(np.float64)) and not unfs[2].any()
    with pytest.raises(ValueError):
        P.poisson_blend_batch(trg, gx[:, :, :-1], gy, hole)
    with pytest.raises(RuntimeError):
        P.poisson_blend_batch(trg, gx, gy, hole, device="cpu")


@pytest.mark.gpu
def test_gpu_poisson_full_size_properties():
    """BASELINE size (240x432, 10 frames): size-independent properties instead of the (slow) oracle —
    every system stops by scipy's rule (istop 1 or 2), the result satisfies the normal equations to the LSQR
    tolerance (|A^T r| small relative to |A||r|, evaluated with the oracle's operator on one frame), pixels outside
    the holes are untouched, and a second run reproduces the first to summation-order noise."""
    from fgt_b200 import poisson as P
    trg, gx, gy, hole, gm = synth.poisson_inputs(seed=13, F=10, H=240, W=432)
    out, unf, istop, itn = P.poisson_blend_batch(trg, gx, gy, hole, gm, return_info=True)
    out2, unf2 = P.poisson_blend_batch(trg, gx, gy, hole, gm)
    assert torch.equal(unf, unf2) and (out - out2).abs().max().item() < 1e-6
    out = out.cpu().numpy()
    assert np.array_equal(out[~hole], trg.astype(np.float64)[~hole])
    live = istop[:-1]
    assert ((live == 1) | (live == 2)).all() and (itn[:-1] > 50).all() and (itn[:-1] < 5000).all()
    assert (istop[-1] == 0).all() and (itn[-1] == 0).all()
    f = 0
    code = PO.equation_codes(hole[f], gm[f], np.zeros_like(hole[f]))
    b = PO.rhs(code, trg[f], gx[f].astype(np.float64), gy[f].astype(np.float64))
    for c in range(3):
        x = out[f, :, :, c] * hole[f]
        r = b[..., c] - PO.matvec(code, x)
        atr = PO.rmatvec(code, r)
        assert np.linalg.norm(atr) <= 2e-5 * np.sqrt(8.0 * code.astype(bool).sum()) * np.linalg.norm(r)
    assert np.array_equal(unf[f].cpu().numpy(), PO.unfilled_mask(hole[f], gm[f]))
