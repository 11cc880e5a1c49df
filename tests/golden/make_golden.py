"""Generates the golden fixtures in this directory by running the UNMODIFIED reference modules
imported from /root/reference (only available in the build container, never on the GPU box).

    python tests/golden/make_golden.py

Weights and inputs come from fgt_b200.synth (seeded, order-independent), so tests can regenerate
the identical inputs anywhere and only the reference OUTPUTS are stored here. The reference has no
tests or golden vectors of its own; these files are the pin for oracle/ and, through
it, for the CUDA path.
"""
import importlib
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
# import order mirrors tool/video_inpainting.py:4-6 (FGT/ and LAFC/ both own a `models` package)
sys.path[:0] = [REF, os.path.join(REF, "FGT"), os.path.join(REF, "LAFC")]
sys.path.insert(0, ROOT)

from fgt_b200 import synth  # noqa: E402

VERSIONS = dict(torch=torch.__version__, numpy=np.__version__)


def fgt_case(name, H, W, t, regime, res, seed, sample=None):
    M = importlib.import_module("FGT.models.model")
    cfg = dict(synth.CFG_A)
    cfg["input_resolution"] = res
    sd = synth.make_state_dict(synth.fgt_param_shapes(cfg), seed=seed, regime=regime)
    torch.manual_seed(0)
    model = M.Model(cfg)
    model.load_state_dict(sd)
    fr, fl, mk = synth.fgt_inputs(seed=seed + 2, t=t, H=H, W=W)
    with torch.no_grad():
        out = model(fr, fl, mk)
    meta = dict(H=H, W=W, t=t, regime=regime, res=list(res), seed=seed, **VERSIONS)
    arrs = dict(meta=np.array(repr(meta)), l2=np.float64(out.double().norm().item()),
                mean=np.float64(out.double().mean().item()), std=np.float64(out.double().std().item()))
    if sample is None:
        arrs["out"] = out.numpy().astype(np.float32)
    else:
        g = torch.Generator().manual_seed(1234)
        idx = torch.randperm(out.numel(), generator=g)[:sample]
        arrs["idx"] = idx.numpy().astype(np.int64)
        arrs["val"] = out.reshape(-1)[idx].numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print(name, "out std", out.std().item(), "saved")


def module_cases():
    """Module-level goldens at BASELINE config 1 (single 432x240 frame, spatial MHSA) and friends."""
    AF = importlib.import_module("FGT.models.transformer_base.attention_flow")
    AB = importlib.import_module("FGT.models.transformer_base.attention_base")
    FF = importlib.import_module("FGT.models.transformer_base.ffn_base")
    shapes = synth.fgt_param_shapes(synth.CFG_A)
    sd = synth.make_state_dict(shapes, seed=7, regime="scaled")
    out = {}
    g = torch.Generator().manual_seed(99)
    x = torch.randn(2, 720, 512, generator=g)
    f = torch.randn(2, 720, 256, generator=g)
    # SWMHSA, constructor geometry 20x36 (config 1) ------------------------------------------
    pre = "net.first_s_transformer.attention."
    m = AF.SWMHSA_depthGlobalWindowConcatLN_qkFlow_reweightFlow([20, 36], 8, 4, 512, 256, 4, p=0)
    m.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
    with torch.no_grad():
        out["swmhsa"] = m(x, f, 2).numpy()
    # TMHSA --------------------------------------------------------------------------------------
    pre = "net.first_t_transformer.attention."
    m = AB.TMHSA([20, 36], 2, 512, 4, p=0)
    m.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
    with torch.no_grad():
        out["tmhsa"] = m(x, 2).numpy()
    # FusionFeedForward --------------------------------------------------------------------------
    pre = "net.first_t_transformer.ffn."
    t2t = {'kernel_size': (7, 7), 'stride': (3, 3), 'padding': (3, 3), 'output_size': (60, 108)}
    m = FF.FusionFeedForward(512, 40, 720, t2t, p=0)
    m.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
    with torch.no_grad():
        out["ffn"] = m(x).numpy()
    g2 = torch.Generator().manual_seed(1234)
    idx = torch.randperm(2 * 720 * 512, generator=g2)[:16384]
    arrs = {"idx": idx.numpy().astype(np.int64), "meta": np.array(repr(dict(seed=7, **VERSIONS)))}
    for k, v in out.items():
        arrs[k] = v.reshape(-1)[idx.numpy()].astype(np.float32)
        arrs[k + "_l2"] = np.float64(np.linalg.norm(v.astype(np.float64)))
    np.savez_compressed(os.path.join(HERE, "fgt_modules.npz"), **arrs)
    print("fgt_modules saved")


def lafc_case(name, H, W, regime, seed, sample=None):
    M = importlib.import_module("LAFC.models.lafc")
    sd = synth.make_state_dict(synth.lafc_param_shapes(), seed=seed, regime=regime)
    model = M.Model(synth.CFG_LAFC)
    model.load_state_dict(sd)
    fl, mk = synth.lafc_inputs(seed=seed + 1, H=H, W=W)
    with torch.no_grad():
        flow, edge = model(fl, mk)
    meta = dict(H=H, W=W, regime=regime, seed=seed, **VERSIONS)
    arrs = dict(meta=np.array(repr(meta)), flow_l2=np.float64(flow.double().norm().item()),
                edge_l2=np.float64(edge.double().norm().item()))
    if sample is None:
        arrs["flow"] = flow.numpy().astype(np.float32)
        arrs["edge"] = edge.numpy().astype(np.float32)
    else:
        g = torch.Generator().manual_seed(1234)
        fi = torch.randperm(flow.numel(), generator=g)[:sample]
        ei = torch.randperm(edge.numel(), generator=g)[:sample]
        arrs.update(flow_idx=fi.numpy(), flow_val=flow.reshape(-1)[fi].numpy(), edge_idx=ei.numpy(),
                    edge_val=edge.reshape(-1)[ei].numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print(name, "flow std", flow.std().item(), "saved")


def raft_case(name, H, W, iters, seed, sample=None):
    import argparse
    R = importlib.import_module("RAFT")
    model = R.RAFT(argparse.Namespace(small=False, mixed_precision=False, alternate_corr=False)).eval()
    sd = synth.raft_state_dict(seed=seed)
    model.load_state_dict(sd)
    im1, im2 = synth.raft_inputs(seed=seed + 1, H=H, W=W)
    with torch.no_grad():
        lo, up = model(im1, im2, iters=iters, test_mode=True)
    meta = dict(H=H, W=W, iters=iters, seed=seed, **VERSIONS)
    arrs = dict(meta=np.array(repr(meta)), lo_l2=np.float64(lo.double().norm().item()),
                up_l2=np.float64(up.double().norm().item()))
    if sample is None:
        arrs["lo"] = lo.numpy().astype(np.float32)
        arrs["up"] = up.numpy().astype(np.float32)
    else:
        g = torch.Generator().manual_seed(1234)
        ui = torch.randperm(up.numel(), generator=g)[:sample]
        arrs.update(lo=lo.numpy().astype(np.float32), up_idx=ui.numpy(), up_val=up.reshape(-1)[ui].numpy())
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print(name, "|flow|max", lo.abs().max().item(), "saved")


def raft_real_weights_check():
    """Pins the RAFT oracle with the REAL raft-things.pth (cannot travel to the GPU box, so assert-only)."""
    import argparse
    from oracle import raft_oracle as RO
    R = importlib.import_module("RAFT")
    model = R.RAFT(argparse.Namespace(small=False, mixed_precision=False, alternate_corr=False)).eval()
    sd = torch.load(os.path.join(REF, "LAFC/flowCheckPoint/raft-things.pth"), map_location="cpu")
    sd = {k[len("module."):]: v for k, v in sd.items()}
    model.load_state_dict(sd)
    im1, im2 = synth.raft_inputs(seed=1, H=128, W=192)
    with torch.no_grad():
        lo, up = model(im1, im2, iters=20, test_mode=True)
        olo, oup = RO.raft_forward(sd, im1, im2, iters=20)
    err = ((oup - up).norm() / up.norm()).item()
    assert err < 1e-6, err
    print("raft oracle vs reference with raft-things.pth: rel", err, "mean flow", up.mean(dim=(0, 2, 3)).tolist())


def prop_case(name, H, W, N, thres, seed):
    import argparse
    import contextlib
    import io
    sys.path.insert(0, os.path.join(REF, "tool"))
    from get_flowNN_gradient import get_flowNN_gradient
    gx, gy, mask, ff, fb = synth.prop_inputs(seed=seed, H=H, W=W, N=N)
    args = argparse.Namespace(Nonlocal=False, consistencyThres=thres, alpha=0.1)
    with contextlib.redirect_stdout(io.StringIO()):
        rx, ry, rm = get_flowNN_gradient(args, gx.copy(), gy.copy(), mask.copy(), mask.copy(), ff, fb, None, None)
    import cv2
    meta = dict(H=H, W=W, N=N, thres=thres, seed=seed, cv2=cv2.__version__, **VERSIONS)
    hole = np.repeat(mask[:, :, None, :], 3, axis=2)
    assert np.array_equal(rx[~hole], gx[~hole]) and np.array_equal(ry[~hole], gy[~hole])  # only holes change
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=np.array(repr(meta)), gx_hole=rx[hole],
                        gy_hole=ry[hole], tofill=np.packbits(rm))
    print(name, "holes", int(mask.sum()), "tofill", int(rm.sum()), "saved")


def regionfill_case(name, B, H, W, seed):
    sys.path.insert(0, os.path.join(REF, "tool"))
    from utils.region_fill import regionfill
    from oracle import regionfill_oracle as RO
    img, mask = synth.regionfill_inputs(seed=seed, B=B, H=H, W=W)
    ref = np.stack([regionfill(img[b], mask[b]) for b in range(B)])
    ora = np.stack([RO.regionfill(img[b], mask[b]) for b in range(B)])
    err = np.abs(ref - ora).max()
    assert err < 1e-9, err
    import scipy
    meta = dict(B=B, H=H, W=W, seed=seed, scipy=scipy.__version__, **VERSIONS)
    assert np.array_equal(ref[~mask], img.astype(np.float64)[~mask])  # only holes change
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=np.array(repr(meta)), out_hole=ref[mask])
    print(name, "holes", int(mask.sum()), "oracle vs reference max abs", err, "saved")


def poisson_case(name, F, H, W, seed, with_edge):
    sys.path.insert(0, os.path.join(REF, "tool"))
    from utils.Poisson_blend_img import Poisson_blend_img
    from oracle import poisson_oracle as PO
    inp = synth.poisson_inputs(seed=seed, F=F, H=H, W=W, with_edge=with_edge)
    trg, gx, gy, hole, gm = inp[:5]
    blends, unfs, err = [], [], 0.0
    for f in range(F):
        if not hole[f].any():                      # the driver skips frames without a hole (video_inpainting.py:647)
            blends.append(trg[f].astype(np.float64)); unfs.append(hole[f]); continue
        kw = dict(edge=inp[5][f]) if with_edge else {}
        rb, ru = Poisson_blend_img(trg[f], gx[f], gy[f], hole[f], gm[f], **kw)
        ob, ou = PO.poisson_blend(trg[f], gx[f], gy[f], hole[f], gm[f], kw.get("edge"))
        assert rb.dtype == np.float64 and ru.dtype == np.bool_ and np.array_equal(ru, ou)
        assert np.array_equal(rb[~hole[f]], trg[f].astype(np.float64)[~hole[f]])      # only holes change
        err = max(err, np.abs(rb - ob).max())
        blends.append(rb); unfs.append(ru)
    assert err < 5e-6, err
    import scipy
    blend, unf = np.stack(blends), np.stack(unfs)
    meta = dict(F=F, H=H, W=W, seed=seed, with_edge=with_edge, scipy=scipy.__version__, **VERSIONS)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), meta=np.array(repr(meta)), blend_hole=blend[hole],
                        unfilled=np.packbits(unf))
    print(name, "holes", int(hole.sum()), "unfilled", int(unf.sum()), "oracle vs reference max abs", err, "saved")


if __name__ == "__main__":
    if "--extra-only" in sys.argv:     # the driver's default working size (imgH=256, :829) on a 240x432 checkpoint; RAFT at 720p
        fgt_case("fgt_driver_256x432_t6", 256, 432, 6, "scaled", (240, 432), seed=6, sample=16384)
        raft_case("raft_720p_i20", 720, 1280, 20, seed=7, sample=8192)
        sys.exit(0)
    if "--flowwarp-only" in sys.argv:  # LAFC/models/utils/flow_warp.py (dead code in the reference)
        FW = importlib.import_module("LAFC.models.utils.flow_warp")
        from oracle import flow_warp_oracle as FO
        feat, flow = synth.flow_warp_inputs(seed=11)
        outs = {}
        for mode in ("forward", "backward"):
            ref = FW.flow_prop(feat, flow, mode)
            err = (FO.flow_prop(feat, flow, mode) - ref).abs().max().item()
            assert err < 1e-5, err
            outs[mode] = ref.numpy().astype(np.float32)
            print("flow_warp", mode, "oracle vs reference max abs", err)
        np.savez_compressed(os.path.join(HERE, "flow_warp.npz"), meta=np.array(repr(dict(seed=11, **VERSIONS))), **outs)
        sys.exit(0)
    if "--flo-only" in sys.argv:       # 3x5 flow written by the reference's writer (tests/test_io.py)
        sys.path.insert(0, os.path.join(REF, "RAFT"))
        from utils import frame_utils as FU
        from tests.test_io import _flow
        FU.writeFlow(os.path.join(HERE, "flo_ref.flo"), _flow())
        assert np.array_equal(FU.readFlow(os.path.join(HERE, "flo_ref.flo")), _flow())
        sys.exit(0)
    if "--poisson-only" in sys.argv:
        poisson_case("poisson_small", 3, 64, 96, seed=7, with_edge=False)
        poisson_case("poisson_edge", 2, 48, 64, seed=8, with_edge=True)
        poisson_case("poisson_mid", 2, 120, 216, seed=9, with_edge=False)
        sys.exit(0)
    if "--regionfill-only" in sys.argv:
        regionfill_case("regionfill_small", 4, 48, 64, seed=5)
        regionfill_case("regionfill_mid", 3, 120, 216, seed=6)
        sys.exit(0)
    if "--prop-only" in sys.argv:
        prop_case("prop_small", 64, 96, 6, 5.0, seed=2)
        prop_case("prop_thres1", 48, 64, 4, 1.0, seed=3)
        prop_case("prop_mid", 120, 160, 8, 5.0, seed=4)
        sys.exit(0)
    if "--raft-only" in sys.argv:
        raft_real_weights_check()
        raft_case("raft_small_i6", 128, 192, 6, seed=3)
        raft_case("raft_small_i20", 128, 192, 20, seed=3)
        raft_case("raft_full_i20", 480, 864, 20, seed=4, sample=8192)
        sys.exit(0)
    lafc_case("lafc_small_scaled", 64, 96, "scaled", seed=4)
    lafc_case("lafc_small_kaiming", 64, 96, "kaiming", seed=4)
    lafc_case("lafc_full", 240, 432, "scaled", seed=5, sample=8192)
    if "--lafc-only" in sys.argv:
        sys.exit(0)
    fgt_case("fgt_small_scaled", 64, 96, 3, "scaled", (64, 96), seed=1)
    fgt_case("fgt_small_default", 64, 96, 3, "default", (64, 96), seed=1)
    fgt_case("fgt_runtime_geo", 72, 100, 2, "scaled", (64, 96), seed=2)
    fgt_case("fgt_full_t10", 240, 432, 10, "scaled", (240, 432), seed=1, sample=16384)
    module_cases()
