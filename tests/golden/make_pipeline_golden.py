"""Runs the UNMODIFIED reference driver `tool/video_inpainting.py::video_inpainting(args)` end to end on the CPU
(object-removal mode) on a small synthetic clip and records what crosses two stage boundaries:

* every `Poisson_blend_img` call (inputs after real propagation + binary_fill_holes, outputs)  -> pipeline_poisson.npz
* the FGT stage (tool/video_inpainting.py:686-745): the three `np2tensor(..., near="t")` inputs (frameBlends,
  mask, completed forward flows) and the frames handed to the video writer                    -> pipeline_clip.npz

    python tests/golden/make_pipeline_golden.py

Only available in the build container (/root/reference). The driver imports three packages that are not in
this image (cvbase: .flo visualisation, imageio: video writer, skimage.feature.canny: unused in this mode); they
are stubbed, `imageio.mimwrite` being the hook that captures the final frames. All three networks use the
seeded synthetic weights of fgt_b200.synth (the FGT / LAFC checkpoints are not in the reference repository, and
the real raft-things.pth cannot travel to the GPU box), written to temporary checkpoint files in the layout
`initialize_RAFT/LAFC/FGT` expect. Also recorded for stage-wise diagnosis: RAFT flows, completed flows and the
propagation's mask                                                                                  -> pipeline_stages.npz
"""
import argparse
import contextlib
import io
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch
import yaml
from PIL import Image

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from fgt_b200 import synth  # noqa: E402

H, W, N = 64, 96, 7
CAPTURE = {}
# second case (--second): 12 frames -> three windows (f = 0, 5, 10), frames visited up to three times
# (0.25 / 0.25 / 0.5 averaging), a reference frame outside the neighbourhood, no frame dilation (uint8 mask branch,
# :556-561), other seeds. Only the final frames and the propagation mask are stored (pipeline12.npz).
SECOND = dict(N=12, clip_seed=9, fgt_seed=41, lafc_seed=42, raft_seed=43, flow_mask_dilates=2, frame_dilates=0)
# other driver modes (--watermark / --extrapolation): 7 frames; watermark: RGB mask files multiplied into the frames before
# resizing; extrapolation: 64x96 clip on an 80x120 canvas (H_scale = W_scale = 1.25; RAFT needs >= 128 px on its
# short side or its coarsest correlation level degenerates to one row and the reference divides by zero). Stored like
# the second case.
MODES = dict(watermark=dict(mode="watermark_removal", out="pipeline_watermark", clip_seed=13, fgt_seed=51, lafc_seed=52,
                            raft_seed=53, flow_mask_dilates=2, frame_dilates=0, consistencyThres=1.0),
             extrapolation=dict(mode="video_extrapolation", out="pipeline_extrapolation", clip_seed=14, fgt_seed=61,
                                lafc_seed=62, raft_seed=63, flow_mask_dilates=0, frame_dilates=0, consistencyThres=5.0,
                                H=64, W=96, scale=1.25, canvas=(80, 120)))


def main(second=False, other=None):
    global N, H, W
    cs, fs, ls, rs, fmd, fd = 5, 31, 32, 33, 3, 1
    mode, thres, scale, cfg_hw = "object_removal", 5.0, 2.0, None
    if other:
        o = MODES[other]
        mode, thres = o["mode"], o["consistencyThres"]
        cs, fs, ls, rs, fmd, fd = o["clip_seed"], o["fgt_seed"], o["lafc_seed"], o["raft_seed"], o["flow_mask_dilates"], o["frame_dilates"]
        if "scale" in o:
            scale, cfg_hw = o["scale"], o["canvas"]       # the FGT checkpoint is configured for the canvas size
            H, W = o["H"], o["W"]
    if second:
        N = SECOND["N"]
        cs, fs, ls, rs = SECOND["clip_seed"], SECOND["fgt_seed"], SECOND["lafc_seed"], SECOND["raft_seed"]
        fmd, fd = SECOND["flow_mask_dilates"], SECOND["frame_dilates"]
    for m in ("cvbase", "imageio", "skimage", "skimage.feature"):
        sys.modules.setdefault(m, types.ModuleType(m))
    sys.modules["skimage.feature"].canny = None
    sys.modules["imageio"].mimwrite = lambda path, frames, **kw: CAPTURE.__setitem__("comp_frames", [np.array(f) for f in frames])
    sys.path.insert(0, os.path.join(REF, "tool"))       # the script directory comes first when the driver is run
    import video_inpainting as VI

    tmp = tempfile.mkdtemp(prefix="fgt_pipeline_")
    try:
        frames, masks = synth.pipeline_clip(seed=cs, N=N, H=H, W=W)
        for d in ("frames", "masks", "fgt_ckpt", "lafc_ckpt", "out"):
            os.makedirs(os.path.join(tmp, d))
        for i, (fr, m) in enumerate(zip(frames, masks)):
            Image.fromarray(fr).save(os.path.join(tmp, "frames", "%05d.png" % i))
            Image.fromarray(np.repeat(m[..., None], 3, -1) if mode == "watermark_removal" else m).save(os.path.join(tmp, "masks", "%05d.png" % i))
        cfg = dict(synth.CFG_A)
        cfg["input_resolution"] = cfg_hw or (H, W)
        fgt_sd = synth.make_state_dict(synth.fgt_param_shapes(cfg), seed=fs)
        torch.save({"model_state_dict": fgt_sd}, os.path.join(tmp, "fgt_ckpt", "fgt.tar"))
        ycfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}
        ycfg["model"] = "model"
        with open(os.path.join(tmp, "fgt_ckpt", "config.yaml"), "w") as fh:
            yaml.safe_dump(ycfg, fh)
        lafc_sd = synth.make_state_dict(synth.lafc_param_shapes(synth.CFG_LAFC), seed=ls)
        torch.save({"model_state_dict": lafc_sd}, os.path.join(tmp, "lafc_ckpt", "lafc.tar"))
        with open(os.path.join(tmp, "lafc_ckpt", "config.yaml"), "w") as fh:
            yaml.safe_dump(dict(synth.CFG_LAFC), fh)
        raft_sd = synth.raft_state_dict(seed=rs)
        torch.save({"module." + k: v for k, v in raft_sd.items()}, os.path.join(tmp, "raft.pth"))
        opt = os.path.join(tmp, "opt.yaml")
        with open(opt, "w") as fh:
            yaml.safe_dump(dict(mode=mode, consistencyThres=thres, alpha=0.1, flow_mask_dilates=fmd, frame_dilates=fd), fh)
        args = argparse.Namespace(
            opt=opt, mode=mode, path=os.path.join(tmp, "frames"), path_mask=os.path.join(tmp, "masks"),
            outroot=os.path.join(tmp, "out"), consistencyThres=thres, alpha=0.1, Nonlocal=False,
            raft_model=os.path.join(tmp, "raft.pth"), small=False, mixed_precision=False,
            alternate_corr=False, lafc_ckpts=os.path.join(tmp, "lafc_ckpt"), fgt_ckpts=os.path.join(tmp, "fgt_ckpt"),
            H_scale=scale, W_scale=scale, imgH=H, imgW=W, flow_mask_dilates=fmd, frame_dilates=fd, gpu=0, step=10, num_ref=-1,
            neighbor_stride=5, vis_flows=False, vis_completed_flows=False, vis_prop=False, vis_frame=False)

        # hooks: record, then call the unmodified function
        t_calls, p_calls = [], []
        ref_np2tensor, ref_poisson = VI.np2tensor, VI.Poisson_blend_img

        def np2tensor(array, near="c"):
            if near == "t":
                t_calls.append(np.stack(array, 0).copy() if isinstance(array, list) else np.array(array))
            return ref_np2tensor(array, near)

        def poisson(*a, **kw):
            out = ref_poisson(*a, **kw)
            p_calls.append(([np.array(x) for x in a], [np.array(o) for o in out]))
            return out

        stages = {}
        ref_calc, ref_comp, ref_prop = VI.calculate_flow, VI.complete_flow, VI.get_flowNN_gradient

        def calc(a, model, video, mode):
            out = ref_calc(a, model, video, mode)
            stages["flow_" + mode[0]] = np.array(out)
            return out

        def comp_flow(config, model, flows, masks, mode, device):
            out = ref_comp(config, model, flows, masks, mode, device)
            stages["done_" + mode[0]] = VI.tensor2np(out)
            return out

        def prop(*a):
            out = ref_prop(*a)
            stages["mask_gradient"] = np.array(out[2])
            return out

        VI.calculate_flow, VI.complete_flow, VI.get_flowNN_gradient = calc, comp_flow, prop
        VI.np2tensor, VI.Poisson_blend_img = np2tensor, poisson
        with contextlib.redirect_stdout(io.StringIO()):
            VI.video_inpainting(args)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)

    import cv2
    import scipy
    meta = dict(mode=mode, consistencyThres=thres, scale=scale, cfg_hw=list(cfg_hw or (H, W)), H=H, W=W, N=N, fgt_seed=fs, lafc_seed=ls, raft_seed=rs, clip_seed=cs, flow_mask_dilates=fmd, frame_dilates=fd, torch=torch.__version__, numpy=np.__version__,
                scipy=scipy.__version__, cv2=cv2.__version__)
    frame_blends, mask, flow_f = t_calls            # the three near="t" conversions of the FGT stage, in order
    comp = np.stack(CAPTURE["comp_frames"])
    oh, ow = comp.shape[1:3]
    assert frame_blends.shape == (N, oh, ow, 3) and mask.shape == (N, oh, ow, 1) and flow_f.shape == (N, oh, ow, 2)
    assert comp.dtype == np.uint8 and ((oh, ow) == (H, W) or mode == "video_extrapolation")
    if second or other:
        np.savez_compressed(os.path.join(HERE, (MODES[other]["out"] if other else "pipeline12") + ".npz"), meta=np.array(repr(meta)), comp=comp,
                            mask_gradient=np.packbits(stages["mask_gradient"]), mask_final=np.packbits(mask.astype(bool)))
        print(MODES[other]["out"] if other else "pipeline12", "saved: comp", comp.shape, "mean", comp.mean(), "final holes", int(mask.sum()))
        return
    # frame_blends as recorded are already RGB (the driver flips in place before np2tensor, :688-689)
    np.savez_compressed(os.path.join(HERE, "pipeline_clip.npz"), meta=np.array(repr(meta)),
                        frames_rgb=frame_blends, mask=np.packbits(mask.astype(bool)), flow_f=flow_f[:-1].astype(np.float32),
                        comp=comp)
    trg = np.stack([a[0] for a, _ in p_calls]); gx = np.stack([a[1] for a, _ in p_calls]); gy = np.stack([a[2] for a, _ in p_calls])
    hole = np.stack([a[3] for a, _ in p_calls]); gm = np.stack([a[4] for a, _ in p_calls])
    blend = np.stack([o[0] for _, o in p_calls]); unf = np.stack([o[1] for _, o in p_calls])
    assert trg.dtype in (np.float32, np.float64) and gx.dtype == np.float32 and hole.dtype == np.bool_
    np.savez_compressed(os.path.join(HERE, "pipeline_poisson.npz"), meta=np.array(repr(meta)), trg=trg, gx=gx, gy=gy,
                        hole=np.packbits(hole), gmask=np.packbits(gm), blend_hole=blend[hole], unfilled=np.packbits(unf))
    np.savez_compressed(os.path.join(HERE, "pipeline_stages.npz"), meta=np.array(repr(meta)),
                        flow_f=stages["flow_f"].astype(np.float16), flow_b=stages["flow_b"].astype(np.float16),
                        done_f=stages["done_f"].astype(np.float16), done_b=stages["done_b"].astype(np.float16),
                        mask_gradient=np.packbits(stages["mask_gradient"]))
    print("flow magnitudes: raw", np.abs(stages["flow_f"]).mean(), "completed", np.abs(stages["done_f"]).mean())
    print("pipeline goldens saved:", len(p_calls), "Poisson calls, holes", int(hole.sum()), "unfilled", int(unf.sum()),
          "| comp mean", comp.mean(), "| blend dtype", blend.dtype, "trg dtype", trg.dtype)


if __name__ == "__main__":
    main(second="--second" in sys.argv,
         other="watermark" if "--watermark" in sys.argv else ("extrapolation" if "--extrapolation" in sys.argv else None))
