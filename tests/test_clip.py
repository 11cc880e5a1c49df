"""The FGT stage of the driver (window loop + compositing, tool/video_inpainting.py:686-745).

Golden: tests/golden/pipeline_clip.npz — the stage's inputs and the frames handed to the video writer, recorded
from ONE RUN OF THE UNMODIFIED REFERENCE DRIVER on the CPU (tests/golden/make_pipeline_golden.py: RAFT with the
real checkpoint -> LAFC -> propagation -> Poisson -> FGT, seeded synthetic FGT / LAFC weights).
Tolerances: gather / compose arithmetic is bit-exact (asserted with an exactly reproducible stand-in model);
with the real models the uint8 frames may differ by 1 level where the model output (1e-6 oracle, 1e-4 CUDA
relative error) straddles a truncation boundary."""
import numpy as np
import pytest
import torch

from fgt_b200 import synth
from fgt_b200.parallel import window_schedule
from oracle import clip_oracle as CO
from oracle import fgt_oracle as O
from tests.util import load_golden


def _golden():
    g = load_golden("pipeline_clip")
    N, H, W = g["meta"]["N"], g["meta"]["H"], g["meta"]["W"]
    mask = np.unpackbits(g["mask"])[:N * H * W].reshape(N, H, W).astype(bool)
    frame_blends = [np.ascontiguousarray(g["frames_rgb"][i][:, :, ::-1]) for i in range(N)]   # the stage receives BGR
    flow_f = np.ascontiguousarray(np.moveaxis(g["flow_f"], 0, -1))                            # [H,W,2,N-1]
    return g, frame_blends, np.ascontiguousarray(np.moveaxis(mask, 0, -1)), flow_f


def _fgt_weights(g):
    cfg = dict(synth.CFG_A)
    cfg["input_resolution"] = (g["meta"]["H"], g["meta"]["W"])
    return cfg, synth.make_state_dict(synth.fgt_param_shapes(cfg), seed=g["meta"]["fgt_seed"])


def _exact_model(frames, flows, masks):
    """Stand-in with the model's signature whose arithmetic is exactly reproducible on any device: scalings by
    powers of two and single IEEE additions."""
    return (frames[0] * 0.5 + flows[0, :, :1] * 0.25 + masks[0] * 0.125).clamp(-1, 1)


def test_oracle_helpers_and_schedule():
    assert [(f, n, r) for f, n, r in window_schedule(7)] == [
        (f, [i for i in range(max(0, f - 5), min(7, f + 6))],
         CO.get_ref_index(f, [i for i in range(max(0, f - 5), min(7, f + 6))], 7, 10, -1)) for f in (0, 5)]
    for f in range(0, 40, 5):
        nb = list(range(max(0, f - 5), min(40, f + 6)))
        for num_ref in (-1, 2, 4):
            assert window_schedule(40, 5, 10, num_ref)[f // 5][2] == CO.get_ref_index(f, nb, 40, 10, num_ref)
    for n in (1, 2, 7, 10, 12, 23, 80):                       # the product's schedule vs the oracle's own statement
        for num_ref in (-1, 0, 2, 5):
            for stride, step in ((5, 10), (3, 7)):
                want = CO.windows(n, step, num_ref, stride)
                assert [(nb, ref) for _, nb, ref in window_schedule(n, stride, step, num_ref)] == want, (n, num_ref, stride)
    fl = torch.randn(1, 3, 2, 4, 5)
    nf = CO.norm_flows(fl)
    assert torch.equal(nf[0, 1, 0], fl[0, 1, 0] / fl[0, 1, 0].max())
    assert CO.np2tensor([np.zeros((4, 5, 3))] * 2, near="t").shape == (1, 2, 3, 4, 5)


def test_oracle_stage_matches_reference_driver_golden():
    g, frame_blends, mask, flow_f = _golden()
    _, sd = _fgt_weights(g)
    inner = O.strip_net(sd)
    comp = np.stack(CO.fgt_stage(lambda a, b, c: O.fgt_forward(inner, a, b, c), frame_blends, mask, flow_f))
    assert comp.dtype == np.uint8 and comp.shape == g["comp"].shape
    diff = np.abs(comp.astype(np.int16) - g["comp"].astype(np.int16))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3, (diff.max(), (diff != 0).mean())
    hole = np.moveaxis(mask, -1, 0)
    assert (comp[~hole] == g["comp"][~hole]).all()      # outside the holes the input frame passes through


@pytest.mark.gpu
def test_gpu_stage_arithmetic_is_bit_exact():
    from fgt_b200 import clip as C
    g, frame_blends, mask, flow_f = _golden()
    want = np.stack(CO.fgt_stage(_exact_model, frame_blends, mask, flow_f))
    got = np.stack(C.inpaint_clip(_exact_model, frame_blends, mask, flow_f))
    assert got.dtype == np.uint8 and np.array_equal(got, want)
    # a longer clip: three visits per frame (0.25 / 0.25 / 0.5 weights), reference frames, num_ref != -1
    gen = torch.Generator().manual_seed(3)
    N, H, W = 23, 32, 48
    fb = [torch.rand(H, W, 3, generator=gen).double().numpy() for _ in range(N)]
    mk = (torch.rand(H, W, N, generator=gen) < 0.3).numpy()
    ff = (torch.randn(H, W, 2, N - 1, generator=gen) * 3).numpy().astype(np.float32)
    for num_ref in (-1, 2):
        want = np.stack(CO.fgt_stage(_exact_model, fb, mk, ff, num_ref=num_ref))
        got = np.stack(C.inpaint_clip(_exact_model, fb, mk, ff, num_ref=num_ref))
        assert np.array_equal(got, want), num_ref
    with pytest.raises(RuntimeError):
        C.inpaint_clip(_exact_model, fb, mk, ff, device="cpu")


@pytest.mark.gpu
def test_gpu_stage_with_cuda_model_vs_reference_driver_golden():
    from fgt_b200 import clip as C
    from fgt_b200.fgt_model import Model
    g, frame_blends, mask, flow_f = _golden()
    cfg, sd = _fgt_weights(g)
    model = Model(cfg)
    model.load_state_dict(sd)
    model = model.to("cuda:0")
    got = np.stack(C.inpaint_clip(model, frame_blends, mask, flow_f))
    diff = np.abs(got.astype(np.int16) - g["comp"].astype(np.int16))
    assert diff.max() <= 1 and (diff != 0).mean() < 5e-3, (diff.max(), (diff != 0).mean())
    hole = np.moveaxis(mask, -1, 0)
    assert (got[~hole] == g["comp"][~hole]).all()
