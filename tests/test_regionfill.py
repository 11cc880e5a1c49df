"""Region fill / flow diffusion: oracle vs the reference goldens on the CPU, CUDA
batched CG vs oracle and goldens on the GPU. fp64 throughout; tolerance 1e-7 absolute on values of
magnitude ~10 (the CG stops at a relative residual of 1e-12)."""
import numpy as np
import pytest

from fgt_b200 import synth
from oracle import regionfill_oracle as RO
from tests.util import load_golden

CASES = {"regionfill_small": (4, 48, 64, 5), "regionfill_mid": (3, 120, 216, 6)}


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    B, H, W, seed = CASES[name]
    g = load_golden(name)
    assert (g["meta"]["B"], g["meta"]["H"], g["meta"]["W"], g["meta"]["seed"]) == (B, H, W, seed)
    img, mask = synth.regionfill_inputs(seed=seed, B=B, H=H, W=W)
    out = np.stack([RO.regionfill(img[b], mask[b]) for b in range(B)])
    assert np.array_equal(out[~mask], img.astype(np.float64)[~mask])
    assert np.abs(out[mask] - g["out_hole"]).max() < 1e-10
    # harmonic: every hole value is the mean of its in-image neighbours
    b, y, x = [v[0] for v in np.nonzero(mask)]
    nb = [out[b, yy, xx] for yy, xx in ((y - 1, x), (y + 1, x), (y, x - 1), (y, x + 1)) if 0 <= yy < H and 0 <= xx < W]
    assert abs(out[b, y, x] - np.mean(nb)) < 1e-9


def test_oracle_diffusion_shapes():
    img, mask = synth.regionfill_inputs(seed=1, B=2, H=24, W=32)
    flows = np.stack([img, img[::-1]], -1)
    out = RO.diffusion(flows, mask[..., None])
    assert len(out) == 2 and out[0].shape == (24, 32, 2) and out[0].dtype == np.float64


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_gpu_regionfill_vs_golden_and_oracle(name):
    from fgt_b200 import regionfill as RF
    B, H, W, seed = CASES[name]
    g = load_golden(name)
    img, mask = synth.regionfill_inputs(seed=seed, B=B, H=H, W=W)
    out, iters = RF.regionfill_batch(img, mask, return_iters=True)
    out = out.cpu().numpy()
    assert out.dtype == np.float64 and 0 < iters < 10000
    assert np.array_equal(out[~mask], img.astype(np.float64)[~mask])          # untouched outside the holes
    assert np.abs(out[mask] - g["out_hole"]).max() < 1e-7, "vs reference golden"
    ora = np.stack([RO.regionfill(img[b], mask[b]) for b in range(B)])
    assert np.abs(out - ora).max() < 1e-7, "vs oracle"


@pytest.mark.gpu
def test_gpu_regionfill_api_forms():
    from fgt_b200 import regionfill as RF
    img, mask = synth.regionfill_inputs(seed=9, B=4, H=40, W=56)
    one = RF.regionfill(img[1], mask[1])
    assert one.dtype == np.float64 and np.abs(one - RO.regionfill(img[1], mask[1])).max() < 1e-7
    assert np.array_equal(RF.regionfill(img[3], mask[3]), img[3])                  # empty mask: copy of the input
    flows = np.stack([img, img[::-1].copy()], -1)
    got = RF.diffusion(flows, mask[..., None])
    ref = RO.diffusion(flows, mask[..., None])
    assert len(got) == 4 and all(np.abs(a - b).max() < 1e-7 for a, b in zip(got, ref))
    with pytest.raises(ValueError):
        RF.regionfill(img[0], mask[0], factor=0.5)
    with pytest.raises(ValueError):
        RF.regionfill_batch(img[:1], np.ones_like(mask[:1]))                      # mask covers the image
    with pytest.raises(RuntimeError):
        RF.regionfill_batch(img, mask, device="cpu")                              # no CPU fallback
