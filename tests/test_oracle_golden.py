"""CPU: the oracle restatement (oracle/fgt_oracle.py) against golden outputs of the unmodified
reference (tests/golden/*.npz, produced by tests/golden/make_golden.py in the build container)."""
import pytest
import torch

from fgt_b200 import synth
from oracle import fgt_oracle as O
from tests.util import assert_close, load_golden

ORACLE_TOL = 2e-5  # fp32 CPU vs fp32 CPU: only summation-order noise is allowed


def _run_oracle(meta, return_intermediates=False):
    cfg = dict(synth.CFG_A)
    cfg["input_resolution"] = tuple(meta["res"])
    sd = synth.make_state_dict(synth.fgt_param_shapes(cfg), seed=meta["seed"], regime=meta["regime"])
    fr, fl, mk = synth.fgt_inputs(seed=meta["seed"] + 2, t=meta["t"], H=meta["H"], W=meta["W"])
    with torch.no_grad():
        return O.fgt_forward(O.strip_net(sd), fr, fl, mk, return_intermediates=return_intermediates)


@pytest.mark.parametrize("name", ["fgt_small_scaled", "fgt_small_default", "fgt_runtime_geo"])
def test_oracle_full_model_small(name):
    g = load_golden(name)
    out = _run_oracle(g["meta"])
    assert tuple(out.shape) == tuple(g["out"].shape)
    assert_close(out, g["out"], ORACLE_TOL, name)


def test_oracle_full_model_t10_sampled():
    """BASELINE config 2 (432x240, T=10): reference output pinned by 16384 sampled pixels + norms."""
    g = load_golden("fgt_full_t10")
    out = _run_oracle(g["meta"])
    assert tuple(out.shape) == (10, 3, 240, 432)
    assert_close(out.reshape(-1)[torch.from_numpy(g["idx"])], g["val"], ORACLE_TOL, "fgt_full_t10 samples")
    assert abs(out.double().norm().item() - float(g["l2"])) / float(g["l2"]) < 1e-5


def test_oracle_modules():
    g = load_golden("fgt_modules")
    sd = O.strip_net(synth.make_state_dict(synth.fgt_param_shapes(synth.CFG_A), seed=7, regime="scaled"))
    gen = torch.Generator().manual_seed(99)
    x = torch.randn(2, 720, 512, generator=gen)
    f = torch.randn(2, 720, 256, generator=gen)
    idx = torch.from_numpy(g["idx"])
    with torch.no_grad():
        sw = O.swmhsa(x, f, sd, "first_s_transformer.attention.", 20, 36)
        tm = O.tmhsa(x, sd, "first_t_transformer.attention.", 2, 20, 36)
        ff = O.fusion_ffn(x, sd, "first_t_transformer.ffn.", 720, (60, 108))
    for name, val in (("swmhsa", sw), ("tmhsa", tm), ("ffn", ff)):
        assert_close(val.reshape(-1)[idx], g[name], ORACLE_TOL, name)
        assert abs(val.double().norm().item() - float(g[name + "_l2"])) / float(g[name + "_l2"]) < 1e-5


def test_oracle_driver_default_size_256x432():
    """Runtime geometry of the driver's default 256x432 working size on a 240x432 checkpoint."""
    g = load_golden("fgt_driver_256x432_t6")
    out = _run_oracle(g["meta"])
    assert tuple(out.shape) == (6, 3, 256, 432)
    assert_close(out.reshape(-1)[torch.from_numpy(g["idx"])], g["val"], ORACLE_TOL, "fgt_driver_256x432 samples")
    assert abs(out.double().norm().item() - float(g["l2"])) / float(g["l2"]) < 1e-5
