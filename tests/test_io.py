"""File formats. The `.flo` bytes are checked against a file written by the reference's
RAFT/utils/frame_utils.py::writeFlow (tests/golden/flo_ref.flo, 3x5 flow with a known pattern, generated in the
build container) and the reference's readFlow semantics."""
import os

import numpy as np
import pytest

from fgt_b200 import io as IO
from tests.util import GOLDEN


def _flow():
    h, w = 3, 5
    u = np.arange(h * w, dtype=np.float32).reshape(h, w) * 0.5 - 2
    v = -np.arange(h * w, dtype=np.float32).reshape(h, w) * 0.25 + 1
    return np.stack([u, v], -1)


def test_flo_bytes_match_reference_writer(tmp_path):
    p = str(tmp_path / "a.flo")
    IO.write_flo(p, _flow())
    with open(p, "rb") as a, open(os.path.join(GOLDEN, "flo_ref.flo"), "rb") as b:
        assert a.read() == b.read()
    q = str(tmp_path / "b.flo")
    IO.write_flo(q, _flow()[..., 0], _flow()[..., 1])                    # (u, v) form of writeFlow
    assert open(q, "rb").read() == open(p, "rb").read()
    back = IO.read_flo(os.path.join(GOLDEN, "flo_ref.flo"))
    assert back.dtype == np.float32 and np.array_equal(back, _flow())


def test_flo_rejects_bad_files(tmp_path):
    p = str(tmp_path / "bad.flo")
    with open(p, "wb") as fh:
        fh.write(b"\x00" * 12)
    with pytest.raises(ValueError):
        IO.read_flo(p)
    IO.write_flo(p, _flow())
    with open(p, "r+b") as fh:
        fh.truncate(20)
    with pytest.raises(ValueError):
        IO.read_flo(p)
    with pytest.raises(ValueError):
        IO.write_flo(p, np.zeros((3, 4, 3)))


def test_frame_directory_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    frames = [rng.integers(0, 256, (12, 16, 3), dtype=np.uint8) for _ in range(3)]
    written = IO.write_frames(str(tmp_path), frames, mp4=False)
    assert [os.path.basename(p) for p in written] == ["00000.png", "00001.png", "00002.png"]
    back = IO.read_frames(str(tmp_path / "frames"))
    assert all(np.array_equal(a, b) for a, b in zip(frames, back))
    masks = IO.read_masks(str(tmp_path / "frames"))
    assert masks[0].shape == (12, 16) and masks[0].dtype == np.uint8
    with pytest.raises(FileNotFoundError):
        IO.read_frames(str(tmp_path / "nothing"))


def test_flow_extract_writes_the_reference_layout(tmp_path):
    """tool/flow_extract.py's output tree, with a stand-in for RAFT (flow = mean colour difference, 2 channels)."""
    import torch
    from fgt_b200 import flow_extract as FE
    rng = np.random.default_rng(1)
    frames = [rng.integers(0, 256, (20, 30, 3), dtype=np.uint8) for _ in range(4)]

    def fake_raft(a, b, iters):
        d = (b - a).mean(1, keepdim=True)
        return torch.cat([d, -d], 1).numpy()

    fwd, bwd = FE.extract_video(frames, fake_raft, str(tmp_path / "vid"), width=32, height=24)
    assert fwd.shape == (3, 24, 32, 2) and np.allclose(fwd, -bwd)
    for mode, ref in (("forward_flo", fwd), ("backward_flo", bwd)):
        files = sorted(os.listdir(tmp_path / "vid" / mode))
        assert files == ["00000.flo", "00001.flo", "00002.flo"]
        assert np.array_equal(IO.read_flo(str(tmp_path / "vid" / mode / files[1])), ref[1])
    with pytest.raises(ValueError):
        FE.extract_video(frames, fake_raft, str(tmp_path / "bad"), width=30, height=20)
