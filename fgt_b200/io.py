"""File formats at the two ends of the driver, host side: Middlebury `.flo` flow files
(RAFT/utils/frame_utils.py:12-32 `readFlow`, :70-99 `writeFlow`; the driver writes them with `--vis_flows`,
tool/video_inpainting.py:278-281, and reads them back in `read_flow`, :388-399), the frame / mask directories the
driver globs (:446-449,:539-546) and the result frames (`save_results`, :410-417). The MP4 writer of the driver is
`imageio.mimwrite` (:746); imageio is not part of this image, so clips are written as PNG frames (and, when
OpenCV's writer has a codec, an `.mp4` next to them). Pure I/O — no arithmetic on the hot path lives here.
"""
import glob
import os

import numpy as np
from PIL import Image

FLO_TAG = np.float32(202021.25)        # 'PIEH' as a little-endian float


def write_flo(path, uv, v=None):
    """writeFlow: header (tag, int32 width, int32 height) + interleaved float32 (u, v) rows."""
    if v is None:
        uv = np.asarray(uv)
        if uv.ndim != 3 or uv.shape[2] != 2:
            raise ValueError(f"write_flo: flow {uv.shape} must be [H,W,2]")
        u, v = uv[:, :, 0], uv[:, :, 1]
    else:
        u, v = np.asarray(uv), np.asarray(v)
    if u.shape != v.shape or u.ndim != 2:
        raise ValueError("write_flo: u and v must be [H,W] arrays of equal shape")
    h, w = u.shape
    body = np.empty((h, w, 2), dtype="<f4")
    body[..., 0], body[..., 1] = u, v
    with open(path, "wb") as fh:
        fh.write(np.array([FLO_TAG], dtype="<f4").tobytes())
        fh.write(np.array([w, h], dtype="<i4").tobytes())
        fh.write(body.tobytes())


def read_flo(path):
    """readFlow: [H,W,2] float32; raises on a bad tag or a truncated file (the reference prints and returns None
    for the former and silently tiles the data for the latter)."""
    with open(path, "rb") as fh:
        head = fh.read(12)
        if len(head) != 12 or np.frombuffer(head[:4], dtype="<f4")[0] != FLO_TAG:
            raise ValueError(f"{path}: not a Middlebury .flo file (bad magic number)")
        w, h = (int(x) for x in np.frombuffer(head[4:], dtype="<i4"))
        data = np.frombuffer(fh.read(), dtype="<f4")
    if w <= 0 or h <= 0 or data.size != 2 * w * h:
        raise ValueError(f"{path}: header says {w}x{h} but the file holds {data.size} values")
    return data.reshape(h, w, 2).astype(np.float32)


def list_images(directory):
    """The driver's file order: *.png then *.jpg, each sorted together (:446-449 + sorted(), :474)."""
    return sorted(glob.glob(os.path.join(directory, "*.png")) + glob.glob(os.path.join(directory, "*.jpg")))


def read_frames(directory):
    """RGB uint8 frames as the driver loads them (np.array(Image.open(f)).astype(np.uint8), :476)."""
    files = list_images(directory)
    if not files:
        raise FileNotFoundError(f"no *.png / *.jpg frames in {directory}")
    return [np.array(Image.open(f).convert("RGB")).astype(np.uint8) for f in files]


def read_masks(directory, rgb=False):
    """Greyscale masks as the driver loads them (Image.open(f).convert('L'), :543); rgb=True keeps the file's own
    channels, which is what the watermark mode multiplies the frames with (:466-471)."""
    files = list_images(directory)
    if not files:
        raise FileNotFoundError(f"no *.png / *.jpg masks in {directory}")
    return [np.array(Image.open(f)) if rgb else np.array(Image.open(f).convert("L")) for f in files]


def write_frames(outdir, frames, mp4=True, fps=30):
    """save_results (:410-417): frames/%05d.png; plus result.mp4 when cv2 can encode it. Returns the paths written."""
    fdir = os.path.join(outdir, "frames")
    os.makedirs(fdir, exist_ok=True)
    written = []
    for i, fr in enumerate(frames):
        p = os.path.join(fdir, "%05d.png" % i)
        Image.fromarray(np.asarray(fr).astype(np.uint8)).save(p)
        written.append(p)
    if mp4 and len(frames):
        import cv2
        h, w = np.asarray(frames[0]).shape[:2]
        path = os.path.join(outdir, "result.mp4")
        vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (w, h))
        if vw.isOpened():
            for fr in frames:
                vw.write(np.ascontiguousarray(np.asarray(fr).astype(np.uint8)[:, :, ::-1]))
            vw.release()
            written.append(path)
    return written
