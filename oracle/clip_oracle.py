"""CPU oracle for the FGT stage of the driver (TEST INFRASTRUCTURE — see oracle/fgt_oracle.py).

What /root/reference/tool/video_inpainting.py:686-745 computes between Poisson blending and the video writer,
around an arbitrary `model` callable, restated as three steps:

* `clip_tensors`  — frames / masks / flows as the float32 tensors the driver builds (`np2tensor(..., near='t')`
  :55-66,:691-694; BGR->RGB flip :688-689; last flow repeated :702-704; `norm_flows` :402-407: every flow channel of
  every frame divided by its own signed maximum);
* `windows`       — the sliding windows: 2*stride+1 neighbours around every stride-th frame plus reference frames every
  `step` frames that are not neighbours (`get_ref_index` :103-117, loop :709-717);
* `fgt_stage`     — per window: masked = (frames*2-1)*(1-mask) (:695,:719-722), model, output mapped to 0..255 in
  float32 and truncated to uint8, blended with the input frame by the mask as float32 (:725-733), first visit stored,
  later visits averaged 0.5/0.5 with what is stored (:734-741), final truncation to uint8 (:744-745).

The enclosing function cannot be called piecewise, so this file is pinned by a full run of the unmodified driver
(tests/golden/make_pipeline_golden.py -> pipeline_clip.npz: inputs of the stage and the frames given to the writer).
"""
import numpy as np
import torch


def clip_tensors(frame_blends, mask, video_flow_f):
    """frame_blends: N arrays [H,W,3] BGR in [0,1]; mask [H,W,N]; video_flow_f [H,W,2,N-1] ->
    frames [1,N,3,H,W], masks [1,N,1,H,W], flows [1,N,2,H,W] (float32, flows normalised)."""
    rgb = np.stack([np.ascontiguousarray(f[:, :, ::-1]) for f in frame_blends], 0)            # [N,H,W,3]
    frames = torch.from_numpy(rgb.transpose(0, 3, 1, 2)).float()[None]
    m = np.moveaxis(np.asarray(mask), -1, 0)[:, None]                                         # [N,1,H,W]
    masks = torch.from_numpy(np.ascontiguousarray(m)).float()[None]
    fl = np.moveaxis(np.asarray(video_flow_f), -1, 0)                                         # [N-1,H,W,2]
    fl = np.concatenate([fl, fl[-1:]], 0).transpose(0, 3, 1, 2)
    flows = torch.from_numpy(np.ascontiguousarray(fl)).float()[None]
    peak = flows.flatten(3).max(dim=-1, keepdim=True)[0].unsqueeze(-1)                        # per frame and channel
    return frames, masks, flows / peak


def windows(n_frames, step=10, num_ref=-1, neighbor_stride=5):
    """[(neighbour ids, reference ids)] in the driver's order."""
    out = []
    for centre in range(0, n_frames, neighbor_stride):
        near = list(range(max(0, centre - neighbor_stride), min(n_frames, centre + neighbor_stride + 1)))
        if num_ref == -1:
            ref = [i for i in range(0, n_frames, step) if i not in near]
        else:
            lo = max(0, centre - step * (num_ref // 2))
            hi = min(n_frames, centre + step * (num_ref // 2))
            ref = []
            for i in range(lo, hi + 1, step):
                if i in near:
                    continue
                if len(ref) > num_ref:
                    break
                ref.append(i)
        out.append((near, ref))
    return out


def get_ref_index(f, neighbor_ids, length, ref_length, num_ref):
    """get_ref_index (:103-117) for one window, via `windows` (kept as a separate entry point for the tests)."""
    if num_ref == -1:
        return [i for i in range(0, length, ref_length) if i not in neighbor_ids]
    ref = []
    for i in range(max(0, f - ref_length * (num_ref // 2)), min(length, f + ref_length * (num_ref // 2)) + 1, ref_length):
        if i not in neighbor_ids:
            if len(ref) > num_ref:
                break
            ref.append(i)
    return ref


def norm_flows(flows):
    """norm_flows (:402-407) on a [b,t,c,H,W] tensor."""
    assert flows.dim() == 5
    return flows / flows.flatten(3).max(dim=-1, keepdim=True)[0].unsqueeze(-1)


def np2tensor(array, near="c"):
    """np2tensor (:55-66): list / array [t,h,w,c] -> [1,c,t,h,w] (near='c') or [1,t,c,h,w] (near='t'), float32."""
    a = np.stack(array, 0) if isinstance(array, list) else array
    order = {"c": (3, 0, 1, 2), "t": (0, 3, 1, 2)}
    if near not in order:
        raise ValueError(f"Unknown near type: {near}")
    return torch.from_numpy(np.transpose(a, order[near])).unsqueeze(0).float()


def _to_u8_levels(x01):
    """float32 image in [0,1] (or model output mapped there) -> uint8 levels the way the driver does it: multiply by
    255 in float32, truncate."""
    return (x01 * np.float32(255)).astype(np.uint8)


def fgt_stage(model, frame_blends, mask, video_flow_f, step=10, num_ref=-1, neighbor_stride=5):
    """model(masked_frames [1,t,3,H,W], flows [1,t,2,H,W], masks [1,t,1,H,W]) -> [t,3,H,W] in [-1,1].
    Returns the list of N uint8 [H,W,3] RGB frames."""
    frames, masks, flows = clip_tensors(frame_blends, mask, video_flow_f)
    normed = frames * 2 - 1
    n = frames.shape[1]
    known = [_to_u8_levels(frames[0, i].permute(1, 2, 0).numpy()) for i in range(n)]       # input frames as levels
    hole = [masks[0, i].permute(1, 2, 0).numpy() for i in range(n)]                         # float32 [H,W,1]
    result = [None] * n
    for near, ref in windows(n, step, num_ref, neighbor_stride):
        ids = near + ref
        sel_masks = masks[:, ids]
        with torch.no_grad():
            out = model(normed[:, ids] * (1 - sel_masks), flows[:, ids], sel_masks)
        levels = ((out + 1) / 2).cpu().permute(0, 2, 3, 1).numpy() * 255                   # float32, :725-726
        for k, idx in enumerate(near):                                                      # reference frames are inputs only
            blended = levels[k].astype(np.uint8) * hole[idx] + known[idx] * (1 - hole[idx])  # float32
            if result[idx] is None:
                result[idx] = blended
            else:
                result[idx] = result[idx].astype(np.float32) * 0.5 + blended.astype(np.float32) * 0.5
    return [r.astype(np.uint8) for r in result]
