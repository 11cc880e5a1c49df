"""CPU oracle for the FGT stage of the driver (TEST INFRASTRUCTURE — see oracle/fgt_oracle.py).

Restates /root/reference/tool/video_inpainting.py:686-745 — the lines of `video_inpainting()` between Poisson
blending and the video writer — around an arbitrary `model` callable: np2tensor (:55-66), norm_flows (:402-407),
get_ref_index (:103-117), the window loop (:709-741) and the final uint8 cast (:744-745). The enclosing function
cannot be imported in this image (cvbase / imageio / skimage are missing), so the loop body is restated line
by line with the same numpy / torch operations and dtypes; the three helpers are checked against the reference's
own (imported with the missing modules stubbed) when tests/golden/make_golden.py generates clip_stage.npz.
"""
import numpy as np
import torch


def np2tensor(array, near="c"):
    """video_inpainting.py:55-66."""
    if isinstance(array, list):
        array = np.stack(array, axis=0)
    if near == "c":
        return torch.from_numpy(np.transpose(array, (3, 0, 1, 2))).unsqueeze(0).float()
    if near == "t":
        return torch.from_numpy(np.transpose(array, (0, 3, 1, 2))).unsqueeze(0).float()
    raise ValueError(f"Unknown near type: {near}")


def norm_flows(flows):
    """video_inpainting.py:402-407: every flow channel of every frame divided by its own (signed) maximum."""
    assert flows.dim() == 5
    flow_max = torch.max(flows.flatten(3), dim=-1, keepdim=True)[0]
    return flows / flow_max.unsqueeze(-1)


def get_ref_index(f, neighbor_ids, length, ref_length, num_ref):
    """video_inpainting.py:103-117."""
    ref_index = []
    if num_ref == -1:
        for i in range(0, length, ref_length):
            if i not in neighbor_ids:
                ref_index.append(i)
    else:
        start_idx = max(0, f - ref_length * (num_ref // 2))
        end_idx = min(length, f + ref_length * (num_ref // 2))
        for i in range(start_idx, end_idx + 1, ref_length):
            if i not in neighbor_ids:
                if len(ref_index) > num_ref:
                    break
                ref_index.append(i)
    return ref_index


def fgt_stage(model, frame_blends, mask, video_flow_f, step=10, num_ref=-1, neighbor_stride=5):
    """frame_blends: list of N [H,W,3] BGR float arrays in [0,1]; mask [H,W,N] bool; video_flow_f [H,W,2,N-1]
    float32; model(masked_frames, flows, masks) -> [t,3,H,W]. Returns the list of N uint8 [H,W,3] frames."""
    frame_blends = [fb[:, :, ::-1] for fb in frame_blends]                               # :688-689
    video_length = len(frame_blends)
    frames_first = np2tensor([np.ascontiguousarray(fb) for fb in frame_blends], near="t")  # :691
    mask = np.moveaxis(np.asarray(mask), -1, 0)[:, :, :, np.newaxis]                     # :692-693
    masks = np2tensor(mask, near="t")                                                    # :694
    normed_frames = frames_first * 2 - 1                                                 # :695
    comp_frames = [None] * video_length
    flow = np.moveaxis(np.asarray(video_flow_f), -1, 0)                                  # :702
    flow = np.concatenate([flow, flow[-1:, ...]], axis=0)                                # :704
    flows = norm_flows(np2tensor(flow, near="t"))                                        # :706-707
    for f in range(0, video_length, neighbor_stride):                                    # :709
        neighbor_ids = [i for i in range(max(0, f - neighbor_stride), min(video_length, f + neighbor_stride + 1))]
        ref_ids = get_ref_index(f, neighbor_ids, video_length, step, num_ref)
        selected_frames = normed_frames[:, neighbor_ids + ref_ids]
        selected_masks = masks[:, neighbor_ids + ref_ids]
        masked_frames = selected_frames * (1 - selected_masks)
        selected_flows = flows[:, neighbor_ids + ref_ids]
        with torch.no_grad():
            filled_frames = model(masked_frames, selected_flows, selected_masks)
        filled_frames = (filled_frames + 1) / 2
        filled_frames = filled_frames.cpu().permute(0, 2, 3, 1).numpy() * 255
        for i in range(len(neighbor_ids)):
            idx = neighbor_ids[i]
            valid_frame = frames_first[0, idx].cpu().permute(1, 2, 0).numpy() * 255.0
            valid_mask = masks[0, idx].cpu().permute(1, 2, 0).numpy()
            comp = np.array(filled_frames[i]).astype(np.uint8) * valid_mask + \
                np.array(valid_frame).astype(np.uint8) * (1 - valid_mask)
            if comp_frames[idx] is None:
                comp_frames[idx] = comp
            else:
                comp_frames[idx] = comp_frames[idx].astype(np.float32) * 0.5 + comp.astype(np.float32) * 0.5
    return [c.astype(np.uint8) for c in comp_frames]                                     # :744-745
