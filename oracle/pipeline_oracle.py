"""CPU backend for fgt_b200.pipeline.video_inpainting (TEST INFRASTRUCTURE — see oracle/fgt_oracle.py).

Implements the stage interface of `fgt_b200.pipeline.GpuBackend` with the CPU oracles (each pinned to the
reference by its own goldens): RAFT (oracle/raft_oracle.py), region fill (regionfill_oracle), LAFC (lafc_oracle),
gradient propagation (prop_oracle), Poisson blending (poisson_oracle), the FGT stage (clip_oracle + fgt_oracle).
Plugging it into the pipeline glue and comparing with tests/golden/pipeline_*.npz — recorded from a full run of the
unmodified reference driver tool/video_inpainting.py::video_inpainting — verifies the glue without a GPU.
"""
import torch

from . import clip_oracle, fgt_oracle, lafc_oracle, poisson_oracle, prop_oracle, raft_oracle, regionfill_oracle


class OracleBackend:
    def __init__(self, raft_sd, lafc_sd, fgt_sd):
        """State dicts without the wrappers' prefixes ('module.' for RAFT, 'net.' for LAFC / FGT)."""
        self.raft_sd, self.lafc_sd, self.fgt_sd = raft_sd, lafc_sd, fgt_sd
        self.fgt_model = lambda a, b, c: fgt_oracle.fgt_forward(self.fgt_sd, a, b, c)   # swapped by ShardedBackend

    def raft_pairs(self, img1, img2, iters):
        with torch.no_grad():
            return torch.cat([raft_oracle.raft_forward(self.raft_sd, img1[i:i + 1], img2[i:i + 1], iters=iters)[1]
                              for i in range(img1.shape[0])], 0).numpy()

    def diffusion(self, flows, masks):
        return regionfill_oracle.diffusion(flows, masks)

    def lafc_complete(self, flows, masks, diffused, triplets, pivot):
        fl, mk, df = (torch.from_numpy(a).unsqueeze(0) for a in (flows, masks, diffused))
        out = []
        with torch.no_grad():
            for idx in triplets:
                cand_masks = mk[:, :, idx]
                res = lafc_oracle.lafc_forward(self.lafc_sd, df[:, :, idx], cand_masks)[0]
                pm = cand_masks[:, :, pivot]
                out.append(res * pm + fl[:, :, idx][:, :, pivot] * (1 - pm))
        return torch.cat(out, 0).numpy()

    def propagate(self, args, gx, gy, mask, mask_gradient, flow_f, flow_b):
        # the driver passes the hole mask as mask_RGB and the dilated (gradient) mask as mask
        # (tool/video_inpainting.py:623-633); the oracle propagates on the latter, like the reference
        return prop_oracle.get_flownn_gradient(gx, gy, mask_gradient, flow_f, flow_b, float(args.consistencyThres),
                                               float(args.alpha))

    def poisson_frames(self, trg, gx, gy, hole, gmask):
        return [poisson_oracle.poisson_blend(t, a, b, h, g) for t, a, b, h, g in zip(trg, gx, gy, hole, gmask)]

    def fgt_stage(self, frame_blends, mask, flow_f, step, num_ref, neighbor_stride):
        return clip_oracle.fgt_stage(self.fgt_model, frame_blends, mask, flow_f, step, num_ref, neighbor_stride)
