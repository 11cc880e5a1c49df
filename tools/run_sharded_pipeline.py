"""Multi-GPU pipeline check + timing under torchrun (one process per GPU, NCCL): fgt_b200.pipeline.ShardedBackend
around the GPU backend — parity of the final frames with the reference driver's golden (and bit-equality across
ranks), then wall time of a 20-frame 240x432 clip sharded vs on one GPU (rank 0 alone). Prints one JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/run_sharded_pipeline.py
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fgt_b200 import pipeline as PL, synth  # noqa: E402
from tests.util import load_golden  # noqa: E402
from tools.bench_pipeline import models  # noqa: E402  (module-level work there is guarded below)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = {"world": world}
g = load_golden("pipeline_clip")
m = g["meta"]
frames, masks = synth.pipeline_clip(seed=m["clip_seed"], N=m["N"], H=m["H"], W=m["W"])
args = PL.make_args(imgH=m["H"], imgW=m["W"], flow_mask_dilates=m["flow_mask_dilates"], frame_dilates=m["frame_dilates"])
inner, _ = models(m["H"], m["W"], dev=dev)
comp = np.stack(PL.video_inpainting(frames, masks, PL.ShardedBackend(inner), args))
diff = np.abs(comp.astype(np.int16) - g["comp"].astype(np.int16))
t = torch.from_numpy(comp.astype(np.int32)).to(dev)
ref = t.clone()
dist.broadcast(ref, 0)
same = torch.tensor([int(torch.equal(t, ref))], device=dev)
dist.all_reduce(same, op=dist.ReduceOp.MIN)
res["parity_clip_7x64x96"] = dict(max_level_diff=int(diff.max()), mean_level_diff=float(diff.mean()), identical_on_all_ranks=bool(same.item()))

N, H, W = 20, 240, 432
frames, masks = synth.pipeline_clip(seed=5, N=N, H=H, W=W)
args = PL.make_args(imgH=H, imgW=W, flow_mask_dilates=3, frame_dilates=1)
inner, _ = models(H, W, dev=dev)
sh = PL.ShardedBackend(inner)


def timed(backend, reps=2):
    PL.video_inpainting(frames, masks, backend, args)
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        PL.video_inpainting(frames, masks, backend, args)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


dist.barrier()
t_sh = timed(sh)
tt = torch.tensor([t_sh], device=dev, dtype=torch.float64)
dist.all_reduce(tt, op=dist.ReduceOp.MAX)
dist.barrier()
t_one = timed(inner) if rank == 0 else 0.0
dist.barrier()
res["clip_20x240x432"] = dict(seconds_sharded_max_over_ranks=tt.item(), seconds_one_gpu=t_one,
                              note="wall time incl. the host glue every rank repeats (TELEA, morphology, resizes)")
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
