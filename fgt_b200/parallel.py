"""Host-side sharding logic for multi-GPU inference.

Two ways to spread the FGT stage over ranks:

* window data parallelism — the driver's loop over clip windows (tool/video_inpainting.py:710-740): every
  window is an independent Model.forward, so windows shard across ranks with NO data-path collective; the
  only communication is the final gather of the composed frames (or, in bench.py, the MAX-reduce of the
  step time). RAFT pairs and LAFC calls shard the same way (independent per frame pair / per index).
* frame sharding inside one window (`FGT.enable_frame_sharding`) — everything in FGT.forward is per frame
  except TMHSA, whose (zone, head) attention spans all T frames (attention_base.py:76-106). Each rank keeps
  its contiguous frames; per temporal layer the LayerNorm'd zone rows are all-gathered (`allgather_zone_rows`,
  T*n*512*4 B in total), K / V are projected for all frames locally and Q only for the rank's own frames.

`get_flowNN_gradient` is sequential over frames: replicas only.
"""
import torch
import torch.distributed as dist


def window_schedule(video_length, neighbor_stride=5, ref_length=10, num_ref=-1):
    """The driver's window list: [(f, neighbor_ids, ref_ids)] — restates tool/video_inpainting.py:710-717
    and get_ref_index (:103-117)."""
    out = []
    for f in range(0, video_length, neighbor_stride):
        neighbor_ids = list(range(max(0, f - neighbor_stride), min(video_length, f + neighbor_stride + 1)))
        ref_ids = []
        if num_ref == -1:
            ref_ids = [i for i in range(0, video_length, ref_length) if i not in neighbor_ids]
        else:
            start = max(0, f - ref_length * (num_ref // 2))
            end = min(video_length, f + ref_length * (num_ref // 2))
            for i in range(start, end + 1, ref_length):
                if i not in neighbor_ids:
                    if len(ref_ids) > num_ref:
                        break
                    ref_ids.append(i)
        out.append((f, neighbor_ids, ref_ids))
    return out


def shard_items(n_items, rank, world, costs=None):
    """Indices of the work items this rank owns. Without costs: contiguous, sizes differ by <= 1.
    With costs (e.g. frames per window): greedy longest-processing-time balancing, deterministic."""
    if costs is None:
        base, extra = divmod(n_items, world)
        start = rank * base + min(rank, extra)
        return list(range(start, start + base + (1 if rank < extra else 0)))
    order = sorted(range(n_items), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    owner = [0] * n_items
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        owner[i] = r
        load[r] += costs[i]
    return [i for i in range(n_items) if owner[i] == rank]


def max_over_ranks(value, device="cpu"):
    """MAX-reduce of a scalar (the step time every rank measured on its own device)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def gather_frames(local_frames, local_ids, total, device="cpu"):
    """All ranks contribute their inpainted frames [k, 3, H, W] with global frame ids; every rank gets
    the dense [total, 3, H, W] stack and a coverage count (frames covered by two windows are averaged
    0.5/0.5 by the driver, video_inpainting.py:736-740 — callers combine using the counts)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    shape = local_frames.shape[1:]
    acc = torch.zeros((total,) + tuple(shape), dtype=local_frames.dtype, device=device)
    cnt = torch.zeros(total, dtype=torch.float32, device=device)
    if len(local_ids):
        idx = torch.as_tensor(local_ids, device=device)
        acc.index_add_(0, idx, local_frames.to(device))
        cnt.index_add_(0, idx, torch.ones(len(local_ids), device=device))
    if world > 1:
        dist.all_reduce(acc)
        dist.all_reduce(cnt)
    return acc, cnt


def frame_counts(total_frames, world):
    """Frames per rank for contiguous frame sharding (sizes differ by at most one)."""
    return [len(shard_items(total_frames, r, world)) for r in range(world)]


def allgather_zone_rows(local, counts, rows_per_frame, group=None, work=None):
    """The exchange step of frame-sharded TMHSA.

    local : [P, Z, tmax*rows_per_frame, d] — this rank's LayerNorm'd tokens in zone-major order (P = 2 bf16
            planes of the split format, Z zones), frame-major inside a zone and zero-padded to
            tmax = max(counts) frames so that every rank contributes the same number of bytes.
    counts: frames owned by each rank (rank order = frame order).
    Returns [P, Z, sum(counts)*rows_per_frame, d]: every zone's rows of ALL frames in frame order.
    Rows travel as raw bytes (exact on every backend); with the gloo backend CUDA tensors are staged
    through the host (used by the single-GPU two-process parity test), with NCCL they stay on the device.
    """
    world = len(counts)
    P, Z, lr, d = local.shape
    tmax, T = max(counts), sum(counts)
    assert lr == tmax * rows_per_frame, (lr, tmax, rows_per_frame)
    key = ("zone_rows", tuple(local.shape), local.dtype, str(local.device), T)
    if work is not None and key in work:
        gathered, out = work[key]
    else:
        gathered = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
        out = torch.empty((P, Z, T * rows_per_frame, d), dtype=local.dtype, device=local.device)
        if work is not None:
            work[key] = (gathered, out)
    if world == 1:
        gathered[0].copy_(local)
    else:
        as_int = lambda t: t.view(torch.uint8)  # noqa: E731  (raw bytes: exact on every backend)
        src = as_int(local.contiguous())
        if dist.get_backend(group) == "gloo" and local.is_cuda:
            host = [torch.empty(src.shape, dtype=src.dtype) for _ in range(world)]
            dist.all_gather(host, src.cpu(), group=group)
            for q in range(world):
                as_int(gathered[q]).copy_(host[q])
        else:
            dist.all_gather_into_tensor(as_int(gathered).view(world * P, Z, lr, -1), src, group=group)
    off = 0
    for q, cnt in enumerate(counts):
        out[:, :, off * rows_per_frame:(off + cnt) * rows_per_frame].copy_(gathered[q][:, :, :cnt * rows_per_frame])
        off += cnt
    return out
