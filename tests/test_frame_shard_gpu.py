"""Frame-sharded FGT forward (TMHSA is the one exchange step) against the unsharded forward.

Two processes share cuda:0 and talk over gloo (NCCL refuses two ranks on one device; the helper stages
the all-gather through the host for gloo), so the full sharded code path — Q for own frames, K/V for all
frames from the all-gathered LayerNorm rows, uneven 3+2 split — runs on the single-GPU test box, for both
exchanges: the torch.distributed all-gather and the fused P2P one (LayerNorm storing into the peers'
buffers through CUDA IPC mappings + the device-side barrier; two processes on one GPU are time-sliced, so
the barrier's spin simply waits for the other process's slice). With two or more GPUs the same test also
runs over NCCL / NVLink, one rank per device."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

T, H, W = 5, 72, 100   # 72x100: temporal zones need padding (token grid 6x9 -> zones of 3x5 with pad rows)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, backend, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    from fgt_b200 import parallel, synth
    from fgt_b200.fgt_model import Model
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        cfg = dict(synth.CFG_A)
        cfg["input_resolution"] = (H, W)
        sd = synth.make_state_dict(synth.fgt_param_shapes(cfg), seed=12)
        model = Model(cfg)
        model.load_state_dict(sd)
        model = model.to(dev)
        fr, fl, mk = [t.to(dev) for t in synth.fgt_inputs(seed=31, t=T, H=H, W=W)]
        with torch.no_grad():
            full = model(fr, fl, mk)                                  # unsharded reference on every rank
            mine = parallel.shard_items(T, rank, world)
            sl = slice(mine[0], mine[-1] + 1)
            errs, same = {}, True
            for exchange in ("nccl", "p2p"):   # "nccl" = torch.distributed all-gather (gloo here when backend is gloo)
                model.net.enable_frame_sharding(T, exchange=exchange)
                part = model(fr[:, sl], fl[:, sl], mk[:, sl])
                part2 = model(fr[:, sl], fl[:, sl], mk[:, sl])        # cached workspaces / second call
                errs[exchange] = (part - full[sl]).abs().max().item() / full.abs().max().item()
                same = same and bool(torch.equal(part, part2))
            # the P2P exchange is kernels only -> the whole sharded forward replays as a CUDA graph
            model.net.enable_cuda_graph(True)
            for _ in range(3):
                gpart = model(fr[:, sl], fl[:, sl], mk[:, sl])
            same = same and bool(torch.equal(gpart, part))
            model.net.enable_cuda_graph(False)
            model.net.enable_frame_sharding(None)
            again = model(fr, fl, mk)
        q.put((rank, mine, tuple(part.shape), errs, same, bool(torch.equal(again, full))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("backend", ["gloo", "nccl"])
def test_frame_sharded_forward_matches_unsharded(backend):
    world = 2
    if backend == "nccl" and torch.cuda.device_count() < 2:
        pytest.skip("NCCL variant needs two GPUs (the gloo variant covers the same code path on one)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [[0, 1, 2], [3, 4]]
    for rank, mine, shape, errs, same, restored in res:
        assert shape == (len(mine), 3, H, W)
        # same kernels and the same per-row arithmetic; only tile boundaries of the split problem differ
        for exchange, err in errs.items():
            assert err < 1e-5, f"rank {rank} ({exchange}): sharded vs unsharded max/max = {err:.3e}"
        assert same and restored
