"""CPU (gloo, world_size 2): the N>1 host logic — window schedule, work sharding, MAX-reduce of step
times and the frame gather — without any GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fgt_b200 import parallel


def test_window_schedule_matches_driver_examples():
    """N=10 -> windows with t=6 and t=10; N=80 -> 16 windows, t in {13,17,18}."""
    w10 = parallel.window_schedule(10)
    assert [len(n) + len(r) for _, n, r in w10] == [6, 10]
    w80 = parallel.window_schedule(80)
    assert len(w80) == 16
    assert set(len(n) + len(r) for _, n, r in w80) == {13, 17, 18}
    f, nb, ref = w80[3]
    assert f == 15 and nb == list(range(10, 21)) and ref == [0, 30, 40, 50, 60, 70]
    assert all(set(nb) & set(ref) == set() for _, nb, ref in w80)


def test_shard_items_partition_and_balance():
    for n, world in ((16, 8), (16, 3), (5, 8), (80, 8)):
        parts = [parallel.shard_items(n, r, world) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(n))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    costs = [13, 17, 18, 18, 18, 18, 18, 18, 18, 18, 18, 18, 18, 18, 17, 13]
    parts = [parallel.shard_items(16, r, 8, costs) for r in range(8)]
    assert sorted(sum(parts, [])) == list(range(16))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sched = parallel.window_schedule(20)  # 4 windows
        mine = parallel.shard_items(len(sched), rank, world)
        # each rank "inpaints" its windows: frame value = frame id (+0.5 noise on rank 1 to test averaging counts)
        ids, frames = [], []
        for wi in mine:
            _, nb, _ = sched[wi]
            ids += nb
            frames.append(torch.tensor(nb, dtype=torch.float32).view(-1, 1, 1, 1).expand(-1, 3, 2, 2).clone())
        acc, cnt = parallel.gather_frames(torch.cat(frames), ids, 20)
        mx = parallel.max_over_ranks(10.0 + rank)
        dist.barrier()
        q.put((rank, mine, acc[:, 0, 0, 0].tolist(), cnt.tolist(), mx))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather_and_max():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sched = parallel.window_schedule(20)
    cover = torch.zeros(20)
    for _, nb, _ in sched:
        cover[nb] += 1
    for rank, mine, acc0, cnt, mx in res:
        assert mine == parallel.shard_items(len(sched), rank, world)
        assert cnt == cover.tolist()                      # every rank sees the full coverage counts
        assert all(abs(a - c * i) < 1e-6 for i, (a, c) in enumerate(zip(acc0, cnt)))  # sum over covering windows
        assert mx == pytest.approx(11.0)                  # MAX over ranks of (10 + rank)
    assert sorted(res[0][1] + res[1][1]) == list(range(len(sched)))


def _zone_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T, Z, rpf, d = 5, 4, 6, 8                       # 5 frames over 2 ranks: 3 + 2 (uneven -> padding path)
        counts = parallel.frame_counts(T, world)
        full = torch.arange(2 * Z * T * rpf * d, dtype=torch.float32).reshape(2, Z, T * rpf, d).to(torch.bfloat16)
        mine = parallel.shard_items(T, rank, world)
        local = torch.zeros(2, Z, max(counts) * rpf, d, dtype=torch.bfloat16)
        local[:, :, :len(mine) * rpf] = full[:, :, mine[0] * rpf:(mine[-1] + 1) * rpf]
        work = {}
        out = parallel.allgather_zone_rows(local, counts, rpf, None, work)
        out2 = parallel.allgather_zone_rows(local, counts, rpf, None, work)   # cached workspace path
        q.put((rank, counts, bool(torch.equal(out, full)), bool(torch.equal(out2, full)), len(work)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_allgather_zone_rows():
    """The exchange step of frame-sharded TMHSA: every rank ends up with all frames' zone rows in frame
    order, bit-exact, also when the frame counts are uneven (3 + 2)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_zone_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, counts, ok, ok2, nwork in res:
        assert counts == [3, 2] and ok and ok2 and nwork == 1


def test_allgather_zone_rows_world1_is_identity():
    x = torch.randn(2, 3, 8, 4).to(torch.bfloat16)
    assert torch.equal(parallel.allgather_zone_rows(x, [2], 4), x)
