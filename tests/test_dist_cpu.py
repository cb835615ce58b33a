"""CPU (gloo, world_size 2): host-side logic of the multi-GPU ingest - frame-batch sharding,
the packed all-gather layout and the deterministic rank-order merge."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from holoagent_b200.ingest import packed_layout, shard_batches


def test_sharding_covers_every_frame_once():
    for (F, FB, world) in [(100, 32, 1), (100, 32, 2), (10000, 32, 8), (33, 8, 4), (5, 8, 2)]:
        seen = []
        for r in range(world):
            allb, mine = shard_batches(F, FB, world, r)
            seen += [f for (b0, n) in mine for f in range(b0, b0 + n)]
            assert sum(n for _, n in allb) == F
        assert sorted(seen) == list(range(F))
        part, stride = packed_layout(1000, 512, F, FB, world, 32)
        nmax = max(sum(n for _, n in shard_batches(F, FB, world, r)[1]) for r in range(world))
        assert part == 1000 * 513 and stride >= part + nmax * 32 * 512


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_nodes, d, F, FB, M = 50, 128, 40, 8, 3
    _, mine = shard_batches(F, FB, world, rank)
    n_local = sum(n for _, n in mine)
    part, stride = packed_layout(n_nodes, d, F, FB, world, M)
    g = torch.Generator().manual_seed(100 + rank)
    send = torch.zeros(stride)
    send[:part] = torch.rand(part, generator=g)
    send[part:part + n_local * M * d] = float(rank + 1)
    bufs = [torch.empty(stride) for _ in range(world)]
    dist.all_gather(bufs, send)
    gathered = torch.stack(bufs)
    merged = torch.zeros(part)
    for r in range(world):          # rank order, like k_merge_partials
        merged += gathered[r, :part]
    q.put((rank, merged.numpy(), [float(gathered[r, part]) for r in range(world)], n_local))
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_merge_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank ends with the same merged partial and sees every rank's F_p block
    assert np.array_equal(res[0][1], res[1][1])
    assert res[0][2] == [1.0, 2.0] and res[1][2] == [1.0, 2.0]
    assert res[0][3] + res[1][3] == 40
    exp = torch.zeros(50 * 129)
    for r in range(2):
        exp += torch.rand(50 * 129, generator=torch.Generator().manual_seed(100 + r))
    assert np.array_equal(res[0][1], exp.numpy())
