"""N>1 host logic on CPU: world_size-2 gloo run of the stream partition + aggregate-rate reduction
that bench.py uses (the data path itself has no collective: streams are independent)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from motcpp_b200 import sharding


def test_partition_covers_every_stream_once():
    for n in (1, 7, 64, 65, 296):
        for w in (1, 2, 3, 8):
            parts = [sharding.partition_streams(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            for (b0, e0), (b1, e1) in zip(parts, parts[1:]):
                assert e0 == b1
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1
            for s in (0, n // 2, n - 1):
                r = sharding.owner_of(s, n, w)
                assert parts[r][0] <= s < parts[r][1]
    with pytest.raises(ValueError):
        sharding.partition_streams(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_streams, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = sharding.partition_streams(n_streams, world, rank)
    # stand-in for the per-rank engine: every stream "produces" (stream_id, n_frames) after T frames
    T = 5 + rank
    local = [(s, T) for s in range(b, e)]
    dist.barrier()
    elapsed = 0.5 * (rank + 1)                       # rank 1 is the slow one
    rate = sharding.aggregate_rate(len(local) * T, elapsed)
    allres = sharding.gather_results(local)
    q.put((rank, rate, allres))
    dist.destroy_process_group()


def test_two_rank_gloo_aggregate():
    world, n_streams = 2, 9
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    frames = 5 * 5 + 4 * 6                           # rank 0: 5 streams x 5 frames, rank 1: 4 streams x 6 frames
    for rank, rate, allres in res:
        assert abs(rate - frames / 1.0) < 1e-9       # sum of frames / MAX elapsed (1.0 s on rank 1)
        assert [s for s, _ in allres] == list(range(n_streams))
