"""N>1 host logic on CPU: stream sharding + the optional result gather over gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest

from avdsp_b200.sharding import shard_range


def test_shard_range_partitions_exactly():
    for n in (1, 2, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            got = [shard_range(n, r, world) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == n
            for (f0, c0), (f1, _) in zip(got, got[1:]):
                assert f1 == f0 + c0
            assert max(c for _, c in got) - min(c for _, c in got) <= 1


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from avdsp_b200.sharding import gather_outputs, shard_range as sr
    n = 5
    first, cnt = sr(n, rank, world)
    full = torch.arange(n * 3 * 2, dtype=torch.int32).reshape(n, 3, 2)
    got = gather_outputs(full[first:first + cnt].clone(), n)
    q.put((rank, bool(torch.equal(got, full))))
    dist.destroy_process_group()


def test_gather_outputs_gloo_world2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
