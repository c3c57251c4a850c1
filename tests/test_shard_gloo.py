"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous sharding of the instances and the
all-gather of the moves reproduce the unsharded ordering, including ragged shards."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_total, ncol, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import mpc_b200
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = mpc_b200.shard_range(n_total, rank, world)
    full = torch.arange(n_total * ncol, dtype=torch.float64).reshape(n_total, ncol)  # instance-major "Ztilde"
    local = full[lo:hi].clone() * 2.0  # the rank's own results
    got = mpc_b200.gather_moves(local, n_total)
    ok = bool(torch.equal(got, full * 2.0))
    got2 = mpc_b200.gather_moves(local)  # n_total inferred by an all-reduce
    ok = ok and bool(torch.equal(got2, full * 2.0))
    q.put((rank, lo, hi, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_shard_and_gather_two_ranks(n_total):
    world, ncol = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, ncol, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[3] for r in res)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n_total


def test_shard_range_properties():
    import mpc_b200
    for n in (1, 7, 4096, 65537):
        for w in (1, 2, 3, 8):
            r = [mpc_b200.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
