"""CPU, world_size 2, gloo: the N>1 host logic (block partition + the single
all-gather of per-eta results) gives the 1-rank answer."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from scintools_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    etas = np.logspace(-1, 1, n)
    calls = []

    def fake_sweep(block):
        calls.append(len(block))
        out = np.sqrt(block) * 3.0
        out[block > 9.0] = np.nan        # NaN slots survive the gather
        return out

    full = sharding.sharded_eta_sweep(fake_sweep, etas)
    seeds = sharding.sharded_items(list(range(7)))
    rows = sharding.sharded_map(lambda i: np.arange(6) + 10.0 * i, list(range(5)), 6)
    q.put((rank, full, calls, seeds, rows))
    dist.barrier()
    dist.destroy_process_group()


def test_block_range_covers_everything():
    for n in (0, 1, 7, 1024, 1025):
        for world in (1, 2, 3, 8):
            got = []
            for r in range(world):
                lo, hi = sharding.block_range(n, r, world)
                got += list(range(lo, hi))
            assert got == list(range(n))


def test_sharded_sweep_world2_matches_single():
    world, n = 2, 101
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    etas = np.logspace(-1, 1, n)
    want = np.sqrt(etas) * 3.0
    want[etas > 9.0] = np.nan
    seeds_all = []
    want_rows = np.arange(6)[None, :] + 10.0 * np.arange(5)[:, None]
    for rank, full, calls, seeds, rows in sorted(res, key=lambda t: t[0]):
        assert np.array_equal(rows, want_rows)       # retrieval-chunk style gather: 3 + 2 rows
        assert np.array_equal(np.isnan(full), np.isnan(want))
        assert np.allclose(full[~np.isnan(want)], want[~np.isnan(want)], rtol=0, atol=0)
        assert calls == [51] if rank == 0 else calls == [50]
        seeds_all += seeds
    assert seeds_all == list(range(7))
