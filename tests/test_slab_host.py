"""
Host-side logic of the slab decomposition (no GPU): partitioning, stored ranges,
global->local stimulus boxes, and the neighbour hand-shake over torch.distributed
(gloo, world_size 2, spawned processes on 127.0.0.1).
"""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from finitewave_b200 import slab


@pytest.mark.parametrize("n,world", [(10, 1), (10, 3), (1024, 8), (7, 7), (4097, 4)])
def test_partition_covers_without_overlap(n, world):
    parts = slab.partition(n, world)
    assert parts[0][0] == 0 and parts[-1][1] == n
    for (a, b), (c, d) in zip(parts, parts[1:]):
        assert b == c and b > a
    sizes = [b - a for a, b in parts]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slab.partition(2, 3)


def test_stored_range_and_owned_view():
    n = 20
    parts = slab.partition(n, 3)
    for r, own in enumerate(parts):
        lo, hi, halo = slab.stored_range(own, n)
        assert halo == (r > 0, r < 2)
        assert lo == own[0] - halo[0] and hi == own[1] + halo[1]
        arr = np.arange(lo, hi)
        assert slab.owned_view(arr, halo).tolist() == list(range(*own))


def test_global_box_to_local_matches_numpy_slicing():
    n = 30
    marks = np.arange(n)
    for x1, x2 in [(0, 30), (5, 12), (-4, 100), (12, 12), (25, 3), (None, 7), (9, None)]:
        want = set(marks[x1:x2].tolist())
        got = set()
        for own in slab.partition(n, 4):
            lo, hi, halo = slab.stored_range(own, n)
            a, b = slab.global_box_to_local(x1, x2, n, lo, hi - lo)
            stored = np.arange(lo, hi)[a:b]
            got |= set(stored.tolist()) & set(range(*own))
            # a slab applies the stimulus to its ghost slices too
            assert set(stored.tolist()) == want & set(range(lo, hi))
        assert got == want


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 64
    own = slab.partition(n, world)[rank]
    lo, hi, halo = slab.stored_range(own, n)
    mine = dict(u0=b"u0-%d" % rank, u1=b"u1-%d" % rank, flags=b"f-%d" % rank, slices=hi - lo)
    everyone = slab.exchange_exports(mine, world, dist)
    ok = len(everyone) == world
    if halo[0]:
        ok &= everyone[rank - 1]["u0"] == b"u0-%d" % (rank - 1)
    if halo[1]:
        ok &= everyone[rank + 1]["flags"] == b"f-%d" % (rank + 1)
        ok &= everyone[rank + 1]["slices"] == slab.stored_range(slab.partition(n, world)[rank + 1], n)[1] - \
            slab.stored_range(slab.partition(n, world)[rank + 1], n)[0]
    out.put((rank, bool(ok), halo))
    dist.barrier()
    dist.destroy_process_group()


def test_neighbour_handshake_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, (False, True)), (1, True, (True, False))]
