"""
Two processes, two GPUs: slabs connected through CUDA IPC peer mappings (the real
multi-GPU path: in-kernel peer stores over NVLink + flags) must reproduce the
single-GPU undivided run bit for bit.  Skipped on boxes with one GPU -- the same
kernels and protocol are covered there by tests/test_gpu_slab.py.
"""
import os
import socket

import numpy as np
import pytest

from tests.cases import random_fibers, random_fibrosis

pytestmark = pytest.mark.gpu
SHAPE = (64, 16, 64)
STEPS = 150


def _inputs():
    from oracle import oracle
    mesh = oracle.apply_boundaries(random_fibrosis(SHAPE, 0.2, 41))
    return mesh, random_fibers(SHAPE, 42)


def _sim(fw, mesh, fibers, **kw):
    from finitewave_b200.devrun import DeviceSimulation
    m = fw.MitchellSchaeffer3D()
    m.dt, m.dr, m.prog_bar = 0.01, 0.25, False
    s = DeviceSimulation(m, mesh, fibers=fibers, **kw)
    s.add_stim(fw.StimVoltageCoord3D(0, 1.0, 0, SHAPE[0], 0, SHAPE[1], 0, 5))
    s.add_stim(fw.StimCurrentCoord3D(0.4, 4.0, 0.3, 20, 44, 2, 14, 30, 40))
    return s


def _worker(rank, world, port, outdir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device(f"cuda:{rank}"))
    import finitewave_b200 as fw
    from finitewave_b200 import slab
    mesh, fibers = _inputs()
    a, b = slab.partition(SHAPE[0], world)[rank]
    lo, hi, halo = slab.stored_range((a, b), SHAPE[0])
    s = _sim(fw, mesh[lo:hi].copy(), fibers[lo:hi].copy(), halo=halo, slow_offset=lo,
             global_slices=SHAPE[0])
    peers = slab.connect_distributed(s, rank, world, dist)
    for n in (1, 49, 100):                      # several run() calls, all ranks in lockstep
        s.run(n)
    s.synchronize()
    dist.barrier()
    np.save(os.path.join(outdir, f"u_{rank}.npy"),
            slab.owned_view(s.u_device(), s.halo).cpu().numpy())
    np.save(os.path.join(outdir, f"h_{rank}.npy"),
            slab.owned_view(torch.from_numpy(s.state_host("h")), s.halo).numpy())
    dist.barrier()
    slab.close_peers(peers)
    dist.destroy_process_group()


def test_two_gpu_slabs_match_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs on the box")
    import torch.multiprocessing as mp
    import finitewave_b200 as fw
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)

    mesh, fibers = _inputs()
    full = _sim(fw, mesh, fibers)
    full.run(STEPS)
    u_ref, h_ref = full.u_host(), full.state_host("h")
    u = np.concatenate([np.load(tmp_path / f"u_{r}.npy") for r in range(2)])
    h = np.concatenate([np.load(tmp_path / f"h_{r}.npy") for r in range(2)])
    assert np.array_equal(u, u_ref)
    assert np.array_equal(h, h_ref)
    assert np.ptp(u_ref) > 0.5
