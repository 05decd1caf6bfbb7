"""
CPU-side checks of the measurement plumbing (no GPU): the `config` object both bench arms
emit, the numba reference arm (the unmodified reference package, /root/reference in the build
container or the oracle/_ref copy made by oracle/make_ref.py) and the device selection of
`CardiacModel.run()` on several GPUs.
"""
import importlib.util
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location("fwb_bench", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_both_arms_emit_the_same_config_object():
    b = _bench()
    for w in b.WORKLOADS:
        for world in (1, 2, 8):
            c = b.workload_config(w, world)
            assert set(c) == {"workload", "shape_per_gpu", "stencil_points",
                              "bytes_per_node_update", "parallelism"}
            assert c == b.workload_config(w, world)
    # algorithmic bytes per node-update
    assert b.WORKLOADS["c5"][3] == 8 + 8 + 1 + 8 * 19 + 8 * (2 * 17 + 1 + 1)
    assert b.WORKLOADS["c2"][3] == 8 + 8 + 1 + 8 * 9 + 8 * (2 * 2)


def test_reference_arm_times_the_numba_reference():
    """`bench.py --impl reference` runs the reference's own CardiacModel.run() (numba) on the
    host cores and prints one JSON line with kind == "reference" and the b200 arm's config."""
    from oracle import ref_numba
    if not ref_numba.available():
        pytest.skip("no copy of the reference here (oracle/make_ref.py needs /root/reference)")
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload",
                        "c2", "--steps", "3", "--warmup", "3"], capture_output=True, text=True,
                       timeout=600, cwd=str(ROOT))
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "reference"
    assert line["config"] == _bench().workload_config("c2", 1)
    assert line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "CardiacModel.run(initialize=False)" in line["cpu_baseline"]["sample"]


def test_reference_copy_is_the_unmodified_package():
    ref = Path("/root/reference/finitewave")
    cp = ROOT / "oracle" / "_ref" / "finitewave"
    if not ref.is_dir() or not cp.is_dir():
        pytest.skip("needs both /root/reference and oracle/_ref")
    files = sorted(p.relative_to(ref) for p in ref.rglob("*.py"))
    assert files and files == sorted(p.relative_to(cp) for p in cp.rglob("*.py"))
    for f in files:
        assert (ref / f).read_bytes() == (cp / f).read_bytes(), f


def test_device_selection_for_multi_gpu_runs(monkeypatch):
    from finitewave_b200 import multi

    class M:
        devices = None
        stim_sequence = None
        tracker_sequence = None
    m = M()
    monkeypatch.delenv("FWB_DEVICES", raising=False)
    assert multi.requested_devices(m) is None
    monkeypatch.setenv("FWB_DEVICES", "0,1,3")
    assert multi.requested_devices(m) == [0, 1, 3]
    monkeypatch.setenv("FWB_DEVICES", "2")
    assert multi.requested_devices(m) is None              # one device = the plain engine
    m.devices = [1, 0]
    assert multi.requested_devices(m) == [1, 0]            # the attribute wins
    ok, why = multi.supported(m, (64, 32, 32), [0, 1])
    if not ok:
        assert "visible" in why                            # no GPU in this container
    ok, why = multi.supported(m, (64, 32, 30), [0, 0])
    assert not ok


def test_slab_rows_of_the_multi_engine_cover_the_tissue():
    from finitewave_b200 import slab
    for n, world in ((24, 2), (25, 3), (128, 8)):
        owned = slab.partition(n, world)
        seen = np.zeros(n, dtype=int)
        for r, (a, b) in enumerate(owned):
            lo, hi, halo = slab.stored_range((a, b), n)
            assert halo == (r > 0, r < world - 1)
            assert lo == a - int(halo[0]) and hi == b + int(halo[1])
            seen[a:b] += 1
        assert (seen == 1).all()


# ---------------------------------------------------------------------------
# bench.py's halo_check (N > 1): over gloo, world size 3, with CPU tensors standing in for the
# slabs' device buffers -- consistent slabs pass, one stale ghost slice is reported.
# ---------------------------------------------------------------------------
class _FakeSlab:
    def __init__(self, u, halo):
        self._u, self.halo = u, halo

    def u_device(self):
        return self._u


def _halo_worker(rank, world, port, corrupt, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from finitewave_b200 import slab
    n0 = 17
    g = torch.from_numpy(np.random.default_rng(3).normal(size=(n0, 5, 6)))      # the whole tissue
    lo, hi, halo = slab.stored_range(slab.partition(n0, world)[rank], n0)
    u = g[lo:hi].clone()
    if corrupt and rank == 1:
        u[0, 2, 3] += 1e-13                     # rank 1's lower ghost slice is stale by one ulp-ish
    res = _bench().halo_check(_FakeSlab(u, halo), rank, world, dist, torch.device("cpu"))
    out.put((rank, res["ok"], res["interfaces"], res["mismatches"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("corrupt", [False, True])
def test_halo_check_over_gloo_world3(corrupt):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_halo_worker, args=(r, 3, port, corrupt, out)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (not corrupt, 2, 1 if corrupt else 0)
    assert [r[1:] for r in res] == [want] * 3, res
