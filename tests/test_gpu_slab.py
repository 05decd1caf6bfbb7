"""
Slab decomposition on ONE GPU: a tissue is cut into 2-3 slabs that live in the same
process and exchange halos through the same in-kernel peer stores + flags as the
multi-GPU path (raw pointers instead of IPC mappings).  Stepping the slabs in
lockstep must reproduce the undivided run bit for bit (same per-node arithmetic),
including stimuli that straddle a cut and the activation-time tracker.
"""
import numpy as np
import pytest

from tests.cases import random_fibers, random_fibrosis

pytestmark = pytest.mark.gpu


def _model(fw, name, dim):
    cls = {"fenton_karma": "FentonKarma", "mitchell_schaeffer": "MitchellSchaeffer",
           "tp06": "TP06", "aliev_panfilov": "AlievPanfilov"}[name]
    m = getattr(fw, f"{cls}{dim}D")()
    m.dt, m.dr, m.prog_bar = 0.01, 0.25, False
    return m


def _stims(fw, dim, shape, value):
    if dim == 2:
        return [fw.StimVoltageCoord2D(0, value, 0, shape[0], 0, 5),
                fw.StimCurrentCoord2D(0.3, 3 * value, 0.2, shape[0] // 3, 2 * shape[0] // 3, 8, 20)]
    return [fw.StimVoltageCoord3D(0, value, 0, shape[0], 0, shape[1], 0, 4),
            fw.StimCurrentCoord3D(0.3, 3 * value, 0.2, shape[0] // 3, 2 * shape[0] // 3,
                                  2, 9, 5, 20)]


@pytest.mark.parametrize("model,shape,world,aniso", [
    ("fenton_karma", (48, 64), 2, True),
    ("aliev_panfilov", (41, 96), 3, False),
    ("mitchell_schaeffer", (23, 12, 32), 3, True),
    ("tp06", (16, 8, 32), 2, True),
], ids=str)
def test_slabs_equal_undivided_run(model, shape, world, aniso):
    import torch
    import finitewave_b200 as fw
    from finitewave_b200 import slab
    from finitewave_b200.devrun import DeviceSimulation
    from oracle import oracle

    dim = len(shape)
    mesh = oracle.apply_boundaries(random_fibrosis(shape, 0.2, 31))
    fibers = random_fibers(shape, 32) if aniso else None
    value = -20.0 if model == "tp06" else 1.0
    n_steps = 120

    def tracker():
        tr = fw.ActivationTime3DTracker() if dim == 3 else fw.ActivationTime2DTracker()
        tr.threshold = -40 if model == "tp06" else 0.5
        tr.step = 3
        return tr

    full = DeviceSimulation(_model(fw, model, dim), mesh, fibers=fibers)
    for st in _stims(fw, dim, shape, value):
        full.add_stim(st)
    full_tr = full.add_tracker(tracker(), 0)
    full.run(n_steps)
    full.collect()
    u_full = full.u_host()

    parts = slab.partition(shape[0], world)
    sims, trs = [], []
    for (a, b) in parts:
        lo, hi, halo = slab.stored_range((a, b), shape[0])
        s = DeviceSimulation(_model(fw, model, dim), mesh[lo:hi].copy(),
                             fibers=None if fibers is None else fibers[lo:hi].copy(),
                             halo=halo, slow_offset=lo, global_slices=shape[0])
        for st in _stims(fw, dim, shape, value):
            s.add_stim(st)
        trs.append(s.add_tracker(tracker(), 0))
        sims.append(s)
    slab.connect_in_process(sims)
    for _ in range(n_steps):
        for s in sims:
            s.run(1, halo_sync=False)       # one stream: program order is the sync
    torch.cuda.synchronize()
    assert sum(s.n_myo for s in sims) == full.n_myo
    for (a, b), s, tr in zip(parts, sims, trs):
        s.collect()
        u = slab.owned_view(s.u_device(), s.halo).cpu().numpy()
        assert np.array_equal(u, u_full[a:b]), f"u differs on slab [{a},{b})"
        for name in s.state_names:
            got = slab.owned_view(torch.from_numpy(s.state_host(name)), s.halo).numpy()
            assert np.array_equal(got, full.state_host(name)[a:b]), name
        act = slab.owned_view(torch.from_numpy(tr.act_t), s.halo).numpy()
        assert np.array_equal(act, full_tr.act_t[a:b])
    assert (u_full != u_full.flat[0]).any(), "nothing propagated"


def test_slab_ecg_sums_to_undivided_trace():
    import torch
    import finitewave_b200 as fw
    from finitewave_b200 import slab
    from finitewave_b200.devrun import DeviceSimulation
    from oracle import oracle

    shape = (30, 10, 32)
    mesh = oracle.apply_boundaries(random_fibrosis(shape, 0.1, 5))
    coords = np.array([[15.0, 5.0, 40.0], [-3.0, 2.0, 7.0]])

    def ecg():
        tr = fw.ECG3DTracker()
        tr.measure_coords = coords
        tr.step = 4
        return tr

    def stim():
        return fw.StimVoltageCoord3D(0, 1.0, 0, 30, 0, 10, 0, 4)

    full = DeviceSimulation(_model(fw, "mitchell_schaeffer", 3), mesh)
    full.add_stim(stim())
    ftr = full.add_tracker(ecg(), 40)
    full.run(100)
    full.collect()
    ref = np.array(ftr.output)

    sims, trs = [], []
    for (a, b) in slab.partition(shape[0], 2):
        lo, hi, halo = slab.stored_range((a, b), shape[0])
        s = DeviceSimulation(_model(fw, "mitchell_schaeffer", 3), mesh[lo:hi].copy(), halo=halo,
                             slow_offset=lo, global_slices=shape[0])
        s.add_stim(stim())
        trs.append(s.add_tracker(ecg(), 40))
        sims.append(s)
    slab.connect_in_process(sims)
    for _ in range(100):
        for s in sims:
            s.run(1, halo_sync=False)
    torch.cuda.synchronize()
    total = 0
    for s, tr in zip(sims, trs):
        s.collect()
        total = total + np.array(tr.output)
    assert total.shape == ref.shape == (25, 2)
    assert np.max(np.abs(total - ref)) <= 1e-12 * np.max(np.abs(ref))


def test_worklist_covers_every_tissue_chunk_once():
    from finitewave_b200.engine import Engine
    from oracle import oracle
    for shape, halo in (((9, 11, 64), (False, False)), ((12, 7, 32), (True, True)),
                        ((40, 96), (True, False)), ((13, 37), (False, False))):
        mesh = oracle.apply_boundaries(random_fibrosis(shape, 0.5, 3))
        if halo[0]:
            mesh[0] = mesh[1]
        if halo[1]:
            mesh[-1] = mesh[-2]
        eng = Engine(shape)
        eng.set_tissue(mesh, halo=halo)
        wl = eng.worklist[:eng.n_work].cpu().numpy()
        ids = wl[wl >= 0]
        assert len(ids) == len(set(ids.tolist()))
        tissue = eng.tissue.cpu().numpy().copy()
        if halo[0]:
            tissue[0] = 0
        if halo[1]:
            tissue[-1] = 0
        flat = tissue.ravel()
        pad = (-len(flat)) % 32
        chunks = np.flatnonzero(np.pad(flat, (0, pad)).reshape(-1, 32).any(axis=1))
        assert sorted(ids.tolist()) == chunks.tolist()
        assert eng.n_work % 8 == 0
        if any(halo):
            slice_chunks = int(np.prod(shape[1:])) // 32
            n_lo, n_hi = eng.halo_blocks
            lo_ids = wl[:n_lo * 8]
            assert all(i < 0 or slice_chunks <= i < 2 * slice_chunks for i in lo_ids)
            hi_ids = wl[n_lo * 8:(n_lo + n_hi) * 8]
            last = shape[0] - 2
            assert all(i < 0 or last * slice_chunks <= i < (last + 1) * slice_chunks for i in hi_ids)
