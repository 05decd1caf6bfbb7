"""
Parity at BASELINE.json's FULL sizes through size-independent properties (the CPU
oracle cannot run 512^3 ... 1e8-node tissues in test time):

  * locality   -- an explicit stencil moves information one node per step, so after s
                  steps a window further than s nodes from every edge of a sub-domain
                  equals the same window of a small oracle run on that sub-domain;
  * invariance -- a tissue and stimulus that do not depend on one axis give a solution
                  that does not depend on it (bit-identical lines away from the ends);
  * determinism.
The tissues come from finitewave_b200.workloads (the bench's device-side builders).
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cfg(m, dt=0.01, dr=0.25):
    m.dt, m.dr, m.prog_bar = dt, dr, False
    return m


def test_c3_512cube_centre_equals_small_oracle_run():
    """C3: Mitchell-Schaeffer 3D 512^3 iso-7, focal stimulus, activation tracker."""
    import torch
    import finitewave_b200 as fw
    from finitewave_b200 import workloads
    from finitewave_b200.devrun import DeviceSimulation
    from oracle import oracle

    n, steps, R = 512, 40, 20
    dev = torch.device("cuda")
    sim = DeviceSimulation(_cfg(fw.MitchellSchaeffer3D()), workloads.fibrosis_mesh((n, n, n), 0.0, 0, dev))
    c = n // 2
    sim.add_stim(fw.StimVoltageCoord3D(0, 1, c - 5, c + 5, c - 5, c + 5, c - 5, c + 5))
    tr = fw.ActivationTime3DTracker()
    tr.threshold, tr.step = 0.5, 1
    sim.add_tracker(tr, 0)
    assert sim.n_myo == 510 ** 3
    sim.run(steps)
    sim.collect()
    win = (slice(c - R, c + R),) * 3
    u_big = sim.u_device()[win].cpu().numpy()
    h_big = torch.from_numpy(sim.state_host("h"))[win].numpy() if False else None
    act_big = tr._dev[win].cpu().numpy()

    m = 128
    cs = m // 2
    case = dict(model="mitchell_schaeffer", shape=[m, m, m], dt=0.01, dr=0.25, t_max=steps * 0.01 - 0.005,
                stims=[dict(kind="voltage_coord", t=0, value=1,
                            box=[cs - 5, cs + 5, cs - 5, cs + 5, cs - 5, cs + 5])],
                trackers=[dict(kind="activation_time", threshold=0.5, step=1)])
    ref = oracle.simulate(case)
    assert int(ref["step"]) == steps
    wref = (slice(cs - R, cs + R),) * 3
    assert np.array_equal(u_big, ref["u"][wref])          # MS has no libm call: bit-exact
    assert np.array_equal(act_big, ref["tracker0"][wref])
    assert u_big.max() > 0.5 and h_big is None


def test_c2_4096sq_window_equals_small_oracle_run():
    """C2: Fenton-Karma 2D 4096^2, 9-point anisotropic, 30 % random fibrosis."""
    import torch
    import finitewave_b200 as fw
    from finitewave_b200 import workloads
    from finitewave_b200.devrun import DeviceSimulation
    from oracle import oracle

    n, steps, m = 4096, 40, 160
    dev = torch.device("cuda")
    mesh = workloads.fibrosis_mesh((n, n), 0.30, 2, dev)
    sim = DeviceSimulation(_cfg(fw.FentonKarma2D()), mesh,
                           fibers=workloads.uniform_fibers_2d((n, n), 0.25 * math.pi, dev))
    o = 2000                                        # sub-domain [o, o+m)^2
    sim.add_stim(fw.StimVoltageCoord2D(0, 1, o + 70, o + 90, o + 70, o + 90))
    frac = sim.n_myo / (n - 2) ** 2
    assert 0.69 < frac < 0.71
    sim.run(steps)
    sub = mesh[o:o + m, o:o + m].cpu().numpy().copy()
    f = np.empty((m, m, 2))
    f[..., 0], f[..., 1] = math.cos(0.25 * math.pi), math.sin(0.25 * math.pi)
    case = dict(model="fenton_karma", shape=[m, m], dt=0.01, dr=0.25, t_max=steps * 0.01 - 0.005,
                mesh=sub, fibers=f,
                stims=[dict(kind="voltage_coord", t=0, value=1, box=[70, 90, 70, 90])])
    ref = oracle.simulate(case)
    k = steps + 2                                   # the sub-domain's ring + s steps of influence
    win = (slice(k, m - k),) * 2
    big = (slice(o + k, o + m - k),) * 2
    u = sim.u_device()[big].cpu().numpy()
    assert np.max(np.abs(u - ref["u"][win])) <= 1e-12 * np.max(np.abs(ref["u"]))
    for name in ("v", "w"):
        got = sim.state_host(name)[big]
        assert np.max(np.abs(got - ref[name][win])) <= 1e-12
    assert u.max() > 0.3


def test_c5_tp06_slab_invariance_and_oracle_line():
    """C5 (one GPU's 128 x 1024 x 1024 share): TP06, 19-point, fibres rotating along axis 0
    (the slab axis), stimulus on the axis-0 face.  Nothing depends on j or k, so interior
    lines along axis 0 are bit-identical; one of them is compared, after 240 steps, with an
    oracle run on a 128 x 12 x 32 strip with the same fibre rotation."""
    import torch
    import finitewave_b200 as fw
    from finitewave_b200 import workloads
    from finitewave_b200.devrun import DeviceSimulation
    from oracle import oracle

    steps = 240
    dev = torch.device("cuda")
    shape = (128, 1024, 1024)
    sim = DeviceSimulation(_cfg(fw.TP063D()), workloads.fibrosis_mesh(shape, 0.0, 0, dev),
                           fibers=workloads.rotating_fibers_3d(shape, dev))
    sim.add_stim(fw.StimVoltageCoord3D(0, -20, 0, 5, 0, 1024, 0, 1024))
    sim.run(steps)
    u = sim.u_device()
    a, b = u[:, 300, 400].cpu().numpy(), u[:, 700, 520].cpu().numpy()
    assert np.array_equal(a, b)
    # fibres have no component along axis 0, so the lateral fluxes of a laterally uniform
    # potential vanish at the strip's lateral (no-flux) boundaries exactly as in the interior
    # of the full tissue: the strip may be narrow however many steps are taken
    small = (128, 12, 32)
    phi = np.linspace(-np.pi / 3, np.pi / 2, small[0] - 2)
    f = np.zeros((*small, 3))
    f[1:-1, :, :, 1] = np.cos(phi)[:, None, None]
    f[1:-1, :, :, 2] = np.sin(phi)[:, None, None]
    case = dict(model="tp06", shape=list(small), dt=0.01, dr=0.25, t_max=steps * 0.01 - 0.005,
                fibers=f, stims=[dict(kind="voltage_coord", t=0, value=-20,
                                      box=[0, 5, 0, small[1], 0, small[2]])])
    ref = oracle.simulate(case)
    got, want = a, ref["u"][:, 6, 16]
    assert np.max(np.abs(got - want)) <= 1e-9 * np.max(np.abs(want))
    for name in ("m", "cass", "Ki"):
        s = sim.state_host(name)[:, 300, 400]
        w = ref[name][:, 6, 16]
        assert np.max(np.abs(s - w)) <= 1e-9 * max(np.max(np.abs(w)), 1e-300), name
    assert np.ptp(got) > 20


def test_c4_ventricle_window_equals_small_oracle_run():
    """C4: TP06 on the ventricle-shaped shell in a 512^3 box (helix fibres, 19-point stencil;
    the tissue fills its tiles poorly, so the tile kernel runs in PACKED mode).  A window at
    the apex -- where the stimulus sits -- equals the same window of an oracle run on a 72^3
    sub-box cut out of the same mesh and fibre field (locality: 24 steps, 26 nodes of margin)."""
    import torch
    import finitewave_b200 as fw
    from finitewave_b200 import workloads
    from finitewave_b200.devrun import DeviceSimulation
    from oracle import oracle

    n, steps, m, marg = 512, 24, 72, 26
    dev = torch.device("cuda")
    mesh, fib = workloads.ventricle_shell((n, n, n), dev)
    sim = DeviceSimulation(_cfg(fw.TP063D()), mesh, fibers=fib)
    sim.add_stim(fw.StimVoltageCoord3D(0, -20, 0, n, 0, n, 0, 40))   # its edge crosses the window
    assert sim.n_myo / (sim.engine.n_work * 32) < 0.85          # -> packed mode
    sim.run(steps)
    o = (220, 220, 8)
    box = tuple(slice(a, a + m) for a in o)
    sub_mesh = mesh[box].cpu().numpy().copy()
    sub_fib = fib[box].cpu().numpy().copy()
    del fib
    case = dict(model="tp06", shape=[m, m, m], dt=0.01, dr=0.25, t_max=steps * 0.01 - 0.005,
                mesh=sub_mesh, fibers=sub_fib,
                stims=[dict(kind="voltage_coord", t=0, value=-20, box=[0, m, 0, m, 0, 40 - o[2]])])
    ref = oracle.simulate(case)
    assert int(ref["step"]) == steps
    win = (slice(marg, m - marg),) * 3
    big = tuple(slice(a + marg, a + m - marg) for a in o)
    tissue = sub_mesh[win] == 1
    assert tissue.sum() > 1000
    u = sim.u_device()[big].cpu().numpy()
    assert np.max(np.abs(u - ref["u"][win])) <= 1e-9 * np.max(np.abs(ref["u"]))
    for name in ("m", "h", "cass", "xs", "Ki"):
        got = sim.state_host(name)[big]
        want = ref[name][win]
        assert np.max(np.abs(got - want)) <= 1e-9 * max(np.max(np.abs(want)), 1e-300), name
    assert np.ptp(u[tissue]) > 10                                # the wave is inside the window


def test_full_size_run_is_deterministic():
    import torch
    from finitewave_b200 import workloads
    dev = torch.device("cuda")
    outs = []
    for _ in range(2):
        sim, info = workloads.build("c2", dev)
        sim.run(60)
        outs.append(sim.u_device().clone())
        del sim
        torch.cuda.empty_cache()
    assert torch.equal(outs[0], outs[1])
