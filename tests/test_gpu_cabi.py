"""
GPU tests of the individual C-ABI entry points against the CPU oracle:
index structures, dense<->compact conversions, the four stencil-weight kernels
(bit-exact), the diffusion apply (bit-exact, plus exact linearity), stimuli and
trackers edge cases (empty ROI, all-fibrotic chunks, ragged sizes).
"""
import ctypes

import numpy as np
import pytest

from tests.cases import random_fibers, random_fibrosis

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    from finitewave_b200 import _lib
    from finitewave_b200.engine import Engine
    return dict(torch=torch, L=_lib.lib(), Engine=Engine, lib=_lib)


SHAPES = [(7, 9), (33, 31), (40, 64), (5, 6, 7), (12, 9, 32), (9, 10, 65)]


@pytest.mark.parametrize("shape", SHAPES, ids=str)
def test_chunks_and_compaction_roundtrip(env, shape):
    from oracle import oracle
    torch = env["torch"]
    mesh = oracle.apply_boundaries(random_fibrosis(shape, 0.3, 5))
    eng = env["Engine"](shape)
    eng.set_tissue(mesh)
    idx = np.flatnonzero(mesh == 1)
    assert eng.n_myo == len(idx)
    flat = np.arange(mesh.size)
    cidx = eng.compact_index(flat)
    assert np.array_equal(np.flatnonzero(cidx >= 0), idx)
    # compact indices are a permutation of 0..n_myo-1 (tile order, see fwb_order_compact)
    assert np.array_equal(np.sort(cidx[idx]), np.arange(len(idx)))
    eng.allocate(1)
    rng = np.random.default_rng(0)
    a = rng.random(shape)
    eng.upload_state(0, a)
    back = np.empty(shape)
    eng.download_state(0, back, -7.0)
    eng.synchronize()
    expect = np.where(mesh == 1, a, -7.0)
    assert np.array_equal(back, expect)
    assert np.array_equal(eng.state[0].cpu().numpy()[cidx[idx]], a.ravel()[idx])
    # every tile (8 work-list entries) owns one contiguous compact range
    tb = eng.tile_base.cpu().numpy().astype(np.int64)
    wl = eng.worklist[:eng.n_work].cpu().numpy().reshape(-1, 8)
    assert tb[0] == 0 and tb[-1] == eng.n_myo and np.all(np.diff(tb) >= 0)
    for t in (0, len(wl) // 2, len(wl) - 1):
        nodes = np.concatenate([np.arange(c * 32, c * 32 + 32) for c in wl[t] if c >= 0])
        nodes = nodes[nodes < mesh.size]
        got = np.sort(cidx[nodes][cidx[nodes] >= 0])
        assert np.array_equal(got, np.arange(tb[t], tb[t + 1]))
    del torch


def test_empty_and_full_tissue(env):
    eng = env["Engine"]((6, 40))
    mesh = np.zeros((6, 40), dtype=np.int8)
    eng.set_tissue(mesh)
    assert eng.n_myo == 0
    mesh = np.ones((6, 40), dtype=np.int8)
    eng.set_tissue(mesh)            # the engine re-applies the empty ring
    assert eng.n_myo == 4 * 38


@pytest.mark.parametrize("shape,kind", [((23, 37), "iso"), ((23, 37), "aniso"),
                                         ((40, 64), "aniso"), ((9, 11, 13), "iso"),
                                         ((9, 11, 13), "aniso"), ((8, 8, 32), "aniso")], ids=str)
def test_weights_bit_exact(env, shape, kind):
    from oracle import oracle
    rng = np.random.default_rng(1)
    mesh = oracle.apply_boundaries(random_fibrosis(shape, 0.25, 9))
    cond = 0.2 + rng.random(shape)
    fibers = random_fibers(shape, 3) if kind == "aniso" else None
    ref = oracle.compute_weights(mesh, cond, fibers, kind, 0.154, 0.01, 0.25)
    eng = env["Engine"](shape)
    eng.set_tissue(mesh)
    eng.compute_weights(0 if kind == "iso" else 1, cond, fibers, 1, 1 / 9, 0.154, 0.01, 0.25)
    got = eng.weights_dense()
    myo = mesh == 1
    assert np.array_equal(got[myo], ref[myo])
    # non-updated nodes: the reference leaves (0,..,1,..,0); they are never read.
    assert np.array_equal(got[~myo], ref[~myo])


@pytest.mark.parametrize("shape,kind", [((21, 19), "iso"), ((21, 64), "aniso"),
                                         ((7, 9, 11), "iso"), ((6, 10, 32), "aniso")], ids=str)
def test_diffuse_bit_exact_and_linear(env, shape, kind):
    from oracle import oracle
    from finitewave_b200.hostcall import diffuse_host
    rng = np.random.default_rng(2)
    mesh = oracle.apply_boundaries(random_fibrosis(shape, 0.2, 4))
    fibers = random_fibers(shape, 6) if kind == "aniso" else None
    w = oracle.compute_weights(mesh, 1.0, fibers, kind, 1.0, 0.01, 0.25)
    idx = np.flatnonzero(mesh == 1).astype(np.int64)
    u = rng.normal(size=shape)
    ref = np.full(shape, 3.25)
    oracle.diffuse(ref, u, w, idx, oracle.flat_offsets(kind, shape))
    got = np.full(shape, 3.25)
    diffuse_host(len(shape), 0 if kind == "iso" else 1, got, u, w, idx)
    assert np.array_equal(got, ref)
    # linearity under an exact scaling (power of two)
    got4 = np.full(shape, 3.25)
    diffuse_host(len(shape), 0 if kind == "iso" else 1, got4, 4.0 * u, w, idx)
    myo = mesh == 1
    assert np.array_equal(got4[myo], 4.0 * got[myo])
    assert np.all(got4[~myo] == 3.25)


def test_stim_box_clipping_and_empty_roi(env):
    import finitewave_b200 as fw
    tissue = fw.CardiacTissue2D([20, 20])
    tissue.mesh[5:8, 5:8] = 2
    model = fw.AlievPanfilov2D()
    model.dt, model.dr, model.t_max, model.prog_bar = 0.01, 0.25, 0.01, False
    model.cardiac_tissue = tissue
    seq = fw.StimSequence()
    seq.add_stim(fw.StimVoltageCoord2D(0, 1.0, 3, 400, -4, 10))     # clipped like numpy slices
    seq.add_stim(fw.StimVoltageCoord2D(0, 2.0, 10, 10, 0, 20))      # empty ROI
    model.stim_sequence = seq
    model.initialize()
    u0 = np.zeros((20, 20))
    sl = (slice(3, 400), slice(-4, 10))
    u0[sl][tissue.mesh[sl] == 1] = 1.0
    model.t_max = 0.0                                                # zero steps: only upload/download
    model.run(initialize=False)
    assert np.array_equal(model.u, np.zeros((20, 20)))
    model.t_max = 0.01
    model.run(initialize=False)
    # after one step the stimulated pattern has diffused; compare with the oracle
    from oracle import oracle
    case = dict(model="aliev_panfilov", shape=[20, 20], dt=0.01, dr=0.25, t_max=0.01,
                mesh=tissue.mesh.copy(),
                stims=[dict(kind="voltage_coord", t=0, value=1.0, box=[3, 400, -4, 10]),
                       dict(kind="voltage_coord", t=0, value=2.0, box=[10, 10, 0, 20])])
    ref = oracle.simulate(case)
    assert np.array_equal(model.u, ref["u"])
    assert np.array_equal(model.v, ref["v"])


def test_unknown_model_or_missing_fibers_raise(env):
    import finitewave_b200 as fw
    tissue = fw.CardiacTissue2D([10, 10])
    model = fw.Barkley2D()
    model.dt, model.dr, model.t_max, model.prog_bar = 0.01, 0.25, 0.05, False
    model.cardiac_tissue = tissue
    model.stencil = fw.AsymmetricStencil2D()
    with pytest.raises(ValueError):
        model.run()
    lib = env["lib"]
    L = env["L"]
    sim = ctypes.c_void_p(0)
    shape = (ctypes.c_int64 * 2)(10, 10)
    one = ctypes.c_void_p(8)
    rc = L.fwb_sim_create(ctypes.byref(sim), 2, shape, 17, 0, one, one, one, 0, 32, one, 8, one,
                          one, one, one, None, 0, 0.01, ctypes.c_void_p(0))
    assert rc == -1
    with pytest.raises(lib.FwbError):
        lib.check(rc, "fwb_sim_create")


def test_public_api_runs_the_tma_kernels(monkeypatch):
    """The host API lays the tissue out tile-ordered, so the step must take the TMA paths:
    the persistent ring for the HBM-bound models (variant 3; tissues of a few thousand nodes
    run many steps per launch in the cluster kernel instead, variant 6), the compact-lane
    tile kernel for the FP64-bound ones (4; 5 when the line length is a multiple of 32 nodes
    and the u brick comes by tensor TMA).  Guards against silently falling back to the
    plain-load kernel (variant 0)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import finitewave_b200 as fw
    from finitewave_b200 import _lib
    from tests.cases import build_model, case_by_name
    L = _lib.lib()
    for name, want in (("c2_fk2d_aniso_fib", 6), ("c3_ms3d_iso_focal", 6),
                       ("tp06_3d_iso_current", 4), ("lr91_2d_iso", 4),
                       ("court2d_iso_current", 4), ("c5_tp06_3d_aniso_slab", 4)):
        case = dict(case_by_name(name))
        case["t_max"] = 0.2
        model, _ = build_model(fw, case)
        model.run()
        assert L.fwb_last_step_variant() == want, (name, L.fwb_last_step_variant())
    # without the multi-step kernel the light models take the persistent ring
    monkeypatch.setenv("FWB_NO_SMALL_KERNEL", "1")
    for name in ("c2_fk2d_aniso_fib", "c3_ms3d_iso_focal"):
        case = dict(case_by_name(name))
        case["t_max"] = 0.2
        model, _ = build_model(fw, case)
        model.run()
        assert L.fwb_last_step_variant() in (3, 7), (name, L.fwb_last_step_variant())
    # a tiled grid (line length 64): TP06 with the u brick by tensor TMA
    monkeypatch.setenv("FWB_PACKED", "0")
    tissue = fw.CardiacTissue3D([6, 12, 64])
    model = fw.TP063D()
    model.dt, model.dr, model.t_max, model.prog_bar = 0.01, 0.25, 0.05, False
    model.cardiac_tissue = tissue
    seq = fw.StimSequence()
    seq.add_stim(fw.StimVoltageCoord3D(0, -20, 1, 5, 1, 11, 1, 6))
    model.stim_sequence = seq
    model.run()
    assert L.fwb_last_step_variant() == 5, L.fwb_last_step_variant()


def test_device_math_helpers_accuracy():
    """What the GPU computes for the fast-path helpers of csrc/fexp.cuh (through fwb_devmath):
    fexp / fexp_fast <= 2 ulp, flog <= 2 ulp (|log| >= 1/2), frcp <= 1 ulp, frcp3 <= 1.5 ulp."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from finitewave_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(17)

    def run(op, x):
        xd = torch.from_numpy(np.ascontiguousarray(x)).cuda()
        yd = torch.empty_like(xd)
        rc = L.fwb_devmath(op, ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(yd.data_ptr()),
                           xd.numel(), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        return yd.cpu().numpy()

    x = np.concatenate([rng.uniform(-690, 690, 2_000_000), rng.uniform(-40, 40, 2_000_000),
                        rng.uniform(-1e-3, 1e-3, 100_000)])
    ref = np.exp(x)
    for op in (0, 1):
        ulp = np.abs(run(op, x) - ref) / np.spacing(ref)
        assert ulp.max() <= 2.0, (op, ulp.max())
    xl = np.concatenate([np.exp(rng.uniform(-690, 690, 2_000_000)), rng.uniform(1e-5, 200, 2_000_000)])
    out, ref = run(2, xl), np.log(xl)
    big = np.abs(ref) >= 0.5
    assert (np.abs(out - ref)[big] / np.spacing(np.abs(ref[big]))).max() <= 2.0
    assert np.abs(out - ref)[~big].max() <= 3e-16
    xr = np.concatenate([np.exp(rng.uniform(-600, 600, 4_000_000)), rng.uniform(1, 2, 4_000_000),
                         -np.exp(rng.uniform(-50, 50, 1_000_000))])
    ref = 1.0 / xr
    for op, bound in ((3, 1.0), (4, 1.5)):
        ulp = np.abs(run(op, xr) - ref) / np.spacing(np.abs(ref))
        print("frcp op", op, "max ulp", ulp.max())
        assert ulp.max() <= bound, (op, ulp.max())
    xs = np.concatenate([np.exp(rng.uniform(-600, 600, 4_000_000)), rng.uniform(1, 4, 4_000_000)])
    ref = np.sqrt(xs)
    ulp = np.abs(run(5, xs) - ref) / np.spacing(ref)
    print("fsqrt max ulp", ulp.max())
    assert ulp.max() <= 1.0, ulp.max()
