"""
Parity of the compact-lane tile kernel (LR91 / TP06 / Courtemanche) and of the tile-granular
work list on TILED grids -- line length a multiple of 32 nodes, where a block owns one
spatial tile (3D: 2 planes x 4 rows x 32 nodes), its u neighbourhood arrives as a brick by
tensor TMA, sparse tissue leaves tiles partially filled (compact lanes, shadow lanes, warps
that skip) and a tile with a node outside the fast path's domain is handed to the
reference-statement kernel.  Everything is compared with the CPU oracle on the same inputs
(bar: 1e-9 relative on u and every state variable, activation maps within one dt, ECG 1e-9),
with and without the brick (FWB_NO_BRICK=1 -> plain loads).
"""
import numpy as np
import pytest

from tests.cases import (build_and_run, max_rel_err, random_fibers, random_fibrosis,
                         ventricle_shell)
from tests.test_gpu_parity import _check_outputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fw():
    import finitewave_b200
    return finitewave_b200


def _tp06_tiled(shape, t_max, fibrosis=0.2):
    return dict(name="tp06_tiled", model="tp06", shape=shape, dt=0.01, dr=0.25, t_max=t_max,
                mesh=random_fibrosis(shape, fibrosis, 31), fibers=random_fibers(shape, 32),
                stims=[dict(kind="voltage_coord", t=0, value=-20,
                            box=[0, 4] + sum(([0, s] for s in shape[1:]), []))],
                trackers=[dict(kind="activation_time", threshold=-40, step=1),
                          dict(kind="ecg", coords=[[4, 8, 70], [1, 1, 1]], step=7)])


@pytest.mark.parametrize("mode", ["brick", "plain-loads", "packed"])
def test_tp06_tracker_ecg_on_a_tiled_shape(fw, mode, monkeypatch):
    """TP06 + ActivationTime (every step) + ECG on 8 x 16 x 64 with 20 % fibrosis and random
    fibres (19-point stencil): real tiles, most of them partially filled, 400 steps -- in tile
    mode with the u brick by tensor TMA, in tile mode with plain loads, and in packed mode
    (256 consecutive compact nodes per block: what a tissue this sparse gets by default)."""
    from oracle import oracle
    monkeypatch.setenv("FWB_PACKED", "1" if mode == "packed" else "0")
    if mode == "plain-loads":
        monkeypatch.setenv("FWB_NO_BRICK", "1")
    case = _tp06_tiled([8, 16, 64], 4.0)
    ref = oracle.simulate(case)
    out = build_and_run(fw, case)
    _check_outputs(case, out, ref, "oracle")
    from finitewave_b200 import _lib
    assert _lib.lib().fwb_last_step_variant() == (5 if mode == "brick" else 4)


def test_tp06_iso7_and_2d_tiled(fw):
    from oracle import oracle
    for shape in ([6, 12, 32], [24, 96]):
        case = dict(name="tp06_iso_tiled", model="tp06", shape=shape, dt=0.01, dr=0.25, t_max=2.0,
                    mesh=random_fibrosis(shape, 0.3, 41),
                    stims=[dict(kind="current_coord", t=0.1, value=80, duration=0.4,
                                box=[0, 5] + sum(([0, s] for s in shape[1:]), []))],
                    trackers=[dict(kind="activation_time", threshold=-40, step=3)])
        ref = oracle.simulate(case)
        out = build_and_run(fw, case)
        _check_outputs(case, out, ref, "oracle")


def test_lr91_and_courtemanche_tiled(fw):
    from oracle import oracle
    for model, shape, fib in (("luo_rudy91", [40, 64], True), ("luo_rudy91", [6, 10, 32], False),
                              ("courtemanche", [20, 32], True), ("courtemanche", [5, 9, 32], False)):
        case = dict(name="heavy_tiled", model=model, shape=shape, dt=0.01, dr=0.25, t_max=2.0,
                    mesh=random_fibrosis(shape, 0.25, 51),
                    stims=[dict(kind="voltage_coord", t=0, value=-20,
                                box=[0, 4] + sum(([0, s] for s in shape[1:]), []))],
                    trackers=[dict(kind="activation_time", threshold=-40, step=1)])
        if fib:
            case["fibers"] = random_fibers(shape, 52)
        ref = oracle.simulate(case)
        out = build_and_run(fw, case)
        _check_outputs(case, out, ref, "oracle")


def test_tile_with_ineligible_node_goes_to_the_reference_statement_kernel(fw):
    """A stimulus of +400 mV puts a box of nodes outside the fast path's |u| < 300 mV: their
    tiles are handed to step_kernel_tile<SLOW> (reference statement, library exp) until the
    potential has relaxed; everything else stays on the fast path.  The oracle evaluates the
    reference statement everywhere."""
    from oracle import oracle
    shape = [6, 12, 64]
    case = dict(name="tp06_defer", model="tp06", shape=shape, dt=0.01, dr=0.25, t_max=1.0,
                fibers=random_fibers(shape, 61),
                stims=[dict(kind="voltage_coord", t=0, value=400, box=[2, 4, 3, 9, 10, 40]),
                       dict(kind="voltage_coord", t=0.3, value=-350, box=[1, 3, 2, 6, 30, 60])],
                trackers=[dict(kind="activation_time", threshold=-40, step=1),
                          dict(kind="ecg", coords=[[3, 6, 70]], step=5)])
    ref = oracle.simulate(case)
    out = build_and_run(fw, case)
    _check_outputs(case, out, ref, "oracle")


@pytest.mark.parametrize("small", [True, False], ids=["cluster-kernel", "ring-kernel"])
def test_light_models_on_sparse_tiled_tissue(fw, small, monkeypatch):
    """Tile-granular work list with empty slots (a ventricle-like shell in a 3D box, 30 %
    fibrosis in 2D) through the persistent ring kernel and through the multi-step cluster
    kernel: bit-exact for models without transcendentals."""
    from oracle import oracle
    if not small:
        monkeypatch.setenv("FWB_NO_SMALL_KERNEL", "1")
    mesh, fib = ventricle_shell([24, 24, 32])
    cases = [dict(name="ms_shell", model="mitchell_schaeffer", shape=[24, 24, 32], dt=0.01,
                  dr=0.25, t_max=3.0, mesh=mesh, fibers=fib,
                  stims=[dict(kind="voltage_coord", t=0, value=1, box=[0, 24, 0, 24, 0, 6])],
                  trackers=[dict(kind="activation_time", threshold=0.5, step=1)]),
             dict(name="ap_fib", model="aliev_panfilov", shape=[48, 96], dt=0.01, dr=0.25,
                  t_max=3.0, mesh=random_fibrosis([48, 96], 0.3, 71),
                  stims=[dict(kind="voltage_coord", t=0, value=1, box=[0, 48, 0, 5]),
                         dict(kind="current_coord", t=1.0, value=3, duration=0.3,
                              box=[10, 20, 40, 60])],
                  trackers=[dict(kind="activation_time", threshold=0.5, step=10)])]
    for case in cases:
        ref = oracle.simulate(case)
        out = build_and_run(fw, case)
        _check_outputs(case, out, ref, "oracle")


def test_readme_quick_start_in_one_launch(fw):
    """C1 (README.rst:417-451): after the stimulus step the remaining 999 steps run inside ONE
    launch of the cluster kernel; result bit-identical to the live-reference fixture."""
    from pathlib import Path
    from tests.cases import build_model, case_by_name, collect_outputs
    g = np.load(Path(__file__).resolve().parent / "golden" / "c1_ap2d_readme.npz")
    case = dict(case_by_name("c1_ap2d_readme"))
    model, trackers = build_model(fw, case)
    model.run()
    out = collect_outputs(case, model, trackers)
    assert np.array_equal(out["u"], g["u"]) and np.array_equal(out["v"], g["v"])
    assert np.array_equal(out["tracker0"], g["tracker0"])
    assert model.gpu_steps == 1000 and model.gpu_launches <= 8, model.gpu_launches
