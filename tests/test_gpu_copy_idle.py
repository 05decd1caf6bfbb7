"""
Whole-sector stores of ``u_new`` in the ring kernel (``fwb_sim_set_copy_idle``): the idle lanes
of a listed chunk store the value ``u`` holds there, so that a warp writes its 256 bytes with one
instruction and L2 never fetches the rest of a partially written 32-byte sector from DRAM.

The switch may change nothing in the results: with it on and off the potential, the state and the
activation map are BIT-identical (and equal the CPU oracle to the parity bar), nodes the solver
does not update keep their values, and the host refuses the switch when ``u`` and ``u_new``
disagree on such a node (the reference leaves both alone:
/root/reference/finitewave/core/model/cardiac_model.py:164-189 only ever writes
``u_new`` through the kernels' myo_indexes loop).
"""
import numpy as np
import pytest

from tests.cases import build_model, collect_outputs, random_fibers, random_fibrosis
from tests.test_gpu_parity import _check_outputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fw():
    import finitewave_b200
    return finitewave_b200


def _case(model, shape, fibrosis):
    c = dict(name="copy_idle", model=model, shape=shape, dt=0.01, dr=0.25, t_max=3.0,
             mesh=random_fibrosis(shape, fibrosis, 71),
             stims=[dict(kind="voltage_coord", t=0, value=1.0,
                         box=[0, 5] + sum(([0, s] for s in shape[1:]), []))],
             trackers=[dict(kind="activation_time", threshold=0.5, step=1)])
    c["fibers"] = random_fibers(shape, 72)
    return c


def _run(fw, case, monkeypatch, copy_idle):
    monkeypatch.setenv("FWB_NO_SMALL_KERNEL", "1")       # the per-step ring kernel, not the cluster one
    monkeypatch.setenv("FWB_COPY_IDLE", "1" if copy_idle else "0")
    model, trackers = build_model(fw, case)
    model.run()
    from finitewave_b200 import _lib
    variant = _lib.lib().fwb_last_step_variant()
    return collect_outputs(case, model, trackers), model, variant


@pytest.mark.parametrize("model,shape", [("fenton_karma", [96, 160]),
                                         ("mitchell_schaeffer", [8, 24, 64])])
def test_whole_sector_stores_change_nothing(fw, monkeypatch, model, shape):
    from oracle import oracle
    case = _case(model, shape, 0.3)
    on, m_on, v_on = _run(fw, case, monkeypatch, True)
    off, m_off, v_off = _run(fw, case, monkeypatch, False)
    assert v_on == v_off and v_on in (3, 7), (v_on, v_off)      # both took the ring kernel
    for key in on:
        a, b = np.asarray(on[key]), np.asarray(off[key])
        assert a.shape == b.shape and np.array_equal(a, b, equal_nan=True), key
    ref = oracle.simulate(case)
    _check_outputs(case, on, ref, "oracle")
    # nodes the solver does not update still hold what they held before the run (rest = 0)
    mesh = np.asarray(m_on.cardiac_tissue.mesh)
    assert np.all(np.asarray(m_on.u)[mesh != 1] == 0.0)
    assert np.all(np.asarray(m_on.u_new)[mesh != 1] == 0.0)


def test_switch_is_refused_when_the_buffers_disagree_on_an_idle_node(fw, monkeypatch):
    """``u`` and ``u_new`` differ on a fibrotic node (a user wrote there): copying u -> u_new
    would change ``u_new``; the engine must leave the switch off and the run must keep both
    values, as the reference does."""
    monkeypatch.setenv("FWB_NO_SMALL_KERNEL", "1")
    monkeypatch.setenv("FWB_COPY_IDLE", "1")
    case = _case("fenton_karma", [96, 160], 0.3)
    case["t_max"] = 0.5
    model, trackers = build_model(fw, case)
    model.initialize()
    mesh = np.asarray(model.cardiac_tissue.mesh)
    i, j = np.argwhere(mesh[2:-2, 2:-2] == 2)[0] + 2
    model.u[i, j] = 0.25
    model.u_new[i, j] = -0.75
    model.run(initialize=False)
    got = sorted([float(np.asarray(model.u)[i, j]), float(np.asarray(model.u_new)[i, j])])
    assert got == [-0.75, 0.25], got
