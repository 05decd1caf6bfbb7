"""
The one-call fixtures taken from the live reference on the CPU side (tests/golden/
make_weights_golden.py, make_stim_golden.py, make_tracker_golden.py, make_scenario_golden.py)
through the CUDA kernels: fwb_compute_weights with conductivity arrays + fibres and
non-default D_al / D_ac, the device face of every stimulus class, fwb_tip_scan on smooth
random fields, the twenty usage scenarios through the real engine, and the hand-callable
track_tip_line() / cross_threshold().  First run on a B200: 40 passed
(profiles/r1_final_fixture_tests.log).
"""
import types
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


def _weight_cases():
    from tests.golden.make_weights_golden import weight_cases
    return weight_cases()


@pytest.mark.parametrize("name", ["iso2d", "aniso2d", "sym2d", "iso3d", "aniso3d"])
def test_device_weights_match_reference_on_wide_inputs(name):
    """fwb_compute_weights with a conductivity array + fibres and non-default D_al / D_ac."""
    from finitewave_b200.engine import Engine
    from oracle import oracle
    c = next(x for x in _weight_cases() if x["name"] == name)
    want = np.load(GOLDEN / f"weights_{name}.npz")["weights"]
    mesh = oracle.apply_boundaries(c["mesh"].copy())
    eng = Engine(mesh.shape)
    eng.set_tissue(mesh)
    kind = {"iso": 0, "aniso": 1, "sym": 2}[c["kind"]]
    eng.compute_weights(kind, c["conductivity"], c["fibers"], c["D_al"], c["D_ac"], c["D_model"],
                        c["dt"], c["dr"])
    got = eng.weights_dense()
    assert np.array_equal(got[mesh == 1], want[mesh == 1])
    assert np.array_equal(got, want)


def _stim_specs():
    from tests.golden.make_stim_golden import stim_specs
    return stim_specs()


@pytest.mark.parametrize("spec", _stim_specs(), ids=[s[0] for s in _stim_specs()])
def test_device_stimuli_match_reference_one_call(spec):
    """The native (device) face of every stimulus class: after ONE step the buffer that was
    `u` during the step holds the stimulated field (the stimulus edits `u` in place before the
    diffusion reads it)."""
    import finitewave_b200 as fw
    from tests.golden.make_stim_golden import DT, apply, fields
    key, dim, cls, args, kwargs, calls = spec
    mesh, u0 = fields(dim)
    want = apply(fw, (key, dim, cls, args, kwargs, 1))       # host statement, one call;
    if calls == 1:                                           # == the live reference's fixture
        assert np.array_equal(want, np.load(GOLDEN / "stim_onecall.npz")[key])
    tissue = (fw.CardiacTissue2D if dim == 2 else fw.CardiacTissue3D)(list(mesh.shape))
    tissue.mesh = mesh.copy()
    model = (fw.AlievPanfilov2D if dim == 2 else fw.AlievPanfilov3D)()
    model.dt, model.dr, model.t_max, model.prog_bar = DT, 0.25, 0.5 * DT, False
    model.cardiac_tissue = tissue
    seq = fw.StimSequence()
    seq.add_stim(getattr(fw, cls)(*args, **kwargs))
    model.stim_sequence = seq
    model.initialize()
    model.u[...] = u0
    model.run(initialize=False)
    assert model.step == 1
    assert np.array_equal(model.u_new, want)                 # the pre-swap `u`, post-stimulus


def test_device_tip_scan_matches_reference_on_random_fields():
    """fwb_tip_scan on the smooth random field pairs (53 tips) of tracker_onecall.npz."""
    import torch
    import finitewave_b200 as fw
    from finitewave_b200.engine import Engine
    from tests.golden.make_tracker_golden import tip_inputs
    g = np.load(GOLDEN / "tracker_onecall.npz")
    for k in range(4):
        a, b, thr = tip_inputs(k)
        eng = Engine(a.shape)
        tr = fw.SpiralWaveCore2DTracker()
        tr.threshold = thr
        tr.initialize(types.SimpleNamespace(u=a))
        rows = tr._scan(eng, torch.from_numpy(np.ascontiguousarray(b)).to(eng.device))
        assert np.array_equal(rows[:, :2], g[f"tips{k}"]), k


def _scenarios():
    from tests.golden.make_scenario_golden import SCENARIOS
    return SCENARIOS


@pytest.mark.parametrize("scenario", _scenarios(), ids=[f.__name__ for f in _scenarios()])
def test_usage_scenarios_on_the_device(scenario):
    """tests/golden/scenarios.npz through the real engine: bit-exact for the models without
    transcendental functions, 1e-9 relative (north-star bar) for FK / LR91 / TP06."""
    import finitewave_b200 as fw
    from tests.cases import max_rel_err
    from tests.golden.make_scenario_golden import DEVICE_EXACT, outputs
    g = np.load(GOLDEN / "scenarios.npz")
    m, ap = scenario(fw)
    for k, v in outputs(m, ap).items():
        want = g[f"{scenario.__name__}.{k}"]
        assert np.shape(v) == want.shape, k
        if scenario.__name__ in DEVICE_EXACT or k in ("step", "t"):
            assert np.array_equal(v, want), k
        else:
            assert max_rel_err(v, want) <= 1e-9, k


def test_public_track_tip_line_and_cross_threshold():
    """The two hand-callable device methods: SpiralWaveCore2DTracker.
    track_tip_line(u, u_new, threshold) against the live-reference tips, and
    LocalActivationTime2DTracker.cross_threshold() against the reference's statement
    (local_activation_time_2d_tracker.py:71-90) over a sequence of fields."""
    import finitewave_b200 as fw
    from tests.golden.make_tracker_golden import tip_inputs
    g = np.load(GOLDEN / "tracker_onecall.npz")
    tr = fw.SpiralWaveCore2DTracker()
    for k in range(4):
        a, b, thr = tip_inputs(k)
        tips = np.array(tr.track_tip_line(a, b, thr), dtype=np.float64).reshape(-1, 2)
        assert np.array_equal(tips, g[f"tips{k}"]), k

    tissue = fw.CardiacTissue2D([20, 16])
    model = fw.AlievPanfilov2D()
    model.dt, model.dr, model.t_max, model.prog_bar = 0.01, 0.25, 0.01, False
    model.cardiac_tissue = tissue
    lat = fw.LocalActivationTime2DTracker()
    lat.threshold = 0.4
    seq = fw.TrackerSequence()
    seq.add_tracker(lat)
    model.tracker_sequence = seq
    model.initialize()
    rng = np.random.default_rng(9)
    activated = np.zeros((20, 16), dtype=bool)
    for _ in range(6):
        model.u[...] = rng.random((20, 16))
        want = (model.u >= 0.4) & ~activated
        activated = np.where(want, True, activated)
        activated = np.where((model.u < 0.4) & activated, False, activated)
        assert np.array_equal(lat.cross_threshold(), want)


def test_calc_ecg_by_hand_matches_live_reference():
    """ECG{2,3}DTracker.calc_ecg() called by hand (stand-alone C-ABI entry fwb_ecg) on the
    final state of the two ECG cases, against the live reference's own calc_ecg() on the
    same potential (tests/golden/make_calc_ecg_golden.py): lead values to 1e-12 (the
    reference's reduction order is unspecified), u_tr = W u bit for bit."""
    import finitewave_b200 as fw
    from tests.cases import build_model, case_by_name, max_rel_err
    from tests.golden.make_calc_ecg_golden import CASES as ECG_CASES
    g = np.load(GOLDEN / "calc_ecg.npz")
    for name in ECG_CASES:
        case = dict(case_by_name(name))
        model, trackers = build_model(fw, case)
        model.initialize()
        tr = [t for _, t in trackers if hasattr(t, "calc_ecg")][0]
        model.u[...] = g[name + "__u"]
        ecg = tr.calc_ecg()
        assert max_rel_err(ecg, g[name + "__ecg"]) <= 1e-12, name
        want = g[name + "__u_tr"]
        myo = model.cardiac_tissue.mesh == 1
        assert np.array_equal(tr.u_tr[myo], want[myo]), name
        tr._track()                      # by hand: appends one sample
        assert len(tr.output) == 1 and max_rel_err(tr.output[0], g[name + "__ecg"]) <= 1e-12
