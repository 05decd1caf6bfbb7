"""
Host logic of the drop-in boundary, without a GPU (`-m "not gpu"`).

`CardiacModel.run()` is executed with the test double of tests/host_engine.py standing in
for the CUDA engine, so what is under test is everything the product does on the host: the
step order of the reference loop (cardiac_model.py:164-189), the planning of device segments
around host hooks, the stimuli's host statements and firing rule (stim_sequence.py:65-77,
stim.py:52-56), tracker gating (tracker.py:70-84), commands, state savers / loaders and the
Python-float accumulation of `t`.  Expected values are the golden fixtures generated from the
live reference (tests/golden/make_golden.py).
"""
from pathlib import Path

import numpy as np
import pytest

import finitewave_b200 as fw
from tests.cases import build_model, case_by_name, collect_outputs, make_cases
from tests.host_engine import host_faces_only, use_oracle_engine

GOLDEN = Path(__file__).resolve().parent / "golden"

# cases whose trackers all have a host statement (ECG and the LAT / period / spiral-core
# trackers are evaluated by device kernels only)
_HOST_TRACKERS = {"activation_time", "action_potential", "multi_variable"}
CASES = [c for c in make_cases()
         if all(t["kind"] in _HOST_TRACKERS for t in c.get("trackers", []))
         and int(np.prod(c["shape"])) * c["t_max"] / c["dt"] <= 4e6]


@pytest.fixture(autouse=True)
def _engine(monkeypatch):
    use_oracle_engine(monkeypatch)


def _run(case, prepare=None):
    model, trackers = build_model(fw, case)
    host_faces_only(model)
    if prepare:
        prepare(model)
    model.run()
    return model, collect_outputs(case, model, trackers)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_host_loop_reproduces_reference(case):
    """Stimuli and trackers as host hooks around oracle-stepped segments == live reference,
    bit for bit (same arithmetic as tests/test_oracle_golden.py, but the ORDER of stimulus,
    step, tracker sample, time increment and swap is now the product's run())."""
    g = np.load(GOLDEN / (case["name"] + ".npz"))
    model, out = _run(case)
    for k in g.files:
        if k in ("checksum", "weights"):
            continue
        assert out[k].shape == g[k].shape, k
        assert np.array_equal(out[k], g[k]), k
    assert np.array_equal(np.asarray(model.weights), g["weights"])


def test_segments_are_cut_only_where_a_hook_is_due():
    """No hooks between two samples of a host tracker -> the steps in between run as ONE
    device segment; the sampled step runs alone (its pre-swap buffers are the sample)."""
    case = dict(case_by_name("c1_ap2d_readme"), t_max=2.0)
    case["trackers"] = [dict(kind="action_potential", cell_ind=[50, 50], step=50)]
    model, out = _run(case)
    eng = model._engine
    assert model.step == 200 and sum(eng.runs) == 200
    # steps 0, 50, 100, 150 carry the tracker; steps 0 carries the voltage stimulus as well
    assert eng.runs == [1, 49, 1, 49, 1, 49, 1, 49]
    assert len(out["tracker0"]) == 4
    assert model.gpu_launches == 200


@pytest.mark.parametrize("t0,dur,dt,expect", [(0, 0.5, 0.01, 51), (0, 1, 0.01, 101),
                                              (5, 0.5, 0.0015, 334)])
def test_current_stimulus_pulse_count(t0, dur, dt, expect):
    """a current stimulus fires on every step with stim.t <= t < stim.t +
    duration plus the one on which it is marked passed, with `t` accumulated as a Python
    float (0, 0.5, 0.01) -> 51 applications, (0, 1, 0.01) -> 101, (5, 0.5, 0.0015) -> 334."""
    fired = []

    class Counting(fw.StimCurrentCoord2D):
        def stimulate(self, model):
            fired.append(model.t)
            super().stimulate(model)

    tissue = fw.CardiacTissue2D([8, 8])
    model = fw.AlievPanfilov2D()
    model.dt, model.dr, model.t_max, model.prog_bar = dt, 0.25, t0 + dur + 20 * dt, False
    model.cardiac_tissue = tissue
    seq = fw.StimSequence()
    st = Counting(t0, 0.1, dur, 2, 5, 2, 5)
    st._native = False
    seq.add_stim(st)
    model.stim_sequence = seq
    model.run()
    assert len(fired) == expect
    assert st.passed


def test_time_accumulates_like_a_python_float():
    model, _ = _run(dict(case_by_name("c1_ap2d_readme")))
    t = 0
    for _ in range(1000):
        t += 0.01
    assert model.step == 1000 and model.t == t == 9.999999999999831


def test_tracker_gate_window_and_stride():
    tr = fw.ActionPotential2DTracker()
    tr.start_time, tr.end_time, tr.step = 1.0, 2.0, 5
    assert not tr.gate(0.99, 0)          # before the window
    assert tr.gate(1.0, 100) and tr.gate(2.0, 200)   # both ends inclusive
    assert not tr.gate(2.0000001, 205)
    assert not tr.gate(1.5, 151)         # off stride


def test_command_sees_swapped_buffers_and_may_edit_state():
    """Commands run after `t += dt` and the swap (cardiac_model.py:180-183): at time c.t the
    model.u a command sees is the freshly computed potential; what it writes is what the next
    step starts from."""
    seen = {}

    class Kick(fw.Command):
        def execute(self, model):
            seen["t"], seen["step"] = model.t, model.step
            seen["u"] = model.u.copy()
            model.u[10:20, 10:20] = 0.75
            model.v[...] = 0.0

    case = dict(case_by_name("c1_ap2d_readme"), t_max=1.0, trackers=[])

    def prepare(model):
        seq = fw.CommandSequence()
        seq.add_command(Kick(0.5))
        model.command_sequence = seq

    model, _ = _run(case, prepare)
    # first step whose incremented time reaches 0.5 (float accumulation: 0.5000000000000002)
    assert seen["step"] == 50 and seen["t"] >= 0.5
    # the same edit applied by hand to a run that stops after those 50 steps
    ref, _ = _run(dict(case, t_max=seen["t"] - 1e-9))
    assert ref.step == 50
    assert np.array_equal(seen["u"], ref.u)       # the command saw the post-step potential
    ref.u[10:20, 10:20] = 0.75
    ref.v[...] = 0.0
    ref.t_max = 1.0
    ref.run(initialize=False)
    assert model.step == ref.step == 100
    assert np.array_equal(model.u, ref.u) and np.array_equal(model.v, ref.v)


def test_state_saver_and_loader_round_trip(tmp_path):
    """StateSaver at a mid-run time + at the end (time = -1), StateLoader restart: the
    restarted run ends bit-identical to the uninterrupted one (core/state/*.py)."""
    case = dict(case_by_name("fk2d_iso_current"), trackers=[], t_max=2)
    full, _ = _run(case)
    assert full.step == 200

    def with_savers(model):
        coll = fw.StateSaverCollection()
        coll.savers.append(fw.StateSaver(str(tmp_path / "mid"), time=1.5))
        coll.savers.append(fw.StateSaver(str(tmp_path / "end")))
        model.state_saver = coll

    first, _ = _run(case, with_savers)
    assert np.array_equal(first.u, full.u)
    for name in ("u", "v", "w"):
        assert np.array_equal(np.load(tmp_path / "end" / f"{name}.npy"), getattr(full, name))
    mid_u = np.load(tmp_path / "mid" / "u.npy")
    assert mid_u.shape == full.u.shape and not np.array_equal(mid_u, full.u)

    # restart from the mid-run state (the stimulus is over by then; the loader restores the
    # arrays only, state_loader.py:49-64, and time restarts at 0)
    t, k = 0, 0
    while t < 1.5:
        t += case["dt"]
        k += 1
    rest = dict(case, stims=[], t_max=(200 - k) * case["dt"] - 0.5 * case["dt"])
    model, _tr = build_model(fw, rest)
    host_faces_only(model)
    model.state_loader = fw.StateLoader(str(tmp_path / "mid"))
    model.run()
    assert model.step == 200 - k
    for name in ("u", "v", "w"):
        assert np.array_equal(getattr(model, name), getattr(full, name)), name


def test_end_of_run_saver_needs_t_to_reach_t_max(tmp_path):
    """Reference quirk kept (state_saver.py: `time < 0 and model.t < model.t_max -> return`):
    with dt = 0.01 and t_max = 3 the accumulated time ends at 2.99999999999998 < t_max, the
    run stops after ceil(t_max / dt) steps and the end-of-run saver never writes."""
    case = dict(case_by_name("fk2d_iso_current"), trackers=[], stims=[], t_max=3)

    def with_saver(model):
        model.state_saver = fw.StateSaver(str(tmp_path / "end"))

    model, _ = _run(case, with_saver)
    assert model.step == 300 and model.t < 3
    assert not (tmp_path / "end" / "u.npy").exists()


def test_user_defined_stimulus_and_tracker_run_as_host_hooks():
    """Classes the library has never seen are honoured through the host-hook mechanism and
    see plain numpy arrays."""
    samples = []

    class Ramp(fw.StimVoltage):
        def stimulate(self, model):
            assert isinstance(model.u, np.ndarray)
            model.u[1:4, 1:-1] = self.volt_value

    class Probe(fw.Tracker):
        def initialize(self, model):
            self.model = model

        def _track(self):
            samples.append((self.model.step, float(self.model.u[2, 5]), float(self.model.v[2, 5])))

    tissue = fw.CardiacTissue2D([24, 12])
    model = fw.AlievPanfilov2D()
    model.dt, model.dr, model.t_max, model.prog_bar = 0.01, 0.25, 0.5, False
    model.cardiac_tissue = tissue
    sseq = fw.StimSequence()
    sseq.add_stim(Ramp(0.1, 1.0))
    model.stim_sequence = sseq
    tseq = fw.TrackerSequence()
    probe = Probe()
    probe.step = 10
    tseq.add_tracker(probe)
    model.tracker_sequence = tseq
    model.run()
    assert [s[0] for s in samples] == [0, 10, 20, 30, 40]
    assert samples[0][1] == 0.0 and samples[1][1] == 0.0      # stimulus not yet due at 0.1-
    assert samples[2][1] > 0.5                                 # fired on the step t >= 0.1
    assert sseq.sequence[0].passed


def _stim_specs():
    from tests.golden.make_stim_golden import stim_specs
    return stim_specs()


@pytest.mark.parametrize("spec", _stim_specs(), ids=[s[0] for s in _stim_specs()])
def test_stimulus_host_statements_match_reference(spec):
    """Every stimulus class, one or two ``stimulate()`` calls on a random field (fixtures of
    tests/golden/make_stim_golden.py from the live reference: negative / out-of-range boxes,
    empty boxes, float matrices, biting u_max clamps, duplicate coordinates, voltage lists):
    the product's host statement AND the oracle's restatement, bit for bit."""
    from oracle import oracle
    from tests.golden.make_stim_golden import DT, apply, fields
    key, dim, cls, args, kwargs, calls = spec
    want = np.load(GOLDEN / "stim_onecall.npz")[key]
    assert np.array_equal(apply(fw, spec), want)

    kind = {"StimVoltageCoord": "voltage_coord", "StimCurrentCoord": "current_coord",
            "StimVoltageMatrix": "voltage_matrix", "StimCurrentMatrix": "current_matrix",
            "StimCurrentArea": "current_area",
            "StimVoltageListMatrix": "voltage_list_matrix"}[cls[:-2]]
    st = dict(kind=kind, value=args[1], u_max=kwargs.get("u_max"))
    if kind.endswith("_coord"):
        st["box"] = list(args[3:] if kind == "current_coord" else args[2:])
    elif kind == "current_area":
        st["coords"] = kwargs["coords"]
    else:
        st["matrix"] = np.asarray(args[-1])
    mesh, u = fields(dim)
    counters = {}
    for _ in range(calls):
        oracle._stimulate(st, u, mesh, DT, counters)
    assert np.array_equal(u, want)


@pytest.mark.parametrize("spec", _stim_specs(), ids=[s[0] for s in _stim_specs()])
def test_native_stimulus_descriptors_select_the_reference_nodes(spec):
    """The descriptor a built-in stimulus hands to the device runner -- a normalised index box
    (`_norm_box`) or a flat node list (`_flat_nodes`) -- applied the way the stim kernels
    document it (value / += dt * value / clamp, on mesh == 1 nodes of the box; on every listed
    node) gives the live reference's field.  The kernels themselves are covered on the GPU
    (tests/test_gpu_cabi.py, tests/test_gpu_parity.py)."""
    import types
    from tests.golden.make_stim_golden import DT, fields
    key, dim, cls, args, kwargs, calls = spec
    want = np.load(GOLDEN / "stim_onecall.npz")[key]
    mesh, u = fields(dim)
    model = types.SimpleNamespace(u=u, dt=DT, t=0.0, cardiac_tissue=types.SimpleNamespace(mesh=mesh))
    st = getattr(fw, cls)(*args, **kwargs)
    st.initialize(model)
    if hasattr(st, "_box"):
        b = st._norm_box(mesh.shape)
        sel = np.zeros(mesh.shape, dtype=bool)
        sel[tuple(slice(b[2 * d], b[2 * d + 1]) for d in range(dim))] = True
        flat = np.flatnonzero(sel & (mesh == 1))
    else:
        flat = np.asarray(st._flat_nodes(model))
        assert len(np.unique(flat)) == len(flat)
    uf = u.reshape(-1)
    for k in range(calls):
        if cls.startswith("StimVoltageList"):
            uf[flat] = st.volt_value[k]
        elif cls.startswith("StimVoltage"):
            uf[flat] = st.volt_value
        else:
            uf[flat] += DT * st.curr_value
            if st.u_max is not None:
                uf[flat] = np.minimum(uf[flat], st.u_max)
    assert np.array_equal(u, want)


def test_clone_continues_like_the_original():
    """``model.clone()`` is a deep copy without the device side (reference
    cardiac_model.py clone = deepcopy); ``run(initialize=False)`` on the clone rebuilds it from
    the copied host state -- arrays, weights, stimulus / tracker status -- and continues bit
    for bit like the original."""
    case = dict(case_by_name("fk2d_iso_current"), t_max=0.75, trackers=[
        dict(kind="action_potential", cell_ind=[3, 16], step=5)])
    model, trackers = build_model(fw, case)
    host_faces_only(model)
    model.run()                                   # stops in the middle of the current pulse
    twin = model.clone()
    assert twin._engine is None and isinstance(twin.weights, np.ndarray)
    assert np.array_equal(twin.weights, np.asarray(model.weights))
    for m in (model, twin):
        m.t_max = 2.0
        m.run(initialize=False)
    assert twin.step == model.step == 200
    for name in ("u", "v", "w"):
        assert np.array_equal(getattr(twin, name), getattr(model, name)), name
    a = model.tracker_sequence.sequence[0].output
    b = twin.tracker_sequence.sequence[0].output
    assert len(a) == 40 and np.array_equal(a, b)
    # and the uninterrupted run agrees with both
    whole, _ = _run(dict(case, t_max=2.0))
    assert np.array_equal(whole.u, model.u)


def _scenarios():
    from tests.golden.make_scenario_golden import SCENARIOS
    return SCENARIOS


@pytest.mark.parametrize("scenario", _scenarios(), ids=[f.__name__ for f in _scenarios()])
def test_usage_scenarios_match_reference(scenario, monkeypatch):
    """The less obvious ways of using a model -- Commands that move ``t_max``, parameters /
    ``dt`` / arrays edited between ``run()`` and ``run(initialize=False)``, tracker windows, a
    second full run, a Command that edits the mesh and recomputes the weights -- give what the
    live reference gives (tests/golden/make_scenario_golden.py), bit for bit."""
    from finitewave_b200 import stimulation, tracker
    from tests.golden.make_scenario_golden import outputs
    for mod in (stimulation, tracker):                  # built-ins on their host statements
        for name in dir(mod):
            cls = getattr(mod, name)
            if isinstance(cls, type) and cls.__dict__.get("_native", False):
                monkeypatch.setattr(cls, "_native", False)
    g = np.load(GOLDEN / "scenarios.npz")
    m, ap = scenario(fw)
    for k, v in outputs(m, ap).items():
        want = g[f"{scenario.__name__}.{k}"]
        assert np.shape(v) == want.shape, k
        assert np.array_equal(v, want), k
