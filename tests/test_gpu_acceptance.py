"""
The reference's own public-API acceptance tests for the on-path classes, run against
finitewave_b200 ("reused as the drop-in acceptance suite").  Same
set-ups and the same physiological bands as
  /root/reference/tests/test_models_2d.py:8-268, test_models_3d.py (6 on-path models),
  tests/test_trackers_2d.py:8-57,60-93,123-146,183-203,261-278,
  tests/test_basics.py:8-75,
plus regressions of the host-hook machinery (second run(), mesh edited by a Command,
user-defined tracker / stimulus classes as host hooks).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fw():
    import finitewave_b200
    return finitewave_b200


def prepare_model(fw, model_class, dim, curr_value, curr_dur, t_calc, t_prebeats):
    ni, nj = 5, 3
    if dim == 2:
        tissue = fw.CardiacTissue2D([ni, nj])
        box = (0, 2, 0, nj)
        stim = fw.StimCurrentCoord2D
    else:
        tissue = fw.CardiacTissue3D([ni, nj, 3])
        box = (0, 2, 0, nj, 0, 3)
        stim = fw.StimCurrentCoord3D
    seq = fw.StimSequence()
    for k in range(4):
        seq.add_stim(stim(k * t_prebeats, curr_value, curr_dur, *box))
    model = model_class()
    model.dt, model.dr, model.prog_bar = 0.01, 0.25, False
    model.t_max = 3 * t_prebeats + t_calc
    model.cardiac_tissue = tissue
    model.stim_sequence = seq
    return model


def run_model(fw, model, dim):
    tracker = fw.ActionPotential2DTracker() if dim == 2 else fw.ActionPotential3DTracker()
    tracker.cell_ind = [3, 1] if dim == 2 else [3, 1, 1]
    tracker.step = 1
    seq = fw.TrackerSequence()
    seq.add_tracker(tracker)
    model.tracker_sequence = seq
    model.run()
    return tracker.output


def calculate_apd(u, dt, threshold, beat_index=3):
    up_idx = np.where((u[:-1] < threshold) & (u[1:] >= threshold))[0]
    down_idx = np.where((u[:-1] > threshold) & (u[1:] <= threshold))[0]
    if len(up_idx) <= beat_index or len(down_idx) == 0:
        return None
    ap_start = up_idx[beat_index]
    ends = down_idx[down_idx > ap_start]
    return None if len(ends) == 0 else (ends[0] - ap_start) * dt


#            class            value dur  t_calc prebeats  max(abs tol)  min          thr   APD band
MODELS = [("AlievPanfilov", 5, 0.5, 80, 60, (1.0, 0.1), (0.0, 0.01), 0.1, (20, 25)),
          ("Barkley", 5, 0.1, 80, 60, (1.0, 0.1), (0.0, 0.01), 0.1, (1, 4)),
          ("MitchellSchaeffer", 5, 0.5, 1000, 1000, (0.95, 0.1), (0.0, 0.01), 0.1, (250, 350)),
          ("FentonKarma", 5, 0.5, 1000, 1000, (1.0, 0.1), (0.0, 0.01), 0.1, (100, 200)),
          ("BuenoOrovio", 5, 0.5, 1000, 1000, (1.4, 0.1), (0.0, 0.01), 0.1, (200, 300)),
          ("LuoRudy91", 100, 1, 1000, 1000, None, None, -70, (350, 400)),
          ("TP06", 100, 1, 1000, 1000, None, None, -70, (280, 320)),
          ("Courtemanche", 100, 1, 1000, 1000, None, None, -70, (200, 300))]


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("spec", MODELS, ids=[m[0] for m in MODELS])
def test_model_action_potential_bands(fw, spec, dim):
    name, val, dur, t_calc, t_pre, mx, mn, thr, band = spec
    model = prepare_model(fw, getattr(fw, f"{name}{dim}D"), dim, val, dur, t_calc, t_pre)
    u = run_model(fw, model, dim)
    assert len(u) == model.step
    if mx is not None:
        assert np.max(u) == pytest.approx(mx[0], abs=mx[1])
        assert np.min(u) == pytest.approx(mn[0], abs=mn[1])
    else:
        assert np.max(u) > (10 if name == "Courtemanche" else 20) and np.min(u) < -80
    apd = calculate_apd(u, model.dt, threshold=thr)
    assert apd is not None and band[0] <= apd <= band[1], f"{name}{dim}D APD {apd}"


def cable_model(fw):
    tissue = fw.CardiacTissue2D([12, 3])
    seq = fw.StimSequence()
    seq.add_stim(fw.StimCurrentCoord2D(0, 5, 0.5, 0, 5, 0, 3))
    model = fw.AlievPanfilov2D()
    model.dt, model.dr, model.t_max, model.prog_bar = 0.01, 0.25, 3, False
    model.cardiac_tissue, model.stim_sequence = tissue, seq
    return model


def test_activation_time_wave_speed(fw):
    model = cable_model(fw)
    tr = fw.ActivationTime2DTracker()
    tr.threshold, tr.step, tr.start_time = 0.5, 1, 0
    seq = fw.TrackerSequence()
    seq.add_tracker(tr)
    model.tracker_sequence = seq
    model.run()
    ats = tr.output
    speed = 5 * model.dr / ats[10, 1]
    assert 1.5 <= speed <= 2


def test_multi_variable_and_variable_trackers(fw):
    model = cable_model(fw)
    mv = fw.MultiVariable2DTracker()
    mv.cell_ind = [10, 1]
    mv.var_list = ["u", "v"]
    var = fw.Variable2DTracker()
    var.var_name, var.cell_ind = "v", [10, 1]
    seq = fw.TrackerSequence()
    seq.add_tracker(mv)
    seq.add_tracker(var)
    model.tracker_sequence = seq
    model.t_max = 30
    model.run()
    assert len(mv.output["u"]) == len(mv.output["v"]) == model.step
    assert np.max(mv.output["u"]) == pytest.approx(1.0, abs=0.1)
    assert np.max(mv.output["v"]) == pytest.approx(2, abs=0.1)     # test_trackers_2d.py:202
    assert np.array_equal(np.squeeze(var.output), mv.output["v"])


def test_ecg_2d_tracker(fw):
    tissue = fw.CardiacTissue2D([50, 5])
    seq = fw.StimSequence()
    seq.add_stim(fw.StimCurrentCoord2D(5, 5, 0.5, 0, 5, 0, 5))
    model = fw.AlievPanfilov2D()
    model.dt, model.dr, model.t_max, model.prog_bar = 0.0015, 0.25, 15, False
    model.cardiac_tissue, model.stim_sequence = tissue, seq
    tr = fw.ECG2DTracker()
    tr.start_time, tr.step = 0, 10
    tr.measure_coords = np.array([[25, 2, 0]])
    ts = fw.TrackerSequence()
    ts.add_tracker(tr)
    model.tracker_sequence = ts
    model.run()
    ecg = tr.output.T[0]
    assert len(ecg) == 1000
    assert ecg.max() > 0.001 and ecg.min() < -0.001 and np.argmax(ecg) > 100


def test_state_saving_and_loading(fw, tmp_path):
    n = 5

    def make():
        m = fw.FentonKarma2D()
        m.dt, m.dr, m.t_max, m.prog_bar = 0.01, 0.25, 5, False
        m.cardiac_tissue = fw.CardiacTissue2D([n, n])
        return m
    model = make()
    seq = fw.StimSequence()
    seq.add_stim(fw.StimCurrentCoord2D(0, 10, 0.5, 1, n // 2, n // 2 + 1, n // 2, n // 2 + 1))
    model.stim_sequence = seq
    savers = fw.StateSaverCollection()
    savers.savers.append(fw.StateSaver(str(tmp_path / "state_0"), time=3))
    model.state_saver = savers
    model.run()
    before = [model.u.copy(), model.v.copy(), model.w.copy()]
    for var in ("u", "v", "w"):
        assert (tmp_path / "state_0" / f"{var}.npy").exists()

    # the saved state is exactly the device state at t = 3: a run stopped there equals it
    ref = make()
    ref.t_max = 3
    seq2 = fw.StimSequence()
    seq2.add_stim(fw.StimCurrentCoord2D(0, 10, 0.5, 1, n // 2, n // 2 + 1, n // 2, n // 2 + 1))
    ref.stim_sequence = seq2
    ref.run()
    saved = np.load(tmp_path / "state_0" / "u.npy")
    assert ref.step in (300, 301)
    assert np.allclose(saved, ref.u, atol=1e-3)

    model = make()
    model.state_loader = fw.StateLoader(str(tmp_path / "state_0"))
    model.run()
    after = [model.u.copy(), model.v.copy(), model.w.copy()]
    for b, a in zip(before, after):
        assert np.allclose(b, a, atol=1e-5)            # the reference's own assertion


def test_command_edits_u(fw):
    n = 5
    model = fw.FentonKarma2D()
    model.dt, model.dr, model.t_max, model.prog_bar = 0.01, 0.25, 10, False

    class ExcitationCommand(fw.Command):
        def execute(self, model):
            model.u[1:-1, 1:-1] = 1
    cs = fw.CommandSequence()
    cs.add_command(ExcitationCommand(5))
    model.cardiac_tissue = fw.CardiacTissue2D([n, n])
    model.stim_sequence = fw.StimSequence()
    model.command_sequence = cs
    model.run()
    assert np.mean(model.u[1:-1, 1:-1]) > 0.5


# ---- regressions of the host-hook machinery ------------------------------------------
def _spiral(fw, n=64):
    tissue = fw.CardiacTissue2D([n, n])
    seq = fw.StimSequence()
    seq.add_stim(fw.StimVoltageCoord2D(0, 1, 0, n, 0, 3))
    model = fw.Barkley2D()
    model.dt, model.dr, model.t_max, model.prog_bar = 0.01, 0.25, 4, False
    model.cardiac_tissue, model.stim_sequence = tissue, seq
    return model


def test_second_run_reinitialises_cleanly(fw):
    model = _spiral(fw)
    model.run()
    a = model.u.copy()
    model.run()                       # re-initialises: new index structures on the device
    assert np.array_equal(a, model.u)
    model.t_max = 6
    model.run(initialize=False)       # continues
    # t accumulates as a Python float (3.9999...): ceil((6 - t) / dt) = 201, as in the reference
    assert model.step in (600, 601) and not np.array_equal(a, model.u)


def test_command_that_edits_the_mesh_and_recomputes_weights(fw):
    """Tutorials/SpiralWaves2D.ipynb `UpdateMesh`: a Command blocks part of the tissue
    mid-run and calls model.compute_weights()."""
    from oracle import oracle

    class Block(fw.Command):
        def execute(self, model):
            model.cardiac_tissue.mesh[20:44, 30:34] = 2
            model.compute_weights()

    model = _spiral(fw)
    act = fw.ActivationTime2DTracker()
    act.threshold, act.step = 0.5, 1
    ts = fw.TrackerSequence()
    ts.add_tracker(act)
    model.tracker_sequence = ts
    cs = fw.CommandSequence()
    cs.add_command(Block(1.5))
    model.command_sequence = cs
    model.run()

    # the same thing in two oracle phases
    case = dict(model="barkley", shape=[64, 64], dt=0.01, dr=0.25, t_max=1.5 - 0.005,
                stims=[dict(kind="voltage_coord", t=0, value=1, box=[0, 64, 0, 3])])
    p1 = oracle.simulate(case)
    assert int(p1["step"]) == 150
    mesh2 = np.ones((64, 64), dtype=np.int8)
    mesh2[20:44, 30:34] = 2
    case2 = dict(model="barkley", shape=[64, 64], dt=0.01, dr=0.25, t_max=2.5 - 0.005, mesh=mesh2,
                 load_state=dict(u=p1["u"], v=p1["v"]))
    p2 = oracle.simulate(case2)
    assert int(p2["step"]) == 250 and model.step == 400
    assert np.array_equal(model.u, p2["u"]) and np.array_equal(model.v, p2["v"])
    assert (act.output[20:44, 30:34][:, :] < 1.5).all()      # blocked nodes never activate later
    assert act.output[50, 50] > 0


def test_user_defined_tracker_and_stim_are_host_hooks(fw):
    class Peak(fw.Tracker):                       # not native: sees numpy arrays
        def initialize(self, model):
            self.model = model
            self.peaks = []

        def _track(self):
            assert isinstance(self.model.u, np.ndarray)
            self.peaks.append((self.model.t, float(self.model.u.max()), float(self.model.v.max())))

        @property
        def output(self):
            return np.array(self.peaks)

    class Poke(fw.StimVoltage):                   # not native: edits model.u with numpy
        def stimulate(self, model):
            model.u[30:34, 30:34] = self.volt_value

    model = _spiral(fw)
    model.stim_sequence.add_stim(Poke(1.0, 0.9))
    tr = Peak()
    tr.step = 50
    ts = fw.TrackerSequence()
    ts.add_tracker(tr)
    model.tracker_sequence = ts
    model.run()
    out = tr.output
    assert len(out) == 8 and out[0][0] == 0 and abs(out[-1][0] - 3.5) < 1e-9
    assert out[1][1] > 0.9

    from oracle import oracle
    case = dict(model="barkley", shape=[64, 64], dt=0.01, dr=0.25, t_max=4,
                stims=[dict(kind="voltage_coord", t=0, value=1, box=[0, 64, 0, 3]),
                       dict(kind="voltage_coord", t=1.0, value=0.9, box=[30, 34, 30, 34])])
    ref = oracle.simulate(case)
    assert np.array_equal(model.u, ref["u"]) and np.array_equal(model.v, ref["v"])


# ---- the "next" rows, with the reference's own assertions -----------------------
def test_local_activation_time_2d_tracker(fw):
    """tests/test_trackers_2d.py:149-181: two waves -> two LAT layers, speeds in [1.5, 2]."""
    model = cable_model(fw)
    tracker = fw.LocalActivationTime2DTracker()
    tracker.threshold, tracker.step, tracker.start_time = 0.5, 1, 0
    seq = fw.TrackerSequence()
    seq.add_tracker(tracker)
    model.tracker_sequence = seq
    model.stim_sequence.add_stim(fw.StimVoltageCoord2D(45, 1, 0, 5, 0, 10))
    model.t_max = 50
    model.run()
    lats = tracker.output
    assert lats is not None and len(lats) == 2, "Every cell should have two LAT values"
    lat1, lat2 = lats[:, 10, 1]
    assert lat1 < lat2
    assert 1.5 <= 5 * model.dr / lat1 <= 2
    assert 1.5 <= 5 * model.dr / (lat2 - 45) <= 2


def test_local_activation_time_3d_tracker(fw):
    """tests/test_trackers_3d.py (cable along the first axis): same assertions in 3D."""
    tissue = fw.CardiacTissue3D([12, 3, 3])
    seq = fw.StimSequence()
    seq.add_stim(fw.StimCurrentCoord3D(0, 5, 0.5, 0, 5, 0, 3, 0, 3))
    seq.add_stim(fw.StimVoltageCoord3D(45, 1, 0, 5, 0, 3, 0, 3))
    model = fw.AlievPanfilov3D()
    model.dt, model.dr, model.t_max, model.prog_bar = 0.01, 0.25, 50, False
    model.cardiac_tissue, model.stim_sequence = tissue, seq
    tracker = fw.LocalActivationTime3DTracker()
    tracker.threshold, tracker.step = 0.5, 1
    ts = fw.TrackerSequence()
    ts.add_tracker(tracker)
    model.tracker_sequence = ts
    model.run()
    lats = tracker.output
    assert len(lats) == 2
    lat1, lat2 = lats[:, 10, 1, 1]
    assert 1.5 <= 5 * model.dr / lat1 <= 2 and 1.5 <= 5 * model.dr / (lat2 - 45) <= 2


def test_spiral_wave_period_2d_tracker(fw):
    """tests/test_trackers_2d.py:25-40, :234-258: Barkley spiral, detector periods 3.5 +- 0.2."""
    ni = nj = 100
    tissue = fw.CardiacTissue2D([ni, nj])
    stims = fw.StimSequence()
    stims.add_stim(fw.StimVoltageCoord2D(0, 1, 0, ni, 0, 3))
    stims.add_stim(fw.StimVoltageCoord2D(5, 1, 0, ni // 2, 0, nj))
    model = fw.Barkley2D()
    model.dt, model.dr, model.t_max, model.prog_bar = 0.01, 0.25, 20, False
    model.cardiac_tissue, model.stim_sequence = tissue, stims
    tracker = fw.Period2DTracker()
    tracker.cell_ind = np.array([[80, 80], [20, 70], [40, 10], [25, 90]])
    tracker.threshold, tracker.start_time, tracker.step = 0.5, 10, 10
    seq = fw.TrackerSequence()
    seq.add_tracker(tracker)
    model.tracker_sequence = seq
    model.run()
    periods = tracker.output
    assert periods is not None and len(periods) > 0
    mean = np.nanmean(np.array([np.mean(x) if len(x) > 0 else np.nan for x in periods]))
    assert mean == pytest.approx(3.5, abs=0.2)


def test_fibrosis_patterns_reference_assertions(fw):
    """tests/test_fibrosis_2d.py, test_fibrosis_3d.py: shape, values in {1, 2}, density of the
    sub-region within 0.01 (diffuse) / 0.05 (structural)."""
    shape, (x1, x2, y1, y2) = (1000, 1000), (100, 900, 200, 800)
    res = fw.Diffuse2DPattern(density=0.3, x1=x1, x2=x2, y1=y1, y2=y2).generate(shape=shape)
    assert res.shape == shape and np.all(np.isin(res, [1, 2]))
    assert abs(np.mean(res[x1:x2, y1:y2] == 2) - 0.3) < 0.01
    res = fw.Structural2DPattern(density=0.4, length_i=5, length_j=4, x1=x1, x2=x2, y1=y1,
                                 y2=y2).generate(shape=shape)
    assert res.shape == shape and np.all(np.isin(res, [1, 2]))
    assert abs(np.mean(res[x1:x2, y1:y2] == 2) - 0.4) < 0.05
    shape3, box = (100, 100, 100), (10, 90, 20, 80, 30, 70)
    res = fw.Diffuse3DPattern(*box, 0.3).generate(shape=shape3)
    sub = res[box[0]:box[1], box[2]:box[3], box[4]:box[5]]
    assert res.shape == shape3 and np.all(np.isin(res, [1, 2])) and abs(np.mean(sub == 2) - 0.3) < 0.01
    res = fw.Structural3DPattern(*box, 0.4, 5, 4, 3).generate(shape=shape3)
    sub = res[box[0]:box[1], box[2]:box[3], box[4]:box[5]]
    assert np.all(np.isin(res, [1, 2])) and abs(np.mean(sub == 2) - 0.4) < 0.05


def test_spiral_wave_core_2d_tracker(fw):
    """tests/test_trackers_2d.py:205-231: the tip of the Barkley spiral stays in a 6 x 6 box."""
    ni = nj = 100
    tissue = fw.CardiacTissue2D([ni, nj])
    stims = fw.StimSequence()
    stims.add_stim(fw.StimVoltageCoord2D(0, 1, 0, ni, 0, 3))
    stims.add_stim(fw.StimVoltageCoord2D(5, 1, 0, ni // 2, 0, nj))
    model = fw.Barkley2D()
    model.dt, model.dr, model.t_max, model.prog_bar = 0.01, 0.25, 20, False
    model.cardiac_tissue, model.stim_sequence = tissue, stims
    tracker = fw.SpiralWaveCore2DTracker()
    tracker.threshold, tracker.start_time, tracker.step = 0.5, 12, 10
    seq = fw.TrackerSequence()
    seq.add_tracker(tracker)
    model.tracker_sequence = seq
    model.run()
    core = tracker.output
    x, y = core["x"], core["y"]
    assert len(x) > 0 and len(y) > 0
    assert np.min(x) >= 32 and np.max(x) <= 38
    assert np.min(y) >= 47 and np.max(y) <= 53
    assert list(core.columns) == ["x", "y", "time", "step"]
