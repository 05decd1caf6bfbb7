"""
Mid-run checkpoints: StateSaver / StateSaverCollection
(finitewave/core/state/state_saver.py) fire during a run without stalling the step loop --
device snapshot on the compute stream, D2H on a copy stream, .npy writes on a thread -- and
the files are identical to the ones the synchronous path writes.
"""
import numpy as np
import pytest

from tests.cases import build_model, case_by_name

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fw():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import finitewave_b200
    return finitewave_b200


def _run(fw, name, root, async_mode, times):
    case = dict(case_by_name(name))
    model, _ = build_model(fw, case)
    coll = fw.StateSaverCollection()
    for i, t in enumerate(times):
        coll.savers.append(fw.StateSaver(str(root / f"ck{i}"), time=t))
    model.state_saver = coll
    model.async_checkpoints = async_mode
    model.run()
    return model


@pytest.mark.parametrize("name", ["c2_fk2d_aniso_fib", "c5_tp06_3d_aniso_slab"])
def test_async_checkpoints_equal_synchronous_ones(fw, tmp_path, name):
    # (a `time = -1` saver fires only if the accumulated t reaches t_max, which it does not
    # for dt = 0.01 -- reference behaviour, state_saver.py:49-53 -- so use explicit times)
    times = [1.0, 2.5, 4.0]
    m_async = _run(fw, name, tmp_path / "a", True, times)
    m_sync = _run(fw, name, tmp_path / "s", False, times)
    assert getattr(m_async, "_ckpt_writer", None) is not None, "asynchronous path was not taken"
    assert getattr(m_sync, "_ckpt_writer", None) is None
    for i in range(len(times)):
        for var in m_async.state_vars:
            a = np.load(tmp_path / "a" / f"ck{i}" / f"{var}.npy")
            s = np.load(tmp_path / "s" / f"ck{i}" / f"{var}.npy")
            assert a.shape == tuple(m_async.u.shape)
            assert np.array_equal(a, s), (i, var)
    # the three checkpoints are three different instants
    u0 = np.load(tmp_path / "a" / "ck0" / "u.npy")
    u1 = np.load(tmp_path / "a" / "ck1" / "u.npy")
    u2 = np.load(tmp_path / "a" / "ck2" / "u.npy")
    assert not np.array_equal(u0, u1) and not np.array_equal(u1, u2)


def test_checkpoint_restart_reproduces_the_run(fw, tmp_path):
    """StateLoader of a mid-run checkpoint + the remaining steps == the uninterrupted run
    (same stimuli-free tail), bit for bit."""
    case = dict(case_by_name("c3_ms3d_iso_focal"))
    model, _ = build_model(fw, case)
    model.state_saver = fw.StateSaver(str(tmp_path / "mid"), time=3.0)
    model.run()
    u_full = np.array(model.u)
    # the saver fired on the first step whose accumulated t reached 3.0
    t, k = 0.0, 0
    while t < 3.0:
        t += case["dt"]
        k += 1
    remaining = int(model.step) - k
    assert remaining > 100
    # restart: t runs from 0 again; the stimulus (t = 0, voltage) must not fire again
    case2 = dict(case, stims=[], trackers=[], t_max=(remaining - 0.5) * case["dt"])
    m2, _ = build_model(fw, case2)
    m2.state_loader = fw.StateLoader(str(tmp_path / "mid"))
    m2.run()
    assert m2.step == remaining
    assert np.array_equal(m2.u, u_full)
    assert np.array_equal(m2.h, model.h)


def test_animation_trackers_stream_frames(fw, tmp_path):
    """Animation2D / AnimationSlice3D trackers (reference tests/test_trackers_2d.py:96-120,
    test_trackers_3d.py): one .npy per sample, streamed while the run continues; the frames
    equal what a host-hook tracker sees at the same steps."""
    case = dict(case_by_name("c2_fk2d_aniso_fib"), t_max=4.0, trackers=[])
    model, _ = build_model(fw, case)
    anim = fw.Animation2DTracker()
    anim.path, anim.dir_name, anim.step, anim.variable_name = str(tmp_path), "frames_u", 50, "u"
    anim_v = fw.Animation2DTracker()
    anim_v.path, anim_v.dir_name, anim_v.step, anim_v.variable_name = str(tmp_path), "frames_v", 100, "v"
    anim_v.frame_type = "float32"

    class Probe(fw.Tracker):                       # host hook: downloads the arrays
        def initialize(self, model):
            self.model, self.seen = model, []

        def _track(self):
            self.seen.append((self.model.u.copy(), self.model.v.copy()))
    probe = Probe()
    probe.step = 100
    seq = fw.TrackerSequence()
    for tr in (anim, anim_v, probe):
        seq.add_tracker(tr)
    model.tracker_sequence = seq
    model.run()
    n_u = len(list((tmp_path / "frames_u").glob("*.npy")))
    n_v = len(list((tmp_path / "frames_v").glob("*.npy")))
    assert n_u == model.step // 50 and n_v == model.step // 100 == len(probe.seen)
    for i, (u, v) in enumerate(probe.seen):
        assert np.array_equal(np.load(tmp_path / "frames_u" / f"{2 * i}.npy"), u)
        fv = np.load(tmp_path / "frames_v" / f"{i}.npy")
        assert fv.dtype == np.float32 and np.array_equal(fv, v.astype(np.float32))
    assert any(np.any(np.load(tmp_path / "frames_u" / f"{i}.npy") > 0) for i in range(n_u))

    case3 = dict(case_by_name("c3_ms3d_iso_focal"), t_max=2.0, trackers=[])
    m3, _ = build_model(fw, case3)
    sl = fw.AnimationSlice3DTracker()
    sl.path, sl.dir_name, sl.step, sl.slice_z = str(tmp_path), "slices", 40, 12
    full = fw.Animation3DTracker()
    full.path, full.dir_name, full.step = str(tmp_path), "vol", 40
    seq3 = fw.TrackerSequence()
    seq3.add_tracker(sl)
    seq3.add_tracker(full)
    m3.tracker_sequence = seq3
    m3.run()
    for i in range(m3.step // 40):
        vol = np.load(tmp_path / "vol" / f"{i}.npy")
        assert vol.shape == (24, 24, 24)
        assert np.array_equal(np.load(tmp_path / "slices" / f"{i}.npy"), vol[:, :, 12])
    with pytest.raises(ValueError):
        bad = fw.AnimationSlice3DTracker()
        bad.initialize(m3)


def test_period_animation_tracker(fw, tmp_path):
    """PeriodAnimation2DTracker (reference period_animation_2d_tracker.py:49-60): frame i is the
    map of the latest up-crossing time after sample i; the last frame equals the per-node
    maximum over the LocalActivationTime layers sampled at the same steps."""
    case = dict(case_by_name("barkley2d_lat_period"), trackers=[])
    model, _ = build_model(fw, case)
    pa = fw.PeriodAnimation2DTracker()
    pa.path, pa.dir_name, pa.threshold, pa.step, pa.overwrite = str(tmp_path), "pm", 0.5, 20, True
    lat = fw.LocalActivationTime2DTracker()
    lat.threshold, lat.step = 0.5, 20
    seq = fw.TrackerSequence()
    seq.add_tracker(pa)
    seq.add_tracker(lat)
    model.tracker_sequence = seq
    model.run()
    n = model.step // 20
    frames = [np.load(tmp_path / "pm" / f"{i}.npy") for i in range(n)]
    assert len(list((tmp_path / "pm").glob("*.npy"))) == n
    last = np.max(lat.output, axis=0)
    assert np.array_equal(frames[-1], last)
    assert np.array_equal(pa.output, last)
    assert (frames[0] == -1).sum() > (frames[-1] == -1).sum()
    for a, b in zip(frames, frames[1:]):
        assert np.all(b >= a)
