"""
``CardiacModel.run()`` on several GPUs of one process (finitewave_b200/multi.py): the same
user script, ``model.devices = [...]`` (or FWB_DEVICES) as the only difference, must give
the single-GPU result bit for bit -- potential, every state variable, activation map; ECG
lead sums to 1e-12 (the slabs' partial sums are added in slab order).

With one visible GPU the slabs share it (devices = [0, 0]: same kernels, same peer stores
and flags, two host threads); with two or more they sit on different GPUs and the ghost
slices travel over NVLink.  The multi-GPU variant appends what it saw to
profiles/r2_multidevice_test.log when the directory is writable.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _build(fw, model_name, shape, devices, t_max):
    dim = len(shape)
    sfx = f"{dim}D"
    tissue = getattr(fw, "CardiacTissue" + sfx)(list(shape))
    rng = np.random.default_rng(5)
    mesh = np.ones(shape, dtype=np.int8)
    mesh[rng.random(shape) < 0.15] = 2
    tissue.mesh = mesh
    f = rng.normal(size=(*shape, dim))
    tissue.fibers = f / np.linalg.norm(f, axis=-1, keepdims=True)
    model = getattr(fw, model_name + sfx)()
    model.dt, model.dr, model.t_max, model.prog_bar = 0.01, 0.25, t_max, False
    model.cardiac_tissue = tissue
    model.devices = devices
    volt = -20 if model_name in ("TP06", "LuoRudy91") else 1
    seq = fw.StimSequence()
    if dim == 3:
        seq.add_stim(fw.StimVoltageCoord3D(0, volt, 0, shape[0], 0, 3, 0, shape[2]))
        seq.add_stim(fw.StimCurrentCoord3D(0.5, 5 * volt, 0.3, shape[0] // 2 - 3, shape[0] // 2 + 3,
                                           2, 6, 4, 20))
    else:
        seq.add_stim(fw.StimVoltageCoord2D(0, volt, 0, shape[0], 0, 3))
        seq.add_stim(fw.StimCurrentCoord2D(0.5, 5 * volt, 0.3, shape[0] // 2 - 3, shape[0] // 2 + 3,
                                           4, 20))
    model.stim_sequence = seq
    ts = fw.TrackerSequence()
    act = getattr(fw, "ActivationTime" + sfx + "Tracker")()
    act.threshold, act.step = (-40 if volt < 0 else 0.5), 1
    ecg = getattr(fw, "ECG" + sfx + "Tracker")()
    ecg.measure_coords = np.array([[shape[0] / 2, shape[1] / 2, 40.0], [1.0, 2.0, 3.0]])
    ecg.step = 5
    ap = getattr(fw, "ActionPotential" + sfx + "Tracker")()
    ap.cell_ind = [shape[0] // 2, 3, 8][:dim]
    ap.step = 4
    for tr in (act, ecg, ap):
        ts.add_tracker(tr)
    model.tracker_sequence = ts

    class Kick(fw.Command):
        def execute(self, m):
            m.u[shape[0] // 2 - 2:shape[0] // 2 + 2] *= 0.5

    cs = fw.CommandSequence()
    cs.add_command(Kick(0.8))
    model.command_sequence = cs
    return model, (act, ecg, ap)


def _run_pair(fw, model_name, shape, devices, t_max=1.5):
    ref, rtr = _build(fw, model_name, shape, None, t_max)
    ref.run()
    got, gtr = _build(fw, model_name, shape, devices, t_max)
    got.run()
    assert getattr(got._engine, "multi", False), "the slab-decomposed engine did not engage"
    assert got.gpu_steps == ref.gpu_steps
    for name in got.state_vars:
        assert np.array_equal(got.__dict__[name], ref.__dict__[name]), name
    assert np.array_equal(gtr[0].act_t, rtr[0].act_t)
    a, b = np.array(gtr[1].output), np.array(rtr[1].output)
    assert a.shape == b.shape and len(a) > 0
    assert np.max(np.abs(a - b)) <= 1e-12 * max(1.0, np.max(np.abs(b)))
    assert np.array_equal(np.asarray(gtr[2].output), np.asarray(rtr[2].output))
    return got


@pytest.mark.parametrize("model_name,shape", [("AlievPanfilov", (24, 10, 32)),
                                              ("TP06", (16, 8, 32)),
                                              ("FentonKarma", (40, 64))])
def test_slabs_of_one_process_match_single_gpu(model_name, shape):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import finitewave_b200 as fw
    _run_pair(fw, model_name, shape, [0, 0])
    _run_pair(fw, model_name, shape, [0, 0, 0])


def test_slabs_on_different_gpus_match_single_gpu():
    import torch
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    if n < 2:
        pytest.skip("needs two GPUs")
    import finitewave_b200 as fw
    devices = list(range(min(n, 4)))
    lines = []
    for model_name, shape in (("TP06", (32, 8, 32)), ("AlievPanfilov", (48, 10, 32))):
        m = _run_pair(fw, model_name, shape, devices)
        lines.append(f"{model_name} {shape} on devices {devices}: bit-identical to one GPU, "
                     f"{m.gpu_steps} steps, {m.gpu_launches} launches")
    try:
        from pathlib import Path
        log = Path(__file__).resolve().parent.parent / "gpurun_out" / "r2_multidevice_test.log"
        log.parent.mkdir(exist_ok=True)
        log.write_text("\n".join(lines) + "\n")
    except OSError:
        pass
