"""
Usage-scenario fixtures from the LIVE reference (build container only): what the reference
does when a model is used in the less obvious ways its API allows -- a Command that shortens
or lengthens ``t_max`` mid-run, parameters / ``dt`` / arrays edited between ``run()`` and
``run(initialize=False)``, a tracker window, a second full ``run()``, Commands that edit the
mesh or the conductivity and recompute the weights, a clone that continues, stimuli / trackers
added or removed between runs, a user-defined Stencil, two models on one tissue, a StateSaver
that fires before the run is continued.

    python tests/golden/make_scenario_golden.py

Writes tests/golden/scenarios.npz (u, v, the action-potential trace, step and t per
scenario).  The scenario functions take the package as argument and are shared with the
tests: tests/test_host_loop.py runs them on the CPU test double, tests/test_gpu_fixtures.py
on the device.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))


def _base(fw, t_max=2.0):
    tissue = fw.CardiacTissue2D([24, 14])
    m = fw.AlievPanfilov2D()
    m.dt, m.dr, m.t_max, m.prog_bar = 0.01, 0.25, t_max, False
    m.cardiac_tissue = tissue
    stims = fw.StimSequence()
    stims.add_stim(fw.StimVoltageCoord2D(0, 1, 0, 4, 0, 14))
    stims.add_stim(fw.StimCurrentCoord2D(0.8, 5, 0.4, 10, 14, 0, 14))
    m.stim_sequence = stims
    ap = fw.ActionPotential2DTracker()
    ap.cell_ind, ap.step = [12, 7], 2
    seq = fw.TrackerSequence()
    seq.add_tracker(ap)
    m.tracker_sequence = seq
    return m, ap


def command_shortens_t_max(fw):
    m, ap = _base(fw)

    class Stop(fw.Command):
        def execute(self, model):
            model.t_max = 0.6
    cs = fw.CommandSequence()
    cs.add_command(Stop(0.5))
    m.command_sequence = cs
    m.run()
    return m, ap


def command_lengthens_t_max(fw):
    m, ap = _base(fw, 1.0)

    class More(fw.Command):
        def execute(self, model):
            model.t_max = 3.0          # the iteration count was fixed at loop entry
    cs = fw.CommandSequence()
    cs.add_command(More(0.5))
    m.command_sequence = cs
    m.run()
    return m, ap


def parameters_changed_between_runs(fw):
    m, ap = _base(fw, 1.0)
    m.run()
    m.a, m.k, m.t_max = 0.15, 7.0, 2.0
    m.run(initialize=False)
    return m, ap


def dt_changed_between_runs(fw):
    m, ap = _base(fw, 1.0)
    m.run()
    m.dt, m.t_max = 0.005, 1.5         # the weights keep the old dt (no recompute)
    m.run(initialize=False)
    return m, ap


def tracker_window(fw):
    m, ap = _base(fw, 2.0)
    ap.start_time, ap.end_time = 0.5, 1.25
    m.run()
    return m, ap


def second_full_run(fw):
    m, ap = _base(fw, 1.0)
    m.run()
    m.run()                            # re-initialises the arrays; the trace keeps growing
    return m, ap


def arrays_edited_between_runs(fw):
    m, ap = _base(fw, 1.0)
    m.run()
    m.u[5:9, 3:6] = 0.9
    m.v[...] *= 0.5
    m.t_max = 1.5
    m.run(initialize=False)
    return m, ap


def command_edits_mesh(fw):
    m, ap = _base(fw, 1.5)

    class Scar(fw.Command):
        def execute(self, model):
            model.cardiac_tissue.mesh[8:12, 2:12] = 2
            model.compute_weights()
    cs = fw.CommandSequence()
    cs.add_command(Scar(0.7))
    m.command_sequence = cs
    m.run()
    return m, ap


def clone_continues(fw):
    m, _ = _base(fw, 0.9)
    m.run()
    twin = m.clone()
    twin.t_max = 1.6
    twin.run(initialize=False)
    return twin, twin.tracker_sequence.sequence[0]


def stimulus_added_between_runs(fw):
    m, ap = _base(fw, 1.0)
    m.run()
    m.stim_sequence.add_stim(fw.StimVoltageCoord2D(1.2, 1, 18, 22, 0, 14))
    m.t_max = 1.6
    m.run(initialize=False)
    return m, ap


def stimuli_and_trackers_removed_between_runs(fw):
    m, ap = _base(fw, 0.9)
    m.run()
    m.stim_sequence.remove_stim()
    m.tracker_sequence.remove_trackers()
    m.t_max = 1.5
    m.run(initialize=False)
    return m, ap


def command_changes_conductivity(fw):
    """examples/basics/2D/change_conductivity_2d.py"""
    m, ap = _base(fw, 1.5)

    class Slow(fw.Command):
        def execute(self, model):
            c = np.ones(model.cardiac_tissue.mesh.shape)
            c[:, 7:] = 0.3
            model.cardiac_tissue.conductivity = c
            model.compute_weights()
    cs = fw.CommandSequence()
    cs.add_command(Slow(0.4))
    m.command_sequence = cs
    m.run()
    return m, ap


def user_defined_stencil(fw):
    m, ap = _base(fw, 1.0)

    class Leaky(fw.IsotropicStencil2D):
        def compute_weights(self, model, tissue):
            w = np.array(super().compute_weights(model, tissue)).copy()
            w[..., 2] -= 0.01
            return w
    m.stencil = Leaky()
    m.run()
    return m, ap


def two_models_on_one_tissue(fw):
    m, ap = _base(fw, 0.8)
    m.run()
    other = fw.Barkley2D()
    other.dt, other.dr, other.t_max, other.prog_bar = 0.01, 0.25, 0.5, False
    other.cardiac_tissue = m.cardiac_tissue
    stims = fw.StimSequence()
    stims.add_stim(fw.StimVoltageCoord2D(0, 1, 0, 4, 0, 14))
    other.stim_sequence = stims
    other.run()
    m.t_max = 1.2
    m.run(initialize=False)
    m.v = m.v + other.v                # both models end up in the compared arrays
    return m, ap


def saver_fires_then_run_continues(fw):
    import os
    import tempfile
    d = tempfile.mkdtemp()
    m, ap = _base(fw, 1.0)
    m.state_saver = fw.StateSaver(os.path.join(d, "s"), time=0.5)
    m.run()
    m.t_max = 1.5
    m.run(initialize=False)
    m.v = m.v + np.load(os.path.join(d, "s", "u.npy"))
    return m, ap


def _with_trace(fw, m, cell, dim=2):
    ap = (fw.ActionPotential2DTracker if dim == 2 else fw.ActionPotential3DTracker)()
    ap.cell_ind = cell
    seq = fw.TrackerSequence()
    seq.add_tracker(ap)
    m.tracker_sequence = seq
    return ap


def special_boundaries_and_area_stimulus_3d(fw):
    tissue = fw.CardiacTissue3D([10, 9, 8])
    sb = np.zeros((10, 9, 8), dtype=np.int8)
    sb[4:6, 3:5, 2:6] = 1
    tissue.special_boundaries = sb
    m = fw.MitchellSchaeffer3D()
    m.dt, m.dr, m.t_max, m.prog_bar = 0.01, 0.25, 1.0, False
    m.cardiac_tissue = tissue
    st = fw.StimCurrentArea3D(0.1, 6, 0.3, u_max=0.8)
    st.add_stim_point([3, 3, 3], tissue.mesh, size=2.5)
    stims = fw.StimSequence()
    stims.add_stim(st)
    m.stim_sequence = stims
    ap = _with_trace(fw, m, [[3, 3, 3], [5, 4, 4]], dim=3)
    m.run()
    m.v = m.h
    return m, ap


def luo_rudy_with_edited_parameters(fw):
    m = fw.LuoRudy912D()
    m.dt, m.dr, m.t_max, m.prog_bar = 0.01, 0.25, 2.0, False
    m.cardiac_tissue = fw.CardiacTissue2D([12, 8])
    m.gna, m.init_u, m.init_m, m.D_model = 20.0, -80.0, 0.01, 0.2
    stims = fw.StimSequence()
    stims.add_stim(fw.StimCurrentCoord2D(0, 40, 1.0, 0, 4, 0, 8))
    m.stim_sequence = stims
    ap = _with_trace(fw, m, [6, 4])
    m.run()
    m.v = m.m
    return m, ap


def odd_inputs(fw):
    """Float, non-contiguous mesh; fibres that are not unit vectors; scalar conductivity;
    t_max that is no multiple of dt; a stimulus in the past, one never reached, and a current
    stimulus of zero duration."""
    grid = np.ones((9, 14))
    grid[3:5, 2:7] = 2.0
    tissue = fw.CardiacTissue2D([14, 9])
    tissue.mesh = grid.T
    rng = np.random.default_rng(5)
    f = rng.normal(size=(14, 9, 2))
    tissue.fibers = 1.3 * f / np.linalg.norm(f, axis=-1, keepdims=True)
    tissue.conductivity = 0.5
    m = fw.FentonKarma2D()
    m.dt, m.dr, m.t_max, m.prog_bar = 0.01, 0.25, 0.333, False
    m.cardiac_tissue = tissue
    stims = fw.StimSequence()
    stims.add_stim(fw.StimVoltageCoord2D(-1.0, 1, 0, 3, 0, 9))
    stims.add_stim(fw.StimVoltageCoord2D(5.0, 1, 0, 3, 0, 9))
    stims.add_stim(fw.StimCurrentCoord2D(0.1, 3.0, 0.0, 8, 12, 0, 9))
    m.stim_sequence = stims
    ap = _with_trace(fw, m, [7, 4])
    m.run()
    return m, ap


def stimulus_on_fibrotic_nodes_only(fw):
    tissue = fw.CardiacTissue2D([12, 10])
    tissue.mesh[2:5, :] = 2
    m = fw.Barkley2D()
    m.dt, m.dr, m.t_max, m.prog_bar = 0.01, 0.25, 0.5, False
    m.cardiac_tissue = tissue
    stims = fw.StimSequence()
    stims.add_stim(fw.StimVoltageCoord2D(0, 1, 2, 5, 0, 10))
    m.stim_sequence = stims
    ap = _with_trace(fw, m, [3, 3])
    m.run()
    return m, ap


def tp06_with_edited_parameters(fw):
    m = fw.TP062D()
    m.dt, m.dr, m.t_max, m.prog_bar = 0.01, 0.25, 1.0, False
    m.cardiac_tissue = fw.CardiacTissue2D([10, 6])
    m.gkr, m.gks, m.init_cai, m.ko = 0.1, 0.5, 0.0001, 4.0
    stims = fw.StimSequence()
    stims.add_stim(fw.StimVoltageCoord2D(0, -20, 0, 3, 0, 6))
    m.stim_sequence = stims
    ap = _with_trace(fw, m, [5, 3])
    m.run()
    m.v = m.cai + m.Ki
    return m, ap


SCENARIOS = [command_shortens_t_max, command_lengthens_t_max, parameters_changed_between_runs,
             dt_changed_between_runs, tracker_window, second_full_run,
             arrays_edited_between_runs, command_edits_mesh, clone_continues,
             stimulus_added_between_runs, stimuli_and_trackers_removed_between_runs,
             command_changes_conductivity, user_defined_stencil, two_models_on_one_tissue,
             saver_fires_then_run_continues, special_boundaries_and_area_stimulus_3d,
             luo_rudy_with_edited_parameters, odd_inputs, stimulus_on_fibrotic_nodes_only,
             tp06_with_edited_parameters]
# bit-exact on the device as well (no transcendental functions in the model)
DEVICE_EXACT = {f.__name__ for f in SCENARIOS} - {
    "luo_rudy_with_edited_parameters", "odd_inputs", "tp06_with_edited_parameters"}


def outputs(m, ap):
    return dict(u=np.array(m.u), v=np.array(m.v), ap=np.array(ap.output, dtype=np.float64),
                step=np.int64(m.step), t=np.float64(m.t))


def main():
    from make_golden import import_reference
    fw = import_reference()
    out = {}
    for f in SCENARIOS:
        m, ap = f(fw)
        for k, v in outputs(m, ap).items():
            out[f"{f.__name__}.{k}"] = v
        print(f"{f.__name__:34s} step={m.step} samples={len(ap.output)} u.sum={m.u.sum():.12g}")
    np.savez_compressed(HERE / "scenarios.npz", **out)


if __name__ == "__main__":
    main()
