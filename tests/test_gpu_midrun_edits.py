"""
Stimulus / tracker sequences edited WHILE a run is in progress (by a Command), and stimulus
state that must survive a continued run -- the reference re-reads its sequences every step
(core/stimulation/stim_sequence.py:65-77, core/tracker/tracker_sequence.py:59-64) and keeps
`StimVoltageListMatrix3D.step` in the stimulus object (stim_voltage_list_matrix_3d.py:60-72);
the device registration has to follow.  Each case is compared, bit for bit, with an equivalent
run in which nothing is edited on the way.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _model(fw, shape=(40, 36)):
    tissue = fw.CardiacTissue2D(list(shape))
    model = fw.AlievPanfilov2D()
    model.dt, model.dr, model.t_max, model.prog_bar = 0.01, 0.25, 3.0, False
    model.cardiac_tissue = tissue
    return model


def test_command_adds_and_rearms_stimuli_mid_run():
    import finitewave_b200 as fw
    # (a) a Command at t = 1 adds stimulus B (fires at 1.5) and re-arms A for t = 2
    a = _model(fw)
    sa = fw.StimVoltageCoord2D(0, 1, 1, 39, 1, 4)
    seq = fw.StimSequence()
    seq.add_stim(sa)
    a.stim_sequence = seq
    ap = fw.ActionPotential2DTracker()
    ap.cell_ind = [20, 18]
    ts = fw.TrackerSequence()
    ts.add_tracker(ap)
    a.tracker_sequence = ts

    class Edit(fw.Command):
        def execute(self, model):
            model.stim_sequence.add_stim(fw.StimVoltageCoord2D(1.5, 1, 10, 20, 20, 30))
            sa.passed = False
            sa.t = 2.0
            extra = fw.ActivationTime2DTracker()
            extra.threshold, extra.step = 0.5, 1
            extra.initialize(model)
            model.tracker_sequence.add_tracker(extra)
            model._extra = extra

    cs = fw.CommandSequence()
    cs.add_command(Edit(1.0))
    a.command_sequence = cs
    a.run()

    # (b) the same three firings declared up front
    b = _model(fw)
    seq = fw.StimSequence()
    seq.add_stim(fw.StimVoltageCoord2D(0, 1, 1, 39, 1, 4))
    seq.add_stim(fw.StimVoltageCoord2D(1.5, 1, 10, 20, 20, 30))
    seq.add_stim(fw.StimVoltageCoord2D(2.0, 1, 1, 39, 1, 4))
    b.stim_sequence = seq
    bp = fw.ActionPotential2DTracker()
    bp.cell_ind = [20, 18]
    act = fw.ActivationTime2DTracker()
    act.threshold, act.step, act.start_time = 0.5, 1, 1.0
    ts = fw.TrackerSequence()
    ts.add_tracker(bp)
    ts.add_tracker(act)
    b.tracker_sequence = ts
    b.run()

    assert np.array_equal(a.u, b.u) and np.array_equal(a.v, b.v)
    assert np.array_equal(np.asarray(ap.output), np.asarray(bp.output))
    assert all(s.passed for s in a.stim_sequence.sequence)
    # the tracker added by the Command samples from the step after the Command fired, which
    # is the first step with t >= 1.0 -- the window of the up-front tracker
    assert np.array_equal(a._extra.act_t, act.act_t) and (act.act_t >= 0).any()


def test_voltage_list_index_survives_a_continued_run():
    import finitewave_b200 as fw
    shape = (12, 10, 32)
    mat = np.zeros(shape)
    mat[3:6, 2:8, 4:20] = 1
    volts = list(np.linspace(0.2, 1.0, 60))

    def make():
        tissue = fw.CardiacTissue3D(list(shape))
        m = fw.AlievPanfilov3D()
        m.dt, m.dr, m.prog_bar = 0.01, 0.25, False
        m.cardiac_tissue = tissue
        seq = fw.StimSequence()
        st = fw.StimVoltageListMatrix3D(0.1, volts, 0.4, mat)
        seq.add_stim(st)
        m.stim_sequence = seq
        return m, st

    one, st1 = make()
    one.t_max = 1.0
    one.run()
    two, st2 = make()
    two.t_max = 0.3                       # stops in the middle of the pulse
    two.run()
    mid = st2.step
    assert 0 < mid < st1.step
    two.t_max = 1.0
    two.run(initialize=False)
    assert st2.step == st1.step
    assert np.array_equal(one.u, two.u) and np.array_equal(one.v, two.v)
