"""
Stimulation classes.  Same names, constructor signatures and firing rules as the
reference (finitewave/core/stimulation/{stim,stim_current,stim_voltage,stim_sequence}.py,
finitewave/cpuwave2D/stimulation/*.py, finitewave/cpuwave3D/stimulation/*.py).

Every built-in stimulus has two faces:
  * ``_register(engine, model)`` hands a descriptor to the device runner, which
    fires it with the reference's rule (``t >= stim.t and not passed``; then
    ``passed = t >= stim.t + duration``) and applies it on the GPU;
  * ``stimulate(model)`` is the host-array statement of the same edit, used when a
    sequence also contains user-defined stimuli (then the whole sequence runs as
    a host hook) or when user code calls it directly.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import check


class Stim:
    def __init__(self, time, duration=0.0):
        self.t = time
        self.duration = duration
        self.passed = False

    def stimulate(self, model):
        raise NotImplementedError

    def initialize(self, model):
        self.passed = False

    def update_status(self, model):
        self.passed = model.t >= (self.t + self.duration)

    _native = False


class StimVoltage(Stim):
    def __init__(self, time, volt_value, duration=0.0):
        super().__init__(time, duration)
        self.volt_value = volt_value


class StimCurrent(Stim):
    def __init__(self, time, curr_value, duration):
        super().__init__(time, duration)
        self.curr_value = curr_value


class StimSequence:
    def __init__(self):
        self.sequence = []
        self.model = None

    def initialize(self, model):
        self.model = model
        for stim in self.sequence:
            stim.initialize(model)

    def add_stim(self, stim):
        self.sequence.append(stim)

    def remove_stim(self):
        self.sequence = []

    def stimulate_next(self):
        """Host-side firing (used only when the sequence is not fully native)."""
        for stim in self.sequence:
            if self.model.t >= stim.t and not stim.passed:
                stim.stimulate(self.model)
                stim.update_status(self.model)

    def all_native(self):
        return all(getattr(s, "_native", False) for s in self.sequence)


# ---------------------------------------------------------------------------
def _clamp(u_sel, u_max):
    return np.where(u_sel > u_max, u_max, u_sel)


def _u_max_args(u_max):
    return (0, 0.0) if u_max is None else (1, float(u_max))


class _BoxMixin:
    """Half-open index box x1:x2, y1:y2[, z1:z2] restricted to mesh == 1."""
    _native = True

    def _slices(self):
        b = self._box()
        return tuple(slice(b[2 * d], b[2 * d + 1]) for d in range(len(b) // 2))

    def _norm_box(self, shape):
        out = []
        for d, sl in enumerate(self._slices()):
            lo, hi, _ = sl.indices(shape[d])      # numpy slice semantics (negatives, clipping)
            out += [lo, max(lo, hi)]
        return out

    def _register(self, engine, model):
        box = self._norm_box(engine.shape)
        arr = (ctypes.c_int64 * len(box))(*box)
        has, um = _u_max_args(getattr(self, "u_max", None))
        mode, value = self._mode_value()
        rc = engine.L.fwb_sim_add_stim_box(engine.sim, mode, float(self.t), float(self.duration),
                                           float(value), has, um, arr)
        if rc < 0:
            check(rc, "fwb_sim_add_stim_box")
        return rc


class _NodesMixin:
    """Arbitrary node set restricted to mesh == 1 (matrix > 0 or coordinate list)."""
    _native = True

    def _register(self, engine, model):
        nodes = np.ascontiguousarray(self._flat_nodes(model), dtype=np.int64)
        t_nodes = engine.keep(torch.from_numpy(nodes).to(engine.device))
        has, um = _u_max_args(getattr(self, "u_max", None))
        mode, value = self._mode_value()
        vals_p, n_vals = None, 0
        if mode == _lib.STIM_VOLTAGE_LIST:
            vals = np.ascontiguousarray(self.volt_value, dtype=np.float64)
            vals_p = vals.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
            n_vals = len(vals)
            value = 0.0
        rc = engine.L.fwb_sim_add_stim_nodes(
            engine.sim, mode, float(self.t), float(self.duration), float(value), has, um,
            ctypes.c_void_p(t_nodes.data_ptr()), len(nodes), vals_p, n_vals)
        if rc < 0:
            check(rc, "fwb_sim_add_stim_nodes")
        if mode == _lib.STIM_VOLTAGE_LIST:
            # the list index lives in the stimulus object (`self.step`), as in the reference
            check(engine.L.fwb_sim_set_stim_fired(engine.sim, rc, int(self.step)),
                  "fwb_sim_set_stim_fired")
        return rc


# ---- coordinate boxes --------------------------------------------------------
class StimVoltageCoord2D(_BoxMixin, StimVoltage):
    def __init__(self, time, volt_value, x1, x2, y1, y2):
        StimVoltage.__init__(self, time, volt_value)
        self.x1, self.x2, self.y1, self.y2 = x1, x2, y1, y2

    def _box(self):
        return [self.x1, self.x2, self.y1, self.y2]

    def _mode_value(self):
        return _lib.STIM_VOLTAGE, self.volt_value

    def stimulate(self, model):
        sl = self._slices()
        sel = model.cardiac_tissue.mesh[sl] == 1
        model.u[sl][sel] = self.volt_value


class StimVoltageCoord3D(StimVoltageCoord2D):
    def __init__(self, time, volt_value, x1, x2, y1, y2, z1, z2):
        StimVoltage.__init__(self, time, volt_value)
        self.x1, self.x2, self.y1, self.y2, self.z1, self.z2 = x1, x2, y1, y2, z1, z2

    def _box(self):
        return [self.x1, self.x2, self.y1, self.y2, self.z1, self.z2]


class StimCurrentCoord2D(_BoxMixin, StimCurrent):
    def __init__(self, time, curr_value, duration, x1, x2, y1, y2, u_max=None):
        StimCurrent.__init__(self, time, curr_value, duration)
        self.x1, self.x2, self.y1, self.y2 = x1, x2, y1, y2
        self.u_max = u_max

    def _box(self):
        return [self.x1, self.x2, self.y1, self.y2]

    def _mode_value(self):
        return _lib.STIM_CURRENT, self.curr_value

    def stimulate(self, model):
        sl = self._slices()
        sel = model.cardiac_tissue.mesh[sl] == 1
        model.u[sl][sel] += model.dt * self.curr_value
        if self.u_max is not None:
            model.u[sl][sel] = _clamp(model.u[sl][sel], self.u_max)


class StimCurrentCoord3D(StimCurrentCoord2D):
    def __init__(self, time, curr_value, duration, x1, x2, y1, y2, z1, z2, u_max=None):
        StimCurrent.__init__(self, time, curr_value, duration)
        self.x1, self.x2, self.y1, self.y2, self.z1, self.z2 = x1, x2, y1, y2, z1, z2
        self.u_max = u_max

    def _box(self):
        return [self.x1, self.x2, self.y1, self.y2, self.z1, self.z2]


# ---- matrices ----------------------------------------------------------------
class _MatrixMixin(_NodesMixin):
    def _mask(self, model):
        return (np.asarray(self.matrix) > 0) & (model.cardiac_tissue.mesh == 1)

    def _flat_nodes(self, model):
        return np.flatnonzero(self._mask(model))


class StimVoltageMatrix2D(_MatrixMixin, StimVoltage):
    def __init__(self, time, volt_value, matrix):
        StimVoltage.__init__(self, time, volt_value)
        self.matrix = matrix

    def _mode_value(self):
        return _lib.STIM_VOLTAGE, self.volt_value

    def stimulate(self, model):
        model.u[self._mask(model)] = self.volt_value


class StimVoltageMatrix3D(StimVoltageMatrix2D):
    pass


class StimCurrentMatrix2D(_MatrixMixin, StimCurrent):
    def __init__(self, time, curr_value, duration, matrix, u_max=None):
        StimCurrent.__init__(self, time, curr_value, duration)
        self.matrix = matrix
        self.u_max = u_max

    def _mode_value(self):
        return _lib.STIM_CURRENT, self.curr_value

    def stimulate(self, model):
        sel = self._mask(model)
        model.u[sel] += model.dt * self.curr_value
        if self.u_max is not None:
            model.u[sel] = _clamp(model.u[sel], self.u_max)


class StimCurrentMatrix3D(StimCurrentMatrix2D):
    pass


class StimVoltageListMatrix3D(_MatrixMixin, StimVoltage):
    """Voltage taken from a list, one entry per firing (stim_voltage_list_matrix_3d.py)."""

    def __init__(self, time, volt_values, duration, matrix):
        StimVoltage.__init__(self, time, volt_values, duration)
        self.matrix = matrix
        self.step = 0

    def initialize(self, model):
        super().initialize(model)
        total_steps = int(self.duration / model.dt) + 1
        if len(self.volt_value) < total_steps:
            raise ValueError("The length of the voltage values array should be "
                             "greater than the total number of steps.")

    def _mode_value(self):
        return _lib.STIM_VOLTAGE_LIST, 0.0

    def _collect(self, engine, sid):
        fired = int(engine.L.fwb_sim_stim_fired(engine.sim, sid))
        if fired >= 0:
            self.step = fired

    def stimulate(self, model):
        model.u[self._mask(model)] = self.volt_value[self.step]
        self.step += 1


# ---- coordinate lists ----------------------------------------------------------
class StimCurrentArea2D(_NodesMixin, StimCurrent):
    def __init__(self, time, curr_value, duration, coords=None, u_max=None):
        StimCurrent.__init__(self, time, curr_value, duration)
        self.coords = coords
        self.u_max = u_max

    def add_stim_point(self, coord, mesh, size=None):
        self._coord = coord
        if size is None:
            self.coords = np.atleast_2d(coord)
            return
        tissue_points = np.argwhere(mesh == 1)
        dist = np.linalg.norm(tissue_points - coord, axis=1)
        self.coords = tissue_points[dist < size]

    def initialize(self, model):
        coords = np.atleast_2d(self.coords)
        healthy = model.cardiac_tissue.mesh[tuple(coords.T)] == 1
        if healthy.sum() == 0:
            raise ValueError("The specified area does not have healthy cells.")
        self._coords = coords[healthy]
        super().initialize(model)

    def _mode_value(self):
        return _lib.STIM_CURRENT, self.curr_value

    def _flat_nodes(self, model):
        # duplicates would be applied twice on the device but once by numpy fancy
        # indexing (`u[inds] += x` is unbuffered) -> deduplicate to match
        flat = np.ravel_multi_index(tuple(self._coords.T), model.cardiac_tissue.mesh.shape)
        return np.unique(flat)

    def stimulate(self, model):
        inds = tuple(self._coords.T)
        model.u[inds] += model.dt * self.curr_value
        if self.u_max is not None:
            model.u[inds] = _clamp(model.u[inds], self.u_max)


class StimCurrentArea3D(StimCurrentArea2D):
    pass
