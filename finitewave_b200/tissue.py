"""
Tissue description (host side).  Same public surface as the reference's
CardiacTissue / CardiacTissue2D / CardiacTissue3D
(finitewave/core/tissue/cardiac_tissue.py:6-100,
 finitewave/cpuwave2D/tissue/cardiac_tissue_2d.py:6-45,
 finitewave/cpuwave3D/tissue/cardiac_tissue_3d.py:6-46):
``mesh`` (0 empty / 1 myocyte / 2 fibrosis, outer ring forced to 0 by the
setter), ``conductivity``, ``fibers``, ``special_boundaries``, ``meta``,
``compute_myo_indexes()``, ``add_boundaries()``, ``add_pattern()``, ``clean()``,
``clone()``.

The device never sees an index list: the engine turns the update mask into a
bit per node (engine.Engine.set_tissue).  ``myo_indexes`` stays available on the
host for user code and is computed lazily.
"""
import copy

import numpy as np


class IncorrectWeightsModeError2D(Exception):
    """Raised for a 2D weights mode other than 'iso' / 'aniso' (reference
    finitewave/cpuwave2D/exception/exceptions_2d.py:1-37: attributes ``mode`` and
    ``message``, and the offending mode appended in ``str()``).  The reference exports the
    class but never raises it (the stencil is chosen from ``tissue.fibers``); kept so that
    user code catching it keeps working."""

    def __init__(self, mode, message="CardiacTissue2D mode attribute must be 'iso' or 'aniso'"):
        super().__init__(message)
        self.mode, self.message = mode, message

    def __str__(self):
        return f"{self.message} (Invalid mode: '{self.mode}')"


class IncorrectWeightsShapeError(Exception):
    """finitewave/core/exception/exceptions.py:1-3 (exported by ``finitewave.core.exception``,
    raised nowhere in the reference)."""


class IncorrectNumberOfWeights(Exception):
    """finitewave/core/exception/exceptions.py:6-39: number of stencil weights that is neither
    of the two expected counts."""

    def __init__(self, number_of_weights, n1, n2):
        self.message = (f"Number of weights provided ({number_of_weights})"
                        f"does not match the expected {n1} or {n2}.")
        super().__init__(self.message)


class CardiacTissue:
    _DIM = None

    def __init__(self, shape):
        self.meta = {"dim": self._DIM, "shape": shape}
        self.special_boundaries = None
        self._myo_indexes = None
        self.mesh = np.ones(shape, dtype=np.int8)
        self.conductivity = 1.0
        self.fibers = None

    # -- mesh with enforced empty outer ring ------------------------------
    @property
    def mesh(self):
        return self._mesh

    @mesh.setter
    def mesh(self, value):
        if value.ndim != self.meta["dim"]:
            raise ValueError("Mesh dimension must match the tissue dimension.")
        self._mesh = value
        self.add_boundaries()

    def add_boundaries(self):
        m = self._mesh
        for axis in range(m.ndim):
            edge = [slice(None)] * m.ndim
            for side in (0, -1):
                edge[axis] = side
                m[tuple(edge)] = 0
        self._myo_indexes = None

    # -- which nodes does the solver update -------------------------------
    def update_mask(self):
        """Boolean (*shape): mesh == 1 and not a special (Dirichlet) node."""
        on = self.mesh == 1
        if self.special_boundaries is not None:
            on &= (self.special_boundaries == 0)
        return on

    def compute_myo_indexes(self):
        self._myo_indexes = None     # recomputed on demand from the current mesh

    @property
    def myo_indexes(self):
        if self._myo_indexes is None:
            self._myo_indexes = np.flatnonzero(self.update_mask())
        return self._myo_indexes

    @myo_indexes.setter
    def myo_indexes(self, value):
        self._myo_indexes = value

    # -- fibrosis helpers -------------------------------------------------
    def add_pattern(self, fibro_pattern):
        fibro_pattern.apply(self)
        self._myo_indexes = None

    def clean(self):
        self.mesh[self.mesh == 2] = 1
        self._myo_indexes = None

    def clone(self):
        return copy.deepcopy(self)


class CardiacTissue2D(CardiacTissue):
    _DIM = 2


class CardiacTissue3D(CardiacTissue):
    _DIM = 3
