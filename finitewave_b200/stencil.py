"""
Stencil classes (host side of the weights path).

Same names and attributes as the reference's Stencil hierarchy
(finitewave/core/stencil/stencil.py:4-46; IsotropicStencil2D/3D
cpuwave2D/stencil/isotropic_stencil_2d.py:7-95, cpuwave3D/stencil/isotropic_stencil_3d.py:10-74;
AsymmetricStencil2D/3D cpuwave2D/stencil/asymmetric_stencil_2d.py:7-161,
cpuwave3D/stencil/asymmetric_stencil_3d.py:11-105), including ``D_al`` / ``D_ac`` and the
slot orders documented there.

``compute_weights(model, tissue)`` runs the device weights kernel
(fwb_compute_weights) and returns a ``DeviceWeights`` handle: the weights stay on
the GPU in compact SoA form and are only materialised as the reference's
``(*shape, K)`` ndarray when host code looks at them (``np.asarray(model.weights)``
or indexing).  A user-defined Stencil may instead return a plain ndarray; the
model packs it onto the device.
"""
import numpy as np

from . import _lib


class DeviceWeights:
    """Lazy host view of device-resident stencil weights."""

    def __init__(self, engine):
        self._engine = engine
        self._host = None

    def _materialise(self):
        if self._host is None:
            self._host = self._engine.weights_dense()
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self._materialise()
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, idx):
        return self._materialise()[idx]

    @property
    def shape(self):
        return (*self._engine.shape, self._engine.K)

    @property
    def dtype(self):
        return np.dtype(np.float64)

    def __repr__(self):
        return f"DeviceWeights(shape={self.shape}, device={self._engine.device})"


class Stencil:
    """Base class: a stencil turns tissue + model numerics into per-node weights."""

    def compute_weights(self, model, cardiac_tissue):
        raise NotImplementedError

    def select_diffusion_kernel(self):
        raise NotImplementedError


class _BuiltinStencil(Stencil):
    _KIND = None
    _DIM = None

    def compute_weights(self, model, cardiac_tissue):
        if np.dtype(model.npfloat) != np.float64:
            raise NotImplementedError("finitewave_b200 computes in float64 only")
        eng = model._engine_for(cardiac_tissue)
        if self._KIND != _lib.STENCIL_ISO and cardiac_tissue.fibers is None:
            raise ValueError("Fibers must be provided for anisotropic diffusion.")
        eng.compute_weights(self._KIND, cardiac_tissue.conductivity,
                            cardiac_tissue.fibers if self._KIND != _lib.STENCIL_ISO else None,
                            getattr(self, "D_al", 1), getattr(self, "D_ac", 1 / 9),
                            model.D_model, model.dt, model.dr)
        return DeviceWeights(eng)

    def select_diffusion_kernel(self):
        """Callable with the reference's signature ``(u_new, u, w, indexes)`` running
        the device diffusion apply (fwb_diffuse) on host arrays."""
        kind, dim = self._KIND, self._DIM
        if kind == _lib.STENCIL_SYM:
            kind = _lib.STENCIL_ANISO

        def diffusion_kernel(u_new, u, w, indexes):
            from .hostcall import diffuse_host
            return diffuse_host(dim, kind, u_new, u, w, indexes)
        return diffusion_kernel


# The three public helpers below exist in the reference as building blocks of its host-side
# weight computation.  Here the weights are computed by the device kernel (csrc/weights.cu,
# which evaluates the same expressions per node); the helpers are kept as plain numpy
# utilities for user code that calls them, and are not used by this package.
class IsotropicStencil2D(_BuiltinStencil):
    """5-point: slots (i-1,j) (i,j-1) (i,j) (i,j+1) (i+1,j)."""
    _KIND, _DIM = _lib.STENCIL_ISO, 2

    def compute_half_step_diffusion(self, mesh, conductivity, num_axes=None):
        """Conductivity at i + 1/2 along every axis, shape ``(num_axes, *mesh.shape)``
        (isotropic_stencil_2d.py:71-95, isotropic_stencil_3d.py)."""
        num_axes = self._DIM if num_axes is None else num_axes
        c = conductivity * np.ones(np.shape(mesh))
        return np.stack([0.5 * (c + np.roll(c, -1, axis=a)) for a in range(num_axes)])


class IsotropicStencil3D(IsotropicStencil2D):
    """7-point: slots (i-1) (j-1) (k-1) centre (k+1) (j+1) (i+1)."""
    _KIND, _DIM = _lib.STENCIL_ISO, 3


class AsymmetricStencil2D(_BuiltinStencil):
    """9-point, row-major 3x3 slot order, centre = 4; D = D_ac*I + (D_al-D_ac) f f^T."""
    _KIND, _DIM = _lib.STENCIL_ANISO, 2

    def __init__(self):
        self.D_al = 1
        self.D_ac = 1 / 9

    def compute_diffusion_components(self, fibers, ind0, ind1, D_al, D_ac):
        """``D_ac * delta + (D_al - D_ac) f_ind0 f_ind1`` per node
        (asymmetric_stencil_2d.py:138-161)."""
        fibers = np.asarray(fibers)
        return D_ac * (ind0 == ind1) + (D_al - D_ac) * fibers[..., ind0] * fibers[..., ind1]

    def compute_half_step_diffusion(self, mesh, conductivity, fibers, axis, num_axes=None):
        """Row ``axis`` of the diffusion tensor times the conductivity, averaged onto the
        interface at i + 1/2 along ``axis``: shape ``(num_axes, *mesh.shape)``
        (asymmetric_stencil_2d.py:97-136)."""
        num_axes = self._DIM if num_axes is None else num_axes
        c = conductivity * np.ones(np.shape(mesh))
        out = []
        for b in range(num_axes):
            d = self.compute_diffusion_components(fibers, axis, b, self.D_al, self.D_ac)
            out.append(0.5 * (d * c + np.roll(d, -1, axis=axis) * np.roll(c, -1, axis=axis)))
        return np.stack(out)


class AsymmetricStencil3D(AsymmetricStencil2D):
    """19-point (faces + 12 edges), slot order of asymmetric_stencil_3d.py:26-44."""
    _KIND, _DIM = _lib.STENCIL_ANISO, 3


class SymmetricStencil2D(AsymmetricStencil2D):
    """Cell-centred 9-point scheme (cpuwave2D/stencil/symmetric_stencil_2d.py:7-66); same
    slot order and apply kernel as AsymmetricStencil2D, only the weights differ.  Never
    auto-selected: assign ``model.stencil = SymmetricStencil2D()``."""
    _KIND, _DIM = _lib.STENCIL_SYM, 2

    def compute_half_step_diffusion(self, mesh, conductivity, fibers, D_al, D_ac):
        """The four tensor components times the conductivity, averaged onto the cell centre
        (i + 1/2, j + 1/2): shape ``(4, *mesh.shape)`` in the order xx, xy, yx, yy
        (symmetric_stencil_2d.py:69-118)."""
        c = conductivity * np.ones(np.shape(mesh))
        out = []
        for k in range(4):
            d = self.compute_diffusion_components(fibers, k // 2, k % 2, D_al, D_ac) * c
            right = np.roll(d, -1, axis=1)
            out.append(0.25 * (d + np.roll(d, -1, axis=0) + right + np.roll(right, -1, axis=0)))
        return np.stack(out)
