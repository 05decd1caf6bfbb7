"""
Synthetic tissues of the BASELINE.json configurations, built on the device.

SURVEY.md section 8d defines the inputs; the small host-side twins used for parity
live in tests/cases.py.  Everything here returns torch tensors on ``device`` so a
512^3 ... 1024^3 tissue never has to exist in host memory.

Algorithmic bytes per myocyte node-update:
    B = 8 (u) + 8 (u_new) + 1 (mask) + 8 K (weights) + 8 (2 S_rw + S_ro + S_wo)
"""
import math

import numpy as np
import torch

from . import model as _m
from .devrun import DeviceSimulation
from .stimulation import StimVoltageCoord2D, StimVoltageCoord3D
from .tracker import ActivationTime3DTracker

BYTES_PER_NODE = {
    ("aliev_panfilov", 5): 73, ("aliev_panfilov", 9): 105,
    ("fenton_karma", 9): 121, ("fenton_karma", 5): 89,
    ("mitchell_schaeffer", 7): 89, ("mitchell_schaeffer", 19): 185,
    ("luo_rudy91", 5): 169, ("luo_rudy91", 7): 185,
    ("tp06", 19): 457, ("tp06", 7): 361,
}


def _ring_zero(mesh, halo=(False, False)):
    """Empty outer ring (CardiacTissue.add_boundaries); a slab keeps its ghost slices."""
    for ax in range(mesh.dim()):
        if not (ax == 0 and halo[0]):
            mesh.select(ax, 0).zero_()
        if not (ax == 0 and halo[1]):
            mesh.select(ax, mesh.shape[ax] - 1).zero_()
    return mesh


def _hash01(idx, seed):
    """Uniform [0, 1) from an int64 tensor of global node ids (murmur3 finaliser): the
    same node gets the same value on whichever rank stores it."""
    m32 = 0xFFFFFFFF
    h = (idx + int(seed) * 0x9E3779B1) & m32
    h = h ^ (h >> 16)
    h = (h * 0x85EBCA6B) & m32
    h = h ^ (h >> 13)
    h = (h * 0xC2B2AE35) & m32
    h = h ^ (h >> 16)
    return h.to(torch.float64) / 4294967296.0


def fibrosis_mesh(shape, density, seed, device, slow_offset=0, halo=(False, False)):
    """mesh with 2 where uniform(0,1) <= density (rule of diffuse_2d_pattern.py:86-87);
    ``shape`` is the stored shape of a slab starting at global slice ``slow_offset``."""
    mesh = torch.ones(shape, dtype=torch.int8, device=device)
    if density > 0:
        per_slice = int(np.prod(shape[1:]))
        step = max(1, (1 << 22) // per_slice)   # a few M nodes at a time: small temporaries
        for i in range(0, shape[0], step):
            cnt = min(step, shape[0] - i) * per_slice
            idx = torch.arange(cnt, dtype=torch.int64, device=device) + \
                (i + slow_offset) * per_slice
            mesh[i:i + step].view(-1)[_hash01(idx, seed) <= density] = 2
    return _ring_zero(mesh, halo)


def uniform_fibers_2d(shape, alpha, device):
    f = torch.empty((*shape, 2), dtype=torch.float64, device=device)
    f[..., 0] = math.cos(alpha)
    f[..., 1] = math.sin(alpha)
    return f


def rotating_fibers_3d(shape, device, s0=0, n_total=None):
    """The rotating-fibre slab of examples/basics/3D/slab_with_fibers_3d.py:64-70 with the
    rotation axis ("z" there) stored as axis 0 -- the slowest axis, along which slabs are cut,
    so that a halo slice is one contiguous block: the in-plane fibre angle goes from -pi/3 to
    pi/2 across the interior slices of the whole tissue, fibres lie in the (axis 1, axis 2)
    plane.  ``s0`` / ``n_total`` select the slab [s0, s0 + shape[0]) of a thicker tissue."""
    n_s, n_j, n_k = shape
    n_total = n_total or n_s
    phi = torch.linspace(-math.pi / 3, math.pi / 2, n_total - 2, dtype=torch.float64, device=device)
    full = torch.zeros(n_total, 2, dtype=torch.float64, device=device)
    full[1:-1, 0] = torch.cos(phi)
    full[1:-1, 1] = torch.sin(phi)
    sl = full[s0:s0 + n_s]
    f = torch.zeros((n_s, n_j, n_k, 3), dtype=torch.float64, device=device)
    f[..., 1] = sl[:, 0].view(-1, 1, 1)
    f[..., 2] = sl[:, 1].view(-1, 1, 1)
    return f


def ventricle_shell(shape, device, slow_offset=0, n_global=None, halo=(False, False)):
    """Half prolate-ellipsoid shell + helix fibres (-60..+60 deg across the wall),
    the geometry of tests/cases.py::ventricle_shell evaluated plane by plane on the
    device (examples/data/*.npy are not in the reference tree)."""
    n_loc, n_j, n_k = shape
    n_i = n_global or n_loc
    mesh = torch.zeros(shape, dtype=torch.int8, device=device)
    fib = torch.zeros((*shape, 3), dtype=torch.float64, device=device)
    ci, cj, ck = n_i / 2, n_j / 2, 0.86 * n_k
    jj, kk = torch.meshgrid(torch.arange(n_j, dtype=torch.float64, device=device),
                            torch.arange(n_k, dtype=torch.float64, device=device), indexing="ij")
    y, z = jj - cj, kk - ck
    for i in range(n_loc):
        x = float(i + slow_offset) - ci
        outer = (x / (0.39 * n_i)) ** 2 + (y / (0.39 * n_j)) ** 2 + (z / (0.82 * n_k)) ** 2
        inner = (x / (0.31 * n_i)) ** 2 + (y / (0.31 * n_j)) ** 2 + (z / (0.74 * n_k)) ** 2
        wall = (outer <= 1.0) & (inner > 1.0) & (kk <= ck)
        if not bool(wall.any()):
            continue
        mesh[i][wall] = 1
        si, so = torch.sqrt(inner), torch.sqrt(outer)
        depth = torch.clamp((si - 1.0) / torch.clamp(si - so, min=1e-9), 0.0, 1.0)
        helix = torch.deg2rad(-60.0 + 120.0 * depth)
        r = torch.sqrt(x * x + y * y) + 1e-12
        f = torch.stack([torch.cos(helix) * (-y / r), torch.cos(helix) * (x / r),
                         torch.sin(helix)], dim=-1)
        f = f / torch.linalg.norm(f, dim=-1, keepdim=True)
        f[~wall] = 0.0
        fib[i] = f
    return _ring_zero(mesh, halo), fib


def _cfg(model, dt=0.01, dr=0.25):
    model.dt, model.dr = dt, dr
    model.prog_bar = False
    return model


def build(name, device, scale=1.0, rank=0, world=1, dist=None):
    """Returns (DeviceSimulation, info dict).  ``scale`` shrinks every axis (tests).
    With world > 1 the tissue is ``world`` times longer along axis 0 (weak scaling) and
    this rank builds and owns slab ``rank`` of it."""
    from . import slab

    def r32(x, lo=32):
        return max(lo, int(round(x * scale)) // 32 * 32)

    if name == "c2":
        n = r32(4096, 64)
        own, rest, dim = n, (n,), 2
    elif name in ("c3", "c4"):
        n = r32(512)
        own, rest, dim = n, (n, n), 3
    elif name in ("c5", "c5t"):
        n = r32(1024)
        own, rest, dim = r32(128), (n, n), 3
    elif name == "lr91":
        # not a BASELINE config: the remaining FP64-bound on-path model, same size as C2
        n = r32(4096, 64)
        own, rest, dim = n, (n,), 2
    else:
        raise ValueError(f"unknown workload {name}")
    n_global = own * world
    a, b = slab.partition(n_global, world)[rank]
    lo, hi, halo = slab.stored_range((a, b), n_global)
    shape = (hi - lo, *rest)
    kw = dict(halo=halo, slow_offset=lo, global_slices=n_global)

    if name == "c2":
        mesh = fibrosis_mesh(shape, 0.30, 2, device, lo, halo)
        sim = DeviceSimulation(_cfg(_m.FentonKarma2D()), mesh,
                               fibers=uniform_fibers_2d(shape, 0.25 * math.pi, device), **kw)
        sim.add_stim(StimVoltageCoord2D(0, 1, 0, n_global, 0, 5))
        sim.add_stim(StimVoltageCoord2D(3.0, 1, 0, n_global // 2, 0, n))
        info = dict(workload=f"C2 Fenton-Karma 2D {n_global}x{n} aniso 9-pt, 30% random fibrosis",
                    model="fenton_karma", K=9)
    elif name == "lr91":
        mesh = fibrosis_mesh(shape, 0.0, 0, device, lo, halo)
        sim = DeviceSimulation(_cfg(_m.LuoRudy912D()), mesh, **kw)
        sim.add_stim(StimVoltageCoord2D(0, -20, 0, n_global, 0, 5))
        info = dict(workload=f"LR91 2D {n_global}x{n} iso 5-pt, planar wave (extra, not a "
                             "BASELINE config)", model="luo_rudy91", K=5)
    elif name == "c3":
        mesh = fibrosis_mesh(shape, 0.0, 0, device, lo, halo)
        sim = DeviceSimulation(_cfg(_m.MitchellSchaeffer3D()), mesh, **kw)
        c, cg = n // 2, n_global // 2
        sim.add_stim(StimVoltageCoord3D(0, 1, cg - 5, cg + 5, c - 5, c + 5, c - 5, c + 5))
        tr = ActivationTime3DTracker()
        tr.threshold, tr.step = 0.5, 1
        sim.add_tracker(tr, 0)
        info = dict(workload=f"C3 Mitchell-Schaeffer 3D {n_global}x{n}x{n} iso 7-pt, focal "
                             "stimulus, activation-time tracker every step",
                    model="mitchell_schaeffer", K=7, tracker_bytes=8)
    elif name == "c4":
        mesh, fib = ventricle_shell(shape, device, lo, n_global, halo)
        sim = DeviceSimulation(_cfg(_m.TP063D()), mesh, fibers=fib, **kw)
        del fib
        sim.add_stim(StimVoltageCoord3D(0, -20, 0, n_global, 0, n, 0, max(4, int(30 * n / 512))))
        info = dict(workload=f"C4 TP06 3D ventricle-shaped shell in {n_global}x{n}x{n}, helix "
                             "fibres, 19-pt", model="tp06", K=19)
    else:
        mesh = fibrosis_mesh(shape, 0.0, 0, device, lo, halo)
        fib = rotating_fibers_3d(shape, device, lo, n_global)
        sim = DeviceSimulation(_cfg(_m.TP063D()), mesh, fibers=fib, **kw)
        del fib
        # a face perpendicular to the cut: the wave runs along every slab at once
        sim.add_stim(StimVoltageCoord3D(0, -20, 0, n_global, 0, 5, 0, n))
        info = dict(workload=f"C5 TP06 3D {n_global}x{n}x{n} aniso 19-pt slab, cut along the "
                             f"fibre-rotation axis ({own} slices per GPU; 1024^3 at 8 GPUs), "
                             "face stimulus", model="tp06", K=19)
        if name == "c5t":
            # TP06 as it runs in practice: with the activation-time map sampled every step
            tr = ActivationTime3DTracker()
            tr.threshold, tr.step = -40, 1
            sim.add_tracker(tr, 0)
            info["workload"] += " + ActivationTime3DTracker every step"
            info["tracker_bytes"] = 8
    info["shape"] = list(sim.shape)
    info["n_myo"] = int(sim.n_myo)
    info["bytes_per_node"] = BYTES_PER_NODE[(info["model"], info["K"])] + info.get("tracker_bytes", 0)
    info["peers"] = None
    if world > 1:
        info["peers"] = slab.connect_distributed(sim, rank, world, dist)
    torch.cuda.empty_cache()
    return sim, info
