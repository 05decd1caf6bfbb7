"""
Bounded CPU samples of the BASELINE.json workloads for bench.py's cpu_baseline and
``--impl reference`` legs (TEST / MEASUREMENT INFRASTRUCTURE, never imported by
finitewave_b200/).  Same models, stencils, fibrosis rule and fibre fields as the
full-size device workloads (finitewave_b200/workloads.py), on
shapes the host holds and finishes in seconds; the CPU rate is size-independent once
the arrays are out of cache.
"""
import numpy as np


def _fibrosis(shape, density, seed):
    rng = np.random.default_rng(seed)
    mesh = np.ones(shape, dtype=np.int8)
    mesh[rng.random(shape) <= density] = 2
    return mesh


def _rotating_fibers(shape):
    n_i, n_j, n_k = shape
    phi = np.linspace(-np.pi / 3, np.pi / 2, n_k - 2)
    f = np.zeros((n_i, n_j, n_k, 3))
    f[:, :, 1:-1, 0] = np.cos(phi)
    f[:, :, 1:-1, 1] = np.sin(phi)
    return f


def cases_for_bench(workload):
    """-> (case dict for oracle.time_steps, human description of the sample)"""
    if workload == "c2":
        n = 2048
        f = np.empty((n, n, 2))
        f[..., 0], f[..., 1] = np.cos(0.25 * np.pi), np.sin(0.25 * np.pi)
        case = dict(model="fenton_karma", shape=[n, n], dt=0.01, dr=0.25,
                    mesh=_fibrosis([n, n], 0.30, 2), fibers=f,
                    stims=[dict(kind="voltage_coord", t=0, value=1, box=[0, n, 0, 5])])
        return case, f"C2 reduced to {n}x{n} (same FK aniso-9 + 30% fibrosis)"
    if workload == "c3":
        n = 192
        c = n // 2
        case = dict(model="mitchell_schaeffer", shape=[n, n, n], dt=0.01, dr=0.25,
                    stims=[dict(kind="voltage_coord", t=0, value=1,
                                box=[c - 5, c + 5, c - 5, c + 5, c - 5, c + 5])])
        return case, f"C3 reduced to {n}^3 (same MS iso-7)"
    if workload in ("c4", "c5"):
        shape = [128, 128, 64]
        case = dict(model="tp06", shape=shape, dt=0.01, dr=0.25,
                    fibers=_rotating_fibers(shape),
                    stims=[dict(kind="voltage_coord", t=0, value=-20,
                                box=[0, 5, 0, shape[1], 0, shape[2]])])
        return case, "TP06 3D aniso-19 reduced to 128x128x64 slab"
    raise ValueError(workload)
