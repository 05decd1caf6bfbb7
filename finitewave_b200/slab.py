"""
Slab decomposition of a tissue across GPUs (one process per GPU).

The reference is single-process; this is the B200 counterpart of
"a bigger machine": the tissue is cut along axis 0 (the slowest-varying axis of
the C-ordered arrays, so a halo slice is one contiguous block), every rank stores
its owned slices plus one ghost slice per neighbour, and the fused step kernel of
the slab-boundary blocks stores ``u_new`` straight into the neighbour's ghost slice
through a CUDA-IPC peer mapping (NVLink) and raises a flag there -- no collective,
no host round trip on the step path (include/finitewave_b200.h, "Slab
decomposition").  ``torch.distributed`` is used only to exchange the IPC handles at
set-up and to reduce small tracker outputs at the end.

To decompose along another axis, store the tissue transposed (axis of the cut
first); the stencils treat all axes alike.
"""
import ctypes

import numpy as np

from . import _lib


def partition(n_slices, world):
    """Owned global slice ranges [a, b) per rank: contiguous, balanced, covering
    [0, n_slices)."""
    if world < 1 or n_slices < world:
        raise ValueError("need at least one slice per rank")
    base, extra = divmod(n_slices, world)
    out, a = [], 0
    for r in range(world):
        b = a + base + (1 if r < extra else 0)
        out.append((a, b))
        a = b
    return out


def stored_range(owned, n_slices):
    """Slices a rank stores: its owned range plus one ghost slice per existing
    neighbour.  -> (lo, hi, (halo_lo, halo_hi))"""
    a, b = owned
    halo = (a > 0, b < n_slices)
    return a - (1 if halo[0] else 0), b + (1 if halo[1] else 0), halo


def owned_view(arr, halo):
    """Strip the ghost slices from a stored-range array."""
    lo = 1 if halo[0] else 0
    hi = arr.shape[0] - (1 if halo[1] else 0)
    return arr[lo:hi]


def connect_in_process(sims):
    """Wire slabs that live in ONE process (same or different devices with peer
    access): raw pointers instead of IPC handles.  Used by the single-GPU tests."""
    ptrs = [s.halo_pointers() if any(s.halo) else None for s in sims]
    for r, s in enumerate(sims):
        if not any(s.halo):
            continue
        s.halo_connect(lo=ptrs[r - 1] if s.halo[0] else None,
                       hi=ptrs[r + 1] if s.halo[1] else None)


def connect_distributed(sim, rank, world, dist):
    """Exchange CUDA IPC handles with the neighbour ranks and wire the halo.
    Returns the opened peer mappings (keep them alive as long as the simulation)."""
    L = _lib.lib()
    everyone = exchange_exports(sim.halo_export() if any(sim.halo) else None, world, dist)
    opened = []

    def open_side(peer):
        out = {}
        for key in ("u0", "u1", "flags"):
            p = ctypes.c_void_p(0)
            _lib.check(L.fwb_ipc_open_handle(everyone[peer][key], ctypes.byref(p)),
                       "fwb_ipc_open_handle")
            out[key] = p.value
            opened.append(p.value)
        out["slices"] = everyone[peer]["slices"]
        return out
    if any(sim.halo):
        sim.halo_connect(lo=open_side(rank - 1) if sim.halo[0] else None,
                         hi=open_side(rank + 1) if sim.halo[1] else None)
    dist.barrier()
    return opened


def exchange_exports(mine, world, dist):
    """all-gather of the per-rank halo descriptions (picklable dicts)."""
    everyone = [None] * world
    dist.all_gather_object(everyone, mine)
    return everyone


def close_peers(opened):
    L = _lib.lib()
    for p in opened:
        L.fwb_ipc_close_handle(ctypes.c_void_p(p))


def gather_owned(sim, arr_local, dist=None):
    """Owned part of a stored-range device array as numpy (rank-local)."""
    return owned_view(arr_local, sim.halo).cpu().numpy()


def global_box_to_local(x1, x2, n_global, offset, n_local):
    """numpy-slice semantics on the global axis, shifted into a slab's stored range."""
    lo, hi, _ = slice(x1, x2).indices(n_global)
    lo, hi = lo - offset, max(lo, hi) - offset
    lo = min(max(lo, 0), n_local)
    return lo, min(max(hi, lo), n_local)


__all__ = ["partition", "stored_range", "owned_view", "connect_in_process",
           "connect_distributed", "close_peers", "global_box_to_local"]
_ = np
