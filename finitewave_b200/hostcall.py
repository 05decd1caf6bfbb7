"""
Host-array entry to the device diffusion apply.

``model.diffusion_kernel(u_new, u, weights, indexes)`` is part of the reference's
public seam (finitewave/core/stencil/stencil.py:37-46, cardiac_model.py:205-211;
ECG trackers and user code call it directly).  Here it stages the host arrays on
the GPU, runs fwb_diffuse and copies ``u_new`` back; nodes outside ``indexes`` keep
their previous ``u_new`` value, as in the reference.  It is a convenience path --
the time loop never goes through it (the fused step kernel does the same sum).
"""
import ctypes

import numpy as np
import torch

from ._lib import check, lib, shape_arr
from .engine import require_cuda
from .stencil import DeviceWeights


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def diffuse_host(dim, kind, u_new, u, w, indexes):
    require_cuda()
    L = lib()
    dev = torch.device(f"cuda:{torch.cuda.current_device()}")
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    shape = tuple(u.shape)
    if len(shape) != dim:
        raise ValueError(f"u must be {dim}-dimensional")
    n_nodes = int(np.prod(shape))
    mask = np.zeros(n_nodes, dtype=np.uint8)
    mask[np.asarray(indexes, dtype=np.int64)] = 1
    d_mask = torch.from_numpy(mask).to(dev)
    n_chunks = (n_nodes + 31) // 32
    bits = torch.empty(n_chunks, dtype=torch.int32, device=dev)
    base = torch.empty(n_chunks, dtype=torch.int32, device=dev)
    n_myo = ctypes.c_int64(0)
    check(L.fwb_build_chunks(_p(d_mask), n_nodes, _p(bits), _p(base), ctypes.byref(n_myo),
                             stream), "fwb_build_chunks")
    ld = max(32, (int(n_myo.value) + 31) // 32 * 32)
    shp = shape_arr(shape)
    cap = int(L.fwb_worklist_capacity(dim, shp))
    worklist = torch.empty(cap, dtype=torch.int32, device=dev)
    n_work = ctypes.c_int64(0)
    check(L.fwb_build_worklist(dim, shp, _p(d_mask), 0, 0, _p(worklist), cap, ctypes.byref(n_work),
                               None, None, stream), "fwb_build_worklist")
    w_host = np.ascontiguousarray(np.asarray(w), dtype=np.float64)
    K = w_host.shape[-1]
    if L.fwb_stencil_k(dim, kind) != K or w_host.shape[:-1] != shape:
        raise ValueError(f"weights must have shape (*{shape}, {L.fwb_stencil_k(dim, kind)})")
    d_w = torch.from_numpy(w_host).to(dev)
    d_ws = torch.empty((K, ld), dtype=torch.float64, device=dev)
    check(L.fwb_weights_pack(_p(d_w), _p(d_ws), K, ld, n_nodes, _p(bits), _p(base), stream),
          "fwb_weights_pack")
    d_u = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float64)).to(dev)
    d_un = torch.from_numpy(np.ascontiguousarray(u_new, dtype=np.float64)).to(dev)
    check(L.fwb_diffuse(dim, kind, shp, _p(bits), _p(base), ld, _p(worklist), int(n_work.value),
                        _p(d_u), _p(d_un),
                        _p(d_ws), stream), "fwb_diffuse")
    u_new[...] = d_un.cpu().numpy()
    return u_new


__all__ = ["diffuse_host", "DeviceWeights"]
