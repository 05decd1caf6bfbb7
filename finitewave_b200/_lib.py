"""
ctypes binding of libfinitewave_b200.so (include/finitewave_b200.h).

There is no CPU fallback: if the CUDA library is missing or fails to load,
`lib()` raises.  Nothing here imports anything from oracle/.
"""
import ctypes
from ctypes import (POINTER, c_char_p, c_double, c_int, c_int64, c_uint8, c_uint32,
                    c_void_p)
from pathlib import Path

import os

_PKG = Path(__file__).resolve().parent
# FWB_LIB selects another build of the same library (A/B measurements of kernel variants,
# scripts/gpu_variants.sh); there is still no fallback if it is missing
LIB_PATH = Path(os.environ["FWB_LIB"]).resolve() if os.environ.get("FWB_LIB") \
    else _PKG / "libfinitewave_b200.so"
_lib = None

MODEL_IDS = {"aliev_panfilov": 0, "barkley": 1, "mitchell_schaeffer": 2, "fenton_karma": 3,
             "luo_rudy91": 4, "tp06": 5, "bueno_orovio": 6,
             "courtemanche": 7}
STENCIL_ISO, STENCIL_ANISO, STENCIL_SYM = 0, 1, 2
STIM_VOLTAGE, STIM_CURRENT, STIM_VOLTAGE_LIST = 0, 1, 2


class FwbError(RuntimeError):
    pass


def _sig(L):
    p = c_void_p
    L.fwb_last_error.restype = c_char_p
    L.fwb_version.restype = c_int
    L.fwb_model_n_state.argtypes = [c_int]
    L.fwb_model_n_params.argtypes = [c_int]
    L.fwb_stencil_k.argtypes = [c_int, c_int]
    L.fwb_model_read_mask.argtypes = [c_int]
    L.fwb_model_read_mask.restype = c_uint32
    L.fwb_model_write_mask.argtypes = [c_int]
    L.fwb_model_write_mask.restype = c_uint32
    L.fwb_build_chunks.argtypes = [p, c_int64, p, p, POINTER(c_int64), p]
    L.fwb_worklist_capacity.argtypes = [c_int, POINTER(c_int64)]
    L.fwb_worklist_capacity.restype = c_int64
    L.fwb_build_worklist.argtypes = [c_int, POINTER(c_int64), p, c_int, c_int, p, c_int64,
                                     POINTER(c_int64), POINTER(c_int64), POINTER(c_int64), p]
    L.fwb_order_compact.argtypes = [p, c_int64, p, c_int64, p, p, p, POINTER(c_int64), p]
    L.fwb_sim_set_tile_base.argtypes = [p, p, p]
    L.fwb_sim_set_tiles.argtypes = [p, p, p]
    L.fwb_sim_set_packed.argtypes = [p, c_int]
    L.fwb_sim_set_copy_idle.argtypes = [p, c_int]
    L.fwb_order_tiles.argtypes = [c_int, POINTER(c_int64), c_int, c_int, c_int64, c_int64, p, p, p,
                                  c_int64, p, p, p, p]
    L.fwb_gather_compact.argtypes = [p, p, c_int64, p, p, p]
    L.fwb_scatter_compact.argtypes = [p, p, c_double, c_int64, p, p, p]
    L.fwb_scatter_compact_keep.argtypes = [p, p, c_int64, p, p, p]
    L.fwb_count_offfill.argtypes = [p, c_double, c_int64, p, p, p]
    L.fwb_weights_pack.argtypes = [p, p, c_int, c_int64, c_int64, p, p, p]
    L.fwb_weights_unpack.argtypes = [p, p, c_int, c_int64, c_int64, p, p, p]
    L.fwb_compute_weights.argtypes = [c_int, c_int, POINTER(c_int64), p, p, c_double, p,
                                      c_double, c_double, c_double, c_double, c_double,
                                      p, p, c_int64, p, p]
    L.fwb_sim_create.argtypes = [POINTER(c_void_p), c_int, POINTER(c_int64), c_int, c_int,
                                 p, p, p, c_int64, c_int64, p, c_int64, p, p, p, p,
                                 POINTER(c_double), c_int, c_double, p]
    L.fwb_sim_destroy.argtypes = [p]
    L.fwb_sim_set_time.argtypes = [p, c_double, c_int64]
    L.fwb_sim_get_time.argtypes = [p, POINTER(c_double), POINTER(c_int64)]
    L.fwb_sim_current_buffer.argtypes = [p]
    L.fwb_sim_set_weights.argtypes = [p, p]
    L.fwb_sim_set_params.argtypes = [p, POINTER(c_double), c_int, c_double]
    L.fwb_sim_clear_stims.argtypes = [p]
    L.fwb_sim_add_stim_box.argtypes = [p, c_int, c_double, c_double, c_double, c_int,
                                       c_double, POINTER(c_int64)]
    L.fwb_sim_add_stim_nodes.argtypes = [p, c_int, c_double, c_double, c_double, c_int,
                                         c_double, p, c_int64, POINTER(c_double), c_int64]
    L.fwb_sim_stim_passed.argtypes = [p, c_int]
    L.fwb_sim_set_stim_passed.argtypes = [p, c_int, c_int]
    L.fwb_sim_stim_fired.argtypes = [p, c_int]
    L.fwb_sim_stim_fired.restype = c_int64
    L.fwb_sim_set_stim_fired.argtypes = [p, c_int, c_int64]
    L.fwb_sim_clear_trackers.argtypes = [p]
    L.fwb_sim_add_tracker_act.argtypes = [p, p, c_double, c_double, c_double, c_int64]
    L.fwb_sim_add_tracker_ecg.argtypes = [p, p, c_int, c_double, c_double, c_double,
                                          c_int64, p, c_int64]
    L.fwb_sim_add_tracker_point.argtypes = [p, p, p, c_int, c_double, c_double, c_int64,
                                            p, c_int64]
    L.fwb_sim_tracker_samples.argtypes = [p, c_int]
    L.fwb_sim_tracker_samples.restype = c_int64
    L.fwb_sim_run.argtypes = [p, c_int64]
    L.fwb_sim_launch_count.argtypes = [p]
    L.fwb_sim_launch_count.restype = c_int64
    L.fwb_sim_device_steps.argtypes = [p]
    L.fwb_sim_device_steps.restype = c_int64
    L.fwb_last_step_variant.restype = c_int
    L.fwb_devmath.argtypes = [c_int, p, p, c_int64, p]
    L.fwb_diffuse.argtypes = [c_int, c_int, POINTER(c_int64), p, p, c_int64, p, c_int64,
                              p, p, p, p]
    L.fwb_ecg.argtypes = [c_int, c_int, POINTER(c_int64), p, p, c_int64, p, c_int64, p, p, p, p,
                          c_int, c_double, p, p]
    L.fwb_dev_alloc.argtypes = [POINTER(c_void_p), c_int64]
    L.fwb_dev_free.argtypes = [p]
    L.fwb_ipc_handle_size.restype = c_int
    L.fwb_ipc_get_handle.argtypes = [p, p]
    L.fwb_ipc_open_handle.argtypes = [p, POINTER(c_void_p)]
    L.fwb_ipc_close_handle.argtypes = [p]
    L.fwb_enable_peer_access.argtypes = [c_int, c_int]
    L.fwb_multi_run.argtypes = [POINTER(c_void_p), c_int, c_int64, c_int64]
    L.fwb_sim_set_halo.argtypes = [p, p, p, p, c_int64, p, c_int64, p, p, c_int64, p, c_int64]
    L.fwb_sim_set_slow_offset.argtypes = [p, c_int64]
    L.fwb_sim_halo_sync.argtypes = [p]
    L.fwb_lat_cross.argtypes = [p, c_int64, c_double, p, p, p, p, p]
    L.fwb_lat_write.argtypes = [p, c_int64, c_double, p, p]
    L.fwb_gather_u8.argtypes = [p, p, c_int64, p, p]
    L.fwb_tip_scan.argtypes = [p, p, c_int, POINTER(c_int64), c_double, p, ctypes.c_uint, p, p]
    L.fwb_pattern_fibrosis.argtypes = [p, c_int, POINTER(c_int64), POINTER(c_int64),
                                       POINTER(c_int64), c_double, ctypes.c_uint64, c_int64, p]


def lib():
    """Load the CUDA library; fail loudly if it is missing (no CPU path)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FwbError(
                f"{LIB_PATH} not found: build it with `python -m finitewave_b200.build` "
                "(needs nvcc; there is no CPU fallback).")
        L = ctypes.CDLL(str(LIB_PATH))
        _sig(L)
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().fwb_last_error().decode(errors="replace")
        raise FwbError(f"{what} failed (rc={rc}): {msg}")


def shape_arr(shape):
    return (c_int64 * len(shape))(*[int(s) for s in shape])
