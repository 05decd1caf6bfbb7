"""
CPU-side checks of the drop-in boundary: libfinitewave_b200.so builds for
sm_100a, loads, and exports every symbol include/finitewave_b200.h declares.
No compute call is made here (there is no GPU in the build container); the
argument-validation paths that return before touching the device are exercised.
"""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "finitewave_b200.h"


@pytest.fixture(scope="module")
def L():
    from finitewave_b200 import build
    build.build()
    from finitewave_b200 import _lib
    return _lib.lib()


def _declared():
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fwb_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported(L):
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/finitewave_b200.h but not exported"


def test_no_undeclared_exports():
    so = ROOT / "finitewave_b200" / "libfinitewave_b200.so"
    out = subprocess.check_output(["nm", "-D", "--defined-only", str(so)], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    extra = {e for e in exported if not e.startswith("fwb_")}
    assert not extra, f"non-ABI symbols leak from the library: {sorted(extra)[:5]}"
    assert exported == set(_declared())


def test_library_is_sm100a_only():
    so = ROOT / "finitewave_b200" / "libfinitewave_b200.so"
    out = subprocess.run(["cuobjdump", "-lelf", str(so)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_model_tables(L):
    assert L.fwb_version() == 100
    assert [L.fwb_model_n_state(m) for m in range(8)] == [1, 1, 1, 2, 7, 19, 3, 21]
    assert [L.fwb_model_n_params(m) for m in range(8)] == [5, 3, 5, 11, 15, 49, 28, 39]
    assert L.fwb_model_n_state(8) < 0 and L.fwb_model_n_state(-1) < 0
    assert [L.fwb_stencil_k(d, s) for d in (2, 3) for s in (0, 1)] == [5, 9, 7, 19]
    assert L.fwb_stencil_k(4, 0) < 0
    # TP06: cai (slot 0) is never written, oo (slot 18) never read
    assert L.fwb_model_write_mask(5) & 1 == 0
    assert L.fwb_model_read_mask(5) & (1 << 18) == 0


def test_argument_errors_do_not_touch_the_device(L):
    null = ctypes.c_void_p(0)
    shape = (ctypes.c_int64 * 2)(8, 8)
    sim = ctypes.c_void_p(0)
    rc = L.fwb_sim_create(ctypes.byref(sim), 2, shape, 0, 0, null, null, null, 0, 32, null, 0,
                          null, null, null, null, None, 0, 0.01, null)
    assert rc < 0 and b"bad argument" in L.fwb_last_error()
    assert L.fwb_sim_run(null, 1) < 0
    assert L.fwb_diffuse(2, 7, shape, null, null, 32, null, 0, null, null, null, null) < 0
    assert L.fwb_build_worklist(2, shape, null, 0, 0, null, 0, None, None, None, null) < 0
    assert L.fwb_worklist_capacity(2, shape) >= 2
    assert L.fwb_sim_set_halo(null, null, null, null, 0, null, 0, null, null, 0, null, 0) < 0
    assert L.fwb_ipc_handle_size() == 64
    assert L.fwb_compute_weights(5, 0, shape, null, null, 1.0, null, 1, 1 / 9, 1, 0.01, 0.0625,
                                 null, null, 32, null, null) < 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from finitewave_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.FwbError):
        _lib.lib()


def test_product_never_imports_the_oracle():
    for py in (ROOT / "finitewave_b200").rglob("*.py"):
        src = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), py
    for cu in (ROOT / "finitewave_b200" / "csrc").iterdir():
        assert "fw_oracle" not in cu.read_text().replace("oracle/", ""), cu


def test_ctypes_signatures_match_the_header(L):
    """Binding drift is the classic FFI bug: every prototype of include/finitewave_b200.h is
    compared with the ctypes ``argtypes`` the package declares for it (finitewave_b200/_lib.py)
    -- same number of parameters, and pointer / double / int64_t / int / unsigned in the same
    positions."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = re.findall(r"FWB_API\s+[\w\s\*]+?\s*\b(fwb_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    assert sorted(n for n, _ in protos) == _declared()

    def kind(param):
        param = " ".join(param.split())
        if param in ("void", ""):
            return None
        if "*" in param or "fwb_stream_t" in param:
            return "ptr"
        return param.rsplit(" ", 1)[0].replace("const ", "").strip()

    def matches(k, a):
        if k == "ptr":
            return a in (ctypes.c_void_p, ctypes.c_char_p) or issubclass(a, ctypes._Pointer)
        return a in {"double": (ctypes.c_double,), "int64_t": (ctypes.c_int64,),
                     "int": (ctypes.c_int,), "unsigned": (ctypes.c_uint, ctypes.c_uint32),
                     "uint32_t": (ctypes.c_uint, ctypes.c_uint32),
                     "uint64_t": (ctypes.c_uint64,),
                     "unsigned long long": (ctypes.c_uint64, ctypes.c_ulonglong)}.get(k, ())

    checked = 0
    for name, params in protos:
        kinds = [k for k in (kind(p) for p in params.split(",")) if k is not None]
        argtypes = getattr(L, name).argtypes
        if argtypes is None:
            assert not kinds, f"{name}: {len(kinds)} parameters but no argtypes declared"
            continue
        assert len(argtypes) == len(kinds), (name, len(argtypes), len(kinds))
        for i, (k, a) in enumerate(zip(kinds, argtypes)):
            assert matches(k, a), (name, i, k, a)
        checked += 1
    assert checked >= 45
