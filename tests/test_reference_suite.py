"""
Drop-in checks against the reference's own public layout and its OWN tests (CPU, no GPU).

* every sub-module import path of the reference resolves inside this package and yields the
  very class the flat namespace exports (finitewave_b200/_compat.py);
* where the reference tree is present (/root/reference, build container only -- skipped
  elsewhere), its own test files run UNMODIFIED, in place, with ``finitewave`` resolving to
  finitewave_b200 on the CPU test double (tests/reference_shim.py).  The default selection
  (state loading, commands, Aliev-Panfilov and Barkley in 2D and 3D, and the nine tracker
  tests whose trackers have a host statement: 15 tests, ~15 s) keeps the CPU suite short;
  FWB_FULL_REFERENCE_SUITE=1 adds the other six models in 2D and 3D: 27 of the reference's
  41 tests, all passing on 2026-10-17 (about half an hour of CPU oracle).  The remaining 14
  need device-only trackers / patterns and are mirrored in tests/test_gpu_acceptance.py.
"""
import importlib
import os
import subprocess
import sys
from pathlib import Path

import pytest

import finitewave_b200 as fw
from finitewave_b200 import _compat

ROOT = Path(__file__).resolve().parent.parent
REF_TESTS = Path("/root/reference/tests")


def test_reference_submodule_paths_resolve():
    assert len(_compat._LAYOUT) >= 95
    for path, (is_package, names) in _compat._LAYOUT.items():
        mod = importlib.import_module(f"finitewave_b200.{path}")
        assert hasattr(mod, "__path__") == is_package, path
        assert names, path
        for name in names:
            assert getattr(mod, name) is getattr(fw, name), (path, name)
    from finitewave_b200.cpuwave2D.fibrosis.diffuse_2d_pattern import Diffuse2DPattern
    from finitewave_b200.cpuwave3D.model.tp06_3d import TP063D
    assert Diffuse2DPattern is fw.Diffuse2DPattern and TP063D is fw.TP063D
    with pytest.raises(ModuleNotFoundError):
        importlib.import_module("finitewave_b200.cpuwave2D.no_such_module")
    # real sub-modules are untouched by the finder
    assert Path(importlib.import_module("finitewave_b200.tracker").__file__).name == "tracker.py"


def test_paths_resolve_under_the_reference_package_name(monkeypatch):
    monkeypatch.setitem(sys.modules, "finitewave", fw)
    try:
        from finitewave.core.state.state_saver import StateSaver
        from finitewave.cpuwave3D.fibrosis.structural_3d_pattern import Structural3DPattern
        assert StateSaver is fw.StateSaver and Structural3DPattern is fw.Structural3DPattern
    finally:
        for k in [k for k in sys.modules if k.startswith("finitewave.")]:
            del sys.modules[k]


def test_exception_classes_of_the_reference():
    e = fw.IncorrectNumberOfWeights(6, 5, 9)
    assert "(6)" in str(e) and "5 or 9" in str(e)
    assert issubclass(fw.IncorrectWeightsShapeError, Exception)
    e = fw.IncorrectWeightsModeError2D("xx")
    assert e.mode == "xx" and "Invalid mode: 'xx'" in str(e)


def _run_reference_tests(tmp_path, files, select=None):
    cmd = [sys.executable, "-m", "pytest", "-p", "tests.reference_shim", "-q", "-x",
           "-p", "no:cacheprovider", "-p", "no:warnings", "--rootdir", str(tmp_path)]
    if select:
        cmd += ["-k", select]
    cmd += [str(REF_TESTS / f) for f in files]
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    return subprocess.run(cmd, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=3000)


# tracker tests of the reference whose trackers have a host statement (the others -- ECG,
# LocalActivationTime, Period, PeriodAnimation, SpiralWaveCore -- are device kernels only)
_HOST_TRACKER_TESTS = ("(activation_time or action_potential or multi_variable or animation) "
                       "and not local and not period")
_ALL = ["test_basics.py", "test_models_2d.py", "test_models_3d.py", "test_trackers_2d.py",
        "test_trackers_3d.py"]


@pytest.mark.skipif(not REF_TESTS.exists(), reason="reference tree not present (build container only)")
def test_reference_own_tests_pass_unmodified(tmp_path):
    full = os.environ.get("FWB_FULL_REFERENCE_SUITE") == "1"
    models = "aliev_panfilov or barkley or mitchell or fenton or bueno or luo_rudy or tp06 or " \
             "courtemanche" if full else "aliev_panfilov or barkley"
    p = _run_reference_tests(tmp_path, _ALL,
                             f"state_loading or commands or {models} or ({_HOST_TRACKER_TESTS})")
    want = 27 if full else 15
    tail = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-500:]
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-1000:]
    assert f"{want} passed" in tail, tail


def test_public_stencil_helpers():
    """`compute_diffusion_components` / `compute_half_step_diffusion` (public building blocks of
    the reference's host-side weights; numpy utilities here, compared bit for bit with the
    live reference when they were written): analytic values on uniform inputs."""
    import numpy as np
    st = fw.AsymmetricStencil2D()
    f = np.zeros((4, 5, 2))
    f[..., 0], f[..., 1] = 0.6, 0.8
    assert np.allclose(st.compute_diffusion_components(f, 0, 0, 1.0, 0.2), 0.2 + 0.8 * 0.36)
    assert np.allclose(st.compute_diffusion_components(f, 0, 1, 1.0, 0.2), 0.8 * 0.48)
    st.D_al, st.D_ac = 1.0, 0.2
    half = st.compute_half_step_diffusion(np.ones((4, 5)), 2.0, f, 1)
    assert half.shape == (2, 4, 5) and np.allclose(half[1], 2.0 * (0.2 + 0.8 * 0.64))
    iso = fw.IsotropicStencil3D().compute_half_step_diffusion(np.ones((3, 4, 5)), np.arange(60.).reshape(3, 4, 5))
    assert iso.shape == (3, 3, 4, 5) and iso[2, 0, 0, 0] == 0.5 and iso[0, 0, 0, 0] == 10.0
    sym = fw.SymmetricStencil2D().compute_half_step_diffusion(np.ones((4, 5)), 1.0, f, 1.0, 0.2)
    assert sym.shape == (4, 4, 5) and np.allclose(sym[1], sym[2]) and np.allclose(sym[3], 0.2 + 0.8 * 0.64)
