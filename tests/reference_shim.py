"""
pytest plugin (TEST INFRASTRUCTURE, CPU only): run the REFERENCE'S OWN test files, unmodified
and where they lie (/root/reference/tests, build container only), against finitewave_b200.

    python -m pytest -p tests.reference_shim /root/reference/tests/test_basics.py

``import finitewave`` resolves to finitewave_b200; the CUDA engine is replaced by the test
double of tests/host_engine.py (time steps by the CPU oracle) and the built-in stimuli /
trackers are put on their host statements, so what the reference's tests exercise is this
package's public API and host logic -- constructor signatures, attribute names, run() /
hook / saver / loader behaviour -- against the reference authors' own assertions.  Tests that
need trackers or patterns which only exist as device kernels (ECG, LocalActivationTime,
Period, SpiralWaveCore, fibrosis patterns) cannot run here; their GPU mirrors are in
tests/test_gpu_acceptance.py.
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import finitewave_b200  # noqa: E402

sys.modules["finitewave"] = finitewave_b200


@pytest.fixture(autouse=True)
def _finitewave_b200_on_the_cpu_double(monkeypatch):
    from finitewave_b200 import model, stimulation, tracker
    from tests.host_engine import use_oracle_engine
    use_oracle_engine(monkeypatch)
    for mod in (stimulation, tracker):
        for name in dir(mod):
            cls = getattr(mod, name)
            if isinstance(cls, type) and cls.__dict__.get("_native", False):
                monkeypatch.setattr(cls, "_native", False)
    # the animation trackers' frame dump has a host statement too (plain np.save)
    monkeypatch.setattr(tracker.Animation2DTracker, "_device_hook", False)
    monkeypatch.setattr(model.CardiacModel, "async_checkpoints", False, raising=False)
