"""Which host call of the upload / download path blocks until the device has caught up?
(1 GPU; host-return time of every call category against the total)"""
import importlib.util
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
spec = importlib.util.spec_from_file_location("fwb_bench", ROOT / "bench.py")
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)
model = b.host_model("c5", 1.0, 32, None)
model.t_max = 0.05
model.run()
eng = model._engine
from finitewave_b200 import engine as E, _lib
acc = {}


def timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter()
        r = fn(*a, **k)
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
        return r
    return w


L = eng.L
for fname in ("fwb_gather_compact", "fwb_count_offfill", "fwb_scatter_compact", "fwb_scatter_compact_keep"):
    setattr(L, fname, timed(fname, getattr(L, fname)))
orig_copy = torch.Tensor.copy_
torch.Tensor.copy_ = timed("Tensor.copy_", orig_copy)
orig_fill = torch.Tensor.fill_
torch.Tensor.fill_ = timed("Tensor.fill_", orig_fill)
orig_set = torch.Tensor.__setitem__
torch.Tensor.__setitem__ = timed("Tensor.__setitem__", orig_set)
for phase, fn in (("upload", model._upload), ("download", model._download)):
    acc.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    print(phase, "total %.3f s;" % tot, ", ".join("%s %.3f" % kv for kv in sorted(acc.items())), flush=True)
