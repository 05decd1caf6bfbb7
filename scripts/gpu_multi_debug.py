"""watchdog-guarded run of one multi-device case: prints where the host is if it hangs"""
import faulthandler
import sys
import time
sys.path.insert(0, ".")
faulthandler.dump_traceback_later(int(sys.argv[2]) if len(sys.argv) > 2 else 40, exit=True)
import numpy as np
import finitewave_b200 as fw
from tests.test_gpu_multidevice import _build, _run_pair
name = sys.argv[1] if len(sys.argv) > 1 else "AlievPanfilov"
shape = {"AlievPanfilov": (24, 10, 32), "TP06": (16, 8, 32), "FentonKarma": (40, 64)}[name]
t0 = time.time()
_run_pair(fw, name, shape, [0, 0])
print(name, "2 slabs ok", round(time.time() - t0, 2), "s", flush=True)
_run_pair(fw, name, shape, [0, 0, 0])
print(name, "3 slabs ok", round(time.time() - t0, 2), "s", flush=True)
