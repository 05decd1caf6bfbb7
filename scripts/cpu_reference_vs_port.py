#!/usr/bin/env python
"""
How conservative is bench.py's CPU baseline?  Times the LIVE reference (numba path,
/root/reference -- build container only, it does not travel to the GPU box) and the oracle
port (oracle/fw_oracle.c, OpenMP) on the same bounded samples bench.py uses
(oracle/bench_cases.py), same host cores, and prints node-updates/s for both.

    python scripts/cpu_reference_vs_port.py

Measured in the build container (8 threads), 2026-10-17:
  C2 sample (FK aniso-9, 30 % fibrosis, 2048^2):  numba reference 6.16e7   port 1.09e8
  C5 sample (TP06 aniso-19, 128x128x64):          numba reference 7.85e6   port 8.41e6
i.e. the port is 1.8x / 1.07x FASTER than the reference it restates, so GPU / port ratios
understate GPU / reference ratios.
"""
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))


def main():
    from make_golden import import_reference
    fw = import_reference()
    import numba
    from oracle import oracle
    from oracle.bench_cases import cases_for_bench
    from tests.cases import build_model
    oracle.build()
    for wl, steps in (("c2", 20), ("c3", 10), ("c5", 10)):
        case, desc = cases_for_bench(wl)
        case = dict(case, t_max=2.5 * case["dt"], trackers=[])
        model, _ = build_model(fw, case)
        model.run()                                   # JIT + 3 warm-up steps
        n_myo = len(model.cardiac_tissue.myo_indexes)
        model.t_max = model.t + (steps - 0.5) * model.dt
        t0 = time.perf_counter()
        model.run(initialize=False)
        ref = n_myo * steps / (time.perf_counter() - t0)
        sec, n2 = oracle.time_steps(case, steps, n_threads=os.cpu_count(), warmup=2)
        assert n2 == n_myo
        print(f"{wl}: {desc}: numba reference {ref:.4g} upd/s ({numba.get_num_threads()} threads)"
              f" | oracle port {n2 * steps / sec:.4g} upd/s ({os.cpu_count()} threads)")


if __name__ == "__main__":
    main()
