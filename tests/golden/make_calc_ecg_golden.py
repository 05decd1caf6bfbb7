"""
`ECG{2,3}DTracker.calc_ecg()` called BY HAND on the live reference (build container only):
the two ECG cases of tests/cases.py are run to their end, then `tracker.calc_ecg()` is called
on the final state -- the reference's second stencil pass (`model.diffusion_kernel` into
`u_tr`) followed by `compute_ecg` (ecg_2d_tracker.py:61-79, ecg_3d_tracker.py:51-69).

    python tests/golden/make_calc_ecg_golden.py

Writes tests/golden/calc_ecg.npz: per case the final `u`, the returned lead values and the
tracker's `u_tr`.  The GPU test (tests/test_gpu_fixtures.py) uploads that `u` into the same
model built on finitewave_b200 and requires `calc_ecg()` to agree to 1e-12 (the reference's
prange reduction order is unspecified) and `u_tr` bit for bit.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

CASES = ("fk2d_iso_current", "ms3d_aniso_random")


def main():
    from oracle import ref_numba
    from tests.cases import build_model, case_by_name
    fw = ref_numba.reference()
    out = {}
    for name in CASES:
        case = dict(case_by_name(name))
        model, trackers = build_model(fw, case)
        model.run()
        tr = [t for _, t in trackers if hasattr(t, "calc_ecg")][0]
        ecg = np.asarray(tr.calc_ecg())
        out[name + "__u"] = np.array(model.u)
        out[name + "__ecg"] = ecg
        out[name + "__u_tr"] = np.array(tr.u_tr)
        print(name, ecg)
    np.savez_compressed(HERE / "calc_ecg.npz", **out)


if __name__ == "__main__":
    main()
