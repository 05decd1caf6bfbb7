#!/usr/bin/env python
"""
Public-surface diff between the LIVE reference and finitewave_b200 (build container only):
for every class both packages export -- public attributes / methods of the class, constructor
parameter names, and (where the class constructs without or with sample arguments) the
instance attributes and their default VALUES (all model parameters and init_* constants,
tracker and stimulus defaults).

    python scripts/api_diff_vs_reference.py

2026-10-17: 75 classes; constructor parameters equal except a trailing optional ``seed`` on
the four fibrosis patterns (device draws are seeded hashes) and ``shape`` on the abstract
CardiacTissue base; default values equal on all 65 constructible classes; public methods
missing here: ECG*Tracker.calc_ecg (the ECG reduction exists only fused into the step
kernel).
"""
import inspect
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))


def sample_args():
    m2, m3 = np.ones((6, 6)), np.ones((5, 5, 5))
    return {
        "StimVoltageCoord2D": (0, 1, 0, 3, 0, 3), "StimVoltageCoord3D": (0, 1, 0, 3, 0, 3, 0, 3),
        "StimCurrentCoord2D": (0, 1, 0.5, 0, 3, 0, 3),
        "StimCurrentCoord3D": (0, 1, 0.5, 0, 3, 0, 3, 0, 3),
        "StimVoltageMatrix2D": (0, 1, m2), "StimVoltageMatrix3D": (0, 1, m3),
        "StimCurrentMatrix2D": (0, 1, 0.5, m2), "StimCurrentMatrix3D": (0, 1, 0.5, m3),
        "StimCurrentArea2D": (0, 1, 0.5), "StimCurrentArea3D": (0, 1, 0.5),
        "StimVoltageListMatrix3D": (0, [1, 2, 3], 0.01, m3),
        "Diffuse2DPattern": (0.2,), "Diffuse3DPattern": (0, 5, 0, 5, 0, 5, 0.2),
        "Structural2DPattern": (0.2, 2, 2, 0, 5, 0, 5),
        "Structural3DPattern": (0, 5, 0, 5, 0, 5, 0.2, 2, 2, 2),
        "CardiacTissue2D": ([5, 5],), "CardiacTissue3D": ([5, 5, 5],),
        "StateSaver": ("x",), "StateLoader": ("x",),
    }


def same_value(v, ov):
    if isinstance(v, type) or isinstance(ov, type):
        return v is ov
    if isinstance(v, (int, float, str, bool, type(None))):
        return v == ov
    try:
        return bool(np.array_equal(np.asarray(v, dtype=object), np.asarray(ov, dtype=object)))
    except Exception:                                            # noqa: BLE001
        return repr(v) == repr(ov)


def main():
    from make_golden import import_reference
    ref = import_reference()
    import finitewave_b200 as ours
    names = sorted(n for n in dir(ref) if not n.startswith("_")
                   and inspect.isclass(getattr(ref, n)) and hasattr(ours, n))
    args = sample_args()
    n_inst = n_diff = 0
    for n in names:
        R, O = getattr(ref, n), getattr(ours, n)
        miss = sorted(a for a in dir(R) if not a.startswith("_") and not hasattr(O, a))
        if miss:
            print(f"{n}: public names missing: {miss}")
            n_diff += 1
        rp = [p for p in inspect.signature(R.__init__).parameters if p != "self"]
        op = [p for p in inspect.signature(O.__init__).parameters if p != "self"]
        if rp != op:
            print(f"{n}: constructor {rp} vs {op}")
        try:
            r, o = R(*args.get(n, ())), O(*args.get(n, ()))
        except TypeError:
            continue
        n_inst += 1
        for k, v in vars(r).items():
            if k.startswith("_") or (callable(v) and not isinstance(v, type)):
                continue
            if not hasattr(o, k):
                print(f"{n}.{k}: missing on the instance")
                n_diff += 1
            elif not same_value(v, getattr(o, k)):
                print(f"{n}.{k}: default {v!r} vs {getattr(o, k)!r}")
                n_diff += 1
    print(f"{len(names)} classes, {n_inst} instantiated, {n_diff} differences")


if __name__ == "__main__":
    main()
