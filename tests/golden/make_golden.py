"""
Generate tests/golden/*.npz from the LIVE reference (finitewave v0.8.5 numba
path at /root/reference).  Runs only in the build container (the reference
does not travel to the GPU box); the fixtures it writes are committed.

    python tests/golden/make_golden.py [case_name ...]

Each fixture holds the reference's outputs for one case of tests/cases.py:
final `u`, every state variable, tracker outputs, `t`, `step`, and the
stencil `weights` the reference computed, plus a checksum of the inputs so a
drifting case definition is detected.
"""
import hashlib
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))


def import_reference():
    """stub the visualisation-only imports."""
    sys.path.insert(0, "/root/reference")
    for n in ["matplotlib", "matplotlib.pyplot", "natsort", "pyvista", "ffmpeg",
              "skimage", "skimage.measure"]:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["natsort"].natsorted = sorted
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    import finitewave as fw
    assert fw.__file__.startswith("/root/reference"), fw.__file__
    return fw


def input_checksum(case):
    h = hashlib.sha256()
    for k in sorted(case):
        v = case[k]
        if isinstance(v, np.ndarray):
            h.update(k.encode() + np.ascontiguousarray(v).tobytes())
        elif isinstance(v, list) and v and isinstance(v[0], dict):
            for d in v:
                for kk in sorted(d):
                    vv = d[kk]
                    h.update(kk.encode())
                    h.update(np.ascontiguousarray(vv).tobytes()
                             if isinstance(vv, np.ndarray) else repr(vv).encode())
        else:
            h.update(k.encode() + repr(v).encode())
    return h.hexdigest()


def main():
    from tests.cases import make_cases, build_model, collect_outputs
    fw = import_reference()
    only = set(sys.argv[1:])
    for case in make_cases():
        if only and case["name"] not in only:
            continue
        model, trackers = build_model(fw, case)
        model.run()
        out = collect_outputs(case, model, trackers)
        out["weights"] = np.array(model.weights)
        out["checksum"] = np.array(input_checksum(case))
        np.savez_compressed(HERE / (case["name"] + ".npz"), **out)
        print(f"{case['name']:28s} step={int(out['step'])} t={float(out['t']):.6f} "
              f"u.sum={out['u'].sum():.12g}")


if __name__ == "__main__":
    main()
