"""
One-call stimulus fixtures from the LIVE reference (build container only): every stimulus
class's ``stimulate(model)`` applied to a random potential field on a fibrotic mesh, with the
arguments the trajectory cases do not reach -- negative and out-of-range box indices (numpy
slice semantics), empty boxes, float-valued matrices, ``u_max`` clamps that bite, coordinate
lists with duplicates and fibrotic entries, a voltage list fired twice.

    python tests/golden/make_stim_golden.py

Writes tests/golden/stim_onecall.npz (one ``u`` per spec).  tests/test_host_loop.py requires
the product's host statements and the oracle's restatement to reproduce each bit for bit.
"""
import sys
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

DT = 0.0125


def stim_specs():
    """-> list of (key, dim, class name, args, kwargs, calls)"""
    rng = np.random.default_rng(5)
    m2 = rng.random((20, 24))
    m2[m2 < 0.6] = 0.0                      # float matrix: > 0 selects
    m3 = (rng.random((10, 12, 9)) > 0.7).astype(np.int8)
    c2 = np.column_stack([rng.integers(0, 20, 50), rng.integers(0, 24, 50)])
    c2 = np.vstack([c2, c2[:12]])           # duplicates
    c3 = np.column_stack([rng.integers(0, 10, 40), rng.integers(0, 12, 40), rng.integers(0, 9, 40)])
    return [
        ("v2_neg", 2, "StimVoltageCoord2D", (0, 0.7, -6, -1, 2, 1000), {}, 1),
        ("v2_empty", 2, "StimVoltageCoord2D", (0, 0.7, 9, 4, 0, 24), {}, 1),
        ("c2_clamp", 2, "StimCurrentCoord2D", (0, 30.0, 1.0, 3, 15, -20, 20), dict(u_max=0.8), 1),
        ("c2_free", 2, "StimCurrentCoord2D", (0, -3.0, 1.0, 0, 20, 0, 24), {}, 2),
        ("vm2", 2, "StimVoltageMatrix2D", (0, -0.25, m2), {}, 1),
        ("cm2", 2, "StimCurrentMatrix2D", (0, 25.0, 1.0, m2), dict(u_max=0.9), 2),
        ("ca2", 2, "StimCurrentArea2D", (0, 18.0, 1.0), dict(coords=c2, u_max=0.85), 2),
        ("v3_neg", 3, "StimVoltageCoord3D", (0, 1.5, 1, -1, -5, 100, -4, -2), {}, 1),
        ("c3_clamp", 3, "StimCurrentCoord3D", (0, 40.0, 1.0, 0, 5, 0, 12, 2, 7), dict(u_max=0.5), 1),
        ("vm3", 3, "StimVoltageMatrix3D", (0, 2.0, m3), {}, 1),
        ("cm3", 3, "StimCurrentMatrix3D", (0, 9.0, 1.0, m3), {}, 1),
        ("ca3", 3, "StimCurrentArea3D", (0, 7.0, 1.0), dict(coords=c3), 1),
        ("vl3", 3, "StimVoltageListMatrix3D", (0, [0.25, -0.5, 0.75], 2 * DT, m3), {}, 2),
    ]


def fields(dim):
    from tests.cases import random_fibrosis
    shape = (20, 24) if dim == 2 else (10, 12, 9)
    rng = np.random.default_rng(100 + dim)
    mesh = random_fibrosis(shape, 0.3, 200 + dim)
    for ax in range(dim):                       # CardiacTissue's empty outer ring
        sl = [slice(None)] * dim
        for side in (0, -1):
            sl[ax] = side
            mesh[tuple(sl)] = 0
    return mesh, rng.uniform(-0.2, 1.1, shape)


def apply(fw, spec):
    key, dim, cls, args, kwargs, calls = spec
    mesh, u = fields(dim)
    model = types.SimpleNamespace(u=u.copy(), dt=DT, t=0.0,
                                  cardiac_tissue=types.SimpleNamespace(mesh=mesh))
    st = getattr(fw, cls)(*args, **kwargs)
    st.initialize(model)
    for _ in range(calls):
        st.stimulate(model)
    return model.u


def main():
    from make_golden import import_reference
    fw = import_reference()
    out = {}
    for spec in stim_specs():
        out[spec[0]] = apply(fw, spec)
        _, u0 = fields(spec[1])
        print(f"{spec[0]:10s} nodes changed: {int(np.count_nonzero(out[spec[0]] != u0))}")
    np.savez_compressed(HERE / "stim_onecall.npz", **out)


if __name__ == "__main__":
    main()
