"""
One-call tracker fixtures from the LIVE reference (build container only): the two tracker
computations with real arithmetic in them, on inputs the trajectory cases do not reach.

  * ``compute_ecg`` (ecg_2d_tracker.py / ecg_3d_tracker.py) on random fields over a fibrotic
    mesh with leads ON a node (d == 0 is skipped), leads inside and far outside the tissue;
  * the spiral-tip finder ``track_tip_line`` (spiral_wave_core_2d_tracker.py) on pairs of smooth
    random fields with dozens of isoline crossings per frame (every branch of the bilinear
    root finder, both roots, rejected roots).

    python tests/golden/make_tracker_golden.py

Writes tests/golden/tracker_onecall.npz.  tests/test_oracle_golden.py requires the oracle to
reproduce the tips bit for bit and the ECG sums to 1e-12 (the reference's prange reduction
order is unspecified).
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))


def ecg_inputs(dim):
    from tests.cases import random_fibrosis
    shape = (26, 31) if dim == 2 else (12, 11, 13)
    rng = np.random.default_rng(300 + dim)
    mesh = random_fibrosis(shape, 0.25, 310 + dim)
    for ax in range(dim):
        sl = [slice(None)] * dim
        for side in (0, -1):
            sl[ax] = side
            mesh[tuple(sl)] = 0
    u, u_tr = rng.normal(size=shape), rng.normal(size=shape)
    if dim == 2:   # (x, y, z): z is the lead's height above the sheet
        coords = np.array([[13.0, 15.0, 0.0], [13.0, 15.0, 2.0], [5.5, 7.25, 1.0],
                           [-8.0, 40.0, 3.0], [25.0, 30.0, 0.5]])
    else:
        coords = np.array([[6.0, 5.0, 6.0], [6.5, 5.5, 6.5], [-3.0, 14.0, 20.0],
                           [11.0, 10.0, 12.0]])
    return mesh, u, u_tr, coords, 0.3


def tip_inputs(k):
    """Smooth random field pairs (sums of a few random plane waves), 48 x 44."""
    rng = np.random.default_rng(400 + k)
    x, y = np.meshgrid(np.arange(48.0), np.arange(44.0), indexing="ij")

    def field():
        f = np.zeros_like(x)
        for _ in range(6):
            kx, ky, ph = rng.uniform(-0.45, 0.45), rng.uniform(-0.45, 0.45), rng.uniform(0, 6.3)
            f += rng.uniform(0.2, 1.0) * np.sin(kx * x + ky * y + ph)
        return 0.5 + 0.25 * f
    return field(), field(), 0.5


def main():
    from make_golden import import_reference
    fw = import_reference()
    from finitewave.cpuwave2D.tracker.ecg_2d_tracker import compute_ecg as ecg2
    from finitewave.cpuwave3D.tracker.ecg_3d_tracker import compute_ecg as ecg3
    out = {}
    for dim, fn in ((2, ecg2), (3, ecg3)):
        mesh, u, u_tr, coords, dr = ecg_inputs(dim)
        idx = np.flatnonzero(mesh == 1)
        out[f"ecg{dim}"] = np.asarray(fn(u_tr, u, coords, dr, idx))
        print(f"ecg{dim}", out[f"ecg{dim}"])
    tr = fw.SpiralWaveCore2DTracker()
    for k in range(4):
        a, b, thr = tip_inputs(k)
        tips = np.array(tr.track_tip_line(a, b, thr), dtype=np.float64).reshape(-1, 2)
        out[f"tips{k}"] = tips
        print(f"tips{k}: {len(tips)} tips")
    np.savez_compressed(HERE / "tracker_onecall.npz", **out)


if __name__ == "__main__":
    main()
