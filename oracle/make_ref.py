#!/usr/bin/env python
"""
oracle/make_ref.py -- RECIPE (test / measurement infrastructure): put an importable copy of the
UNMODIFIED reference package where the GPU box can see it.

    python oracle/make_ref.py            # /root/reference/finitewave -> oracle/_ref/finitewave

`/root/reference` only exists in the build container; `oracle/_ref/` is git-ignored (reference
sources never enter the history) but travels with the repository snapshot, so that
`bench.py --impl reference` and the `cpu_baseline` leg can time the reference's own numba
kernels (finitewave/core/model/cardiac_model.py:130-189 driving cpuwave{2D,3D}) on the GPU
box's host cores.  Only the `.py` files of the package are copied, byte for byte; the five
visualisation-only imports the package pulls in are stubbed at import time
by oracle/ref_numba.py, not by editing the copy.
"""
import shutil
import sys
from pathlib import Path

SRC = Path("/root/reference/finitewave")
DST = Path(__file__).resolve().parent / "_ref" / "finitewave"


def make(force=False):
    if not SRC.is_dir():
        return DST if DST.is_dir() else None
    if DST.is_dir() and not force:
        newest = max(p.stat().st_mtime for p in SRC.rglob("*.py"))
        if (DST / ".stamp").exists() and (DST / ".stamp").stat().st_mtime >= newest:
            return DST
    if DST.exists():
        shutil.rmtree(DST)
    n = 0
    for p in SRC.rglob("*.py"):
        q = DST / p.relative_to(SRC)
        q.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(p, q)
        n += 1
    (DST / ".stamp").write_text(f"{n} files copied from {SRC}\n")
    return DST


if __name__ == "__main__":
    d = make(force="--force" in sys.argv)
    print(d if d else "reference not available here")
