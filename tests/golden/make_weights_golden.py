"""
Weights-only fixtures from the LIVE reference (build container only): the five stencil
classes on inputs the full-run cases of tests/cases.py do not reach -- a conductivity ARRAY
together with fibres (2D and 3D), non-default ``D_al`` / ``D_ac``, heavy fibrosis with
isolated nodes and one-node-wide strands, non-cubic shapes, odd dt / dr / D_model.

    python tests/golden/make_weights_golden.py

Writes tests/golden/weights_<kind>.npz: the inputs' checksum and the reference's
``stencil.compute_weights(model, tissue)`` array.  tests/test_oracle_golden.py requires the
oracle to reproduce every one of them bit for bit.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))


def weight_cases():
    """name -> dict(shape, kind, mesh, conductivity, fibers, D_al, D_ac, D_model, dt, dr)"""
    from tests.cases import random_fibers, random_fibrosis
    out = []
    rng = np.random.default_rng(77)
    for name, shape, kind in [("iso2d", (29, 35), "iso"), ("aniso2d", (31, 27), "aniso"),
                              ("sym2d", (26, 33), "sym"), ("iso3d", (11, 9, 14), "iso"),
                              ("aniso3d", (10, 13, 9), "aniso")]:
        mesh = random_fibrosis(shape, 0.42, int(rng.integers(1 << 30)))
        # a one-node-wide strand and an isolated node inside a fibrotic block
        blk = tuple(slice(3, 8) for _ in shape)
        mesh[blk] = 2
        mesh[tuple(5 for _ in shape)] = 1
        strand = [4] * len(shape)
        strand[-1] = slice(1, shape[-1] - 1)
        mesh[tuple(strand)] = 1
        cond = 0.05 + 1.5 * rng.random(shape)
        fibers = None if kind == "iso" else random_fibers(shape, int(rng.integers(1 << 30)))
        out.append(dict(name=name, shape=shape, kind=kind, mesh=mesh, conductivity=cond,
                        fibers=fibers, D_al=0.83, D_ac=0.27, D_model=0.154, dt=0.0123, dr=0.3))
    return out


def main():
    from make_golden import import_reference
    fw = import_reference()
    for c in weight_cases():
        dim = len(c["shape"])
        tissue = (fw.CardiacTissue2D if dim == 2 else fw.CardiacTissue3D)(list(c["shape"]))
        tissue.mesh = c["mesh"].copy()
        tissue.conductivity = c["conductivity"].copy()
        if c["fibers"] is not None:
            tissue.fibers = c["fibers"].copy()
        model = (fw.AlievPanfilov2D if dim == 2 else fw.AlievPanfilov3D)()
        model.dt, model.dr, model.D_model = c["dt"], c["dr"], c["D_model"]
        model.cardiac_tissue = tissue
        stencil = {("iso", 2): fw.IsotropicStencil2D, ("aniso", 2): fw.AsymmetricStencil2D,
                   ("sym", 2): fw.SymmetricStencil2D, ("iso", 3): fw.IsotropicStencil3D,
                   ("aniso", 3): fw.AsymmetricStencil3D}[(c["kind"], dim)]()
        if c["kind"] != "iso":
            stencil.D_al, stencil.D_ac = c["D_al"], c["D_ac"]
        w = np.array(stencil.compute_weights(model, tissue))
        np.savez_compressed(HERE / f"weights_{c['name']}.npz", weights=w,
                            mesh_sum=np.int64(tissue.mesh.astype(np.int64).sum()))
        print(c["name"], w.shape, float(w.sum()))


if __name__ == "__main__":
    main()
