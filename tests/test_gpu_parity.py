"""
Parity of the CUDA path (through the public finitewave API of finitewave_b200,
i.e. through the C ABI of libfinitewave_b200.so) against

  * the golden fixtures generated from the LIVE reference (tests/golden/*.npz),
  * the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star):
  * u and every state variable: max-abs error relative to max|x| <= 1e-9 after
    the whole run (up to 1000 steps);
  * activation-time maps: identical to within one dt;
  * stencil weights: bit-exact (mul/add/div only, IEEE, no FMA contraction);
  * models without transcendental calls (Aliev-Panfilov, Barkley,
    Mitchell-Schaeffer): u and state bit-exact;
  * ECG traces: 1e-9 relative (reduction order differs).
"""
from pathlib import Path

import numpy as np
import pytest

from tests.cases import build_and_run, build_model, collect_outputs, make_cases, max_rel_err

pytestmark = pytest.mark.gpu

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = make_cases()
TOL = 1e-9
EXACT_MODELS = {"aliev_panfilov", "barkley", "mitchell_schaeffer"}


@pytest.fixture(scope="module")
def fw():
    import finitewave_b200
    return finitewave_b200


def _check_outputs(case, out, ref, label):
    dt = case["dt"]
    kinds = {f"tracker{i}": t["kind"] for i, t in enumerate(case.get("trackers", []))}
    for k in ref:
        if k in ("checksum", "weights") or k.startswith("_"):
            continue
        a, b = np.asarray(out[k]), np.asarray(ref[k])
        assert a.shape == b.shape, f"{label}:{k} shape {a.shape} vs {b.shape}"
        if k in ("t", "step"):
            assert a == b, f"{label}:{k}"
        elif kinds.get(k) == "activation_time":
            assert np.array_equal(a < 0, b < 0), f"{label}:{k} activated sets differ"
            assert np.max(np.abs(a - b)) <= dt * (1 + 1e-9), f"{label}:{k} differs by more than one dt"
        elif case["model"] in EXACT_MODELS and kinds.get(k) != "ecg":
            assert np.array_equal(a, b), f"{label}:{k} not bit-exact ({max_rel_err(a, b):.3e})"
        else:
            err = max_rel_err(a, b)
            assert err <= TOL, f"{label}:{k} rel err {err:.3e}"


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_matches_reference_golden(fw, case):
    g = np.load(GOLDEN / (case["name"] + ".npz"))
    model, trackers = build_model(fw, case)
    model.run()
    assert model.gpu_steps >= int(g["step"]) and model.gpu_launches > 0, \
        "the device step kernels did not run"
    out = collect_outputs(case, model, trackers)
    # (special-boundary nodes are tissue for their neighbours but never updated; the
    # device keeps no weight row for them, the reference computes an unused one)
    w = np.asarray(model.weights)
    keep = np.ones(w.shape[:-1], dtype=bool)
    if case.get("special_boundaries") is not None:
        keep = np.asarray(case["special_boundaries"]) == 0
    assert np.array_equal(w[keep], g["weights"][keep]), \
        f"weights not bit-exact: {np.max(np.abs(w - g['weights'])):.3e}"
    _check_outputs(case, out, {k: g[k] for k in g.files}, "golden")


@pytest.mark.parametrize("name", ["c1_ap2d_readme", "c2_fk2d_aniso_fib", "c3_ms3d_iso_focal",
                                  "c4_tp06_3d_ventricle", "c5_tp06_3d_aniso_slab"])
def test_cuda_matches_oracle(fw, name):
    from oracle import oracle
    case = next(c for c in CASES if c["name"] == name)
    ref = oracle.simulate(case)
    out = build_and_run(fw, case)
    _check_outputs(case, out, ref, "oracle")


def test_run_is_deterministic(fw):
    case = next(c for c in CASES if c["name"] == "tp06_2d_aniso_fib")
    a = build_and_run(fw, case)
    b = build_and_run(fw, case)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_line_multiple_of_32_uses_tiled_mapping(fw):
    """Shapes whose contiguous axis is a multiple of 32 take the tiled block
    mapping; results must equal the oracle exactly like the linear mapping."""
    from oracle import oracle
    from tests.cases import random_fibrosis, random_fibers
    for shape, model in (([40, 64], "barkley"), ([10, 12, 32], "mitchell_schaeffer")):
        case = dict(name="tiled", model=model, shape=shape, dt=0.01, dr=0.25, t_max=2,
                    mesh=random_fibrosis(shape, 0.2, 21), fibers=random_fibers(shape, 22),
                    stims=[dict(kind="voltage_coord", t=0, value=1,
                                box=[0, 6] + sum(([0, s] for s in shape[1:]), []))],
                    trackers=[dict(kind="activation_time", threshold=0.5, step=1)])
        ref = oracle.simulate(case)
        out = build_and_run(fw, case)
        _check_outputs(case, out, ref, "oracle-tiled")
