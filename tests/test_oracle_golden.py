"""
Pins the CPU oracle (oracle/fw_oracle.c + oracle/oracle.py) against the
fixtures generated from the LIVE reference (tests/golden/make_golden.py).

Bar: bit-exact for u, every state variable, activation-time maps, point
samplers and the stencil weights (same operation order, same libm);
ECG traces within 1e-12 relative (the reference's prange reduction order is
unspecified).
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle
from tests.cases import make_cases, max_rel_err

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = make_cases()


@pytest.fixture(scope="module", autouse=True)
def _build():
    oracle.build()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference(case):
    g = np.load(GOLDEN / (case["name"] + ".npz"))
    out = oracle.simulate(case, return_model=True)
    ecg_keys = {f"tracker{i}" for i, t in enumerate(case.get("trackers", []))
                if t["kind"] == "ecg"}
    assert np.array_equal(out["_weights"], g["weights"]), "stencil weights differ"
    for k in g.files:
        if k in ("checksum", "weights"):
            continue
        if k in ecg_keys:
            assert max_rel_err(out[k], g[k]) <= 1e-12, k
        else:
            assert out[k].shape == g[k].shape, k
            assert np.array_equal(out[k], g[k]), f"{k}: {max_rel_err(out[k], g[k]):.3e}"


def _weight_cases():
    from tests.golden.make_weights_golden import weight_cases
    return weight_cases()


@pytest.mark.parametrize("kind", ["iso2d", "aniso2d", "sym2d", "iso3d", "aniso3d"])
def test_oracle_weights_match_reference_on_wide_inputs(kind):
    """Weights-only fixtures of the live reference (tests/golden/make_weights_golden.py):
    conductivity arrays together with fibres, non-default D_al / D_ac, 42 % fibrosis with
    isolated nodes and one-node strands, odd dt / dr / D_model -- bit for bit."""
    c = next(x for x in _weight_cases() if x["name"] == kind)
    g = np.load(GOLDEN / f"weights_{kind}.npz")
    mesh = oracle.apply_boundaries(c["mesh"].copy())
    assert int(g["mesh_sum"]) == int(mesh.astype(np.int64).sum())
    w = oracle.compute_weights(mesh, c["conductivity"], c["fibers"], c["kind"], c["D_model"],
                               c["dt"], c["dr"], D_al=c["D_al"], D_ac=c["D_ac"])
    assert np.array_equal(w, g["weights"])


@pytest.mark.parametrize("model", list(oracle.MODELS))
def test_oracle_ionic_matches_reference_one_step(model):
    """One call of the live reference's own ``run_ionic_kernel()`` on random node states on
    both sides of (and exactly at) every branch threshold
    (tests/golden/make_ionic_golden.py): u_new and every state array, bit for bit."""
    from tests.golden.make_ionic_golden import DT, SHAPE, ionic_inputs
    spec = oracle.MODELS[model]
    g = np.load(GOLDEN / f"ionic_{model}.npz")
    u, u_new, states = ionic_inputs(model)
    idx = np.flatnonzero(oracle.apply_boundaries(np.ones(SHAPE, dtype=np.int8)) == 1)
    states = [np.ascontiguousarray(s) for s in states]
    pvec = np.array([float(v) for v in spec["params"].values()], dtype=np.float64)
    oracle.ionic(model, u_new, u, states, idx.astype(np.int64), DT, pvec)
    assert np.array_equal(u_new, g["u_new"])
    for var, s in zip(spec["state"], states):
        assert np.array_equal(s, g[var]), var


def test_oracle_ecg_and_tip_finder_match_reference_one_call():
    """``compute_ecg`` (leads on a node, inside and far outside the tissue) to 1e-12 and the
    spiral-tip finder on smooth random field pairs (dozens of tips per frame) bit for bit,
    against one-call outputs of the live reference (tests/golden/make_tracker_golden.py)."""
    from tests.golden.make_tracker_golden import ecg_inputs, tip_inputs
    g = np.load(GOLDEN / "tracker_onecall.npz")
    for dim in (2, 3):
        mesh, u, u_tr, coords, dr = ecg_inputs(dim)
        idx = np.flatnonzero(mesh == 1).astype(np.int64)
        got = oracle.ecg(u_tr, u, coords, dr, idx, mesh.shape)
        assert max_rel_err(got, g[f"ecg{dim}"]) <= 1e-12
    n_tips = 0
    for k in range(4):
        a, b, thr = tip_inputs(k)
        tips = np.array(oracle.tip_scan_2d(a, b, thr), dtype=np.float64).reshape(-1, 2)
        assert np.array_equal(tips, g[f"tips{k}"]), k
        n_tips += len(tips)
    assert n_tips > 40


def test_golden_inputs_unchanged():
    """The case definitions still produce the inputs the fixtures were made from."""
    from tests.golden.make_golden import input_checksum
    for case in CASES:
        g = np.load(GOLDEN / (case["name"] + ".npz"))
        assert str(g["checksum"]) == input_checksum(case), case["name"]


def test_reference_fingerprints():
    """SURVEY.md App. C fingerprints of the README quick start."""
    g = np.load(GOLDEN / "c1_ap2d_readme.npz")
    assert float(g["u"].sum()) == 5997.748456816207
    assert float(g["v"].sum()) == 794.2711218120692
    assert float(g["u"].max()) == 0.9899530914614856
    assert float(g["t"]) == 9.999999999999831
    assert np.allclose(g["weights"][50, 50], [0.16, 0.16, 0.36, 0.16, 0.16])
    assert np.allclose(g["weights"][1, 1], [0, 0, 0.36, 0.32, 0.32])


def test_quirks():
    """SURVEY.md App. A.4 quirks present in the fixtures (and so in the oracle)."""
    g = np.load(GOLDEN / "tp06_2d_iso.npz")
    assert np.all(g["cai"] == 0.00007), "TP06 cai must stay frozen"
    g = np.load(GOLDEN / "c3_ms3d_iso_focal.npz")
    act = g["tracker0"]
    assert act[0, 0, 0] == -1 and act[12, 12, 12] == 0.0
    # current stim fires duration/dt + 1 times: (t0=0.5, dur=0.5, dt=0.01) -> 51
    t, n = 0.0, 0
    for _ in range(600):
        if t >= 0.5:
            n += 1
            if t >= 0.5 + 0.5:
                break
        t += 0.01
    assert n == 51
