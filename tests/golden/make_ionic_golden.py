"""
One-step ionic fixtures from the LIVE reference (build container only): every model's own
``run_ionic_kernel()`` (the numba ``ionic_kernel_2d`` with the model's attributes as
arguments) applied ONCE to random node states spread over both sides of every branch
threshold, including values exactly AT the thresholds -- regimes the trajectory fixtures of
make_golden.py visit rarely or never.

    python tests/golden/make_ionic_golden.py

Writes tests/golden/ionic_<model>.npz: ``u_new`` and every state array after the call.  The
inputs are regenerated from the seed by ``ionic_inputs`` (shared with the test), and
tests/test_oracle_golden.py requires the oracle to reproduce the outputs bit for bit.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

SHAPE = (18, 64)                     # 16 x 62 = 992 updated nodes
DT = 0.01
EXACT = {  # u values exactly at the models' branch thresholds
    "aliev_panfilov": [0.0, 0.1, 1.0], "barkley": [0.0, 1.0],
    "mitchell_schaeffer": [0.13, 0.0, 1.0],
    "fenton_karma": [0.13, 0.85, 0.0, 1.0],
    "bueno_orovio": [0.006, 0.13, 0.3, 0.0, 0.65, 0.9087, 1.55],
    "luo_rudy91": [-40.0, -100.0, -84.5, 0.0, -87.0, -23.0, -77.000001, -47.129999, 7.0, -10.0],
    "tp06": [-40.0, -84.5, 0.0, -60.0, -26.0, -35.0, -5.0, 15.000001, -8.0],
    "courtemanche": [-40.0, -47.13, -84.5, 0.0, -10.000001, -30.0, -14.099999, -3.3328, -19.9,
                     3.0],
}
# NOT representable in a fixture: at u exactly -47.13 or -77 (LR91), 15 (TP06), -10 or -14.1
# (Courtemanche) a rate is 0 / 0 and the reference's numba kernels RAISE ZeroDivisionError
# (python error model); the values next to those removable singularities are used instead.


def ionic_inputs(model):
    """-> (u, u_new, [state arrays]) over SHAPE, deterministic."""
    from tests.test_host_models import _random_node_states
    rng = np.random.default_rng(2026)
    n = int(np.prod(SHAPE))
    u, states = _random_node_states(model, n, rng)
    ex = EXACT[model]
    u = u.reshape(SHAPE)
    u[1, 1:1 + len(ex)] = ex
    u_new = u + 0.01 * rng.uniform(-1, 1, SHAPE)
    return np.ascontiguousarray(u), np.ascontiguousarray(u_new), [s.reshape(SHAPE) for s in states]


def main():
    from make_golden import import_reference
    fw = import_reference()
    from oracle import oracle
    from tests.cases import MODEL_CLASS
    for name, spec in oracle.MODELS.items():
        tissue = fw.CardiacTissue2D(list(SHAPE))
        model = getattr(fw, MODEL_CLASS[name] + "2D")()
        model.dt, model.dr, model.t_max, model.prog_bar = DT, 0.25, DT, False
        model.cardiac_tissue = tissue
        model.initialize()
        u, u_new, states = ionic_inputs(name)
        model.u, model.u_new = u.copy(), u_new.copy()
        for var, s in zip(spec["state"], states):
            setattr(model, var, s.copy())
        model.run_ionic_kernel()
        out = {"u_new": model.u_new}
        assert np.array_equal(model.u, u)
        for var in spec["state"]:
            out[var] = getattr(model, var)
        bad = [k for k, v in out.items() if not np.all(np.isfinite(v))]
        np.savez_compressed(HERE / f"ionic_{name}.npz", **out)
        print(f"{name:20s} u_new.sum={out['u_new'].sum():.12g} non-finite: {bad}")


if __name__ == "__main__":
    main()
