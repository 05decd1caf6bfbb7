"""
CPU check of the CUDA kernels' per-node model math.

finitewave_b200/csrc/models.cuh is a __host__ __device__ header; tests/hostcheck
compiles it with g++ (TEST ONLY, not a product path) so the transcription --
operation order, host-hoisted constants, read/write masks -- can be compared
with the oracle bit for bit on random node states without a GPU.  What the GPU
adds on top (CUDA libm vs glibc, <= 2 ulp per call) is covered by the -m gpu
parity tests with the 1e-9 tolerance.
"""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "hostcheck" / "host_models.cpp"
SO = ROOT / "tests" / "hostcheck" / "libhost_models.so"
MODEL_IDS = {"aliev_panfilov": 0, "barkley": 1, "mitchell_schaeffer": 2, "fenton_karma": 3,
             "luo_rudy91": 4, "tp06": 5, "bueno_orovio": 6, "courtemanche": 7}
c_double_p = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def hostlib():
    deps = [SRC, ROOT / "finitewave_b200" / "csrc" / "models.cuh",
            ROOT / "finitewave_b200" / "csrc" / "fexp.cuh"]
    if not SO.exists() or SO.stat().st_mtime < max(d.stat().st_mtime for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-fno-fast-math", "-fvisibility=hidden",
                               "-I/usr/local/cuda/include", "-o", str(SO), str(SRC), "-lm"])
    return ctypes.CDLL(str(SO))


def _random_node_states(model, n, rng):
    """u and state values spread over the ranges a simulation visits (both sides of
    every branch threshold)."""
    spec = oracle.MODELS[model]
    if model in ("luo_rudy91", "tp06", "courtemanche"):
        u = rng.uniform(-95.0, 45.0, n)
        u[: n // 8] = rng.uniform(-41.0, -39.0, n // 8)     # around the h/j branch
    else:
        u = rng.uniform(-0.1, 1.1, n)
        u[: n // 8] = rng.uniform(0.12, 0.14, n // 8)       # around u_gate / u_c / theta_w
        u[n // 8: n // 4] = rng.uniform(0.0, 0.012, n // 8)  # around theta_o / theta_v_m
    states = []
    for name in spec["state"]:
        init = spec["init"][name]
        if model == "courtemanche" and name in ("caup", "carel"):
            s = rng.uniform(0.5, 3.0, n)
        elif model == "courtemanche" and name == "nai":
            s = rng.uniform(9.0, 13.0, n)
        elif model == "courtemanche" and name == "irel":
            s = rng.uniform(0.0, 0.05, n)
        elif name in ("cai", "cass"):
            s = rng.uniform(5e-5, 2e-3, n)
        elif name == "casr":
            s = rng.uniform(0.3, 4.0, n)
        elif name in ("nai",):
            s = rng.uniform(7.0, 11.0, n)
        elif name in ("Ki", "ki"):
            s = rng.uniform(130.0, 142.0, n)
        else:
            s = rng.uniform(0.0, 1.0, n)
        s[0] = init
        states.append(np.ascontiguousarray(s))
    return np.ascontiguousarray(u), states


@pytest.mark.parametrize("model", list(MODEL_IDS))
@pytest.mark.parametrize("dt", [0.01, 0.005])
def test_models_header_matches_oracle_bitwise(hostlib, model, dt):
    rng = np.random.default_rng(11)
    n = 20000
    spec = oracle.MODELS[model]
    u, st = _random_node_states(model, n, rng)
    pvec = np.array([float(v) for v in spec["params"].values()], dtype=np.float64)
    diff = rng.uniform(-1, 1, n) * 0.01 + u                    # a "diffusion result"

    # oracle
    un_o = diff.copy()
    st_o = [s.copy() for s in st]
    idx = np.arange(n, dtype=np.int64)
    oracle.ionic(model, un_o, u, st_o, idx, dt, pvec)

    # the kernels' header
    un_h = diff.copy()
    st_h = [s.copy() for s in st]
    arr = (c_double_p * max(1, len(st_h)))(*[s.ctypes.data_as(c_double_p) for s in st_h])
    rc = hostlib.fwb_host_ionic(MODEL_IDS[model], un_h.ctypes.data_as(c_double_p),
                                u.ctypes.data_as(c_double_p), arr, ctypes.c_int64(n),
                                ctypes.c_double(dt), pvec.ctypes.data_as(c_double_p))
    assert rc == 0
    assert np.array_equal(un_h, un_o), f"u_new differs: {np.max(np.abs(un_h - un_o)):.3e}"
    for name, a, b in zip(spec["state"], st_h, st_o):
        if model == "tp06" and name == "cai":
            assert np.array_equal(a, st[0]), "TP06 cai must stay untouched"
        assert np.array_equal(a, b), f"{name} differs: {np.max(np.abs(a - b)):.3e}"


def test_tp06_fast_path_matches_reference_statement(hostlib):
    """Model<TP06>::ionic_fast (the device's path for |u| < 300 mV: shared exponentials,
    one reciprocal per gate, branch-free divisions) is an algebraic rearrangement of the
    reference statement: one step from the same node states agrees to rounding level."""
    rng = np.random.default_rng(3)
    n = 200000
    spec = oracle.MODELS["tp06"]
    u, st = _random_node_states("tp06", n, rng)
    u[-n // 10:] = rng.uniform(-290.0, 290.0, n // 10)       # the whole guarded range
    pvec = np.array([float(v) for v in spec["params"].values()], dtype=np.float64)
    diff = rng.uniform(-1, 1, n) * 0.01 + u
    for dt in (0.01, 0.001):
        outs = []
        for fn, extra in ((hostlib.fwb_host_ionic, (MODEL_IDS["tp06"],)),
                          (hostlib.fwb_host_tp06_fast, ())):
            un = diff.copy()
            s2 = [s.copy() for s in st]
            arr = (c_double_p * len(s2))(*[s.ctypes.data_as(c_double_p) for s in s2])
            rc = fn(*extra, un.ctypes.data_as(c_double_p), u.ctypes.data_as(c_double_p), arr,
                    ctypes.c_int64(n), ctypes.c_double(dt), pvec.ctypes.data_as(c_double_p))
            assert rc == 0
            outs.append((un, s2))
        (un_r, st_r), (un_f, st_f) = outs
        # the step's increments (what the rearrangement computes) agree to 1e-11 relative
        # to the increment scale; the values themselves to ~1e-14
        du = np.abs(un_r - diff)
        err = np.abs(un_f - un_r) / np.maximum(np.abs(un_r), 1.0)
        assert err.max() < 2e-13, ("u_new", err.max(), int(err.argmax()), u[err.argmax()])
        for name, a, b, s0 in zip(spec["state"], st_f, st_r, st):
            if name == "cai":
                assert np.array_equal(a, s0)
                continue
            err = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
            assert err.max() < 2e-12, (name, err.max(), int(err.argmax()), u[err.argmax()])


def test_lr91_fast_path_matches_reference_statement(hostlib):
    """Model<LR91>::ionic_fast (division-free gates, shared exponentials) against the
    reference statement, one step from the same node states."""
    rng = np.random.default_rng(4)
    n = 200000
    spec = oracle.MODELS["luo_rudy91"]
    u, st = _random_node_states("luo_rudy91", n, rng)
    u[-n // 10:] = rng.uniform(-290.0, 290.0, n // 10)
    u[np.abs(u + 77.0) < 1e-3] += 0.01          # both forms are 0/0 at u = -77, -47.13
    u[np.abs(u + 47.13) < 1e-3] += 0.01
    pvec = np.array([float(v) for v in spec["params"].values()], dtype=np.float64)
    diff = rng.uniform(-1, 1, n) * 0.01 + u
    outs = []
    for fn, extra in ((hostlib.fwb_host_ionic, (MODEL_IDS["luo_rudy91"],)),
                      (hostlib.fwb_host_lr91_fast, ())):
        un = diff.copy()
        s2 = [s.copy() for s in st]
        arr = (c_double_p * len(s2))(*[s.ctypes.data_as(c_double_p) for s in s2])
        rc = fn(*extra, un.ctypes.data_as(c_double_p), u.ctypes.data_as(c_double_p), arr,
                ctypes.c_int64(n), ctypes.c_double(0.01), pvec.ctypes.data_as(c_double_p))
        assert rc == 0
        outs.append((un, s2))
    (un_r, st_r), (un_f, st_f) = outs
    err = np.abs(un_f - un_r) / np.maximum(np.abs(un_r), 1.0)
    assert err.max() < 1e-12, ("u_new", err.max(), u[err.argmax()])
    for name, a, b in zip(spec["state"], st_f, st_r):
        err = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
        assert err.max() < 2e-12, (name, err.max(), u[err.argmax()])


def test_courtemanche_fast_path_matches_reference_statement(hostlib):
    """Model<COURTEMANCHE>::ionic_fast against the reference statement, one step from the
    same node states (away from the removable 0/0 points of the rate functions)."""
    rng = np.random.default_rng(6)
    n = 200000
    spec = oracle.MODELS["courtemanche"]
    u, st = _random_node_states("courtemanche", n, rng)
    u[-n // 10:] = rng.uniform(-290.0, 290.0, n // 10)
    for sing in (-47.13, -14.1, 3.3328, 19.9, -10.0, 7.9):
        u[np.abs(u - sing) < 1e-3] += 0.01
    pvec = np.array([float(v) for v in spec["params"].values()], dtype=np.float64)
    diff = rng.uniform(-1, 1, n) * 0.01 + u
    outs = []
    for fn, extra in ((hostlib.fwb_host_ionic, (MODEL_IDS["courtemanche"],)),
                      (hostlib.fwb_host_court_fast, ())):
        un = diff.copy()
        s2 = [s.copy() for s in st]
        arr = (c_double_p * len(s2))(*[s.ctypes.data_as(c_double_p) for s in s2])
        rc = fn(*extra, un.ctypes.data_as(c_double_p), u.ctypes.data_as(c_double_p), arr,
                ctypes.c_int64(n), ctypes.c_double(0.01), pvec.ctypes.data_as(c_double_p))
        assert rc == 0
        outs.append((un, s2))
    (un_r, st_r), (un_f, st_f) = outs
    err = np.abs(un_f - un_r) / np.maximum(np.abs(un_r), 1.0)
    assert err.max() < 1e-12, ("u_new", err.max(), u[err.argmax()])
    for name, a, b in zip(spec["state"], st_f, st_r):
        err = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
        assert err.max() < 5e-12, (name, err.max(), u[err.argmax()])


def test_param_order_matches_oracle_tables():
    """The device parameter vectors are indexed by position: the attribute order in
    finitewave_b200.model must be the oracle's (= the reference's kernel call order)."""
    from finitewave_b200 import model as m
    for cls in (m.AlievPanfilov2D, m.Barkley2D, m.MitchellSchaeffer2D, m.FentonKarma2D,
                m.BuenoOrovio2D, m.Courtemanche2D, m.LuoRudy912D, m.TP062D):
        spec = oracle.MODELS[cls._MODEL]
        assert list(cls._PARAMS) == list(spec["params"]), cls.__name__
        assert list(cls._STATE) == list(spec["state"]), cls.__name__
        obj = cls()
        for k, v in spec["params"].items():
            assert float(getattr(obj, k)) == float(v), (cls.__name__, k)
        for k, v in spec["init"].items():
            assert float(getattr(obj, "init_" + k.rstrip("_"))) == float(v), (cls.__name__, k)
        assert obj.D_model == spec["D_model"]


def _fexp(hostlib, x, mode):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    hostlib.fwb_host_fexp(x.ctypes.data_as(c_double_p), out.ctypes.data_as(c_double_p),
                          ctypes.c_int64(len(x)), ctypes.c_int(mode))
    return out


def test_fexp_within_two_ulp_of_libm(hostlib):
    """The device's table-driven exp (same source, host build): <= 2 ulp against libm
    over its whole domain |x| < 700; NaN stays NaN; the clamped variants saturate."""
    rng = np.random.default_rng(5)
    x = np.concatenate([rng.uniform(-700, 700, 2_000_000), rng.uniform(-40, 40, 2_000_000),
                        rng.uniform(-1, 1, 1_000_000), rng.uniform(-1e-3, 1e-3, 200_000),
                        np.array([0.0, -0.0, 1.0, -1.0, 699.999, -699.999, 1e-300, -1e-300])])
    out = _fexp(hostlib, x, 0)
    ref = np.exp(x)
    ulp = np.abs(out - ref) / np.spacing(ref)
    assert ulp.max() <= 2.0, ulp.max()
    assert np.isnan(_fexp(hostlib, [np.nan], 0)[0])
    assert np.isnan(_fexp(hostlib, [np.nan], 1)[0]) and np.isnan(_fexp(hostlib, [np.nan], 2)[0])
    lo = _fexp(hostlib, [-1e6, -np.inf, -800.0, -700.0, -3.5], 1)
    assert np.all(lo[:4] == np.exp(-700.0)) and abs(lo[4] / np.exp(-3.5) - 1) < 1e-15
    both = _fexp(hostlib, [-1e300, 1e300, np.inf, 12.0], 2)
    assert both[0] == np.exp(-700.0) and both[1] == both[2] == _fexp(hostlib, [700.0], 0)[0]
    assert abs(both[3] / np.exp(12.0) - 1) < 1e-15
    print(f"fexp max error {ulp.max():.3f} ulp, mean {ulp.mean():.3f}")


def test_flog_accuracy(hostlib):
    """The device's table-driven log (same source, host build) on positive normal arguments:
    <= 2 ulp of the result where |log x| >= 1/2, <= 3e-16 absolute near x = 1."""
    rng = np.random.default_rng(9)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 1_000_000)), rng.uniform(1e-5, 1e-2, 500_000),
                        rng.uniform(1, 200, 500_000), rng.uniform(0.5, 2.0, 500_000),
                        np.array([1.0, 2.0, 0.5, 7e-5, 138.3, 5.4, 1 + 2 ** -52, 2 - 2 ** -52])])
    x = np.ascontiguousarray(x)
    out = np.empty_like(x)
    hostlib.fwb_host_flog(x.ctypes.data_as(c_double_p), out.ctypes.data_as(c_double_p),
                          ctypes.c_int64(len(x)))
    ref = np.log(x)
    big = np.abs(ref) >= 0.5
    ulp = np.abs(out - ref)[big] / np.spacing(np.abs(ref[big]))
    assert ulp.max() <= 2.0, ulp.max()
    assert np.abs(out - ref)[~big].max() <= 3e-16
    print(f"flog max error {ulp.max():.3f} ulp (|log| >= 0.5), "
          f"{np.abs(out - ref)[~big].max():.2e} absolute near 1")


@pytest.mark.parametrize("model", list(MODEL_IDS))
def test_kernel_header_matches_live_reference_one_step(hostlib, model):
    """The kernels' per-node header against the LIVE reference directly (fixtures of
    tests/golden/make_ionic_golden.py: one ``run_ionic_kernel()`` call of the numba code on
    random node states, thresholds included): the reference statement (`Model<>::ionic` as the
    host compiles it) bit for bit; the rearranged device paths of LR91 / TP06 / Courtemanche
    (`ionic_fast`) and Fenton-Karma (`ionic_t<IO, true>`) to rounding level -- the 1e-9-after-1000-steps bar leaves 1e-12 per step."""
    from tests.golden.make_ionic_golden import DT, SHAPE, ionic_inputs
    spec = oracle.MODELS[model]
    g = np.load(ROOT / "tests" / "golden" / f"ionic_{model}.npz")
    u, u_new, states = ionic_inputs(model)
    myo = oracle.apply_boundaries(np.ones(SHAPE, dtype=np.int8)).reshape(-1) == 1
    pvec = np.array([float(v) for v in spec["params"].values()], dtype=np.float64)
    n = u.size

    def run(fn, *extra):
        un = np.ascontiguousarray(u_new.reshape(-1).copy())
        st = [np.ascontiguousarray(s.reshape(-1).copy()) for s in states]
        arr = (c_double_p * max(1, len(st)))(*[s.ctypes.data_as(c_double_p) for s in st])
        rc = fn(*extra, un.ctypes.data_as(c_double_p),
                np.ascontiguousarray(u.reshape(-1)).ctypes.data_as(c_double_p), arr,
                ctypes.c_int64(n), ctypes.c_double(DT), pvec.ctypes.data_as(c_double_p))
        assert rc == 0
        return un, st

    un, st = run(hostlib.fwb_host_ionic, MODEL_IDS[model])
    assert np.array_equal(un[myo], g["u_new"].reshape(-1)[myo])
    for var, s in zip(spec["state"], st):
        if model == "tp06" and var == "oo":
            continue        # write-only in the reference: the header stores it, compared below
        assert np.array_equal(s[myo], g[var].reshape(-1)[myo]), var

    fast = {"tp06": hostlib.fwb_host_tp06_fast, "luo_rudy91": hostlib.fwb_host_lr91_fast,
            "courtemanche": hostlib.fwb_host_court_fast,
            "fenton_karma": hostlib.fwb_host_fk_fast}.get(model)
    if fast is None:
        return
    un, st = run(fast)
    # within 1e-5 mV of a removable 0 / 0 point of a rate function (make_ionic_golden.EXACT)
    # the cancellation costs BOTH forms 6-7 digits: looser bound there
    uf = u.reshape(-1)
    near = np.zeros(n, dtype=bool)
    for sing in (-47.13, -77.0, 15.0, -10.0, -14.1):
        near |= np.abs(uf - sing) < 1e-5
    ref = g["u_new"].reshape(-1)
    err = np.abs(un - ref) / np.maximum(np.abs(ref), 1.0)
    assert err[myo & ~near].max() < 1e-12 and err[myo].max() < 1e-9, ("u_new", err[myo].max())
    for var, s in zip(spec["state"], st):
        r = g[var].reshape(-1)
        err = np.abs(s - r) / np.maximum(np.abs(r), 1e-3)
        assert err[myo & ~near].max() < 2e-12 and err[myo].max() < 1e-9, (var, err[myo].max())


def test_fenton_karma_device_path_matches_reference_statement(hostlib):
    """Model<FK>::ionic_t<IO, true> (the device's path: 1 + tanh(x) as 2 g / (1 + g) with
    g = exp(2x), the same cost on every lane) against the exact value and against the reference
    statement, one step from the same node states."""
    # the helper: a few ulp RELATIVE to the exact 2 e^2x / (1 + e^2x) for any x (1 + tanh(x)
    # itself cancels for x << 0), NaN / huge arguments through the library statement
    x = np.concatenate([np.linspace(-340, 340, 200001), np.random.default_rng(8).normal(0, 2, 200000),
                        [0.0, -0.55, 0.55, 299.999, -299.999, 300.0, -300.0, 1e6, -1e6]])
    out = np.empty_like(x)
    hostlib.fwb_host_one_plus_tanh(x.ctypes.data_as(c_double_p), out.ctypes.data_as(c_double_p),
                                   ctypes.c_int64(len(x)))
    xl = x.astype(np.longdouble)
    g = np.exp(2 * np.clip(xl, -5000, 5000))
    exact = np.where(xl > 300, np.longdouble(2.0), 2 * g / (1 + g))
    inside = np.abs(x) < 300
    rel = np.abs((out - exact) / np.where(exact > 0, exact, 1)).astype(np.float64)
    assert rel[inside].max() < 1.5e-15, rel[inside].max()
    assert np.array_equal(out[~inside], 1 + np.tanh(x[~inside]))
    nan = np.array([np.nan])
    hostlib.fwb_host_one_plus_tanh(nan.ctypes.data_as(c_double_p), nan.ctypes.data_as(c_double_p),
                                   ctypes.c_int64(1))
    assert np.isnan(nan[0])

    rng = np.random.default_rng(12)
    n = 200000
    spec = oracle.MODELS["fenton_karma"]
    u, st = _random_node_states("fenton_karma", n, rng)
    u[-n // 10:] = rng.uniform(0.75, 0.95, n // 10)           # around the tanh's centre
    pvec = np.array([float(v) for v in spec["params"].values()], dtype=np.float64)
    diff = rng.uniform(-1, 1, n) * 0.01 + u
    outs = []
    for fn, extra in ((hostlib.fwb_host_ionic, (MODEL_IDS["fenton_karma"],)),
                      (hostlib.fwb_host_fk_fast, ())):
        un = diff.copy()
        s2 = [s.copy() for s in st]
        arr = (c_double_p * len(s2))(*[s.ctypes.data_as(c_double_p) for s in s2])
        rc = fn(*extra, un.ctypes.data_as(c_double_p), u.ctypes.data_as(c_double_p), arr,
                ctypes.c_int64(n), ctypes.c_double(0.01), pvec.ctypes.data_as(c_double_p))
        assert rc == 0
        outs.append((un, s2))
    (un_r, st_r), (un_f, st_f) = outs
    assert np.max(np.abs(un_f - un_r)) < 1e-16 * 50          # dt * J_si differs by a few ulp
    for a, b in zip(st_f, st_r):
        assert np.array_equal(a, b)                           # v, w do not depend on the tanh
