"""
Shared parity cases.

A case is a plain dict describing tissue, model, stimuli and trackers.  The
same dict drives three things:

  * ``build_and_run(fw, case)``  -- builds the simulation through the PUBLIC
    finitewave API of module ``fw`` and runs it.  ``fw`` is either the live
    reference (tests/golden/make_golden.py, in the build container only) or
    ``finitewave_b200`` (the CUDA product, ``-m gpu`` tests).  Identical driver
    code for both is the drop-in proof.
  * ``oracle.simulate(case)``    -- the CPU restatement.
  * tests/golden/<name>.npz      -- outputs of the live reference.

Inputs are regenerated deterministically from seeds (np.random.default_rng),
never from the reference's global-state generators.
"""
import numpy as np

MODEL_CLASS = {
    "aliev_panfilov": "AlievPanfilov",
    "barkley": "Barkley",
    "mitchell_schaeffer": "MitchellSchaeffer",
    "fenton_karma": "FentonKarma",
    "bueno_orovio": "BuenoOrovio",
    "courtemanche": "Courtemanche",
    "luo_rudy91": "LuoRudy91",
    "tp06": "TP06",
}

STATE_VARS = {
    "aliev_panfilov": ["u", "v"],
    "barkley": ["u", "v"],
    "mitchell_schaeffer": ["u", "h"],
    "fenton_karma": ["u", "v", "w"],
    "bueno_orovio": ["u", "v", "w", "s"],
    "courtemanche": ["u", "nai", "ki", "cai", "caup", "carel", "m", "h", "j_", "d", "f", "oa",
                     "oi", "ua", "ui", "xr", "xs", "fca", "irel", "vrel", "urel", "wrel"],
    "luo_rudy91": ["u", "m", "h", "j", "d", "f", "x", "cai"],
    "tp06": ["u", "cai", "casr", "cass", "nai", "Ki", "m", "h", "j", "xr1", "xr2",
             "xs", "r", "s", "d", "f", "f2", "fcass", "rr", "oo"],
}


# --------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------
def random_fibrosis(shape, density, seed):
    """mesh with `2` where rng.random(shape) <= density (rule of
    diffuse_2d_pattern.py:86-87), boundary ring 0."""
    rng = np.random.default_rng(seed)
    mesh = np.ones(shape, dtype=np.int8)
    mesh[rng.random(shape) <= density] = 2
    return mesh


def uniform_fibers_2d(shape, alpha):
    f = np.zeros((*shape, 2))
    f[..., 0] = np.cos(alpha)
    f[..., 1] = np.sin(alpha)
    return f


def rotating_fibers_3d(shape):
    """examples/basics/3D/slab_with_fibers_3d.py:64-70"""
    n_i, n_j, n_k = shape
    phi_k = np.linspace(-np.pi / 3, np.pi / 2, n_k - 2)
    f = np.zeros((n_i, n_j, n_k, 3))
    for k, phi in enumerate(phi_k):
        f[:, :, k + 1, 0] = np.cos(phi)
        f[:, :, k + 1, 1] = np.sin(phi)
    return f


def random_fibers(shape, seed):
    rng = np.random.default_rng(seed)
    f = rng.normal(size=(*shape, len(shape)))
    f /= np.linalg.norm(f, axis=-1, keepdims=True)
    return f


def ventricle_shell(shape):
    """Half prolate-ellipsoid shell scaled into `shape`:
    centre (n/2, n/2, 0.86 n), outer semi-axes (0.39, 0.39, 0.82) n, inner
    (0.31, 0.31, 0.74) n, open base at k = 0.86 n. Returns (mesh, fibers)."""
    n_i, n_j, n_k = shape
    ii, jj, kk = np.meshgrid(np.arange(n_i), np.arange(n_j), np.arange(n_k), indexing="ij")
    ci, cj, ck = n_i / 2, n_j / 2, 0.86 * n_k
    x, y, z = (ii - ci), (jj - cj), (kk - ck)
    outer = (x / (0.39 * n_i)) ** 2 + (y / (0.39 * n_j)) ** 2 + (z / (0.82 * n_k)) ** 2
    inner = (x / (0.31 * n_i)) ** 2 + (y / (0.31 * n_j)) ** 2 + (z / (0.74 * n_k)) ** 2
    wall = (outer <= 1.0) & (inner > 1.0) & (kk <= ck)
    mesh = np.zeros(shape, dtype=np.int8)
    mesh[wall] = 1
    # transmural coordinate 0 (endo) .. 1 (epi), helix angle -60..+60 deg
    depth = np.clip((np.sqrt(inner) - 1.0) / np.maximum(np.sqrt(inner) - np.sqrt(outer), 1e-9),
                    0.0, 1.0)
    helix = np.deg2rad(-60.0 + 120.0 * depth)
    r = np.sqrt(x ** 2 + y ** 2) + 1e-12
    circ = np.stack([-y / r, x / r, np.zeros_like(r)], axis=-1)
    longi = np.zeros_like(circ)
    longi[..., 2] = 1.0
    f = np.cos(helix)[..., None] * circ + np.sin(helix)[..., None] * longi
    f /= np.linalg.norm(f, axis=-1, keepdims=True)
    f[~wall] = 0.0
    return mesh, f


# --------------------------------------------------------------------------
# the case list
# --------------------------------------------------------------------------
def make_cases():
    cases = []

    # C1: README quick start, exactly (README.rst:417-451)
    cases.append(dict(
        name="c1_ap2d_readme", model="aliev_panfilov", shape=[100, 100],
        dt=0.01, dr=0.25, t_max=10,
        stims=[dict(kind="voltage_coord", t=0, value=1, box=[1, 99, 1, 3])],
        trackers=[dict(kind="activation_time", threshold=0.5, step=100)]))

    # Barkley 2D iso, fibrosis, two voltage stims (spiral protocol of
    # tests/test_trackers_2d.py:24-35), action potential tracker
    cases.append(dict(
        name="barkley2d_iso_fib", model="barkley", shape=[48, 40],
        dt=0.01, dr=0.25, t_max=3.6,
        mesh=random_fibrosis([48, 40], 0.15, 1),
        stims=[dict(kind="voltage_coord", t=0, value=1, box=[0, 48, 0, 5]),
               dict(kind="voltage_coord", t=3.1, value=1, box=[0, 24, 0, 40])],
        trackers=[dict(kind="action_potential", cell_ind=[20, 21], step=3)]))

    # C2 reduced: FK 2D aniso-9 + 30 % fibrosis, S1-S2
    cases.append(dict(
        name="c2_fk2d_aniso_fib", model="fenton_karma", shape=[64, 64],
        dt=0.01, dr=0.25, t_max=10,
        mesh=random_fibrosis([64, 64], 0.30, 2),
        fibers=uniform_fibers_2d([64, 64], 0.25 * np.pi),
        stims=[dict(kind="voltage_coord", t=0, value=1, box=[0, 64, 0, 5]),
               dict(kind="voltage_coord", t=6, value=1, box=[0, 32, 0, 64])],
        trackers=[dict(kind="activation_time", threshold=0.5, step=1)]))

    # FK 2D iso with current stim + u_max clamp + conductivity map
    rng = np.random.default_rng(3)
    cases.append(dict(
        name="fk2d_iso_current", model="fenton_karma", shape=[40, 33],
        dt=0.01, dr=0.25, t_max=6,
        conductivity=0.3 + 0.7 * rng.random([40, 33]),
        stims=[dict(kind="current_coord", t=0.5, value=5, duration=0.5,
                    box=[1, 6, 1, 32], u_max=0.9)],
        trackers=[dict(kind="ecg", coords=[[20, 16, 3], [5, 30, 1]], step=10)]))

    # C3 reduced: MS 3D iso-7 focal + activation time every step
    cases.append(dict(
        name="c3_ms3d_iso_focal", model="mitchell_schaeffer", shape=[24, 24, 24],
        dt=0.01, dr=0.25, t_max=10,
        stims=[dict(kind="voltage_coord", t=0, value=1, box=[10, 14, 10, 14, 10, 14])],
        trackers=[dict(kind="activation_time", threshold=0.5, step=1)]))

    # MS 3D aniso-19, random fibres + fibrosis, matrix stim, ECG
    mat = np.zeros([16, 18, 20])
    mat[2:6, 3:9, 4:12] = 1
    cases.append(dict(
        name="ms3d_aniso_random", model="mitchell_schaeffer", shape=[16, 18, 20],
        dt=0.01, dr=0.25, t_max=4,
        mesh=random_fibrosis([16, 18, 20], 0.2, 4),
        fibers=random_fibers([16, 18, 20], 5),
        stims=[dict(kind="voltage_matrix", t=0.2, value=1, matrix=mat)],
        trackers=[dict(kind="ecg", coords=[[8, 9, 25], [0, 0, 0]], step=5)]))

    # AP 3D aniso rotating fibres (slab_with_fibers_3d.py), current matrix stim
    mat = np.zeros([14, 14, 12])
    mat[5:9, 5:9, :] = 2
    cases.append(dict(
        name="ap3d_aniso_rot", model="aliev_panfilov", shape=[14, 14, 12],
        dt=0.01, dr=0.25, t_max=5,
        fibers=rotating_fibers_3d([14, 14, 12]),
        stims=[dict(kind="current_matrix", t=0, value=5, duration=0.5, matrix=mat)],
        trackers=[dict(kind="action_potential", cell_ind=[7, 7, 6], step=1)]))

    # Barkley 3D iso, special Dirichlet boundaries
    sb = np.zeros([12, 12, 12], dtype=np.int8)
    sb[6, 3:9, 3:9] = 1
    cases.append(dict(
        name="barkley3d_iso_special", model="barkley", shape=[12, 12, 12],
        dt=0.01, dr=0.25, t_max=3, special_boundaries=sb,
        stims=[dict(kind="voltage_coord", t=0, value=1, box=[0, 4, 0, 12, 0, 12])]))

    # LR91 2D iso (test_models_2d.py protocol, current stim)
    cases.append(dict(
        name="lr91_2d_iso", model="luo_rudy91", shape=[20, 12],
        dt=0.01, dr=0.25, t_max=10,
        stims=[dict(kind="current_coord", t=0, value=100, duration=1, box=[1, 4, 1, 11])],
        trackers=[dict(kind="multi_variable", cell_ind=[10, 6], step=10,
                       vars=["u", "m", "cai"])]))

    # LR91 3D aniso + fibrosis
    cases.append(dict(
        name="lr91_3d_aniso_fib", model="luo_rudy91", shape=[12, 10, 10],
        dt=0.01, dr=0.25, t_max=8,
        mesh=random_fibrosis([12, 10, 10], 0.2, 6),
        fibers=rotating_fibers_3d([12, 10, 10]),
        stims=[dict(kind="voltage_coord", t=0, value=-20, box=[0, 4, 0, 10, 0, 10])],
        trackers=[dict(kind="activation_time", step=1)]))

    # TP06 2D iso, SURVEY App. C fingerprint case but 1000 steps
    cases.append(dict(
        name="tp06_2d_iso", model="tp06", shape=[20, 5],
        dt=0.01, dr=0.25, t_max=10,
        stims=[dict(kind="voltage_coord", t=0, value=-20, box=[0, 5, 0, 5])],
        trackers=[dict(kind="action_potential", cell_ind=[2, 2], step=10)]))

    # TP06 2D aniso + fibrosis
    cases.append(dict(
        name="tp06_2d_aniso_fib", model="tp06", shape=[24, 16],
        dt=0.01, dr=0.25, t_max=10,
        mesh=random_fibrosis([24, 16], 0.25, 7),
        fibers=uniform_fibers_2d([24, 16], 0.3 * np.pi),
        stims=[dict(kind="voltage_coord", t=0, value=-20, box=[0, 5, 0, 16])],
        trackers=[dict(kind="activation_time", step=1)]))

    # C5 reduced: TP06 3D aniso-19 slab, rotating fibres, face stimulus
    cases.append(dict(
        name="c5_tp06_3d_aniso_slab", model="tp06", shape=[12, 12, 10],
        dt=0.01, dr=0.25, t_max=10,
        fibers=rotating_fibers_3d([12, 12, 10]),
        stims=[dict(kind="voltage_coord", t=0, value=-20, box=[0, 4, 0, 12, 0, 10])],
        trackers=[dict(kind="activation_time", step=1)]))

    # C4 reduced: TP06 3D ventricle-shaped shell, helix fibres, apex stimulus
    vm, vf = ventricle_shell([20, 20, 24])
    cases.append(dict(
        name="c4_tp06_3d_ventricle", model="tp06", shape=[20, 20, 24],
        dt=0.01, dr=0.25, t_max=10, mesh=vm, fibers=vf,
        stims=[dict(kind="voltage_coord", t=0, value=-20, box=[0, 20, 0, 20, 0, 4])],
        trackers=[dict(kind="activation_time", step=1)]))

    # TP06 3D iso-7 with current stim (test_models_3d.py protocol)
    cases.append(dict(
        name="tp06_3d_iso_current", model="tp06", shape=[9, 5, 5],
        dt=0.01, dr=0.25, t_max=6,
        stims=[dict(kind="current_coord", t=0, value=100, duration=1,
                    box=[1, 3, 1, 4, 1, 4])],
        trackers=[dict(kind="action_potential", cell_ind=[4, 2, 2], step=1)]))

    # Bueno-Orovio 2D iso, current stimulus (test_models_2d.py protocol) + spiral S2
    cases.append(dict(
        name="bo2d_iso_current", model="bueno_orovio", shape=[40, 30],
        dt=0.01, dr=0.25, t_max=10,
        stims=[dict(kind="current_coord", t=0, value=5, duration=0.5, box=[0, 4, 0, 30]),
               dict(kind="voltage_coord", t=6, value=1, box=[0, 20, 0, 30])],
        trackers=[dict(kind="multi_variable", cell_ind=[20, 15], step=5, vars=["u", "v", "w", "s"])]))

    # Bueno-Orovio 3D aniso + fibrosis, non-default parameter
    cases.append(dict(
        name="bo3d_aniso_fib", model="bueno_orovio", shape=[14, 12, 10],
        dt=0.01, dr=0.25, t_max=8, params=dict(tau_si=2.0),
        mesh=random_fibrosis([14, 12, 10], 0.2, 12),
        fibers=rotating_fibers_3d([14, 12, 10]),
        stims=[dict(kind="voltage_coord", t=0, value=1, box=[0, 4, 0, 12, 0, 10])],
        trackers=[dict(kind="activation_time", threshold=0.3, step=1)]))

    # Courtemanche 2D iso, current stimulus (test_models_2d.py protocol)
    cases.append(dict(
        name="court2d_iso_current", model="courtemanche", shape=[16, 9],
        dt=0.01, dr=0.25, t_max=10,
        stims=[dict(kind="current_coord", t=0, value=100, duration=1, box=[0, 3, 0, 9])],
        trackers=[dict(kind="action_potential", cell_ind=[8, 4], step=10)]))

    # Courtemanche 3D aniso + fibrosis
    cases.append(dict(
        name="court3d_aniso_fib", model="courtemanche", shape=[12, 10, 8],
        dt=0.01, dr=0.25, t_max=8,
        mesh=random_fibrosis([12, 10, 8], 0.2, 13),
        fibers=rotating_fibers_3d([12, 10, 8]),
        stims=[dict(kind="voltage_coord", t=0, value=-20, box=[0, 4, 0, 10, 0, 8])],
        trackers=[dict(kind="activation_time", step=1)]))

    # AP 2D aniso, non-default parameters and initial conditions
    cases.append(dict(
        name="ap2d_aniso_params", model="aliev_panfilov", shape=[37, 45],
        dt=0.005, dr=0.2, t_max=4,
        params=dict(a=0.12, k=7.5, mu_1=0.25), init=dict(v=0.05),
        fibers=random_fibers([37, 45], 8),
        stims=[dict(kind="voltage_coord", t=0.1, value=1, box=[1, 36, 1, 4])],
        trackers=[dict(kind="activation_time", threshold=0.5, step=7, start_time=0.5,
                       end_time=3.0)]))
    # LocalActivationTime + Period trackers on a re-entrant Barkley wave:
    # S1-S2 spiral protocol so that nodes activate more than once (several layers)
    cases.append(dict(
        name="barkley2d_lat_period", model="barkley", shape=[40, 40],
        dt=0.01, dr=0.25, t_max=20,
        stims=[dict(kind="voltage_coord", t=0, value=1, box=[0, 40, 0, 5]),
               dict(kind="voltage_coord", t=3.1, value=1, box=[0, 20, 0, 40])],
        trackers=[dict(kind="local_activation_time", threshold=0.5, step=2),
                  dict(kind="period", threshold=0.5, step=1,
                       cell_ind=[[10, 10], [20, 30], [30, 15]])]))
    # the 3D classes, AP focal stimulus fired twice
    cases.append(dict(
        name="ap3d_lat_period", model="aliev_panfilov", shape=[14, 12, 10],
        dt=0.01, dr=0.25, t_max=60,
        stims=[dict(kind="voltage_coord", t=0, value=1, box=[0, 3, 0, 12, 0, 10]),
               dict(kind="voltage_coord", t=35, value=1, box=[0, 3, 0, 12, 0, 10])],
        trackers=[dict(kind="local_activation_time", threshold=0.5, step=5),
                  dict(kind="period", threshold=0.5, step=5, cell_ind=[[7, 6, 5], [12, 3, 8]])]))

    # StimCurrentArea2D / 3D: coordinate lists with duplicates, fibrotic entries and a u_max
    # clamp (numpy applies `u[inds] += x` once per distinct node)
    rng = np.random.default_rng(31)
    pts2 = np.column_stack([rng.integers(2, 12, 60), rng.integers(2, 30, 60)])
    pts2 = np.vstack([pts2, pts2[:10]])
    cases.append(dict(
        name="fk2d_current_area", model="fenton_karma", shape=[36, 32],
        dt=0.01, dr=0.25, t_max=5,
        mesh=random_fibrosis([36, 32], 0.15, 32),
        stims=[dict(kind="current_area", t=0.3, value=6, duration=0.6, coords=pts2, u_max=0.95)],
        trackers=[dict(kind="activation_time", threshold=0.4, step=1)]))
    pts3 = np.column_stack([rng.integers(1, 5, 80), rng.integers(1, 11, 80), rng.integers(1, 9, 80)])
    cases.append(dict(
        name="ms3d_current_area", model="mitchell_schaeffer", shape=[14, 12, 10],
        dt=0.01, dr=0.25, t_max=4,
        stims=[dict(kind="current_area", t=0, value=4, duration=0.4, coords=pts3)],
        trackers=[dict(kind="action_potential", cell_ind=[3, 5, 4], step=2)]))
    # StimVoltageListMatrix3D: a voltage ramp, one list entry per firing
    mat = np.zeros([12, 10, 10])
    mat[1:4, :, :] = 1
    cases.append(dict(
        name="ap3d_voltage_list", model="aliev_panfilov", shape=[12, 10, 10],
        dt=0.01, dr=0.25, t_max=4,
        mesh=random_fibrosis([12, 10, 10], 0.1, 33),
        stims=[dict(kind="voltage_list_matrix", t=0.2, duration=0.3, matrix=mat,
                    value=list(np.linspace(0.1, 1.0, 40)))],
        trackers=[dict(kind="action_potential", cell_ind=[6, 5, 5], step=1)]))

    # SpiralWaveCore trackers: Barkley spiral (the protocol of the
    # reference's tests/test_trackers_2d.py:25-40, smaller), tips every 10 steps
    cases.append(dict(
        name="barkley2d_spiral_core", model="barkley", shape=[100, 100],
        dt=0.01, dr=0.25, t_max=14,
        stims=[dict(kind="voltage_coord", t=0, value=1, box=[0, 100, 0, 3]),
               dict(kind="voltage_coord", t=5, value=1, box=[0, 50, 0, 100])],
        trackers=[dict(kind="spiral_core", threshold=0.5, step=10, start_time=6)]))
    cases.append(dict(
        name="barkley3d_spiral_core", model="barkley", shape=[48, 48, 5],
        dt=0.01, dr=0.25, t_max=9,
        stims=[dict(kind="voltage_coord", t=0, value=1, box=[0, 48, 0, 3, 0, 5]),
               dict(kind="voltage_coord", t=2.4, value=1, box=[0, 24, 0, 48, 0, 5])],
        trackers=[dict(kind="spiral_core", threshold=0.5, step=20, start_time=3)]))

    # SymmetricStencil2D: cell-centred 9-point weights, random fibres,
    # fibrosis and a conductivity map; the apply kernel is the asymmetric 9-point one
    rng = np.random.default_rng(21)
    cases.append(dict(
        name="ap2d_sym_fib", model="aliev_panfilov", shape=[40, 36],
        dt=0.01, dr=0.25, t_max=5, stencil="sym",
        mesh=random_fibrosis([40, 36], 0.2, 22),
        fibers=random_fibers([40, 36], 23),
        conductivity=0.4 + 0.6 * rng.random([40, 36]),
        stims=[dict(kind="voltage_coord", t=0, value=1, box=[0, 40, 0, 5])],
        trackers=[dict(kind="activation_time", threshold=0.5, step=1)]))
    # the same stencil with uniform fibres on FK (scalar conductivity)
    cases.append(dict(
        name="fk2d_sym_uniform", model="fenton_karma", shape=[32, 30],
        dt=0.01, dr=0.25, t_max=4, stencil="sym",
        fibers=uniform_fibers_2d([32, 30], 0.3 * np.pi),
        stims=[dict(kind="voltage_coord", t=0, value=1, box=[0, 5, 0, 30])],
        trackers=[dict(kind="action_potential", cell_ind=[16, 15], step=2)]))
    return cases


def case_by_name(name):
    for c in make_cases():
        if c["name"] == name:
            return c
    raise KeyError(name)


# --------------------------------------------------------------------------
# Build + run through the public finitewave API of module `fw`
# --------------------------------------------------------------------------
def _suffix(case):
    return "2D" if len(case["shape"]) == 2 else "3D"


def build_model(fw, case):
    sfx = _suffix(case)
    tissue = getattr(fw, "CardiacTissue" + sfx)(list(case["shape"]))
    if case.get("mesh") is not None:
        tissue.mesh = np.array(case["mesh"])
    if case.get("conductivity") is not None:
        tissue.conductivity = case["conductivity"]
    if case.get("fibers") is not None:
        tissue.fibers = np.array(case["fibers"])
    if case.get("special_boundaries") is not None:
        tissue.special_boundaries = np.array(case["special_boundaries"])

    model = getattr(fw, MODEL_CLASS[case["model"]] + sfx)()
    model.dt, model.dr, model.t_max = case["dt"], case["dr"], case["t_max"]
    model.prog_bar = False
    for k, v in case.get("params", {}).items():
        setattr(model, k, v)
    for k, v in case.get("init", {}).items():
        setattr(model, "init_" + k.rstrip("_"), v)
    if "D_model" in case:
        model.D_model = case["D_model"]
    model.cardiac_tissue = tissue
    if case.get("stencil") == "sym":
        model.stencil = fw.SymmetricStencil2D()

    seq = fw.StimSequence()
    for s in case.get("stims", []):
        kind = s["kind"]
        if kind == "voltage_coord":
            st = getattr(fw, "StimVoltageCoord" + sfx)(s["t"], s["value"], *s["box"])
        elif kind == "current_coord":
            st = getattr(fw, "StimCurrentCoord" + sfx)(s["t"], s["value"], s["duration"],
                                                       *s["box"], u_max=s.get("u_max"))
        elif kind == "voltage_matrix":
            st = getattr(fw, "StimVoltageMatrix" + sfx)(s["t"], s["value"], s["matrix"])
        elif kind == "current_matrix":
            st = getattr(fw, "StimCurrentMatrix" + sfx)(s["t"], s["value"], s["duration"],
                                                        s["matrix"], u_max=s.get("u_max"))
        elif kind == "voltage_list_matrix":
            st = getattr(fw, "StimVoltageListMatrix" + sfx)(s["t"], list(s["value"]), s["duration"],
                                                            np.array(s["matrix"]))
        elif kind == "current_area":
            st = getattr(fw, "StimCurrentArea" + sfx)(s["t"], s["value"], s["duration"],
                                                      coords=np.array(s["coords"]),
                                                      u_max=s.get("u_max"))
        else:
            raise ValueError(kind)
        seq.add_stim(st)
    model.stim_sequence = seq

    tseq = fw.TrackerSequence()
    trackers = []
    for t in case.get("trackers", []):
        kind = t["kind"]
        if kind == "activation_time":
            tr = getattr(fw, "ActivationTime" + sfx + "Tracker")()
            if "threshold" in t:
                tr.threshold = t["threshold"]
        elif kind == "ecg":
            tr = getattr(fw, "ECG" + sfx + "Tracker")()
            tr.measure_coords = np.array(t["coords"])
        elif kind == "action_potential":
            tr = getattr(fw, "ActionPotential" + sfx + "Tracker")()
            tr.cell_ind = t["cell_ind"]
        elif kind == "multi_variable":
            tr = getattr(fw, "MultiVariable" + sfx + "Tracker")()
            tr.cell_ind = t["cell_ind"]
            tr.var_list = list(t["vars"])
        elif kind == "local_activation_time":
            tr = getattr(fw, "LocalActivationTime" + sfx + "Tracker")()
            if "threshold" in t:
                tr.threshold = t["threshold"]
        elif kind == "spiral_core":
            tr = getattr(fw, "SpiralWaveCore" + sfx + "Tracker")()
            if "threshold" in t:
                tr.threshold = t["threshold"]
        elif kind == "period":
            tr = getattr(fw, "Period" + sfx + "Tracker")()
            tr.cell_ind = t["cell_ind"]
            if "threshold" in t:
                tr.threshold = t["threshold"]
        else:
            raise ValueError(kind)
        for k in ("step", "start_time", "end_time"):
            if k in t:
                setattr(tr, k, t[k])
        tseq.add_tracker(tr)
        trackers.append((t, tr))
    model.tracker_sequence = tseq
    return model, trackers


def collect_outputs(case, model, trackers):
    out = {"t": np.float64(model.t), "step": np.int64(model.step)}
    for v in STATE_VARS[case["model"]]:
        out[v] = np.array(getattr(model, v), dtype=np.float64)
    for i, (t, tr) in enumerate(trackers):
        if t["kind"] == "multi_variable":
            for v in t["vars"]:
                out[f"tracker{i}_{v}"] = np.array(tr.output[v], dtype=np.float64)
        elif t["kind"] == "spiral_core":
            cols = ["x", "y", "time", "step"] if len(case["shape"]) == 2 else \
                ["x", "y", "z", "time", "step"]
            out[f"tracker{i}"] = np.array(tr.output[cols], dtype=np.float64).reshape(-1, len(cols))
        elif t["kind"] == "period":
            # the layers (n_layers, n_cells); `output` is a DataFrame of their differences
            tr.output
            out[f"tracker{i}"] = np.array(tr.act_t, dtype=np.float64)
        else:
            out[f"tracker{i}"] = np.array(tr.output, dtype=np.float64)
    return out


def build_and_run(fw, case):
    model, trackers = build_model(fw, case)
    model.run()
    return collect_outputs(case, model, trackers)


def max_rel_err(a, b):
    """max-abs error relative to max|reference| (north_star's definition)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return np.inf
    if a.size == 0:
        return 0.0
    scale = max(np.max(np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b)) / scale)
