"""
Builds finitewave_b200/libfinitewave_b200.so (sm_100a only) with nvcc.

    python -m finitewave_b200.build [--force] [--verbose]

The library is built IN-TREE so that it travels with the repository snapshot to
the GPU box; it is git-ignored.  No GPU is needed to build (nvcc cross-compiles).
-fmad=false: the kernels keep the reference's IEEE operation order (no FMA
contraction), see DESIGN.md section 5.
"""
import concurrent.futures
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
BUILD = PKG / "_build"
LIB = PKG / "libfinitewave_b200.so"

SOURCES = ["sim.cu", "aux_kernels.cu", "weights.cu", "step_nomodel.cu", "step_ap.cu",
           "step_barkley.cu", "step_ms.cu", "step_fk.cu", "step_bo.cu", "step_lr91.cu", "step_tp06.cu", "step_court.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden",
    "-Xptxas", "-v",
]


def nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (Path(cand).exists() or cand == "nvcc"):
            return cand
    raise RuntimeError("nvcc not found")


def _deps_mtime():
    heads = list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "finitewave_b200.h"]
    return max(h.stat().st_mtime for h in heads)


def _compile(src, verbose):
    obj = BUILD / (src.replace(".cu", ".o"))
    srcp = CSRC / src
    if (obj.exists() and obj.stat().st_mtime >= srcp.stat().st_mtime
            and obj.stat().st_mtime >= _deps_mtime()):
        return obj, ""
    cmd = [nvcc(), *NVCC_FLAGS, *os.environ.get("FWB_EXTRA_FLAGS", "").split(), "-c", str(srcp),
           "-o", str(obj)]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}\n{p.stderr}")
    (BUILD / (src + ".ptxas.log")).write_text(p.stderr)
    return obj, p.stderr if verbose else ""


def build(force=False, verbose=False):
    BUILD.mkdir(exist_ok=True)
    sources = [s for s in SOURCES if (CSRC / s).exists()]
    if force:
        for o in BUILD.glob("*.o"):
            o.unlink()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), sources))
    objs = [str(o) for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    newest = max(Path(o).stat().st_mtime for o in objs)
    if force or not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [nvcc(), "-shared", "-o", str(LIB), *objs,
               "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
               "-Xlinker", "--no-undefined"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB


def build_variant(suffix, flags, sources=("step_tp06.cu",)):
    """A second library for A/B measurements: `sources` recompiled with extra `flags`,
    every other object shared with the default build -> libfinitewave_b200_<suffix>.so
    (select it with FWB_LIB=<path>)."""
    build()
    objs = []
    for src in SOURCES:
        if src in sources:
            obj = BUILD / src.replace(".cu", f"_{suffix}.o")
            cmd = [nvcc(), *NVCC_FLAGS, *flags, "-c", str(CSRC / src), "-o", str(obj)]
            p = subprocess.run(cmd, capture_output=True, text=True)
            if p.returncode != 0:
                raise RuntimeError(f"nvcc failed for {src} [{suffix}]:\n{p.stdout}\n{p.stderr}")
            (BUILD / f"{src}.{suffix}.ptxas.log").write_text(p.stderr)
        else:
            obj = BUILD / src.replace(".cu", ".o")
        objs.append(str(obj))
    lib = PKG / f"libfinitewave_b200_{suffix}.so"
    cmd = [nvcc(), "-shared", "-o", str(lib), *objs,
           "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
           "-Xlinker", "--no-undefined"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return lib


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2].split()))
    else:
        lib = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
        print(lib)
