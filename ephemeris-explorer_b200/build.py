"""Build libee_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

`-fmad=false`: the parity kernels must never contract a*b+c (the reference is Rust, which does not); the throughput
kernels call fma() explicitly, which this flag does not affect.
"""
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libee_b200.so"
SOURCES = ["ee_nbody.cu", "ee_solout.cu", "ee_ships.cu", "ee_small.cu", "ee_capi.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xptxas", "-v",
    # host-side parity arithmetic (solution bounds, sampling schedule, step factors) has a + b*c shapes; the reference is
    # Rust, which never contracts them, so the host compiler must not either (GCC's default on aarch64 would)
    "-Xcompiler", "-ffp-contract=off",
]


def _stale(out: Path, deps) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    headers = list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [HERE.parent / "include" / "ee_b200.h"]
    objs = []
    for src in SOURCES:
        s = CSRC / src
        if not s.exists():
            continue
        o = CSRC / (s.stem + ".o")
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + FLAGS + ["-c", str(s), "-o", str(o)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed for %s" % src)
            (CSRC / (s.stem + ".ptxas.log")).write_text(r.stderr)
        objs.append(str(o))
    if force or _stale(LIB, objs):
        cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB)] + objs + ["-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    build_tools(force)
    return LIB


def build_tools(force: bool = False) -> Path:
    """tools/planner_loop.cpp -> tools/bin/planner_loop: the Prediction Planner's call pattern driven through the C ABI
    (C++ host code over include/ee_b200.hpp), used by bench.py for the C2 end-to-end figure."""
    root = HERE.parent
    src = root / "tools" / "planner_loop.cpp"
    exe = root / "tools" / "bin" / "planner_loop"
    if not src.exists():
        return exe
    exe.parent.mkdir(parents=True, exist_ok=True)
    if force or _stale(exe, [src, root / "include" / "ee_b200.hpp", root / "include" / "ee_b200.h", LIB]):
        cmd = [os.environ.get("HOST_CXX", "/usr/bin/g++"), "-std=c++17", "-O2", "-ffp-contract=off", "-I", str(root / "include"),
               str(src), "-o", str(exe), "-L", str(HERE), "-lee_b200", "-Wl,-rpath,$ORIGIN/../../ephemeris-explorer_b200"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("building tools/planner_loop failed")
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
