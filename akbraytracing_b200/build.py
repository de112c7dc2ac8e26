"""Build libakb_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m akbraytracing_b200.build [--force] [--verbose]

The shared library is a plain C-ABI (include/akb_b200.h): no torch, no pybind.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libakb_b200.so")
SOURCES = ["common.cu", "fresnel.cu", "sharded.cu", "ray.cu", "handoff.cu", "probe.cu"]
HEADERS = [os.path.join(CSRC, "akb_common.cuh"), os.path.join(os.path.dirname(PKG), "include", "akb_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",  # never contract the reference's a*b+c into FMA; fused ops are written explicitly
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-fast-math", "-Xcompiler", "-ffp-contract=off",
    "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA toolkit is required to build libakb_b200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    flags = list(NVCC_FLAGS)
    if os.environ.get("AKB_AB_VARIANTS") == "1":  # extra pair-kernel variants for tools/variant_bench.py
        flags.append("-DAKB_AB_VARIANTS")
    objs = []
    build_dir = os.path.join(PKG, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(build_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if src == "fresnel.cu" and os.environ.get("AKB_FRESNEL_PTXAS"):  # A/B of ptxas options on the pair kernels
            cmd[1:1] = [f"-Xptxas={o}" for o in os.environ["AKB_FRESNEL_PTXAS"].split()]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(out, file=sys.stderr)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
            "-Xcompiler", "-fPIC", "-o", LIB, *objs, "-ldl"]
    subprocess.run(link, check=True)
    return LIB


def ensure_built() -> str:
    """Build if the library is missing or older than its sources; safe to call from several ranks
    at once (an exclusive file lock serialises them, late comers find the library up to date).
    Used by tests, bench.py and __graft_entry__ -- the product import path never builds implicitly."""
    import fcntl
    if not needs_build():
        return LIB
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    with open(os.path.join(PKG, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return build()
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
