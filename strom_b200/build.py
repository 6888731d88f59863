"""In-tree builds: the sm_100a engine (nvcc) and, for tests/bench only, the CPU oracle (gcc).

``python -m strom_b200.build`` builds everything; ``__graft_entry__.build()`` calls ``build_all``.
The shared objects stay in-tree (git-ignored) so they travel to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "strom_b200", "csrc")
ENGINE_LIB = os.path.join(CSRC, "libstrom_b200.so")
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "liboracle_beagle.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall", "--shared", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the engine cannot be built (there is no CPU fallback)")


def _stale(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def engine_sources():
    inc = os.path.join(ROOT, "include")
    return [os.path.join(CSRC, f) for f in ("engine.cu", "schedule.cpp", "kernels.cuh", "schedule.hpp")] + \
           [os.path.join(inc, f) for f in ("strom_beagle.h", "strom_b200.h")]


def build_engine(force: bool = False, verbose: bool = False) -> str:
    srcs = engine_sources()
    if force or _stale(ENGINE_LIB, srcs):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-o", ENGINE_LIB, os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "schedule.cpp")]
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        r = subprocess.run(cmd + ["-ccbin", "/usr/bin/g++"], capture_output=True, text=True, env=env)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed building libstrom_b200.so")
    return ENGINE_LIB


def build_oracle(force: bool = False) -> str:
    src = os.path.join(ROOT, "oracle", "beagle_cpu.c")
    if force or _stale(ORACLE_LIB, [src, os.path.join(ROOT, "include", "strom_beagle.h")]):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")] + (["-B"] if force else []),
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("building the CPU oracle failed")
    return ORACLE_LIB


def build_all(force: bool = False, verbose: bool = False):
    build_engine(force, verbose)
    build_oracle(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(ENGINE_LIB)
    print(ORACLE_LIB)
