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
    "-DSB200_NVTX",           # NVTX ranges around kernels (a), (b), (c) of an evaluation (header-only nvtx3; needs libdl)
    "-ldl",
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
    return [os.path.join(CSRC, f) for f in ("engine.cu", "patterns.cu", "schedule.cpp", "kernels.cuh", "schedule.hpp")] + \
           [os.path.join(inc, f) for f in ("strom_beagle.h", "strom_b200.h")]


def build_engine(force: bool = False, verbose: bool = False) -> str:
    srcs = engine_sources()
    if force or _stale(ENGINE_LIB, srcs):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-o", ENGINE_LIB, os.path.join(CSRC, "engine.cu"), os.path.join(CSRC, "patterns.cu"), os.path.join(CSRC, "schedule.cpp")]
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


HOST_DIR = os.path.join(ROOT, "strom_b200", "host")
HOST_LIB = os.path.join(HOST_DIR, "libstrom_host.so")
HOST_ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "libstrom_host_oracle.so")
GXX = "/usr/bin/g++"


def host_sources():
    hdrs = [os.path.join(HOST_DIR, f) for f in os.listdir(HOST_DIR) if f.endswith((".hpp", ".cpp"))]
    return hdrs + [os.path.join(ROOT, "include", f) for f in ("strom_beagle.h", "strom_b200.h")]


def build_host(force: bool = False) -> str:
    """C++ host mirror of the reference's Likelihood/Model/Data/Tree classes, linked to the CUDA engine."""
    build_engine()
    if force or _stale(HOST_LIB, host_sources() + [ENGINE_LIB]):
        cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-pthread", "-DSTROM_B200_ENGINE",
               "-o", HOST_LIB, os.path.join(HOST_DIR, "capi.cpp"), "-L" + CSRC, "-lstrom_b200", "-Wl,-rpath,$ORIGIN/../csrc"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("building libstrom_host.so failed")
    return HOST_LIB


def build_host_oracle(force: bool = False) -> str:
    """TEST ONLY: the same host classes on the reference's 13-call path, linked to the CPU checker."""
    build_oracle()
    if force or _stale(HOST_ORACLE_LIB, host_sources() + [ORACLE_LIB]):
        cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-pthread",
               "-o", HOST_ORACLE_LIB, os.path.join(HOST_DIR, "capi.cpp"), "-L" + os.path.dirname(ORACLE_LIB),
               "-loracle_beagle", "-Wl,-rpath,$ORIGIN"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("building the oracle-backed host library failed")
    return HOST_ORACLE_LIB


HOST_13CALL_LIB = os.path.join(HOST_DIR, "libstrom_host_13call.so")


def build_host_13call(force: bool = False) -> str:
    """The host classes exactly as the reference drives BeagleLib -- Likelihood::calcLogLikelihood issuing the 13 calls
    of include/strom_beagle.h one by one (no engine-level entry point) -- linked to the CUDA engine: the drop-in case."""
    build_engine()
    if force or _stale(HOST_13CALL_LIB, host_sources() + [ENGINE_LIB]):
        cmd = [GXX, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-pthread",
               "-o", HOST_13CALL_LIB, os.path.join(HOST_DIR, "capi.cpp"), "-L" + CSRC, "-lstrom_b200", "-Wl,-rpath,$ORIGIN/../csrc"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("building libstrom_host_13call.so failed")
    return HOST_13CALL_LIB


def build_all(force: bool = False, verbose: bool = False):
    build_engine(force, verbose)
    build_oracle(force)
    build_host(force)
    build_host_13call(force)
    build_host_oracle(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(ENGINE_LIB)
    print(ORACLE_LIB)
    print(HOST_LIB)
    print(HOST_ORACLE_LIB)
