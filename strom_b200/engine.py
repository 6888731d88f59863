"""Python face of the engine-level C ABI (include/strom_b200.h) plus ``ShardedLikelihood``, the
multi-GPU form of ``Likelihood::calcLogLikelihood`` (reference src/likelihood.hpp:390-448):
site patterns are sharded across ranks (one process per GPU), every rank holds all partials for
its slice, and the only exchange per evaluation is the sum of one double per rank.

PyTorch is plumbing only here: it owns the CUDA stream the engine runs on, the one-element
device tensor the kernels write the scalar into, and the NCCL all-reduce of that scalar.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import beagle
from .sharding import sum_over_ranks

_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)


class StromEvalArgs(C.Structure):
    _fields_ = [("stateFrequencies", _DP), ("eigenVectors", _DP), ("inverseEigenVectors", _DP), ("eigenValues", _DP),
                ("categoryRates", _DP), ("categoryWeights", _DP), ("matrixIndices", _IP), ("edgeLengths", _DP),
                ("edgeCount", C.c_int), ("operations", _IP), ("operationCount", C.c_int),
                ("rootParentBuffer", C.c_int), ("rootChildBuffer", C.c_int), ("rootMatrix", C.c_int),
                ("categoryCount", C.c_int)]


class EngineAPI:
    """sb200* entry points of libstrom_b200.so (raises if the library is missing: no fallback)."""

    def __init__(self):
        self.beagle = beagle.engine()
        lib = self.lib = self.beagle.lib
        sig = {
            "sb200Version": (C.c_char_p, []),
            "sb200SetDevice": (C.c_int, [C.c_int]),
            "sb200SetStream": (C.c_int, [C.c_int, C.c_void_p]),
            "sb200SetTipStatesU8": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
            "sb200SetAllTipStatesU8": (C.c_int, [C.c_int, C.c_void_p]),
            "sb200EvalBatchAsync": (C.c_int, [C.c_int, _IP, C.POINTER(StromEvalArgs)]),
            "sb200ReplayBatchAsync": (C.c_int, [C.c_int, _IP]),
            "sb200Eval": (C.c_int, [C.c_int, C.POINTER(StromEvalArgs), _DP]),
            "sb200EvalAsync": (C.c_int, [C.c_int, C.POINTER(StromEvalArgs), C.c_void_p]),
            "sb200ReplayAsync": (C.c_int, [C.c_int, C.c_void_p]),
            "sb200Finish": (C.c_int, [C.c_int, _DP]),
            "sb200Synchronize": (C.c_int, [C.c_int]),
            "sb200SetProfiling": (C.c_int, [C.c_int, C.c_int]),
            "sb200SetSubtreeScalers": (C.c_int, [C.c_int, C.c_int]),
            "sb200SetTransientPartials": (C.c_int, [C.c_int, C.c_int]),
            "sb200CompressPatterns": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, _DP, _IP]),
            "sb200PartialsTime": (C.c_int, [C.c_int, _DP, C.POINTER(C.c_long)]),
            "sb200LaunchCount": (C.c_long, [C.c_int]),
            "sb200LastTraffic": (C.c_int, [C.c_int, _DP, _DP]),
            "sb200DebugSchedule": (C.c_int, [_IP, C.c_int, C.c_int, C.c_int, _IP, _IP, _IP]),
            "sb200DebugPlan": (C.c_int, [_IP, C.c_int, C.c_int, C.c_int, C.c_int, _IP]),
            "sb200DebugPlanLaunches": (C.c_int, [_IP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _IP]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

    def check(self, code: int, what: str):
        if code != 0:
            raise beagle.BeagleError(what, code)

    def compress_patterns(self, codes: np.ndarray, device: int = 0):
        """Distinct site columns of codes[taxon][site] (uint8 < 32) in lexicographic order, and their counts -- on the GPU."""
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        ntaxa, nsites = codes.shape
        out = np.empty(codes.size, dtype=np.uint8)
        counts = np.empty(nsites, dtype=np.float64)
        npat = C.c_int()
        self.check(self.sb200CompressPatterns(device, ntaxa, nsites, codes.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                              counts.ctypes.data_as(_DP), C.byref(npat)), "compress patterns")
        P = npat.value
        return out[:ntaxa * P].reshape(ntaxa, P).copy(), counts[:P].copy()

    def plan(self, ops: np.ndarray, ntaxa: int, keep_subtree_scalers: bool = False) -> np.ndarray:
        """Planned operations [n][8] = dest, a, b, sa, sb, sOut, flags, launch -- host-only, needs no GPU."""
        ops = np.ascontiguousarray(ops, dtype=np.int32)
        n = ops.size // 7
        out = np.zeros((max(n, 1), 8), dtype=np.int32)
        r = self.sb200DebugPlan(ops.ctypes.data_as(_IP), n, ntaxa, ntaxa - 2, int(keep_subtree_scalers), out.ctypes.data_as(_IP))
        if r < 0:
            raise beagle.BeagleError("plan", r)
        return out[:r]

    def plan_launches(self, ops: np.ndarray, ntaxa: int, groups: int = 4, groups_tips: int = 0, keep_subtree_scalers: bool = False):
        """Launches of a plan [n][5] = chains, groups, tips-only, operations, operations of the longest group -- host-only."""
        ops = np.ascontiguousarray(ops, dtype=np.int32)
        out = np.zeros((64, 5), dtype=np.int32)
        r = self.sb200DebugPlanLaunches(ops.ctypes.data_as(_IP), ops.size // 7, ntaxa, ntaxa - 2, int(keep_subtree_scalers), groups,
                                        groups_tips, 64, out.ctypes.data_as(_IP))
        if r < 0:
            raise beagle.BeagleError("plan launches", r)
        return out[:r]

    def schedule(self, ops: np.ndarray, ntaxa: int):
        """(levels, chains, level of each op, chain of each op) -- host-only, needs no GPU."""
        ops = np.ascontiguousarray(ops, dtype=np.int32)
        n = ops.size // 7
        lev = np.zeros(n, dtype=np.int32)
        ch = np.zeros(n, dtype=np.int32)
        nch = C.c_int()
        r = self.sb200DebugSchedule(ops.ctypes.data_as(_IP), n, ntaxa, max(ntaxa - 2, n), lev.ctypes.data_as(_IP),
                                    ch.ctypes.data_as(_IP), C.byref(nch))
        if r < 0:
            raise beagle.BeagleError("schedule", r)
        return r, nch.value, lev, ch


_api = None


def api() -> EngineAPI:
    global _api
    if _api is None:
        _api = EngineAPI()
    return _api


class EvalArgs:
    """Keeps the numpy arrays behind a StromEvalArgs alive."""

    def __init__(self, model: dict, ops, pmatrix_index, edge_lengths, parent: int, child: int = 0):
        f = lambda a, t: np.ascontiguousarray(a, dtype=t)
        self.arrays = dict(freqs=f(model["freqs"], np.float64), evec=f(model["evec"], np.float64),
                           ivec=f(model["ivec"], np.float64), evals=f(model["evals"], np.float64),
                           rates=f(model["rates"], np.float64), probs=f(model["probs"], np.float64),
                           pidx=f(pmatrix_index, np.int32), elen=f(edge_lengths, np.float64), ops=f(ops, np.int32))
        a = self.arrays
        self.c = StromEvalArgs(a["freqs"].ctypes.data_as(_DP), a["evec"].ctypes.data_as(_DP), a["ivec"].ctypes.data_as(_DP),
                               a["evals"].ctypes.data_as(_DP), a["rates"].ctypes.data_as(_DP), a["probs"].ctypes.data_as(_DP),
                               a["pidx"].ctypes.data_as(_IP), a["elen"].ctypes.data_as(_DP), int(a["pidx"].size),
                               a["ops"].ctypes.data_as(_IP), int(a["ops"].size // 7), int(parent), int(child), int(child),
                               int(a["rates"].size))

    @property
    def h2d_bytes(self) -> int:
        a = self.arrays
        return int(sum(a[k].nbytes for k in ("freqs", "evec", "ivec", "evals", "rates", "probs", "pidx", "elen")))


class ShardedLikelihood:
    """One pattern slice of a (possibly multi-GPU) likelihood.

    ``tips_u8``: uint8 [ntaxa][local patterns]; ``weights``: float64 [local patterns].  With
    ``torch.distributed`` initialised, ``calc_log_likelihood`` all-reduces the per-rank scalars
    (NCCL, one double) before returning; otherwise it is the single-GPU call."""

    def __init__(self, tips_u8: np.ndarray, weights: np.ndarray, ncateg: int, device: int = 0, single: bool = False,
                 distributed: bool = False, transient_partials: bool = False):
        import torch
        self.torch = torch
        self.E = api()
        self.B = self.E.beagle
        self.ntaxa, self.P = tips_u8.shape
        self.C = ncateg
        self.distributed = distributed
        self.device = device
        torch.cuda.set_device(device)
        self.E.check(self.E.sb200SetDevice(device), "set device")
        req = (beagle.BEAGLE_FLAG_PRECISION_SINGLE if single else beagle.BEAGLE_FLAG_PRECISION_DOUBLE) | \
            beagle.BEAGLE_FLAG_SCALING_MANUAL
        n = self.ntaxa
        self.inst = self.B.beagleCreateInstance(n, n - 2, n, 4, self.P, 1, 2 * n - 3, ncateg, n - 1, None, 0,
                                                beagle.BEAGLE_FLAG_PROCESSOR_GPU, req, None)
        if self.inst < 0:
            raise beagle.BeagleError("failed to create instance", self.inst)
        self.stream = torch.cuda.Stream(device=device)
        self.E.check(self.E.sb200SetStream(self.inst, C.c_void_p(self.stream.cuda_stream)), "set stream")
        if transient_partials:
            self.E.check(self.E.sb200SetTransientPartials(self.inst, 1), "transient partials")
        self.out = torch.zeros(1, dtype=torch.float64, device="cuda:%d" % device)
        for t in range(n):
            row = np.ascontiguousarray(tips_u8[t])
            self.E.check(self.E.sb200SetTipStatesU8(self.inst, t, row.ctypes.data_as(C.c_void_p)), "set tip states")
        w = np.ascontiguousarray(weights, dtype=np.float64)
        self.E.check(self.B.beagleSetPatternWeights(self.inst, w.ctypes.data_as(_DP)), "set pattern weights")

    def eval_async(self, args: EvalArgs):
        """Enqueue one evaluation (+ the cross-rank sum) on this instance's stream; no host sync."""
        self.E.check(self.E.sb200EvalAsync(self.inst, C.byref(args.c), C.c_void_p(self.out.data_ptr())), "evaluate")
        if self.distributed:
            with self.torch.cuda.stream(self.stream):
                sum_over_ranks(self.out)

    def replay_async(self):
        self.E.check(self.E.sb200ReplayAsync(self.inst, C.c_void_p(self.out.data_ptr())), "replay")
        if self.distributed:
            with self.torch.cuda.stream(self.stream):
                sum_over_ranks(self.out)

    def calc_log_likelihood(self, args: EvalArgs) -> float:
        if not self.distributed:
            # one GPU: the root kernel writes the scalar into pinned host memory and sb200Finish polls for it
            self.E.check(self.E.sb200EvalAsync(self.inst, C.byref(args.c), None), "evaluate")
            res = C.c_double()
            code = self.E.sb200Finish(self.inst, C.byref(res))
            if code not in (0, -8):
                self.E.check(code, "finish")
            v = res.value
            if code == -8 or v != v:
                raise beagle.BeagleError("failed to calculate edge logLikelihoods", -8)
            return v
        self.eval_async(args)
        with self.torch.cuda.stream(self.stream):
            v = float(self.out.item())      # device -> host read of the scalar, synchronises the stream
        if v != v:
            raise beagle.BeagleError("failed to calculate edge logLikelihoods", -8)
        return v

    def launches(self) -> int:
        return int(self.E.sb200LaunchCount(self.inst))

    def close(self):
        if self.inst >= 0:
            self.E.sb200Synchronize(self.inst)
            self.B.beagleFinalizeInstance(self.inst)
            self.inst = -1
