"""ctypes binding of the 13-call boundary declared in include/strom_beagle.h.

``BeagleAPI(path)`` binds any shared library that exports those symbols; ``engine()`` returns
the binding of the product library (strom_b200/csrc/libstrom_b200.so, hand-written sm_100a
kernels) and raises if it has not been built -- there is no CPU fallback.

``calc_log_likelihood`` issues the calls in the order and with the arguments of
``Likelihood::calcLogLikelihood`` (reference src/likelihood.hpp:390-448).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ENGINE_LIB = os.environ.get("STROM_B200_LIB") or os.path.join(_HERE, "csrc", "libstrom_b200.so")

BEAGLE_FLAG_PRECISION_SINGLE = 1 << 0
BEAGLE_FLAG_PRECISION_DOUBLE = 1 << 1
BEAGLE_FLAG_SCALING_MANUAL = 1 << 6
BEAGLE_FLAG_PROCESSOR_CPU = 1 << 15
BEAGLE_FLAG_PROCESSOR_GPU = 1 << 16
BEAGLE_OP_NONE = -1

ERRORS = {  # reference src/likelihood.hpp:79-87
    0: "success",
    -1: "unspecified error",
    -2: "not enough memory could be allocated",
    -3: "unspecified exception",
    -4: "the instance index is out of range, or the instance has not been created",
    -5: "one of the indices specified exceeded the range of the array",
    -6: "no resource matches requirements",
    -7: "no implementation matches requirements",
    -8: "floating-point range exceeded",
}


class BeagleError(RuntimeError):
    def __init__(self, what: str, code: int):
        super().__init__("%s. BeagleLib error code was %d (%s)" % (what, code, ERRORS.get(code, "?")))
        self.code = code


class BeagleInstanceDetails(C.Structure):
    _fields_ = [("resourceNumber", C.c_int), ("resourceName", C.c_char_p), ("implName", C.c_char_p),
                ("implDescription", C.c_char_p), ("flags", C.c_long)]


class BeagleResource(C.Structure):
    _fields_ = [("name", C.c_char_p), ("description", C.c_char_p), ("supportFlags", C.c_long),
                ("requiredFlags", C.c_long)]


class BeagleResourceList(C.Structure):
    _fields_ = [("list", C.POINTER(BeagleResource)), ("length", C.c_int)]


_IP = C.POINTER(C.c_int)
_DP = C.POINTER(C.c_double)

SIGNATURES = {
    "beagleGetResourceList": (C.POINTER(BeagleResourceList), []),
    "beagleCreateInstance": (C.c_int, [C.c_int] * 9 + [_IP, C.c_int, C.c_long, C.c_long,
                                                      C.POINTER(BeagleInstanceDetails)]),
    "beagleFinalizeInstance": (C.c_int, [C.c_int]),
    "beagleSetTipStates": (C.c_int, [C.c_int, C.c_int, _IP]),
    "beagleSetPatternWeights": (C.c_int, [C.c_int, _DP]),
    "beagleSetStateFrequencies": (C.c_int, [C.c_int, C.c_int, _DP]),
    "beagleSetEigenDecomposition": (C.c_int, [C.c_int, C.c_int, _DP, _DP, _DP]),
    "beagleSetCategoryRates": (C.c_int, [C.c_int, _DP]),
    "beagleSetCategoryWeights": (C.c_int, [C.c_int, C.c_int, _DP]),
    "beagleUpdateTransitionMatrices": (C.c_int, [C.c_int, C.c_int, _IP, _IP, _IP, _DP, C.c_int]),
    "beagleResetScaleFactors": (C.c_int, [C.c_int, C.c_int]),
    "beagleUpdatePartials": (C.c_int, [C.c_int, _IP, C.c_int, C.c_int]),
    "beagleCalculateEdgeLogLikelihoods": (C.c_int, [C.c_int, _IP, _IP, _IP, _IP, _IP, _IP, _IP, _IP, C.c_int,
                                                    _DP, _DP, _DP]),
}


def _ip(a):
    return a.ctypes.data_as(_IP)


def _dp(a):
    return a.ctypes.data_as(_DP)


class BeagleAPI:
    """The 13 entry points of one shared library."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(
                "%s not found: build it first (python -c 'import __graft_entry__ as g; g.build()')" % path)
        self.path = path
        self.lib = C.CDLL(path, mode=C.RTLD_LOCAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.lib, name)          # AttributeError if a symbol is missing
            fn.restype, fn.argtypes = res, args
            setattr(self, name, fn)

    # -- convenience wrappers mirroring the reference's private Likelihood helpers ------------
    def resources(self):
        rl = self.beagleGetResourceList().contents
        return [(rl.list[i].name.decode(), (rl.list[i].description or b"").decode()) for i in range(rl.length)]

    def create(self, ntaxa: int, npatterns: int, ncateg: int, single: bool = False, prefer_gpu: bool = False):
        """Likelihood::initBeagleLib (src/likelihood.hpp:165-226) without the data upload."""
        req = (BEAGLE_FLAG_PRECISION_SINGLE if single else BEAGLE_FLAG_PRECISION_DOUBLE) | BEAGLE_FLAG_SCALING_MANUAL
        pref = BEAGLE_FLAG_PROCESSOR_GPU if prefer_gpu else BEAGLE_FLAG_PROCESSOR_CPU
        details = BeagleInstanceDetails()
        inst = self.beagleCreateInstance(ntaxa, ntaxa - 2, ntaxa, 4, npatterns, 1, 2 * ntaxa - 3, ncateg,
                                         ntaxa - 1, None, 0, pref, req, C.byref(details))
        if inst < 0:
            raise BeagleError("failed to create instance", inst)
        return inst

    def set_data(self, inst: int, data_matrix: np.ndarray, pattern_counts: np.ndarray):
        """setTipStates + setPatternWeights (src/likelihood.hpp:228-264)."""
        for t in range(data_matrix.shape[0]):
            row = np.ascontiguousarray(data_matrix[t], dtype=np.int32)
            code = self.beagleSetTipStates(inst, t, _ip(row))
            if code != 0:
                raise BeagleError("failed to set tip state for taxon %d" % (t + 1), code)
        w = np.ascontiguousarray(pattern_counts, dtype=np.float64)
        code = self.beagleSetPatternWeights(inst, _dp(w))
        if code != 0:
            raise BeagleError("failed to set pattern weights", code)

    def calc_log_likelihood(self, inst: int, model: dict, ops: np.ndarray, pmatrix_index: np.ndarray,
                            edge_lengths: np.ndarray, parent: int, child: int = 0) -> float:
        """Likelihood::calcLogLikelihood (src/likelihood.hpp:409-447), call for call."""
        freqs = np.ascontiguousarray(model["freqs"], dtype=np.float64)
        evec = np.ascontiguousarray(model["evec"], dtype=np.float64)
        ivec = np.ascontiguousarray(model["ivec"], dtype=np.float64)
        evals = np.ascontiguousarray(model["evals"], dtype=np.float64)
        rates = np.ascontiguousarray(model["rates"], dtype=np.float64)
        probs = np.ascontiguousarray(model["probs"], dtype=np.float64)
        ops = np.ascontiguousarray(ops, dtype=np.int32)
        pidx = np.ascontiguousarray(pmatrix_index, dtype=np.int32)
        elen = np.ascontiguousarray(edge_lengths, dtype=np.float64)

        def chk(code, what):
            if code != 0:
                raise BeagleError(what, code)

        # setModelRateMatrix (:277-294)
        chk(self.beagleSetStateFrequencies(inst, 0, _dp(freqs)), "failed to set state frequencies")
        chk(self.beagleSetEigenDecomposition(inst, 0, _dp(evec), _dp(ivec), _dp(evals)),
            "failed to set eigen decomposition")
        chk(self.beagleSetCategoryRates(inst, _dp(rates)), "failed to set among-site rate variation rates")
        chk(self.beagleSetCategoryWeights(inst, 0, _dp(probs)), "failed to set among-site rate variation weights")
        # setDiscreteGammaShape (:266-275) -- the reference uploads them a second time
        chk(self.beagleSetCategoryRates(inst, _dp(rates)), "failed to set category rates")
        chk(self.beagleSetCategoryWeights(inst, 0, _dp(probs)), "failed to set category probabilities")
        # updateTransitionMatrices (:357-370)
        chk(self.beagleUpdateTransitionMatrices(inst, 0, _ip(pidx), None, None, _dp(elen), int(pidx.size)),
            "failed to update transition matrices")
        # calculatePartials (:372-388)
        chk(self.beagleResetScaleFactors(inst, 0), "failed to reset scale factors in calculatePartials")
        chk(self.beagleUpdatePartials(inst, _ip(ops), int(ops.size // 7), 0), "failed to update partials")
        # edge log-likelihood at the root edge (:418-447)
        zero = np.zeros(1, dtype=np.int32)
        par = np.array([parent], dtype=np.int32)
        ch = np.array([child], dtype=np.int32)
        out = np.zeros(1, dtype=np.float64)
        chk(self.beagleCalculateEdgeLogLikelihoods(inst, _ip(par), _ip(ch), _ip(ch), None, None, _ip(zero), _ip(zero),
                                                   _ip(zero), 1, _dp(out), None, None),
            "failed to calculate edge logLikelihoods in CalcLogLikelihood")
        return float(out[0])

    def finalize(self, inst: int):
        code = self.beagleFinalizeInstance(inst)
        if code != 0:
            raise BeagleError("failed to finalize instance", code)


_engine = None


def engine() -> BeagleAPI:
    """The product library.  Raises if the CUDA extension has not been built."""
    global _engine
    if _engine is None:
        _engine = BeagleAPI(ENGINE_LIB)
    return _engine
