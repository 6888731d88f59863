"""Shared test helpers: golden fixtures, named models, whole-call drivers."""
import ctypes as C
import os

import numpy as np

from oracle import strom_host_ref as H

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

JC = ([0.25, 0.25, 0.25, 0.25], [1.0] * 6)
PAUP_PI = [0.309769, 0.163380, 0.121023]
PAUP_PI.append(1.0 - sum(PAUP_PI))
PAUP_R = [1.06923, 4.34568, 0.45898, 2.00465, 3.85915, 1.0]
PAUP_ALPHA = 0.517281

MODEL_NAMES = ["jc", "jc_g4", "gtr_g4", "gtr_i_g4", "gtr_g2", "gtr_g8", "gtr_g3"]


def make_model(freqs, rmat, rates, probs):
    evec, ivec, evals = H.gtr_eigen(freqs, rmat)
    return dict(freqs=np.asarray(freqs, dtype=np.float64), evec=evec, ivec=ivec, evals=evals,
                rates=np.asarray(rates, dtype=np.float64), probs=np.asarray(probs, dtype=np.float64))


def named_model(name):
    if name == "jc":
        return make_model(*JC, *H.gamma_rates(0.5, 1))
    if name == "jc_g4":
        return make_model(*JC, *H.gamma_rates(0.5, 4))
    if name == "gtr_i_g4":
        return make_model(PAUP_PI, PAUP_R, *H.invariant_gamma_rates(PAUP_ALPHA, 4, 0.2))
    if name.startswith("gtr_g"):
        return make_model(PAUP_PI, PAUP_R, *H.gamma_rates(PAUP_ALPHA, int(name[5:])))
    raise KeyError(name)


def load_golden(key):
    z = np.load(os.path.join(GOLDEN, key + ".npz"))
    return {k: z[k] for k in z.files}


def full_eval(api, fx, model, single=False, keep=False):
    """create + upload + Likelihood::calcLogLikelihood sequence; returns logL (and the instance if keep)."""
    data = fx["data"].astype(np.int32)
    n, P = data.shape
    inst = api.create(n, P, len(model["rates"]), single=single)
    api.set_data(inst, data, fx["counts"])
    ll = api.calc_log_likelihood(inst, model, fx["ops"], fx["pmatrix_index"], fx["edge_lengths"], int(fx["parent"]), 0)
    if keep:
        return ll, inst
    api.finalize(inst)
    return ll


def rel_err(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


# ---- random problems (same generator the bench uses) -----------------------------------------
def random_problem(ntaxa, npatterns, seed, missing_frac=0.02, mean_edge=0.05):
    from strom_b200 import synth
    rng = np.random.default_rng(seed)
    nwk = synth.random_newick(ntaxa, rng, mean_edge)
    tree = H.build_tree(nwk)
    ops, pidx, elen = H.define_operations(tree)
    data = rng.integers(0, 4, size=(ntaxa, npatterns)).astype(np.uint8)
    if missing_frac > 0:
        data[rng.random(data.shape) < missing_frac] = 4
    counts = rng.integers(1, 5, size=npatterns).astype(np.float64)
    return dict(data=data, counts=counts, ops=ops, pmatrix_index=pidx, edge_lengths=elen,
                parent=np.int32(tree.preorder[0].number))


class Extra:
    """ctypes access to the engine-level / oracle inspection entry points."""

    def __init__(self, api):
        self.lib = api.lib
        dp = C.POINTER(C.c_double)
        self.is_engine = hasattr(self.lib, "sb200GetPartials")
        pre = "sb200" if self.is_engine else "oracle"
        for name in ("GetPartials", "GetScaleFactors", "GetTransitionMatrix"):
            fn = getattr(self.lib, pre + name)
            fn.restype, fn.argtypes = C.c_int, [C.c_int, C.c_int, dp]
            setattr(self, "_" + name, fn)

    def partials(self, inst, idx, P, Cc):
        out = np.zeros((P, Cc, 4))
        assert self._GetPartials(inst, idx, out.ctypes.data_as(C.POINTER(C.c_double))) == 0
        return out

    def scale(self, inst, idx, P):
        out = np.zeros(P)
        assert self._GetScaleFactors(inst, idx, out.ctypes.data_as(C.POINTER(C.c_double))) == 0
        return out

    def matrix(self, inst, idx, Cc):
        out = np.zeros((Cc, 4, 4))
        assert self._GetTransitionMatrix(inst, idx, out.ctypes.data_as(C.POINTER(C.c_double))) == 0
        return out
