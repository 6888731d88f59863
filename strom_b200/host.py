"""ctypes face of the C++ host library (strom_b200/host/*.hpp): the mirror of the reference's
Data / TreeManip / Model / Likelihood classes, linked to the CUDA engine (libstrom_host.so)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB = os.path.join(_HERE, "host", "libstrom_host.so")

_IP = C.POINTER(C.c_int)
_DP = C.POINTER(C.c_double)


class HostError(RuntimeError):
    pass


class HostAPI:
    def __init__(self, path: str = HOST_LIB):
        if not os.path.exists(path):
            raise FileNotFoundError("%s not found: run python -m strom_b200.build" % path)
        self.lib = L = C.CDLL(path, mode=C.RTLD_LOCAL)
        L.strom_data_from_file.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.strom_data_from_codes.argtypes = [C.c_int, C.c_int, _IP, C.c_int, C.c_char_p, C.c_int]
        L.strom_data_dims.argtypes = [C.c_int, _IP, _IP, _IP]
        L.strom_data_copy.argtypes = [C.c_int, _IP, _DP]
        L.strom_data_free.argtypes = [C.c_int]
        L.strom_tree_ops.argtypes = [C.c_char_p, _IP, _IP, _DP, _IP, _IP, C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        L.strom_model_numbers.argtypes = [_DP, _DP, C.c_double, C.c_int, C.c_double, _DP, _DP, _DP, _DP, _DP, _IP, C.c_char_p, C.c_int]
        L.strom_loglik.argtypes = [C.c_int, C.c_char_p, _DP, _DP, C.c_double, C.c_int, C.c_double, C.c_int, _IP, C.c_int, C.c_int,
                                   _DP, _DP, C.c_char_p, C.c_int]
        L.strom_mcmc.argtypes = [C.c_int, C.c_char_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_double, C.c_uint, C.c_double,
                                 _DP, _DP, C.c_int, _IP, _DP, C.c_int, _DP, C.c_char_p, C.c_int]
        L.strom_mcmc_ex.argtypes = [C.c_int, C.c_char_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_double, C.c_uint,
                                    C.c_double, _DP, _DP, C.c_int, _IP, C.c_uint, _DP, C.c_int, _DP, C.c_char_p, C.c_int]
        L.strom_mcmc_ex2.argtypes = [C.c_int, C.c_char_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_double, C.c_uint,
                                     C.c_double, C.c_double, _DP, _DP, C.c_int, _IP, C.c_uint, _DP, C.c_int, _DP, C.c_char_p, C.c_int]
        L.strom_lot_draws.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _DP]
        L.strom_incremental_rebuild_check.argtypes = [C.c_int, C.c_char_p, _DP, _DP, C.c_double, C.c_int, C.c_int, C.c_uint, _DP, C.c_char_p,
                                                      C.c_int]
        L.strom_incremental_walk.argtypes = [C.c_int, C.c_char_p, _DP, _DP, C.c_double, C.c_int, C.c_double, C.c_int, _IP, C.c_uint,
                                             C.c_int, _DP, _DP, _IP, _DP, _DP, C.c_char_p, C.c_int]
        self._err = C.create_string_buffer(1024)

    def _check(self, rc):
        if rc < 0:
            raise HostError(self._err.value.decode(errors="replace"))
        return rc

    # ---- Data ----------------------------------------------------------------------------
    def data_from_file(self, path: str) -> int:
        return self._check(self.lib.strom_data_from_file(path.encode(), self._err, 1024))

    def data_from_codes(self, codes: np.ndarray, compress=True) -> int:
        """``compress``: False = every site its own pattern; True = Data::compressPatterns (on the GPU from 32768 sites
        up in engine builds); 2 = always on the GPU; 3 = always on the host."""
        codes = np.ascontiguousarray(codes, dtype=np.int32)
        return self._check(self.lib.strom_data_from_codes(codes.shape[0], codes.shape[1], codes.ctypes.data_as(_IP), int(compress),
                                                          self._err, 1024))

    def data_arrays(self, h: int):
        n, p, s = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.strom_data_dims(h, C.byref(n), C.byref(p), C.byref(s)))
        m = np.zeros((n.value, p.value), dtype=np.int32)
        w = np.zeros(p.value)
        self._check(self.lib.strom_data_copy(h, m.ctypes.data_as(_IP), w.ctypes.data_as(_DP)))
        return m, w, s.value

    def data_free(self, h: int):
        self.lib.strom_data_free(h)

    # ---- Tree / operations -----------------------------------------------------------------
    def tree_ops(self, newick: str, max_leaves: int = 4096):
        n = max_leaves
        ops = np.zeros(7 * (n - 2), dtype=np.int32)
        pidx = np.zeros(2 * n - 3, dtype=np.int32)
        elen = np.zeros(2 * n - 3)
        nl, par = C.c_int(), C.c_int()
        out = C.create_string_buffer(64 * n)
        self._check(self.lib.strom_tree_ops(newick.encode(), ops.ctypes.data_as(_IP), pidx.ctypes.data_as(_IP), elen.ctypes.data_as(_DP),
                                            C.byref(nl), C.byref(par), out, 64 * n, self._err, 1024))
        n = nl.value
        return ops[:7 * (n - 2)].reshape(-1, 7).copy(), pidx[:2 * n - 3].copy(), elen[:2 * n - 3].copy(), par.value, out.value.decode()

    def local_move_check(self, newick: str, seed: int = 1, lam: float = 0.5, nsteps: int = 200):
        """Test hook for the LOCAL move: (take-backs that did not restore the tree, topology changes, worst consistency
        error of path scaling / Hastings ratio / tree length, final newick)."""
        self.lib.strom_local_move_check.argtypes = [C.c_char_p, C.c_uint, C.c_double, C.c_int, _IP, _DP, C.c_char_p, C.c_int,
                                                    C.c_char_p, C.c_int]
        ch, worst = C.c_int(), C.c_double()
        out = C.create_string_buffer(1 << 20)
        bad = self._check(self.lib.strom_local_move_check(newick.encode(), seed, lam, nsteps, C.byref(ch), C.byref(worst), out, 1 << 20,
                                                          self._err, 1024))
        return bad, ch.value, worst.value, out.value.decode()

    # ---- Model -----------------------------------------------------------------------------
    def model_numbers(self, freqs, rmat, shape: float, ncat: int, pinvar: float = 0.0) -> dict:
        f = np.ascontiguousarray(freqs, dtype=np.float64)
        r = np.ascontiguousarray(rmat, dtype=np.float64)
        evec, ivec, evals = np.zeros((4, 4)), np.zeros((4, 4)), np.zeros(4)
        rates, probs = np.zeros(ncat + 1), np.zeros(ncat + 1)
        nt = C.c_int()
        self._check(self.lib.strom_model_numbers(f.ctypes.data_as(_DP), r.ctypes.data_as(_DP), shape, ncat, pinvar, evec.ctypes.data_as(_DP),
                                                 ivec.ctypes.data_as(_DP), evals.ctypes.data_as(_DP), rates.ctypes.data_as(_DP),
                                                 probs.ctypes.data_as(_DP), C.byref(nt), self._err, 1024))
        return dict(freqs=f, evec=evec, ivec=ivec, evals=evals, rates=rates[:nt.value].copy(), probs=probs[:nt.value].copy())

    # ---- Likelihood --------------------------------------------------------------------------
    def loglik(self, data_handle: int, newick: str, freqs, rmat, shape: float, ncat: int, pinvar: float = 0.0, devices=(0,),
               single: bool = False, repeats: int = 1, transient_partials: bool = True):
        """Likelihood::calcLogLikelihood through the C++ facade; returns (logL, seconds per evaluation)."""
        f = np.ascontiguousarray(freqs, dtype=np.float64)
        r = np.ascontiguousarray(rmat, dtype=np.float64)
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        out, sec = C.c_double(), C.c_double()
        self._check(self.lib.strom_loglik(data_handle, newick.encode(), f.ctypes.data_as(_DP), r.ctypes.data_as(_DP), shape, ncat, pinvar,
                                          int(dev.size), dev.ctypes.data_as(_IP), int(bool(single)) | (0 if transient_partials else 4), repeats,
                                          C.byref(out), C.byref(sec),
                                          self._err, 1024))
        return out.value, sec.value


    # ---- MCMC (callers of the hot path) ------------------------------------------------------
    def mcmc(self, data_handle: int, newick: str, seed: int = 13579, niter: int = 100, burnin: int = 100, samplefreq: int = 1,
             nchains: int = 1, heatfactor: float = 0.5, ncateg: int = 4, gammashape: float = 0.5, freqs=(0.25,) * 4,
             rmat=(1.0 / 6,) * 6, devices=(0,), per_chain_streams: bool = False, parallel_chains: bool = False,
             incremental: bool = False, lockstep_chains: bool = False, pinvar: float = 0.0, batch_evaluations: bool = True, transient_partials: bool = True, boost_lot: bool = True, time_iterations_only: bool = False):
        """Returns (trace [rows][15] = iteration, lnL, lnPrior, TL, 6 r, 4 pi, alpha; seconds).

        ``parallel_chains``: the heated chains are stepped side by side on host threads (one chain per GPU of
        ``devices``), each with its own random stream; ``per_chain_streams`` alone keeps them in turn;
        ``lockstep_chains``: one host thread, every update begun on all chains before it is finished on any, so the
        chains' likelihoods overlap on the GPU(s) they share -- same trace as ``per_chain_streams``; with
        ``batch_evaluations`` (engine builds) the evaluations the chains of one GPU begin together are ONE set of kernel
        launches.  ``pinvar`` > 0: GTR+I+G with that fixed proportion of invariable sites."""
        f = np.ascontiguousarray(freqs, dtype=np.float64)
        r = np.ascontiguousarray(rmat, dtype=np.float64)
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        max_rows = niter // max(samplefreq, 1) + 2
        rows = np.zeros((max_rows, 15))
        sec = C.c_double()
        flags = (1 if per_chain_streams else 0) | (2 if parallel_chains else 0) | (4 if incremental else 0) | \
            (8 if lockstep_chains else 0) | (0 if batch_evaluations else 16) | (0 if transient_partials else 32) | (0 if boost_lot else 64) | (128 if time_iterations_only else 0)
        n = self._check(self.lib.strom_mcmc_ex2(data_handle, newick.encode(), seed, niter, burnin, samplefreq, nchains, heatfactor,
                                                ncateg, gammashape, float(pinvar), f.ctypes.data_as(_DP), r.ctypes.data_as(_DP),
                                                int(dev.size), dev.ctypes.data_as(_IP), flags, rows.ctypes.data_as(_DP), max_rows,
                                                C.byref(sec), self._err, 1024))
        return rows[:n].copy(), sec.value


    def lot_draws(self, seed: int, kind: str, n: int, a: float = 0.0, b: float = 0.0, boost: bool = True) -> np.ndarray:
        """n variates of the host's random-number class (test hook): kind = uniform | randint | gamma | loguniform."""
        out = np.zeros(n)
        self.lib.strom_lot_draws(seed, int(boost), {"uniform": 0, "randint": 1, "gamma": 2, "loguniform": 3}[kind], a, b, n, out.ctypes.data_as(_DP))
        return out

    def rebuild_check(self, data_handle: int, newick: str, freqs, rmat, shape: float, ncat: int, rounds: int = 10, seed: int = 1) -> float:
        """Incremental against full evaluation over a sequence of trees rebuilt from newick strings (test hook)."""
        return self._rebuild_check(data_handle, newick, freqs, rmat, shape, ncat, rounds, seed)

    def incremental_walk(self, data_handle: int, newick: str, freqs, rmat, shape: float, ncat: int, pinvar: float = 0.0,
                         devices=(0,), seed: int = 1, nsteps: int = 100):
        """Random tree moves evaluated incrementally and in full: (incremental logL[], full logL[], ops listed[],
        seconds per incremental evaluation, seconds per full evaluation)."""
        f = np.ascontiguousarray(freqs, dtype=np.float64)
        r = np.ascontiguousarray(rmat, dtype=np.float64)
        a, b = np.zeros(nsteps), np.zeros(nsteps)
        k = np.zeros(nsteps, dtype=np.int32)
        si, sf = C.c_double(), C.c_double()
        dev = np.ascontiguousarray(devices, dtype=np.int32)
        n = self._check(self.lib.strom_incremental_walk(data_handle, newick.encode(), f.ctypes.data_as(_DP), r.ctypes.data_as(_DP),
                                                        shape, ncat, pinvar, int(dev.size), dev.ctypes.data_as(_IP), seed, nsteps,
                                                        a.ctypes.data_as(_DP),
                                                        b.ctypes.data_as(_DP), k.ctypes.data_as(_IP), C.byref(si), C.byref(sf),
                                                        self._err, 1024))
        return a[:n], b[:n], k[:n], si.value, sf.value


    def _rebuild_check(self, data_handle, newick, freqs, rmat, shape, ncat, rounds, seed):
        f = np.ascontiguousarray(freqs, dtype=np.float64)
        r = np.ascontiguousarray(rmat, dtype=np.float64)
        worst = C.c_double()
        self._check(self.lib.strom_incremental_rebuild_check(data_handle, newick.encode(), f.ctypes.data_as(_DP), r.ctypes.data_as(_DP),
                                                             shape, ncat, rounds, seed, C.byref(worst), self._err, 1024))
        return worst.value


_host = None


def host() -> HostAPI:
    global _host
    if _host is None:
        _host = HostAPI()
    return _host
