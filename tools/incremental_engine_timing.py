"""Where the time of an incremental evaluation goes (engine level, rbcl738 GTR+G4): the same partial list repeated
(plan cached) against two alternating partial lists (plan rebuilt and uploaded every call)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import ctypes as C
from bench import gtr_gamma_model, PAUP_PI, PAUP_R, PAUP_ALPHA
from strom_b200.engine import EvalArgs, ShardedLikelihood
fx = np.load(os.path.join(ROOT, "tests", "golden", "rbcl738.npz"))
model = gtr_gamma_model(PAUP_PI, PAUP_R, PAUP_ALPHA, 4)
L = ShardedLikelihood(np.minimum(fx["data"], 4).astype(np.uint8), fx["counts"], 4, device=0)
L.E.check(L.E.sb200SetSubtreeScalers(L.inst, 1), "scalers")
ops = fx["ops"].reshape(-1, 7); pidx = fx["pmatrix_index"]; el = fx["edge_lengths"].astype(np.float64)
parent = int(fx["parent"])
full = EvalArgs(model, ops, pidx, el, parent, 0)
print("full logL", L.calc_log_likelihood(full))
parent_of = {}
for o in ops:
    parent_of[int(o[3])] = int(o[0]); parent_of[int(o[5])] = int(o[0])
def path_args(tip):
    dirty = set(); v = parent_of[tip]
    while v is not None:
        dirty.add(v); v = parent_of.get(v)
    sub = np.array([o for o in ops if int(o[0]) in dirty], dtype=np.int32)
    e = int(np.where(pidx == tip)[0][0])
    return EvalArgs(model, sub, pidx[e:e + 1], el[e:e + 1], parent, 0), len(sub)
depth = {t: path_args(t)[1] for t in range(1, 738, 37)}
print("path lengths", sorted(depth.values()))
ta, tb = sorted(depth, key=lambda t: depth[t])[-1], sorted(depth, key=lambda t: depth[t])[-2]
A, na = path_args(ta); B, nb = path_args(tb)
def timeit(seq, n=400):
    for a in seq: L.calc_log_likelihood(a)
    t0 = time.perf_counter()
    for i in range(n): L.calc_log_likelihood(seq[i % len(seq)])
    return (time.perf_counter() - t0) / n * 1e6
print("full list, cached plan       : %.1f us" % timeit([full]))
print("path of %d ops, cached plan   : %.1f us" % (na, timeit([A])))
print("paths of %d/%d ops alternating : %.1f us" % (na, nb, timeit([A, B])))
L.E.sb200SetProfiling(L.inst, 1)
for i in range(50): L.calc_log_likelihood(A)
ms = C.c_double(); calls = C.c_long()
L.E.sb200PartialsTime(L.inst, C.byref(ms), C.byref(calls))
print("device time of the pruning launches for the path: %.1f us" % (ms.value / max(calls.value, 1) * 1e3))
