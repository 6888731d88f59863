"""Per-phase clock samples of the steps of ONE pruning launch of an rbcl738 evaluation (block 0, thread SB200_TRACE_TID);
needs a library built with -DSB200_TRACE (tools/build_variant.sh trace -DSB200_TRACE); SB200_TRACE_LEVEL picks the launch."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from bench import gtr_gamma_model, PAUP_PI, PAUP_R, PAUP_ALPHA
from strom_b200.engine import EvalArgs, ShardedLikelihood
fx = np.load(os.path.join(ROOT, "tests", "golden", "rbcl738.npz"))
model = gtr_gamma_model(PAUP_PI, PAUP_R, PAUP_ALPHA, 4)
L = ShardedLikelihood(np.minimum(fx["data"], 4).astype(np.uint8), fx["counts"], 4, device=0)
ea = EvalArgs(model, fx["ops"], fx["pmatrix_index"], fx["edge_lengths"], int(fx["parent"]), 0)
L.E.lib.sb200DebugTrace.argtypes = [C.c_int, C.c_void_p]
L.E.lib.sb200DebugTrace(L.inst, None)
for _ in range(5): L.calc_log_likelihood(ea)
out = np.zeros(480, dtype=np.int64)
L.E.lib.sb200DebugTrace(L.inst, out.ctypes.data_as(C.c_void_p))
t = out.reshape(60, 8)
print("level", os.environ.get("SB200_TRACE_LEVEL"), "ticks 0..7 as deltas (cycles); last column = whole step")
for j in range(60):
    if t[j, 0] == 0: break
    r = t[j]
    nxt = t[j + 1, 0] if j + 1 < 60 and t[j + 1, 0] else r.max()
    d = [int(r[q + 1] - r[q]) if r[q + 1] and r[q] else -1 for q in range(7)]
    print("%3d: " % j + " ".join("%6d" % x for x in d) + " | %6d" % (nxt - r[0]))
