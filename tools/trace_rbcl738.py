import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from bench import gtr_gamma_model, PAUP_PI, PAUP_R, PAUP_ALPHA
from strom_b200.engine import EvalArgs, ShardedLikelihood
fx = np.load(os.path.join(ROOT, "tests", "golden", "rbcl738.npz"))
model = gtr_gamma_model(PAUP_PI, PAUP_R, PAUP_ALPHA, 4)
PINV = float(os.environ.get("PINVAR", "0"))          # > 0: GTR+I+G4 (a fifth category of rate 0)
if PINV > 0:
    from strom_b200 import hostmodel
    model["rates"], model["probs"] = hostmodel.invariant_gamma_rates(PAUP_ALPHA, 4, PINV)
NC = len(model["rates"])
L = ShardedLikelihood(np.minimum(fx["data"], 4).astype(np.uint8), fx["counts"], NC, device=0)
ea = EvalArgs(model, fx["ops"], fx["pmatrix_index"], fx["edge_lengths"], int(fx["parent"]), 0)
L.E.lib.sb200DebugTrace.argtypes = [C.c_int, C.c_void_p]
L.E.lib.sb200DebugTrace(L.inst, None)
for _ in range(5): L.calc_log_likelihood(ea)
out = np.zeros(480, dtype=np.int64)
L.E.lib.sb200DebugTrace(L.inst, out.ctypes.data_as(C.c_void_p))
t = out.reshape(60, 8)
print("step: wait  barrier  issue  factorA  factorB+join  rescale  tail | total   (cycles, last launch = top level, block 0)")
for j in range(60):
    if t[j, 0] == 0: break
    r = t[j]
    nxt = t[j + 1, 0] if j + 1 < 60 and t[j + 1, 0] else r[5]
    print("%3d: %6d %6d %6d %6d %6d %6d %6d | %6d" % (j, r[1] - r[0], r[2] - r[1], r[6] - r[2], r[7] - r[6], r[3] - r[7], r[4] - r[3], r[5] - r[4], nxt - r[0]))
