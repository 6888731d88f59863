"""Runs N evaluations of the rbcl738 fixture on the engine (for ncu launch lists / timing)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bench import gtr_gamma_model, PAUP_PI, PAUP_R, PAUP_ALPHA
from strom_b200.engine import EvalArgs, ShardedLikelihood

n_eval = int(sys.argv[1]) if len(sys.argv) > 1 else 20
fx = np.load(os.path.join(ROOT, "tests", "golden", "rbcl738.npz"))
model = gtr_gamma_model(PAUP_PI, PAUP_R, PAUP_ALPHA, 4)
PINV = float(os.environ.get("PINVAR", "0"))          # > 0: GTR+I+G4 (a fifth category of rate 0)
if PINV > 0:
    from strom_b200 import hostmodel
    model["rates"], model["probs"] = hostmodel.invariant_gamma_rates(PAUP_ALPHA, 4, PINV)
NC = len(model["rates"])
L = ShardedLikelihood(np.minimum(fx["data"], 4).astype(np.uint8), fx["counts"], NC, device=0, transient_partials=bool(os.environ.get("TRANSIENT")))
ea = EvalArgs(model, fx["ops"], fx["pmatrix_index"], fx["edge_lengths"], int(fx["parent"]), 0)
print("logL", L.calc_log_likelihood(ea))
for _ in range(10):
    L.replay_async()
L.stream.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(L.stream):
    e0.record()
    for _ in range(n_eval):
        L.replay_async()
    e1.record()
L.stream.synchronize()
print("device us/eval", e0.elapsed_time(e1) * 1000 / n_eval)
t0 = time.perf_counter()
for _ in range(n_eval):
    L.calc_log_likelihood(ea)
print("fused call us/eval", (time.perf_counter() - t0) * 1e6 / n_eval)
L.close()

# host-side split of the fused call: enqueue (sb200EvalAsync returns) vs completion (sb200Finish)
import ctypes as C
from strom_b200.engine import api
E = api()
L2 = ShardedLikelihood(np.minimum(fx["data"], 4).astype(np.uint8), fx["counts"], NC, device=0, transient_partials=bool(os.environ.get("TRANSIENT")))
E.lib.sb200Finish.argtypes = [C.c_int, C.POINTER(C.c_double)]
out = C.c_double()
for _ in range(20):
    E.sb200EvalAsync(L2.inst, C.byref(ea.c), None); E.lib.sb200Finish(L2.inst, C.byref(out))
te = tf = 0.0
for _ in range(n_eval):
    t0 = time.perf_counter()
    E.sb200EvalAsync(L2.inst, C.byref(ea.c), None)
    t1 = time.perf_counter()
    E.lib.sb200Finish(L2.inst, C.byref(out))
    t2 = time.perf_counter()
    te += t1 - t0; tf += t2 - t1
print("enqueue us %.1f  finish-wait us %.1f  total %.1f  logL %.6f" % (te * 1e6 / n_eval, tf * 1e6 / n_eval, (te + tf) * 1e6 / n_eval, out.value))
L2.close()
