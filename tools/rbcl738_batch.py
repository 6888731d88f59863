"""K evaluations of the rbcl738 fixture (one tree / model per chain) as ONE set of launches: device time per batch,
aggregate evaluations/s, and each chain's value against its own single evaluation."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bench import gtr_gamma_model, PAUP_PI, PAUP_R, PAUP_ALPHA
from strom_b200 import hostmodel
from strom_b200.engine import EvalArgs, ShardedLikelihood, StromEvalArgs, api

K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_eval = int(sys.argv[2]) if len(sys.argv) > 2 else 200
PINV = float(os.environ.get("PINVAR", "0"))
fx = np.load(os.path.join(ROOT, "tests", "golden", "rbcl738.npz"))
E = api()
rng = np.random.default_rng(5)
tips = np.minimum(fx["data"], 4).astype(np.uint8)
Ls, eas = [], []
for c in range(K):
    model = gtr_gamma_model(PAUP_PI, PAUP_R, PAUP_ALPHA * (1.0 + 0.1 * c), 4)
    if PINV > 0:
        model["rates"], model["probs"] = hostmodel.invariant_gamma_rates(PAUP_ALPHA * (1.0 + 0.1 * c), 4, PINV)
    elen = fx["edge_lengths"] * np.exp(0.2 * rng.standard_normal(fx["edge_lengths"].shape)) if c else fx["edge_lengths"]
    L = ShardedLikelihood(tips, fx["counts"], len(model["rates"]), device=0, transient_partials=bool(os.environ.get("TRANSIENT")), single=bool(os.environ.get("FP32")))
    Ls.append(L)
    eas.append(EvalArgs(model, fx["ops"], fx["pmatrix_index"], elen, int(fx["parent"]), 0))
single = [L.calc_log_likelihood(ea) for L, ea in zip(Ls, eas)]
insts = (C.c_int * K)(*[L.inst for L in Ls])
args = (StromEvalArgs * K)(*[ea.c for ea in eas])
out = C.c_double()
def batch_eval():
    E.check(E.sb200EvalBatchAsync(K, insts, args), "batch")
    vals = []
    for L in Ls:
        E.check(E.sb200Finish(L.inst, C.byref(out)), "finish")
        vals.append(out.value)
    return vals
vals = batch_eval()
print("max rel diff batch vs single:", max(abs(a - b) / abs(b) for a, b in zip(vals, single)))
stream = Ls[0].stream
for _ in range(10):
    E.check(E.sb200ReplayBatchAsync(K, insts), "replay")
stream.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(stream):
    e0.record()
    for _ in range(n_eval):
        E.check(E.sb200ReplayBatchAsync(K, insts), "replay")
    e1.record()
stream.synchronize()
us = e0.elapsed_time(e1) * 1000 / n_eval
print("K=%d device us/batch %.1f  aggregate evals/s %.0f" % (K, us, K / us * 1e6))
t0 = time.perf_counter()
for _ in range(n_eval):
    batch_eval()
us = (time.perf_counter() - t0) * 1e6 / n_eval
print("K=%d fused batch call us/batch %.1f  aggregate evals/s %.0f" % (K, us, K / us * 1e6))

# host side of the fused batch call: enqueue (sb200EvalBatchAsync returns) vs completion (sb200Finish of every chain)
te = tf = 0.0
for _ in range(n_eval):
    t0 = time.perf_counter()
    E.check(E.sb200EvalBatchAsync(K, insts, args), "batch")
    t1 = time.perf_counter()
    for L in Ls:
        E.check(E.sb200Finish(L.inst, C.byref(out)), "finish")
    t2 = time.perf_counter()
    te += t1 - t0; tf += t2 - t1
print("K=%d enqueue us %.1f  finish-wait us %.1f" % (K, te * 1e6 / n_eval, tf * 1e6 / n_eval))
