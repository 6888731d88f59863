"""Marginal device cost of the kernels around the pruning launches on rbcl738: evaluations issued through the
13-call boundary with (a) everything, (b) no root-edge call, (c) no matrix update either."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np
import torch
from bench import gtr_gamma_model, PAUP_PI, PAUP_R, PAUP_ALPHA
from strom_b200.engine import EvalArgs, ShardedLikelihood
from strom_b200.beagle import _ip, _dp

fx = np.load(os.path.join(ROOT, "tests", "golden", "rbcl738.npz"))
model = gtr_gamma_model(PAUP_PI, PAUP_R, PAUP_ALPHA, 4)
L = ShardedLikelihood(np.minimum(fx["data"], 4).astype(np.uint8), fx["counts"], 4, device=0)
ea = EvalArgs(model, fx["ops"], fx["pmatrix_index"], fx["edge_lengths"], int(fx["parent"]), 0)
print("logL", L.calc_log_likelihood(ea))
B, inst = L.B, L.inst
ops = np.ascontiguousarray(fx["ops"], dtype=np.int32); pidx = np.ascontiguousarray(fx["pmatrix_index"], dtype=np.int32)
elen = np.ascontiguousarray(fx["edge_lengths"], dtype=np.float64)
zero = np.zeros(1, dtype=np.int32); par = np.array([int(fx["parent"])], dtype=np.int32); ch = np.zeros(1, dtype=np.int32)
out = np.zeros(1)

def run(n, mats, root):
    for _ in range(n):
        if mats:
            B.beagleUpdateTransitionMatrices(inst, 0, _ip(pidx), None, None, _dp(elen), int(pidx.size))
        B.beagleResetScaleFactors(inst, 0)
        B.beagleUpdatePartials(inst, _ip(ops), int(ops.size // 7), 0)
    if root:   # the root call synchronises: issue it once at the end only
        B.beagleCalculateEdgeLogLikelihoods(inst, _ip(par), _ip(ch), _ip(ch), None, None, _ip(zero), _ip(zero), _ip(zero), 1, _dp(out), None, None)

N = 400
for name, mats in (("matrices+tables+prune", True), ("tables+prune", False)):
    run(50, mats, False)
    L.stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(L.stream):
        e0.record()
        run(N, mats, False)
        e1.record()
    L.stream.synchronize()
    print("%-24s %.2f us per evaluation (device)" % (name, e0.elapsed_time(e1) * 1000 / N))
for _ in range(20): L.replay_async()
L.stream.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(L.stream):
    e0.record()
    for _ in range(N): L.replay_async()
    e1.record()
L.stream.synchronize()
print("%-24s %.2f us per evaluation (device)" % ("replay (all + root)", e0.elapsed_time(e1) * 1000 / N))
L.close()
