"""SURVEY.md 8(f) rank 1 measured: rbcl738 GTR+I+G4, random LOCAL-style tree moves evaluated by the incremental
facade and by the facade that recomputes everything (the reference's behaviour).  usage: python tools/incremental_run.py [nsteps]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from strom_b200.host import host
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import PAUP_ALPHA, PAUP_PI, PAUP_R

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
fx = np.load(os.path.join(ROOT, "tests", "golden", "rbcl738.npz"))
cols = np.repeat(np.arange(fx["data"].shape[1]), fx["counts"].astype(int))
api = host()
h = api.data_from_codes(fx["data"].astype(np.int32)[:, cols], compress=True)
api.incremental_walk(h, str(fx["newick"]), PAUP_PI, PAUP_R, PAUP_ALPHA, 4, 0.2, nsteps=50, seed=1)     # warm-up
inc, full, nops, si, sf = api.incremental_walk(h, str(fx["newick"]), PAUP_PI, PAUP_R, PAUP_ALPHA, 4, 0.2, nsteps=nsteps, seed=2)
moves = nops[nops < 736]
print(json.dumps({"workload": "rbcl738 GTR+I+G4, LOCAL-style tree moves (1 in 16 steps changes the model => full list)",
                  "steps": int(len(inc)), "max_rel_diff_vs_full": float(np.max(np.abs(inc - full) / np.abs(full))),
                  "operations_listed_median": float(np.median(moves)), "operations_listed_max": int(moves.max()),
                  "operations_full": 736, "us_per_eval_incremental_mix": si * 1e6, "us_per_eval_full": sf * 1e6,
                  "speedup_mix": sf / si}))
