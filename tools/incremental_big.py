"""Incremental evaluation where a full evaluation is bandwidth-bound: synthetic alignment, random tree, LOCAL-style
moves.  usage: python tools/incremental_big.py [ntaxa] [nsites] [nsteps]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from strom_b200 import synth
from strom_b200.host import host
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import PAUP_ALPHA, PAUP_PI, PAUP_R
ntaxa = int(sys.argv[1]) if len(sys.argv) > 1 else 200
nsites = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
rng = np.random.default_rng(7)
nwk = synth.random_newick(ntaxa, rng, 0.05)
data = rng.integers(0, 4, size=(ntaxa, nsites)).astype(np.int32)       # (random columns: every site its own pattern)
api = host()
h = api.data_from_codes(data, compress=False)
inc, full, nops, si, sf = api.incremental_walk(h, nwk, PAUP_PI, PAUP_R, PAUP_ALPHA, 4, 0.0, nsteps=nsteps, seed=2)
moves = nops[nops < ntaxa - 2]
print(json.dumps({"workload": "synthetic %d taxa x %d patterns GTR+G4, LOCAL-style tree moves" % (ntaxa, nsites), "steps": int(len(inc)),
                  "max_rel_diff_vs_full": float(np.max(np.abs(inc - full) / np.abs(full))),
                  "operations_listed_median": float(np.median(moves)), "operations_full": ntaxa - 2,
                  "us_per_eval_incremental_mix": si * 1e6, "us_per_eval_full": sf * 1e6, "speedup_mix": sf / si}))
