"""End-to-end effect of incremental likelihoods: the MCMC of config C3 (rbcl738, GTR+G4, all updaters, one chain)
with and without Likelihood::setIncremental.  usage: python tools/mcmc_incremental_run.py [niter]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from strom_b200.host import host
niter = int(sys.argv[1]) if len(sys.argv) > 1 else 400
fx = np.load(os.path.join(ROOT, "tests", "golden", "rbcl738.npz"))
cols = np.repeat(np.arange(fx["data"].shape[1]), fx["counts"].astype(int))
api = host()
h = api.data_from_codes(fx["data"].astype(np.int32)[:, cols], compress=True)
out = {"workload": "rbcl738 GTR+G4 MCMC, 1 chain, all updaters", "niter": niter}
for name, inc in (("full", False), ("incremental", True)):
    api.mcmc(h, str(fx["newick"]), niter=20, burnin=0, nchains=1, samplefreq=20, incremental=inc)
    tr, sec = api.mcmc(h, str(fx["newick"]), niter=niter, burnin=0, nchains=1, samplefreq=niter, seed=5, incremental=inc, time_iterations_only=True)
    out[name] = {"seconds": sec, "iterations_per_sec": niter / sec, "final_lnL": float(tr[-1, 1])}
print(json.dumps(out))
