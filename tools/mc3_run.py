"""Config C5: Metropolis-coupled MCMC on rbcl738, one heated chain per GPU.  Prints chain-iterations per second
with the chains stepped in turn (reference order), side by side (host thread per chain) and in lockstep (one host
thread, every update begun on all chains before it is finished on any).
usage: python tools/mc3_run.py [nchains] [niter]   (uses every visible GPU)"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import torch
from strom_b200.host import host

nchains = int(sys.argv[1]) if len(sys.argv) > 1 else 8
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 200
fx = np.load(os.path.join(ROOT, "tests", "golden", "rbcl738.npz"))
cols = np.repeat(np.arange(fx["data"].shape[1]), fx["counts"].astype(int))
api = host()
h = api.data_from_codes(fx["data"].astype(np.int32)[:, cols], compress=True)
devices = tuple(range(torch.cuda.device_count()))
out = {"workload": "rbcl738 GTR+G4 MC3", "nchains": nchains, "niter": niter, "gpus": len(devices)}
for name, kw in (("in_turn", dict(per_chain_streams=True)), ("side_by_side", dict(parallel_chains=True)),
                 ("lockstep", dict(lockstep_chains=True))):
    api.mcmc(h, str(fx["newick"]), niter=10, burnin=0, nchains=nchains, samplefreq=10, devices=devices, **kw)   # warm-up
    tr, sec = api.mcmc(h, str(fx["newick"]), niter=niter, burnin=0, nchains=nchains, samplefreq=niter, devices=devices, seed=5, **kw)
    out[name] = {"seconds": sec, "chain_iterations_per_sec": nchains * niter / sec, "final_lnL": float(tr[-1, 1])}
print(json.dumps(out))
