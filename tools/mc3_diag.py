import os, sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from bench import mc3_data, MC3_KW
from strom_b200.host import host
H = host()
hd, newick = mc3_data(H)
kw = dict(MC3_KW, nchains=8, devices=(0,), niter=400, burnin=200, time_iterations_only=True)
for name, extra in (("lockstep batched", dict(lockstep_chains=True)), ("lockstep unbatched", dict(lockstep_chains=True, batch_evaluations=False)), ("in turn", dict(per_chain_streams=True))):
    H.mcmc(hd, newick, **dict(kw, niter=4, burnin=2), **extra)
    tr, sec = H.mcmc(hd, newick, **kw, **extra)
    print(name, "chain-it/s %.0f" % (8 * 600 / sec))
