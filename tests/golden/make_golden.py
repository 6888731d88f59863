"""Generates tests/golden/*.npz from the reference's own fixture files.

Run in the build container (``/root/reference`` present):  python tests/golden/make_golden.py

For each fixture pair it stores what ``Likelihood`` hands to the 13 BeagleLib calls --
compressed tip states, pattern counts, operation list, matrix indices, edge lengths
(src/likelihood.hpp:228-264, :296-355) -- derived with oracle/strom_host_ref.py, plus the
log-likelihoods the CPU oracle (oracle/beagle_cpu.c) gives for named models.  Two of those are
the reference's own golden values (paup.nex:12, rbcLjc.tre:57); the others are pinned only by
the oracle.  The GPU box has no /root/reference, so GPU tests read these files instead.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import strom_host_ref as H  # noqa: E402
from strom_b200.beagle import BeagleAPI  # noqa: E402
from strom_b200.build import build_oracle  # noqa: E402

REF = "/root/reference"
JC = ([0.25, 0.25, 0.25, 0.25], [1.0] * 6)
PAUP_PI = [0.309769, 0.163380, 0.121023]
PAUP_PI.append(1.0 - sum(PAUP_PI))                       # paup.nex:8
PAUP_R = [1.06923, 4.34568, 0.45898, 2.00465, 3.85915, 1.0]
PAUP_ALPHA = 0.517281

MODELS = {
    # name: (freqs, exchangeabilities, rates, probs)
    "jc": lambda: JC + H.gamma_rates(0.5, 1),
    "jc_g4": lambda: JC + H.gamma_rates(0.5, 4),
    "gtr_g4": lambda: (PAUP_PI, PAUP_R) + H.gamma_rates(PAUP_ALPHA, 4),
    "gtr_i_g4": lambda: (PAUP_PI, PAUP_R) + H.invariant_gamma_rates(PAUP_ALPHA, 4, 0.2),
    "gtr_g2": lambda: (PAUP_PI, PAUP_R) + H.gamma_rates(PAUP_ALPHA, 2),
    "gtr_g8": lambda: (PAUP_PI, PAUP_R) + H.gamma_rates(PAUP_ALPHA, 8),
    "gtr_g3": lambda: (PAUP_PI, PAUP_R) + H.gamma_rates(PAUP_ALPHA, 3),
}

FIXTURES = {
    "rbcL": ("rbcL.nex", "rbcLjc.tre"),
    "rbcl10": ("rbcl10.nex", "rbcl10.tre"),
    "rbcl738": ("rbcl738.nex", "rbcl738nj.tre"),
}


def model_dict(name):
    freqs, rmat, rates, probs = MODELS[name]()
    evec, ivec, evals = H.gtr_eigen(freqs, rmat)
    return dict(freqs=np.asarray(freqs, dtype=np.float64), evec=evec, ivec=ivec, evals=evals, rates=rates, probs=probs)


def main():
    api = BeagleAPI(build_oracle())
    for key, (nex, tre) in FIXTURES.items():
        names, data, counts, nsites = H.read_nexus_data(os.path.join(REF, nex))
        newick = H.read_nexus_trees(os.path.join(REF, tre))[0]
        tree = H.build_tree(newick)
        ops, pidx, elen = H.define_operations(tree)
        n, P = data.shape
        out = dict(data=data.astype(np.uint8), counts=counts, ops=ops, pmatrix_index=pidx, edge_lengths=elen,
                   parent=np.int32(tree.preorder[0].number), nsites=np.int32(nsites), newick=np.array(newick))
        for mname in MODELS:
            m = model_dict(mname)
            inst = api.create(n, P, len(m["rates"]))
            api.set_data(inst, data, counts)
            ll = api.calc_log_likelihood(inst, m, ops, pidx, elen, int(out["parent"]), 0)
            api.finalize(inst)
            out["lnL_" + mname] = np.float64(ll)
            print("%-8s %-9s n=%d P=%d lnL=%.9f" % (key, mname, n, P, ll))
        np.savez_compressed(os.path.join(HERE, key + ".npz"), **out)


if __name__ == "__main__":
    main()
