"""rbcl738 through the C++ Likelihood facade (what strom's Chain / Updaters call): seconds per calcLogLikelihood."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from bench import PAUP_PI, PAUP_R, PAUP_ALPHA
from strom_b200.host import host
fx = np.load(os.path.join(ROOT, "tests", "golden", "rbcl738.npz"))
data = np.minimum(fx["data"], 4).astype(np.uint8)
P = data.shape[1]
H = host()
cols = np.repeat(np.arange(P), fx["counts"].astype(int))
hd = H.data_from_codes(data.astype(np.int32)[:, cols], compress=True)
for pinv in (0.0, 0.2):
    for rep in range(3):
        ll, sec = H.loglik(hd, str(fx["newick"]), PAUP_PI, PAUP_R, PAUP_ALPHA, 4, pinvar=pinv, repeats=2001)
        print("pinvar %.1f  logL %.6f  %.2f us per call  (%.0f evals/s)" % (pinv, ll, sec * 1e6, 1 / sec))
H.data_free(hd)
