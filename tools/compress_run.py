"""SURVEY.md 8(f) rank 2 measured: site-pattern compression of a 1000-taxon alignment on the GPU (sb200CompressPatterns,
host buffers in and out) against the C++ host path (Data::compressPatterns restated: stable sort of site indices).
usage: python tools/compress_run.py [ntaxa] [nsites_device] [nsites_host]"""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from strom_b200.engine import api
from strom_b200.host import host
ntaxa = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nd = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
nh = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
rng = np.random.default_rng(1)
def alignment(nsites):
    base = rng.integers(0, 4, size=(ntaxa, nsites // 2), dtype=np.uint8)
    return np.ascontiguousarray(base[:, rng.integers(0, base.shape[1], size=nsites)])
E, H = api(), host()
small = alignment(nh)
t0 = time.perf_counter(); hh = H.data_from_codes(small.astype(np.int32), compress=3); th = time.perf_counter() - t0
want, wc, _ = H.data_arrays(hh); H.data_free(hh)
E.compress_patterns(small)                                   # warm-up (context, first allocation)
t0 = time.perf_counter(); got, gc = E.compress_patterns(small); tds = time.perf_counter() - t0
assert np.array_equal(got, want) and np.array_equal(gc, wc)
big = alignment(nd)
t0 = time.perf_counter(); pb, cb = E.compress_patterns(big); tdb = time.perf_counter() - t0
assert cb.sum() == nd
print(json.dumps({"ntaxa": ntaxa, "host_path": {"sites": nh, "patterns": int(want.shape[1]), "seconds": th},
                  "device_same_input": {"sites": nh, "seconds": tds, "identical_output": True, "speedup": th / tds},
                  "device_large": {"sites": nd, "patterns": int(pb.shape[1]), "seconds": tdb, "sites_per_sec": nd / tdb}}))
