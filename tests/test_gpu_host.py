"""The C++ Likelihood facade (strom_b200/host/likelihood.hpp) driving the CUDA engine: the call a strom
Chain / Updater makes.  Compared with the golden fixtures (CPU oracle values)."""
import numpy as np
import pytest

from helpers import MODEL_NAMES, load_golden, rel_err
from test_host_cpu import MODEL_ARGS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hostlib():
    from strom_b200.host import host
    return host()            # libstrom_host.so, linked to libstrom_b200.so (no fallback)


def _data(hostlib, fx):
    cols = np.repeat(np.arange(fx["data"].shape[1]), fx["counts"].astype(int))
    return hostlib.data_from_codes(fx["data"].astype(np.int32)[:, cols], compress=True)


@pytest.mark.parametrize("key", ["rbcL", "rbcl10", "rbcl738"])
@pytest.mark.parametrize("model", MODEL_NAMES)
def test_cpp_facade_matches_goldens(hostlib, key, model):
    fx = load_golden(key)
    h = _data(hostlib, fx)
    ll, _ = hostlib.loglik(h, str(fx["newick"]), *MODEL_ARGS[model])
    assert rel_err(ll, float(fx["lnL_" + model])) < 1e-11
    hostlib.data_free(h)


@pytest.fixture(scope="module")
def hostlib13():
    # the same host classes compiled WITHOUT the engine-level entry point: Likelihood::calcLogLikelihood issues the
    # reference's 13 BeagleLib calls one by one (src/likelihood.hpp:409-447) -- and the library behind them is the CUDA engine
    from strom_b200.build import build_host_13call
    from strom_b200.host import HostAPI
    return HostAPI(build_host_13call())


@pytest.mark.parametrize("key", ["rbcL", "rbcl10", "rbcl738"])
@pytest.mark.parametrize("model", ["jc", "gtr_g4", "gtr_i_g4"])
def test_drop_in_13_call_sequence_matches_goldens(hostlib, hostlib13, key, model):
    fx = load_golden(key)
    h13 = _data(hostlib13, fx)
    ll13, _ = hostlib13.loglik(h13, str(fx["newick"]), *MODEL_ARGS[model], repeats=3)
    assert rel_err(ll13, float(fx["lnL_" + model])) < 1e-11
    hostlib13.data_free(h13)
    h = _data(hostlib, fx)
    ll, _ = hostlib.loglik(h, str(fx["newick"]), *MODEL_ARGS[model])
    hostlib.data_free(h)
    assert ll13 == ll                        # the fused engine call runs the very same kernels


@pytest.mark.parametrize("shards", [2, 3, 5])
def test_pattern_shards_sum_to_the_whole(hostlib, shards):
    # several shards (here all on device 0) = independent instances holding a pattern slice each;
    # the host adds one double per shard in rank order
    fx = load_golden("rbcl738")
    h = _data(hostlib, fx)
    one, _ = hostlib.loglik(h, str(fx["newick"]), *MODEL_ARGS["gtr_g4"], devices=(0,))
    many, _ = hostlib.loglik(h, str(fx["newick"]), *MODEL_ARGS["gtr_g4"], devices=tuple([0] * shards))
    assert rel_err(many, one) < 1e-13
    assert rel_err(many, float(fx["lnL_gtr_g4"])) < 1e-11
    hostlib.data_free(h)


def test_repeated_evaluations_are_bitwise_stable(hostlib):
    fx = load_golden("rbcl10")
    h = _data(hostlib, fx)
    a, _ = hostlib.loglik(h, str(fx["newick"]), *MODEL_ARGS["gtr_g4"], repeats=1)
    b, sec = hostlib.loglik(h, str(fx["newick"]), *MODEL_ARGS["gtr_g4"], repeats=50)
    assert a == b and sec > 0
    hostlib.data_free(h)


def test_fp32_facade(hostlib):
    fx = load_golden("rbcl738")
    h = _data(hostlib, fx)
    ll, _ = hostlib.loglik(h, str(fx["newick"]), *MODEL_ARGS["gtr_g4"], single=True)
    assert rel_err(ll, float(fx["lnL_gtr_g4"])) < 2e-5
    hostlib.data_free(h)


@pytest.mark.parametrize("key,ncat,pinvar,nsteps", [("rbcl10", 4, 0.0, 200), ("rbcL", 1, 0.0, 120), ("rbcl738", 4, 0.2, 150)])
def test_incremental_likelihood_follows_tree_moves(key, ncat, pinvar, nsteps):
    """Likelihood::setIncremental: after LOCAL-style moves (with and without NNI), take-backs and model changes the
    incremental facade lists a handful of operations and returns the log-likelihood of a full evaluation."""
    from helpers import PAUP_ALPHA, PAUP_PI, PAUP_R
    from strom_b200.host import host
    api = host()
    fx = load_golden(key)
    cols = np.repeat(np.arange(fx["data"].shape[1]), fx["counts"].astype(int))
    h = api.data_from_codes(fx["data"].astype(np.int32)[:, cols], compress=True)
    inc, full, nops, _, _ = api.incremental_walk(h, str(fx["newick"]), PAUP_PI, PAUP_R, PAUP_ALPHA, ncat, pinvar, nsteps=nsteps, seed=11)
    api.data_free(h)
    ntaxa = fx["data"].shape[0]
    assert len(inc) == nsteps and np.all(np.isfinite(inc))
    assert np.max(np.abs(inc - full) / np.abs(full)) < 1e-12
    assert nops[0] == ntaxa - 2                                         # first call: everything
    if ncat > 1:
        assert (nops[1:] == ntaxa - 2).sum() >= 2                       # and again after every change of the model
    moves = nops[nops < ntaxa - 2]
    assert len(moves) > nsteps // 2
    if ntaxa > 100:
        assert np.median(moves) < (ntaxa - 2) / 10
    assert len(set(np.round(full, 6))) > nsteps // 3                    # the walk really moves


def test_incremental_likelihood_on_pattern_shards():
    """The same walk with the patterns sharded over three engine instances (over the visible GPUs, round robin)."""
    import torch
    from helpers import PAUP_ALPHA, PAUP_PI, PAUP_R
    from strom_b200.host import host
    api = host()
    fx = load_golden("rbcl738")
    cols = np.repeat(np.arange(fx["data"].shape[1]), fx["counts"].astype(int))
    h = api.data_from_codes(fx["data"].astype(np.int32)[:, cols], compress=True)
    ng = torch.cuda.device_count()
    inc, full, nops, _, _ = api.incremental_walk(h, str(fx["newick"]), PAUP_PI, PAUP_R, PAUP_ALPHA, 4, 0.0, nsteps=60, seed=5,
                                                 devices=tuple(i % ng for i in range(3)))
    one, _, _, _, _ = api.incremental_walk(h, str(fx["newick"]), PAUP_PI, PAUP_R, PAUP_ALPHA, 4, 0.0, nsteps=60, seed=5)
    api.data_free(h)
    assert np.max(np.abs(inc - full) / np.abs(full)) < 1e-12 and np.max(np.abs(inc - one) / np.abs(one)) < 1e-12
    assert np.median(nops[nops < 736]) < 60


@pytest.mark.parametrize("ntaxa,nsites,alphabet,seed", [(5, 1, 4, 1), (3, 7, 2, 2), (12, 5000, 2, 3), (40, 70000, 5, 4),
                                                        (738, 1314, 6, 5), (7, 300000, 3, 6), (2, 2049, 20, 7)])
def test_device_pattern_compression_matches_host(ntaxa, nsites, alphabet, seed):
    """sb200CompressPatterns against Data::compressPatterns (reference src/data.hpp:102-161): the same distinct site
    columns, in the same (lexicographic) order, with the same counts -- engine ABI and C++ facade."""
    from strom_b200.engine import api
    from strom_b200.host import host
    rng = np.random.default_rng(seed)
    base = rng.integers(0, alphabet, size=(ntaxa, max(1, nsites // 3))).astype(np.uint8)
    codes = base[:, rng.integers(0, base.shape[1], size=nsites)]          # plenty of repeated columns
    H = host()
    hh = H.data_from_codes(codes.astype(np.int32), compress=3)
    want, wcounts, _ = H.data_arrays(hh)
    H.data_free(hh)
    got, counts = api().compress_patterns(codes)
    assert got.shape == want.shape and np.array_equal(got, want) and np.array_equal(counts, wcounts)
    assert counts.sum() == nsites
    # ... and directly against the oracle's restatement of Data::compressPatterns (sorted unique columns)
    from oracle import strom_host_ref as HR
    owant, ocounts = HR.compress_patterns(codes.astype(np.int32))
    assert np.array_equal(got, owant) and np.array_equal(counts, ocounts)
    hd = H.data_from_codes(codes.astype(np.int32), compress=2)
    m, w, ns = H.data_arrays(hd)
    H.data_free(hd)
    assert ns == nsites and np.array_equal(m, want) and np.array_equal(w, wcounts)


def test_device_pattern_compression_rejects_wide_codes():
    from strom_b200.beagle import BeagleError
    from strom_b200.engine import api
    codes = np.zeros((3, 10), dtype=np.uint8)
    codes[1, 4] = 32
    with pytest.raises(BeagleError):
        api().compress_patterns(codes)


def test_incremental_cache_is_tied_to_the_tree_object_not_its_address():
    """the incremental facade used to recognise "the same tree" by its address.  A tree that is destroyed and
    rebuilt (TreeManip::buildFromNewick makes a new Tree, usually at the address the old one had) must be evaluated afresh."""
    from helpers import PAUP_ALPHA, PAUP_PI, PAUP_R
    from strom_b200.host import host
    H = host()
    fx = load_golden("rbcl10")
    cols = np.repeat(np.arange(fx["data"].shape[1]), fx["counts"].astype(int))
    h = H.data_from_codes(fx["data"].astype(np.int32)[:, cols], compress=True)
    bad = H.rebuild_check(h, str(fx["newick"]), PAUP_PI, PAUP_R, PAUP_ALPHA, 4, rounds=12, seed=3)
    H.data_free(h)
    assert bad < 1e-12
