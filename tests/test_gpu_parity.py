"""Parity of the sm_100a engine against the CPU oracle, through the C ABI (include/strom_beagle.h).

Tolerance: the north-star bar is 1e-9 relative on logL for FP64; the kernels differ from the
oracle only in summation order, FMA contraction and exp/log ULPs, so the tests ask for 1e-11.
Partials, matrices and cumulative log-scalers are compared buffer by buffer."""
import ctypes as C

import numpy as np
import pytest

from helpers import MODEL_NAMES, Extra, full_eval, load_golden, named_model, random_problem, rel_err

pytestmark = pytest.mark.gpu

LOGL_RTOL = 1e-11          # FP64 (north-star requirement: 1e-9)
LOGL_RTOL_FP32 = 2e-5      # FP32 partials/matrices, FP64 scalers and accumulation (stated, measured in test)


@pytest.mark.parametrize("key", ["rbcL", "rbcl10", "rbcl738"])
@pytest.mark.parametrize("model", MODEL_NAMES)
def test_logl_matches_golden_fixture(engine, key, model):
    fx = load_golden(key)
    ll = full_eval(engine, fx, named_model(model))
    assert rel_err(ll, float(fx["lnL_" + model])) < LOGL_RTOL


def test_reference_golden_values_on_gpu(engine):
    # paup.nex:12 and rbcLjc.tre:57, straight from the engine
    assert abs(full_eval(engine, load_golden("rbcl738"), named_model("gtr_g4")) - (-144730.8)) < 0.05
    assert abs(-full_eval(engine, load_golden("rbcL"), named_model("jc")) - 286.9238) < 5e-5


@pytest.mark.parametrize("key,model", [("rbcl10", "jc"), ("rbcl10", "gtr_g4"), ("rbcl10", "gtr_i_g4"), ("rbcL", "gtr_g8"),
                                       ("rbcl738", "gtr_g4"), ("rbcl738", "gtr_g2")])
def test_every_buffer_matches_oracle(engine, oracle, key, model):
    fx = load_golden(key)
    m = named_model(model)
    n, P = fx["data"].shape
    Cc = len(m["rates"])
    ll_o, io = full_eval(oracle, fx, m, keep=True)
    ll_e, ie = full_eval(engine, fx, m, keep=True)
    xo, xe = Extra(oracle), Extra(engine)
    for mi in set(fx["pmatrix_index"].tolist()):
        assert np.allclose(xe.matrix(ie, mi, Cc), xo.matrix(io, mi, Cc), rtol=1e-12, atol=1e-15)
    for op in fx["ops"]:
        a, b = xe.partials(ie, int(op[0]), P, Cc), xo.partials(io, int(op[0]), P, Cc)
        # The fixture trees have edges of 0 and 1e-8 (rbcLjc.tre, rbcl738nj.tre).  There an off-diagonal
        # P_ij(t) ~ q*t is a sum of cancelling eigen terms, accurate to ~2e-16 ABSOLUTE in any implementation
        # (the reference's included), i.e. ~1e-8 relative; two tips that differ across such edges make a whole
        # pattern's partials out of those entries and rescaling lifts them to O(1).  Hence 1e-7 here; the
        # well-conditioned test below holds the kernels to 1e-12.
        assert np.allclose(a, b, rtol=1e-7, atol=1e-13), (int(op[0]), float(np.abs(a - b).max()))
    assert np.allclose(xe.scale(ie, 0, P), xo.scale(io, 0, P), rtol=1e-12, atol=1e-12)
    assert rel_err(ll_e, ll_o) < LOGL_RTOL
    oracle.finalize(io)
    engine.finalize(ie)


@pytest.mark.parametrize("model", ["jc", "gtr_g2", "gtr_g3", "gtr_g4", "gtr_i_g4", "gtr_g8"])
def test_every_buffer_matches_oracle_well_conditioned(engine, oracle, model):
    # edges >= 0.02: every P entry is accurate to a few ulps, so every buffer must agree to rounding
    fx = random_problem(48, 333, 77, missing_frac=0.03)
    fx["edge_lengths"] = fx["edge_lengths"] + 0.02
    m = named_model(model)
    n, P = fx["data"].shape
    Cc = len(m["rates"])
    ll_o, io = full_eval(oracle, fx, m, keep=True)
    ll_e, ie = full_eval(engine, fx, m, keep=True)
    xo, xe = Extra(oracle), Extra(engine)
    for mi in set(fx["pmatrix_index"].tolist()):
        assert np.allclose(xe.matrix(ie, mi, Cc), xo.matrix(io, mi, Cc), rtol=1e-13, atol=1e-16)
    for op in fx["ops"]:
        a, b = xe.partials(ie, int(op[0]), P, Cc), xo.partials(io, int(op[0]), P, Cc)
        # atol: the slowest gamma category of G8 has rate ~1e-4, i.e. an effective edge of ~2e-6
        assert np.allclose(a, b, rtol=1e-12, atol=1e-14), (int(op[0]), float(np.abs(a - b).max()))
    assert np.allclose(xe.scale(ie, 0, P), xo.scale(io, 0, P), rtol=1e-13, atol=1e-13)
    assert rel_err(ll_e, ll_o) < 1e-13
    oracle.finalize(io)
    engine.finalize(ie)


@pytest.mark.parametrize("ntaxa,npat,model,seed", [(4, 1, "jc", 1), (5, 7, "gtr_g4", 2), (33, 129, "gtr_g4", 3),
                                                   (64, 1000, "gtr_i_g4", 4), (200, 257, "gtr_g8", 5),
                                                   (17, 63, "gtr_g3", 6), (100, 4097, "jc", 7), (3, 5, "gtr_g2", 8)])
def test_random_trees_and_ragged_sizes(engine, oracle, ntaxa, npat, model, seed):
    fx = random_problem(ntaxa, npat, seed, missing_frac=0.05)
    m = named_model(model)
    assert rel_err(full_eval(engine, fx, m), full_eval(oracle, fx, m)) < LOGL_RTOL


# Every category count the engine takes, at sizes on both sides of the tile switch (P*C >= 2^18 runs the streaming
# variant: 2 patterns per thread, 256-thread blocks); 3, 5 (+I+G4), 6 and 7 leave 2-4 lanes of each warp idle.
@pytest.mark.parametrize("ntaxa,npat,model,seed", [(14, 70001, "gtr_g4", 21), (14, 60013, "gtr_i_g4", 22), (9, 100003, "gtr_g3", 23),
                                                   (9, 50021, "gtr_g6", 24), (9, 40009, "gtr_g7", 25), (40, 40001, "gtr_g8", 26),
                                                   (30, 1235, "gtr_g6", 27), (30, 97, "gtr_g7", 28), (50, 2000, "gtr_g5", 29),
                                                   (12, 300007, "jc", 30)])
def test_every_category_count_both_tile_sizes(engine, oracle, ntaxa, npat, model, seed):
    fx = random_problem(ntaxa, npat, seed, missing_frac=0.05)
    m = named_model(model)
    ex, ox = Extra(engine), Extra(oracle)
    Cc = len(m["rates"])
    le, ie = full_eval(engine, fx, m, keep=True)
    lo, io = full_eval(oracle, fx, m, keep=True)
    assert rel_err(le, lo) < LOGL_RTOL
    root = int(fx["parent"])
    np.testing.assert_allclose(ex.partials(ie, root, npat, Cc), ox.partials(io, root, npat, Cc), rtol=1e-7, atol=1e-13)
    np.testing.assert_allclose(ex.scale(ie, 0, npat), ox.scale(io, 0, npat), rtol=1e-12, atol=1e-12)
    oracle.finalize(io)
    engine.finalize(ie)


def test_all_missing_column_and_zero_length_edges(engine, oracle):
    fx = random_problem(12, 40, 11, missing_frac=0.0)
    fx["data"][:, 3] = 4                      # a pattern with no information: site likelihood 1
    fx["data"][5, :] = 4                      # a taxon that is all gaps
    # zero-length edges (rbcl738nj.tre has them): P = I.  Only on internal edges and on the all-gap taxon:
    # two conflicting tips joined by zero-length edges would make the site likelihood pure rounding noise
    # (in the reference as well), which no implementation can reproduce.
    internal = np.flatnonzero(fx["pmatrix_index"] >= 12)
    fx["edge_lengths"][internal[::2]] = 0.0
    fx["edge_lengths"][internal[1]] = 1e-12   # Node::_smallest_edge_length
    fx["edge_lengths"][np.flatnonzero(fx["pmatrix_index"] == 5)] = 0.0
    m = named_model("gtr_g4")
    assert rel_err(full_eval(engine, fx, m), full_eval(oracle, fx, m)) < LOGL_RTOL


def test_underflow_needs_rescaling(engine, oracle):
    # long edges on many taxa: unscaled partials would underflow double precision
    fx = random_problem(600, 64, 12, missing_frac=0.0, mean_edge=1.5)
    m = named_model("gtr_g4")
    lo, le = full_eval(oracle, fx, m), full_eval(engine, fx, m)
    assert np.isfinite(lo) and lo < -700 * 64 / 64
    assert rel_err(le, lo) < LOGL_RTOL


def _shaped_problem(newick, npat, seed):
    from oracle import strom_host_ref as H
    rng = np.random.default_rng(seed)
    tree = H.build_tree(newick)
    ops, pidx, elen = H.define_operations(tree)
    n = tree.nleaves
    data = rng.integers(0, 4, size=(n, npat)).astype(np.uint8)
    data[rng.random(data.shape) < 0.02] = 4
    return dict(data=data, counts=rng.integers(1, 4, size=npat).astype(np.float64), ops=ops, pmatrix_index=pidx,
                edge_lengths=elen, parent=np.int32(tree.preorder[0].number))


@pytest.mark.parametrize("model,shape", [("gtr_g4", ""), ("jc", ""), ("gtr_g4", "1"), ("gtr_g4", "2"), ("gtr_i_g4", "4"), ("gtr_g4", "3")])
def test_caterpillar_tree_is_one_long_chain(engine, oracle, model, shape, monkeypatch):
    # 700 taxa in a ladder: ONE chain of 698 operations, longer than the 128 operations a block keeps in shared
    # memory (chunked path, in every thread shape), with enough rescaling for the deferred log to fold several times
    if shape:
        monkeypatch.setenv("SB200_VARIANT", shape)
    else:
        monkeypatch.delenv("SB200_VARIANT", raising=False)
    n = 700
    nwk = "1:0.3"
    for t in range(2, n - 1):
        nwk = "(%s,%d:0.4):0.2" % (nwk, t)
    nwk = "(%s,%d:0.3,%d:0.1)" % (nwk, n - 1, n)
    fx = _shaped_problem(nwk, 70, 5)
    m = named_model(model)
    lo, le = full_eval(oracle, fx, m), full_eval(engine, fx, m)
    assert np.isfinite(lo) and rel_err(le, lo) < LOGL_RTOL


def test_balanced_tree_has_the_most_launches(engine, oracle):
    def sub(lo, hi):
        if hi - lo == 1:
            return "%d:0.07" % (lo + 1)
        mid = (lo + hi) // 2
        return "(%s,%s):0.05" % (sub(lo, mid), sub(mid, hi))
    nwk = "(%s,%s,%s)" % (sub(0, 86), sub(86, 171), sub(171, 256))
    fx = _shaped_problem(nwk, 333, 6)
    m = named_model("gtr_g4")
    assert rel_err(full_eval(engine, fx, m), full_eval(oracle, fx, m)) < LOGL_RTOL


def test_repeat_calls_and_changed_parameters(engine, oracle):
    # one instance, many calls: new edge lengths, new model, new topology (schedule cache must follow)
    fx = random_problem(40, 300, 21)
    fx2 = random_problem(40, 300, 22)
    fx2["data"], fx2["counts"] = fx["data"], fx["counts"]
    m1, m2 = named_model("gtr_g4"), named_model("jc_g4")
    n, P = fx["data"].shape
    ie, io = engine.create(n, P, 4), oracle.create(n, P, 4)
    for api, inst in ((engine, ie), (oracle, io)):
        api.set_data(inst, fx["data"].astype(np.int32), fx["counts"])
    rng = np.random.default_rng(5)
    for it in range(6):
        f = fx if it % 3 != 2 else fx2
        mm = m1 if it % 2 == 0 else m2
        el = f["edge_lengths"] * rng.uniform(0.5, 1.5, size=f["edge_lengths"].size)
        a = engine.calc_log_likelihood(ie, mm, f["ops"], f["pmatrix_index"], el, int(f["parent"]), 0)
        b = oracle.calc_log_likelihood(io, mm, f["ops"], f["pmatrix_index"], el, int(f["parent"]), 0)
        assert rel_err(a, b) < LOGL_RTOL, it
    engine.finalize(ie)
    oracle.finalize(io)


def test_frequencies_and_weights_set_after_the_matrix_update_reach_the_root_call(engine, oracle):
    # BeagleLib's setters take effect at the next call that reads them: new state frequencies / category weights followed
    # directly by beagleCalculateEdgeLogLikelihoods (no matrix update in between) must be the ones the root reduction uses
    fx = random_problem(20, 333, 41)
    m = named_model("gtr_g4")
    new_freqs = np.array([0.1, 0.2, 0.3, 0.4])
    new_probs = np.array([0.4, 0.3, 0.2, 0.1])
    got = []
    for api in (engine, oracle):
        ll0, inst = full_eval(api, fx, m, keep=True)
        assert api.beagleSetStateFrequencies(inst, 0, new_freqs.ctypes.data_as(C.POINTER(C.c_double))) == 0
        assert api.beagleSetCategoryWeights(inst, 0, new_probs.ctypes.data_as(C.POINTER(C.c_double))) == 0
        zero = np.zeros(1, dtype=np.int32)
        par = np.array([int(fx["parent"])], dtype=np.int32)
        out = np.zeros(1)
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
        assert api.beagleCalculateEdgeLogLikelihoods(inst, ip(par), ip(zero), ip(zero), None, None, ip(zero), ip(zero), ip(zero), 1,
                                                     out.ctypes.data_as(C.POINTER(C.c_double)), None, None) == 0
        got.append((ll0, float(out[0])))
        api.finalize(inst)
    assert rel_err(got[0][0], got[1][0]) < LOGL_RTOL
    assert rel_err(got[0][1], got[1][1]) < LOGL_RTOL
    assert abs(got[0][1] - got[0][0]) > 1.0            # the new numbers did change the result


def test_operation_list_in_any_valid_order(engine, oracle):
    # the reference emits reverse level-order; any children-first order must give the same answer
    fx = random_problem(30, 50, 31)
    m = named_model("gtr_g4")
    base = full_eval(oracle, fx, m)
    # postorder by destination number is also children-first (internals are numbered in postorder)
    fx2 = dict(fx)
    fx2["ops"] = fx["ops"][np.argsort(fx["ops"][:, 0])]
    assert rel_err(full_eval(engine, fx2, m), base) < LOGL_RTOL


def test_non_forest_operation_list_runs_sequentially(engine, oracle):
    # a buffer recomputed twice in one list is legal BeagleLib usage; the planner must fall back
    fx = random_problem(10, 33, 41)
    m = named_model("gtr_g4")
    fx2 = dict(fx)
    fx2["ops"] = np.concatenate([fx["ops"][:3], fx["ops"]])
    assert rel_err(full_eval(engine, fx2, m), full_eval(oracle, fx2, m)) < LOGL_RTOL


def test_error_codes(engine):
    fx = load_golden("rbcl10")
    n, P = fx["data"].shape
    m = named_model("jc")
    assert engine.beagleFinalizeInstance(12345) == -4
    assert engine.beagleSetPatternWeights(999, fx["counts"].ctypes.data_as(C.POINTER(C.c_double))) == -4
    inst = engine.create(n, P, 1)
    row = fx["data"][0].astype(np.int32)
    assert engine.beagleSetTipStates(inst, n, row.ctypes.data_as(C.POINTER(C.c_int))) == -5
    # tips not uploaded yet: partials update must refuse
    ops = np.ascontiguousarray(fx["ops"], dtype=np.int32)
    assert engine.beagleUpdatePartials(inst, ops.ctypes.data_as(C.POINTER(C.c_int)), 8, 0) == -5
    engine.set_data(inst, fx["data"].astype(np.int32), fx["counts"])
    bad = ops.copy()
    bad[0, 0] = 500
    assert engine.beagleUpdatePartials(inst, bad.ctypes.data_as(C.POINTER(C.c_int)), 8, 0) == -5
    bad = ops.copy()
    bad[0, 2] = 1            # scale-read is not supported (the reference never uses it)
    assert engine.beagleUpdatePartials(inst, bad.ctypes.data_as(C.POINTER(C.c_int)), 8, 0) == -7
    # NaN log-likelihood -> -8 (BEAGLE_ERROR_FLOATING_POINT): negative frequencies make the site likelihood negative
    m_bad = dict(m)
    m_bad["freqs"] = np.array([-1.0, -1.0, -1.0, -1.0])
    with pytest.raises(Exception) as ei:
        engine.calc_log_likelihood(inst, m_bad, fx["ops"], fx["pmatrix_index"], fx["edge_lengths"], int(fx["parent"]), 0)
    assert getattr(ei.value, "code", None) == -8
    # and the instance still works afterwards
    ll = engine.calc_log_likelihood(inst, m, fx["ops"], fx["pmatrix_index"], fx["edge_lengths"], int(fx["parent"]), 0)
    assert rel_err(ll, float(fx["lnL_jc"])) < LOGL_RTOL
    engine.finalize(inst)
    assert engine.beagleFinalizeInstance(inst) == -4
    # unsupported shapes
    assert engine.beagleCreateInstance(4, 2, 4, 20, 10, 1, 5, 1, 3, None, 0, 0, 1 << 1 | 1 << 6, None) == -7   # 20 states
    assert engine.beagleCreateInstance(4, 2, 4, 4, 10, 1, 5, 9, 3, None, 0, 0, 1 << 1 | 1 << 6, None) == -7    # 9 categories


@pytest.mark.parametrize("key,model", [("rbcl10", "gtr_g4"), ("rbcl738", "gtr_g4"), ("rbcl738", "gtr_i_g4"), ("rbcL", "jc")])
def test_fp32_variant_within_stated_tolerance(engine, key, model):
    fx = load_golden(key)
    ll = full_eval(engine, fx, named_model(model), single=True)
    assert rel_err(ll, float(fx["lnL_" + model])) < LOGL_RTOL_FP32


def test_many_instances_coexist(engine):
    # one Likelihood (instance) per chain, reference src/strom.hpp:249
    fx = load_golden("rbcl10")
    res = []
    insts = []
    for name in ("jc", "jc_g4", "gtr_g4", "gtr_i_g4"):
        m = named_model(name)
        ll, inst = full_eval(engine, fx, m, keep=True)
        insts.append(inst)
        res.append((name, ll))
    assert len(set(insts)) == 4
    for name, ll in res:
        assert rel_err(ll, float(fx["lnL_" + name])) < LOGL_RTOL
    for inst in insts:
        engine.finalize(inst)


@pytest.mark.parametrize("source,model,seed", [(("random", 40, 300), "gtr_g4", 1), (("random", 25, 70), "gtr_g3", 2),
                                               (("random", 60, 2000), "jc", 3), (("golden", "rbcl738"), "gtr_g4", 4),
                                               (("random", 120, 500), "gtr_g8", 5)])
def test_incremental_lists_cover_only_the_changed_path(engine, oracle, source, model, seed):
    """with sb200SetSubtreeScalers the operation list may hold only the nodes above the
    edges that changed (and only the changed matrices); logL must equal a full evaluation of the CPU oracle."""
    fx = random_problem(source[1], source[2], 100 + seed) if source[0] == "random" else load_golden(source[1])
    m = named_model(model)
    ops = np.asarray(fx["ops"], dtype=np.int32).reshape(-1, 7)
    pidx = np.asarray(fx["pmatrix_index"], dtype=np.int32)
    el = np.asarray(fx["edge_lengths"], dtype=np.float64).copy()
    parent = int(fx["parent"])
    parent_of = {}
    for o in ops:
        parent_of[int(o[3])] = int(o[0])
        parent_of[int(o[5])] = int(o[0])
    n, P = fx["data"].shape
    ie, io = engine.create(n, P, len(m["rates"])), oracle.create(n, P, len(m["rates"]))
    for api, inst in ((engine, ie), (oracle, io)):
        api.set_data(inst, fx["data"].astype(np.int32), fx["counts"])
    engine.lib.sb200SetSubtreeScalers.argtypes = [C.c_int, C.c_int]
    assert engine.lib.sb200SetSubtreeScalers(ie, 1) == 0
    a = engine.calc_log_likelihood(ie, m, ops, pidx, el, parent, 0)
    b = oracle.calc_log_likelihood(io, m, ops, pidx, el, parent, 0)
    assert rel_err(a, b) < LOGL_RTOL
    rng = np.random.default_rng(seed)
    sizes = []
    for it in range(12):
        changed = rng.choice(pidx.size, size=int(rng.integers(1, 4)), replace=False)
        if it == 5:
            changed = np.array([pidx.size - 1])                     # the root edge alone: no partial changes at all
        el[changed] *= rng.uniform(0.3, 2.0, size=changed.size)
        dirty = {parent}                                            # the root's partials are always listed
        for e in changed:
            v = parent_of.get(int(pidx[e])) if e != pidx.size - 1 else None
            while v is not None:
                dirty.add(v)
                v = parent_of.get(v)
        sub = np.array([o for o in ops if int(o[0]) in dirty], dtype=np.int32)
        sizes.append(len(sub))
        a = engine.calc_log_likelihood(ie, m, sub, pidx[changed], el[changed], parent, 0)
        b = oracle.calc_log_likelihood(io, m, ops, pidx, el, parent, 0)
        assert rel_err(a, b) < LOGL_RTOL, (it, len(sub))
    assert min(sizes) == 1 and max(sizes) < len(ops)
    # a new model invalidates everything: full list again, then incremental lists keep working
    # (same category count as the instance: 13-call uploads carry no count of their own)
    from helpers import PAUP_PI, PAUP_R, make_model
    from oracle import strom_host_ref as HR
    ncat = len(m["rates"])
    m2 = make_model(PAUP_PI, PAUP_R, *(HR.invariant_gamma_rates(0.9, ncat - 1, 0.3) if model == "gtr_i_g4" else HR.gamma_rates(0.9, ncat)))
    assert len(m2["rates"]) == ncat
    a = engine.calc_log_likelihood(ie, m2, ops, pidx, el, parent, 0)
    b = oracle.calc_log_likelihood(io, m2, ops, pidx, el, parent, 0)
    assert rel_err(a, b) < LOGL_RTOL
    engine.finalize(ie)
    oracle.finalize(io)


# ---- sets of evaluations as one set of launches (sb200EvalBatchAsync) ---------------------------------------------------
def _batch_setup(engine, problems, models, single=False):
    from strom_b200.engine import EvalArgs, StromEvalArgs, api
    E = api()
    insts, eas = [], []
    for pr, m in zip(problems, models):
        n, P = pr["data"].shape
        inst = engine.create(n, P, len(m["rates"]), single=single)
        engine.set_data(inst, pr["data"].astype(np.int32), pr["counts"])
        insts.append(inst)
        eas.append(EvalArgs(m, pr["ops"], pr["pmatrix_index"], pr["edge_lengths"], int(pr["parent"]), 0))
    return E, insts, eas


def _batch_eval(E, insts, eas):
    from strom_b200.engine import StromEvalArgs
    K = len(insts)
    ci = (C.c_int * K)(*insts)
    ca = (StromEvalArgs * K)(*[e.c for e in eas])
    rc = E.sb200EvalBatchAsync(K, ci, ca)
    if rc != 0:
        return rc, None
    out, vals = C.c_double(), []
    for i in insts:
        assert E.sb200Finish(i, C.byref(out)) == 0
        vals.append(out.value)
    return 0, vals


@pytest.mark.parametrize("model,K,ntaxa,npat", [("gtr_g4", 8, 30, 700), ("gtr_i_g4", 3, 17, 95), ("jc", 11, 9, 33), ("gtr_g8", 2, 60, 1500)])
def test_batched_evaluations_match_the_oracle_chain_by_chain(engine, oracle, model, K, ntaxa, npat):
    """K chains -- the same alignment, a tree (own topology and edge lengths) and a model each -- as one set of launches:
    every chain's value against the CPU oracle, every chain's root partials and cumulative scalers too; and bit for bit the
    value of the same chain evaluated on its own."""
    from helpers import PAUP_PI, PAUP_R, make_model
    from oracle import strom_host_ref as HR
    base = random_problem(ntaxa, npat, seed=100 + K)
    problems, models = [], []
    ncat = len(named_model(model)["rates"])
    for c in range(K):
        pr = random_problem(ntaxa, npat, seed=200 + 7 * c)          # another topology ...
        pr["data"], pr["counts"] = base["data"], base["counts"]     # ... on the same alignment
        problems.append(pr)
        if model == "jc":
            models.append(named_model("jc"))
        elif model == "gtr_i_g4":
            models.append(make_model(PAUP_PI, PAUP_R, *HR.invariant_gamma_rates(0.3 + 0.2 * c, 4, 0.1 + 0.05 * c)))
        else:
            models.append(make_model(PAUP_PI, PAUP_R, *HR.gamma_rates(0.3 + 0.2 * c, ncat)))
    E, insts, eas = _batch_setup(engine, problems, models)
    rc, vals = _batch_eval(E, insts, eas)
    assert rc == 0
    xe, xo = Extra(engine), Extra(oracle)
    for c, (pr, m) in enumerate(zip(problems, models)):
        want, io = full_eval(oracle, pr, m, keep=True)
        assert rel_err(vals[c], want) < LOGL_RTOL, c
        P, Cc = pr["data"].shape[1], len(m["rates"])
        root = int(pr["parent"])
        # (tolerances of test_every_buffer_matches_oracle: random trees have edges short enough for ~1e-8 relative noise
        # in off-diagonal P entries, in any implementation)
        assert np.allclose(xe.partials(insts[c], root, P, Cc), xo.partials(io, root, P, Cc), rtol=1e-7, atol=1e-13)
        assert np.allclose(xe.scale(insts[c], 0, P), xo.scale(io, 0, P), rtol=1e-12, atol=1e-12)
        oracle.finalize(io)
    # the same chains one at a time: the same bits
    for c, (inst, ea) in enumerate(zip(insts, eas)):
        out = C.c_double()
        assert E.sb200Eval(inst, C.byref(ea.c), C.byref(out)) == 0
        assert out.value == vals[c], c
    # a second batched call after single calls (plans for both launch shapes are kept)
    rc, again = _batch_eval(E, insts, eas)
    assert rc == 0 and again == vals
    for i in insts:
        engine.finalize(i)


def test_batched_evaluation_rejects_bad_sets_without_enqueueing(engine):
    pr = random_problem(12, 64, seed=5)
    E, insts, eas = _batch_setup(engine, [pr, pr], [named_model("gtr_g4"), named_model("gtr_g4")])
    rc, vals = _batch_eval(E, insts, eas)
    assert rc == 0
    # the same instance twice
    rc, _ = _batch_eval(E, [insts[0], insts[0]], eas)
    assert rc == -5
    # a chain whose category count differs from the first one's
    E2, other, eo = _batch_setup(engine, [pr], [named_model("gtr_g8")])
    rc, _ = _batch_eval(E, [insts[0], other[0]], [eas[0], eo[0]])
    assert rc == -5
    # an out-of-range matrix index in the SECOND chain: nothing of the first one may have been touched
    from strom_b200.engine import EvalArgs
    bad = EvalArgs(named_model("gtr_g4"), pr["ops"], pr["pmatrix_index"] + 1000, pr["edge_lengths"], int(pr["parent"]), 0)
    rc, _ = _batch_eval(E, insts, [eas[0], bad])
    assert rc == -5
    # a model with another category count than the instance was created with
    wrong = EvalArgs(named_model("gtr_g8"), pr["ops"], pr["pmatrix_index"], pr["edge_lengths"], int(pr["parent"]), 0)
    rc, _ = _batch_eval(E, insts, [eas[0], wrong])
    assert rc == -5
    rc, again = _batch_eval(E, insts, eas)
    assert rc == 0 and again == vals
    for i in insts + other:
        engine.finalize(i)


@pytest.mark.parametrize("model", ["gtr_g4", "gtr_i_g4", "jc", "gtr_g3"])
def test_every_thread_shape_gives_the_same_bits(engine, oracle, model, monkeypatch):
    """The launches of a small instance choose between three thread shapes (all categories of a pattern in a thread / a warp
    per category / a thread per state); forcing each one on every launch gives bit-identical partials, scalers and logL."""
    pr = random_problem(25, 300, seed=77)
    m = named_model(model)
    P, Cc = pr["data"].shape[1], len(m["rates"])
    want = full_eval(oracle, pr, m)
    xe = Extra(engine)
    seen = []
    for shape in ("", "1", "2", "3", "4"):
        if shape:
            monkeypatch.setenv("SB200_VARIANT", shape)
        else:
            monkeypatch.delenv("SB200_VARIANT", raising=False)
        ll, inst = full_eval(engine, pr, m, keep=True)                    # (the instance reads the variable when it is created)
        bufs = [xe.partials(inst, 25 + b, P, Cc) for b in range(23)]
        seen.append((ll, bufs, xe.scale(inst, 0, P)))
        engine.finalize(inst)
        assert rel_err(ll, want) < LOGL_RTOL
    for ll, bufs, sc in seen[1:]:
        assert ll == seen[0][0] and np.array_equal(sc, seen[0][2])
        assert all(np.array_equal(a, b) for a, b in zip(bufs, seen[0][1]))


def test_incremental_list_that_reads_a_slot_nobody_wrote_is_rejected(engine):
    """sb200SetSubtreeScalers: the first list after enabling must cover every node; a partial first list would read
    subtree scalers that no call has stored -- now BEAGLE_ERROR_OUT_OF_RANGE instead of a wrong value."""
    pr = random_problem(16, 100, seed=9)
    m = named_model("gtr_g4")
    n, P = pr["data"].shape
    inst = engine.create(n, P, 4)
    engine.set_data(inst, pr["data"].astype(np.int32), pr["counts"])
    engine.lib.sb200SetSubtreeScalers.argtypes = [C.c_int, C.c_int]
    assert engine.lib.sb200SetSubtreeScalers(inst, 1) == 0
    from strom_b200.beagle import BeagleError
    root_only = pr["ops"][-1:]
    with pytest.raises(BeagleError) as e:
        engine.calc_log_likelihood(inst, m, root_only, pr["pmatrix_index"], pr["edge_lengths"], int(pr["parent"]), 0)
    assert e.value.code == -5
    full = engine.calc_log_likelihood(inst, m, pr["ops"], pr["pmatrix_index"], pr["edge_lengths"], int(pr["parent"]), 0)
    again = engine.calc_log_likelihood(inst, m, root_only, pr["pmatrix_index"][-1:], pr["edge_lengths"][-1:], int(pr["parent"]), 0)
    assert rel_err(again, full) < 1e-13
    engine.finalize(inst)


@pytest.mark.parametrize("source,model", [("rbcl738", "gtr_g4"), ("rbcl10", "gtr_i_g4"), ("random", "jc"), ("random", "gtr_g8")])
def test_transient_partials_give_the_same_bits_and_leave_chain_interiors_unwritten(engine, oracle, source, model):
    """sb200SetTransientPartials: only the buffers a later launch reads (the last node of every chain) are written.  Same
    log-likelihood and cumulative scalers bit for bit, same root partials; and the switch goes back."""
    from strom_b200.engine import api
    E = api()
    pr = load_golden(source) if source != "random" else random_problem(40, 500, seed=12)
    m = named_model(model)
    n, P = pr["data"].shape
    Cc = len(m["rates"])
    xe = Extra(engine)
    ll0, inst = full_eval(engine, pr, m, keep=True)
    root = int(pr["parent"])
    want_root, want_scale = xe.partials(inst, root, P, Cc), xe.scale(inst, 0, P)
    first_op_dest = int(pr["ops"][0][0])
    assert E.sb200SetTransientPartials(inst, 1) == 0
    # poison every buffer through a list that writes them with other edge lengths, then evaluate transiently
    ll1 = engine.calc_log_likelihood(inst, m, pr["ops"], pr["pmatrix_index"], pr["edge_lengths"], root, 0)
    assert ll1 == ll0
    assert np.array_equal(xe.partials(inst, root, P, Cc), want_root) and np.array_equal(xe.scale(inst, 0, P), want_scale)
    alg, sch = C.c_double(), C.c_double()
    E.sb200LastTraffic(inst, C.byref(alg), C.byref(sch))
    assert E.sb200SetTransientPartials(inst, 0) == 0
    ll2 = engine.calc_log_likelihood(inst, m, pr["ops"], pr["pmatrix_index"], pr["edge_lengths"] * 1.0, root, 0)
    sch_all = C.c_double()
    E.sb200LastTraffic(inst, C.byref(alg), C.byref(sch_all))
    assert ll2 == ll0 and sch.value < sch_all.value
    assert rel_err(ll0, full_eval(oracle, pr, m)) < LOGL_RTOL
    engine.finalize(inst)
