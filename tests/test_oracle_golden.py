"""Pins the CPU oracle (oracle/beagle_cpu.c + oracle/strom_host_ref.py) to every golden value the
reference holds for this path, and to the committed fixtures."""
import os

import numpy as np
import pytest

from helpers import MODEL_NAMES, full_eval, load_golden, named_model, rel_err
from oracle import strom_host_ref as H

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present")


def test_golden_rbcl738_gtr_g4_paup(oracle):
    # paup.nex:8 parameters, paup.nex:12 "lnL should be -144730.8"
    fx = load_golden("rbcl738")
    ll = full_eval(oracle, fx, named_model("gtr_g4"))
    assert abs(ll - (-144730.8)) < 0.05
    assert rel_err(ll, float(fx["lnL_gtr_g4"])) < 1e-13


def test_golden_rbcL_jc_paup(oracle):
    # rbcLjc.tre:57 "Score of best tree(s) found = 286.9238" (JC69, 23 patterns at :35)
    fx = load_golden("rbcL")
    assert fx["data"].shape == (14, 23)
    ll = full_eval(oracle, fx, named_model("jc"))
    assert abs(-ll - 286.9238) < 5e-5


@pytest.mark.parametrize("key", ["rbcL", "rbcl10", "rbcl738"])
@pytest.mark.parametrize("model", MODEL_NAMES)
def test_oracle_reproduces_fixture_values(oracle, key, model):
    fx = load_golden(key)
    ll = full_eval(oracle, fx, named_model(model))
    assert rel_err(ll, float(fx["lnL_" + model])) < 1e-13


def test_structural_golden_rbcl10_operations():
    # the operation list strom builds for rbcl10.tre
    fx = load_golden("rbcl10")
    expect = [(10, 1, -1, 8, 8, 9, 9), (11, 2, -1, 7, 7, 10, 10), (12, 3, -1, 6, 6, 11, 11), (13, 4, -1, 4, 4, 5, 5),
              (14, 5, -1, 13, 13, 12, 12), (15, 6, -1, 1, 1, 2, 2), (16, 7, -1, 15, 15, 14, 14), (17, 8, -1, 16, 16, 3, 3)]
    assert [tuple(r) for r in fx["ops"].tolist()] == expect
    assert fx["pmatrix_index"].tolist() == [9, 8, 10, 7, 11, 6, 5, 4, 12, 13, 2, 1, 14, 15, 3, 16, 0]
    assert int(fx["parent"]) == 17
    # edge lengths went through std::stof (float32 rounding), tree_manip.hpp:299
    assert fx["edge_lengths"][-1] == float(np.float32(0.594710))
    assert fx["edge_lengths"][-1] != 0.594710


def test_structural_golden_other_trees():
    a = load_golden("rbcL")
    assert a["ops"].shape == (12, 7) and a["pmatrix_index"].size == 25
    assert tuple(a["ops"][-1]) == (25, 12, -1, 24, 24, 13, 13)
    b = load_golden("rbcl738")
    assert b["ops"].shape == (736, 7) and b["pmatrix_index"].size == 1473
    assert tuple(b["ops"][-1]) == (1473, 736, -1, 1472, 1472, 758, 758)
    assert int(b["pmatrix_index"].max()) == 1472 and int(b["pmatrix_index"][-1]) == 0
    assert b["data"].shape == (738, 1235)


@needs_ref
@pytest.mark.parametrize("key,nex,tre", [("rbcL", "rbcL.nex", "rbcLjc.tre"), ("rbcl10", "rbcl10.nex", "rbcl10.tre"),
                                         ("rbcl738", "rbcl738.nex", "rbcl738nj.tre")])
def test_fixtures_match_reference_files(key, nex, tre):
    fx = load_golden(key)
    names, data, counts, nsites = H.read_nexus_data(os.path.join(REF, nex))
    tree = H.build_tree(H.read_nexus_trees(os.path.join(REF, tre))[0])
    ops, pidx, elen = H.define_operations(tree)
    assert np.array_equal(np.minimum(data, 255).astype(np.uint8), fx["data"])
    assert np.array_equal(counts, fx["counts"]) and counts.sum() == nsites
    assert np.array_equal(ops, fx["ops"]) and np.array_equal(pidx, fx["pmatrix_index"])
    assert np.array_equal(elen, fx["edge_lengths"])


def test_jc_transition_probability_closed_form(oracle):
    # JC69: P_ii(t) = 1/4 + 3/4 e^{-4t/3}; two taxa joined at a third gives a hand-computable likelihood
    from helpers import Extra, make_model
    m = make_model([0.25] * 4, [1.0] * 6, [1.0], [1.0])
    inst = oracle.create(3, 1, 1)
    oracle.set_data(inst, np.array([[0], [0], [1]], dtype=np.int32), np.array([1.0]))
    ops = np.array([[3, 1, -1, 1, 1, 2, 2]], dtype=np.int32)
    t = np.array([0.1, 0.2, 0.3])
    ll = oracle.calc_log_likelihood(inst, m, ops, np.array([1, 2, 0], dtype=np.int32), t, 3, 0)
    P = lambda tt: 0.25 + 0.75 * np.exp(-4 * tt / 3) * np.eye(4) - 0.25 * np.exp(-4 * tt / 3) * (1 - np.eye(4)) * 0 + \
        (0.25 - 0.25 * np.exp(-4 * tt / 3)) * (1 - np.eye(4)) - 0.25 * (1 - np.eye(4))
    # brute force over the internal state
    def pjc(tt):
        e = np.exp(-4.0 * tt / 3.0)
        return np.where(np.eye(4) > 0, 0.25 + 0.75 * e, 0.25 - 0.25 * e)
    like = sum(0.25 * pjc(0.3)[x, 0] * pjc(0.1)[x, 0] * pjc(0.2)[x, 1] for x in range(4))
    assert rel_err(ll, np.log(like)) < 1e-13
    got = Extra(oracle).matrix(inst, 1, 1)[0]
    assert np.allclose(got, pjc(0.1), rtol=1e-13, atol=1e-15)
    oracle.finalize(inst)


def test_missing_data_equals_marginalisation(oracle):
    # a tip coded 4 (missing) must give the sum of the likelihoods with that tip set to each state
    from helpers import make_model
    m = named_model("gtr_g4")
    ops = np.array([[4, 1, -1, 1, 1, 2, 2], [5, 2, -1, 4, 4, 3, 3]], dtype=np.int32)
    pidx = np.array([1, 2, 4, 3, 0], dtype=np.int32)
    t = np.array([0.1, 0.2, 0.05, 0.3, 0.15])
    def lik(states):
        inst = oracle.create(4, 1, 4)
        oracle.set_data(inst, np.array(states, dtype=np.int32).reshape(4, 1), np.array([1.0]))
        ll = oracle.calc_log_likelihood(inst, m, ops, pidx, t, 5, 0)
        oracle.finalize(inst)
        return np.exp(ll)
    total = sum(lik([0, s, 2, 3]) for s in range(4))
    assert rel_err(lik([0, 4, 2, 3]), total) < 1e-12
    total0 = sum(lik([s, 1, 2, 3]) for s in range(4))
    assert rel_err(lik([7, 1, 2, 3]), total0) < 1e-12      # any code >= 4 is missing, also at the root tip


@pytest.mark.parametrize("ntaxa,seed,model", [(4, 3, "gtr_g4"), (6, 5, "gtr_g4"), (7, 8, "gtr_i_g4"), (5, 9, "jc")])
def test_brute_force_enumeration_of_internal_states(oracle, ntaxa, seed, model):
    """Independent of pruning, rescaling and the eigen path: the site likelihood as the plain sum over every assignment of
    states to the internal nodes, with P(t r_k) = expm(Q t r_k) from the rate matrix itself (scipy), against the oracle's
    logL on a random tree with missing data."""
    from scipy.linalg import expm
    from helpers import JC, PAUP_PI, PAUP_R, random_problem
    fx = random_problem(ntaxa, 12, seed, missing_frac=0.1, mean_edge=0.2)
    m = named_model(model)
    pi, r = (np.array(JC[0]), JC[1]) if model == "jc" else (np.array(PAUP_PI), PAUP_R)
    R = np.zeros((4, 4))
    R[np.triu_indices(4, 1)] = r
    R = R + R.T
    Q = R * pi[None, :]
    np.fill_diagonal(Q, 0.0)
    np.fill_diagonal(Q, -Q.sum(axis=1))
    Q /= -(pi * np.diag(Q)).sum()                                   # mean rate 1 (reference src/model.hpp:228-229)
    length_of = {int(mi): float(t) for mi, t in zip(fx["pmatrix_index"], fx["edge_lengths"])}
    P = {mi: [expm(Q * t * rk) for rk in m["rates"]] for mi, t in length_of.items()}
    children = {int(op[0]): (int(op[3]), int(op[5])) for op in fx["ops"]}
    internal = sorted(children)
    top, data = int(fx["parent"]), fx["data"]
    total = 0.0
    for p in range(data.shape[1]):
        site = 0.0
        for k, wk in enumerate(m["probs"]):
            acc = 0.0
            for code in range(4 ** len(internal)):
                state = {nd: (code >> (2 * i)) & 3 for i, nd in enumerate(internal)}
                term = pi[state[top]]
                s0 = int(data[0, p])
                term *= P[0][k][state[top], s0] if s0 < 4 else 1.0       # the root tip hangs from `top` by matrix 0
                for nd, kids in children.items():
                    for c in kids:
                        if c in state:
                            term *= P[c][k][state[nd], state[c]]
                        else:
                            s = int(data[c, p])
                            term *= P[c][k][state[nd], s] if s < 4 else 1.0
                acc += term
            site += wk * acc
        total += fx["counts"][p] * np.log(site)
    inst = oracle.create(ntaxa, data.shape[1], len(m["rates"]))
    oracle.set_data(inst, data.astype(np.int32), fx["counts"])
    ll = oracle.calc_log_likelihood(inst, m, fx["ops"], fx["pmatrix_index"], fx["edge_lengths"], top, 0)
    oracle.finalize(inst)
    assert rel_err(ll, total) < 1e-11


def test_high_precision_replay_of_rbcl10(oracle):
    """How far the oracle's double arithmetic is from the exact value: the whole rbcl10 GTR+G4 evaluation repeated with
    40-digit mpmath numbers (P = expm(Q t r_k), plain pruning without rescaling -- nothing underflows at 40 digits)."""
    mp = pytest.importorskip("mpmath")
    from helpers import PAUP_PI, PAUP_R
    mp.mp.dps = 40
    fx = load_golden("rbcl10")
    m = named_model("gtr_g4")
    pi = [mp.mpf(repr(float(x))) for x in PAUP_PI]
    R = mp.zeros(4)
    it = iter(PAUP_R)
    for i in range(4):
        for j in range(i + 1, 4):
            R[i, j] = R[j, i] = mp.mpf(repr(float(next(it))))
    Q = mp.zeros(4)
    for i in range(4):
        for j in range(4):
            if i != j:
                Q[i, j] = R[i, j] * pi[j]
        Q[i, i] = -sum(Q[i, j] for j in range(4) if j != i)
    scale = -sum(pi[i] * Q[i, i] for i in range(4))
    Q = Q / scale
    rates = [mp.mpf(repr(float(x))) for x in m["rates"]]
    probs = [mp.mpf(repr(float(x))) for x in m["probs"]]
    P = {}
    for mi, t in zip(fx["pmatrix_index"], fx["edge_lengths"]):
        P[int(mi)] = [mp.expm(Q * (mp.mpf(repr(float(t))) * rk)) for rk in rates]
    data, counts, n = fx["data"], fx["counts"], fx["data"].shape[0]
    total = mp.mpf(0)
    for p in range(data.shape[1]):
        site = mp.mpf(0)
        for k in range(4):
            part = {}
            def cond(c, i):                       # conditional likelihood of the subtree below node c given state i above its edge
                if c < n:
                    s = int(data[c, p])
                    return P[c][k][i, s] if s < 4 else mp.mpf(1)
                return sum(P[c][k][i, j] * part[c][j] for j in range(4))
            for op in fx["ops"]:
                d, c1, c2 = int(op[0]), int(op[3]), int(op[5])
                part[d] = [cond(c1, i) * cond(c2, i) for i in range(4)]
            top, s0 = int(fx["parent"]), int(data[0, p])
            site += probs[k] * sum(pi[i] * part[top][i] * (P[0][k][i, s0] if s0 < 4 else 1) for i in range(4))
        total += mp.mpf(repr(float(counts[p]))) * mp.log(site)
    inst = oracle.create(n, data.shape[1], 4)
    oracle.set_data(inst, data.astype(np.int32), counts)
    ll = oracle.calc_log_likelihood(inst, m, fx["ops"], fx["pmatrix_index"], fx["edge_lengths"], int(fx["parent"]), 0)
    oracle.finalize(inst)
    assert abs((mp.mpf(repr(ll)) - total) / total) < mp.mpf("1e-14")                          # observed 1.9e-16
    assert abs((mp.mpf(repr(float(fx["lnL_gtr_g4"]))) - total) / total) < mp.mpf("1e-14")     # the committed fixture value too
