"""Host-side planner (strom_b200/csrc/schedule.cpp) through sb200DebugSchedule: no GPU needed."""
import numpy as np
import pytest

from helpers import load_golden, random_problem


@pytest.fixture(scope="module")
def E():
    from strom_b200.engine import api
    return api()


def check_plan(ops, ntaxa, levels, nchains, lev, ch):
    n = ops.shape[0]
    producer = {int(op[0]): i for i, op in enumerate(ops)}
    for i, op in enumerate(ops):
        kids = [producer[int(c)] for c in (op[3], op[5]) if int(c) >= ntaxa]
        same_chain = [k for k in kids if ch[k] == ch[i]]
        assert len(same_chain) <= 1                                   # at most one child is carried in registers
        for k in kids:
            if ch[k] == ch[i]:
                assert lev[k] == lev[i]
            else:
                assert lev[k] < lev[i]                                 # memory operands were finished by an earlier launch
    assert lev.max() + 1 == levels and len(set(ch.tolist())) == nchains
    # Strahler bound: a binary tree with n tips needs at most floor(log2(n)) + 1 launches
    assert levels <= int(np.floor(np.log2(ntaxa))) + 1


@pytest.mark.parametrize("key,expect_levels", [("rbcL", None), ("rbcl10", None), ("rbcl738", 6)])
def test_fixture_trees(E, key, expect_levels):
    fx = load_golden(key)
    ntaxa = fx["data"].shape[0]
    levels, nchains, lev, ch = E.schedule(fx["ops"], ntaxa)
    check_plan(fx["ops"], ntaxa, levels, nchains, lev, ch)
    if expect_levels:
        assert levels == expect_levels                               # vs 35 dependency levels of the raw list


@pytest.mark.parametrize("ntaxa,seed", [(3, 1), (4, 2), (9, 3), (100, 4), (1000, 5)])
def test_random_trees(E, ntaxa, seed):
    fx = random_problem(ntaxa, 4, seed)
    levels, nchains, lev, ch = E.schedule(fx["ops"], ntaxa)
    check_plan(fx["ops"], ntaxa, levels, nchains, lev, ch)


def test_caterpillar_is_one_chain(E):
    # ((((1,2),3),4),5)... : every operation carries its predecessor -> one chain, one launch
    n = 40
    ops = []
    prev = 1
    for k in range(n - 2):
        dest = n + k
        ops.append([dest, k + 1, -1, prev, prev, k + 2, k + 2])
        prev = dest
    ops = np.array(ops, dtype=np.int32)
    levels, nchains, lev, ch = E.schedule(ops, n)
    assert levels == 1 and nchains == 1


def test_balanced_tree_levels(E):
    # perfectly balanced 16-tip tree: 8 cherries, then pairs of equal level meet: 4 launches
    n = 16
    ops, nxt, layer = [], n, list(range(n))
    while len(layer) > 1:
        new = []
        for a, b in zip(layer[::2], layer[1::2]):
            ops.append([nxt, nxt - n + 1, -1, a, a, b, b])
            new.append(nxt)
            nxt += 1
        layer = new
    ops = np.array(ops, dtype=np.int32)
    levels, nchains, lev, ch = E.schedule(ops, n)
    assert levels == 4 and nchains == len(ops)


def test_non_forest_list_is_sequential(E):
    fx = random_problem(8, 4, 3)
    ops = np.concatenate([fx["ops"][:2], fx["ops"]])                  # a destination written twice
    levels, nchains, lev, ch = E.schedule(ops, 8)
    assert levels == len(ops) and nchains == len(ops)


def test_out_of_range_indices_are_rejected(E):
    from strom_b200.beagle import BeagleError
    fx = random_problem(8, 4, 3)
    bad = fx["ops"].copy()
    bad[0, 3] = 10_000
    with pytest.raises(BeagleError):
        E.schedule(bad, 8)
    bad = fx["ops"].copy()
    bad[0, 2] = 1                                                      # scale-read is not supported
    with pytest.raises(BeagleError):
        E.schedule(bad, 8)
