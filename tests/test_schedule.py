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


def _walk_plan(plan, ntaxa, in_list):
    """Replays a plan symbolically: every operand read from memory must be there (a tip, a buffer of an earlier
    launch, or -- incremental plans -- a buffer no operation of the list produces), chains must be contiguous."""
    done_before_launch = {}
    carried = None
    written = set()
    for dest, a, b, sa, sb, s_out, flags, launch in plan.tolist():
        start, end = bool(flags & 2), bool(flags & 4)
        assert (a >= 0) == start                                       # only a chain start reads operand A from memory
        if not start:
            assert carried is not None                                 # ... otherwise A is what the previous step produced
        for operand in ([a] if start else []) + [b]:
            if operand >= ntaxa and operand in in_list:
                assert operand in written and done_before_launch[operand] < launch
        written.add(dest)
        done_before_launch[dest] = launch
        carried = None if end else dest
    return written


@pytest.mark.parametrize("key", ["rbcl10", "rbcL", "rbcl738"])
def test_incremental_plans_keep_every_subtree_scaler(E, key):
    """Planner side of sb200SetSubtreeScalers: slot = buffer for every operation, chain-internal operations store and
    carry on (flag 128), and an internal operand the list does not produce brings its own slot."""
    fx = load_golden(key)
    ops = np.asarray(fx["ops"], dtype=np.int32).reshape(-1, 7)
    ntaxa = fx["data"].shape[0]
    parent_of = {}
    for o in ops:
        parent_of[int(o[3])] = int(o[0])
        parent_of[int(o[5])] = int(o[0])
    # full list
    plan = E.plan(ops, ntaxa, keep_subtree_scalers=True)
    assert len(plan) == len(ops) and _walk_plan(plan, ntaxa, set(ops[:, 0].tolist())) == set(ops[:, 0].tolist())
    for dest, a, b, sa, sb, s_out, flags, launch in plan.tolist():
        assert s_out == dest - ntaxa
        assert bool(flags & 128) != bool(flags & 4)                    # stores-and-continues XOR ends its chain
        if flags & 2 and a >= ntaxa:
            assert sa == a - ntaxa
        if b >= ntaxa:
            assert sb == b - ntaxa
        if b < ntaxa:
            assert sb == -1
    assert sum(1 for r in plan.tolist() if r[6] & 8) == 1              # exactly one chain feeds the cumulative buffer
    # the same list without the switch: only chain ends that are re-read own a slot
    plain = E.plan(ops, ntaxa)
    assert not any(r[6] & 128 for r in plain.tolist())
    assert sum(1 for r in plain.tolist() if r[5] >= 0) < len(ops)
    # a path list: the operations above one tip
    tip = 1 if ntaxa > 3 else 0
    dirty, v = set(), parent_of.get(tip)
    while v is not None:
        dirty.add(v)
        v = parent_of.get(v)
    sub = np.array([o for o in ops if int(o[0]) in dirty], dtype=np.int32)
    path = E.plan(sub, ntaxa, keep_subtree_scalers=True)
    assert len(path) == len(sub) and _walk_plan(path, ntaxa, dirty) == dirty
    outside = 0
    for dest, a, b, sa, sb, s_out, flags, launch in path.tolist():
        assert s_out == dest - ntaxa
        for operand, slot in ((a, sa), (b, sb)):
            if operand >= ntaxa:
                assert slot == operand - ntaxa
                outside += operand not in dirty
    assert outside >= 1 or len(sub) == 1                               # siblings along the path come with their old scalers
    assert len({r[7] for r in path.tolist()}) == 1                     # one path = one chain = one launch


def test_launches_tips_only_flag_and_group_counts(E):
    """Level 0 of a full evaluation reads nothing but tips (its launch may run the kernel variant without partials-operand
    code, with its own group count); an incremental list whose operands are partials of earlier calls does not qualify."""
    fx = load_golden("rbcl738")
    L = E.plan_launches(fx["ops"], 738, groups=44, groups_tips=29)
    assert L.shape == (6, 5) and L[:, 3].sum() == 736
    assert L[0, 2] == 1 and not L[1:, 2].any()                      # only the first launch is tips-only
    assert L[0, 1] == 29 and L[1, 1] == 44                          # the tips-only launch got its own group count
    assert L[-1, 0] == 1 and L[-1, 1] == 1 and L[-1, 4] == L[-1, 3]   # one chain at the top: one group, run in sequence
    assert np.all(L[:, 4] * L[:, 1] >= L[:, 3])                     # the longest group bounds the mean
    same = E.plan_launches(fx["ops"], 738, groups=44)               # 0 = the same count for every launch
    assert same[0, 1] == 44 and same[0, 2] == 1
    tail = E.plan_launches(fx["ops"][-20:], 738, groups=44, groups_tips=29, keep_subtree_scalers=True)
    assert tail[:, 3].sum() == 20 and not tail[:, 2].any()          # operands are buffers an earlier call left: not tips-only
    small = random_problem(4, 3, 1)
    assert E.plan_launches(small["ops"], 4, groups=8, groups_tips=8)[0, 2] == 1
