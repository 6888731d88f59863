"""BASELINE.json's full-size configuration (configs[3]: 1000 taxa x 1,000,000 site patterns, GTR+G4, FP64: 128 GB
of partials on one B200) checked through size-independent properties, through the C ABI (include/strom_b200.h):

* pattern weights enter the log-likelihood linearly (reference src/likelihood.hpp:258-260, 429-442): an alignment
  made of 64 copies of a 15,625-pattern block, each pattern with its own weight, must give what the CPU oracle
  gives for the block alone with the 64 weights of every pattern added up;
* doubling every weight doubles logL exactly (a power of two: bit for bit);
* the pattern axis shards without any exchange but one scalar: the whole equals the sum of
  8 contiguous slices evaluated on their own instances;
* the per-pattern cumulative log-scalers of copy c equal those of copy 0 (same columns, same arithmetic).

The oracle finishes the block in about a second; nothing here reads /root/reference."""
import ctypes as C

import numpy as np
import pytest

from helpers import named_model, rel_err

pytestmark = pytest.mark.gpu

NTAXA, NPAT, COPIES = 1000, 1_000_000, 64
BLOCK = NPAT // COPIES


@pytest.fixture(scope="module")
def problem():
    import torch
    from strom_b200 import synth
    free, _ = torch.cuda.mem_get_info(0)
    if free < 150 * (1 << 30):
        pytest.skip("needs 150 GB of free HBM (one B200)")
    tree = synth.RandomTree(NTAXA, seed=1, mean_edge=0.05)
    model = named_model("gtr_g4")
    block = synth.simulate_alignment(tree, model, 0, BLOCK, seed=3, device="cuda:0", chunk=BLOCK).cpu().numpy()
    rng = np.random.default_rng(5)
    block[rng.random(block.shape) < 0.01] = 4                           # some missing data
    weights = rng.integers(1, 5, size=NPAT).astype(np.float64)
    return dict(tree=tree, model=model, block=block, weights=weights)


def _oracle_block_logl(oracle, pr, weights_block):
    oracle.lib.oracleSetThreads.argtypes = [C.c_int]
    oracle.lib.oracleSetThreads(8)
    t = pr["tree"]
    inst = oracle.create(NTAXA, BLOCK, 4)
    oracle.set_data(inst, pr["block"].astype(np.int32), weights_block)
    ll = oracle.calc_log_likelihood(inst, pr["model"], t.ops, t.pmatrix_index, t.edge_lengths, t.parent_buffer, 0)
    oracle.finalize(inst)
    oracle.lib.oracleSetThreads(1)
    return ll


def test_million_patterns_against_the_oracle_by_weight_linearity(oracle, problem):
    from strom_b200.engine import EvalArgs, ShardedLikelihood, api
    pr, t = problem, problem["tree"]
    tips = np.tile(pr["block"], (1, COPIES))
    L = ShardedLikelihood(tips, pr["weights"], 4, device=0)
    ea = EvalArgs(pr["model"], t.ops, t.pmatrix_index, t.edge_lengths, t.parent_buffer, 0)
    whole = L.calc_log_likelihood(ea)
    want = _oracle_block_logl(oracle, pr, pr["weights"].reshape(COPIES, BLOCK).sum(axis=0))
    assert rel_err(whole, want) < 1e-11, (whole, want)

    # cumulative log-scalers: every copy of the block repeats copy 0's
    E = api()
    scal = np.zeros(NPAT)
    E.lib.sb200GetScaleFactors.restype = C.c_int
    E.lib.sb200GetScaleFactors.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
    assert E.lib.sb200GetScaleFactors(L.inst, 0, scal.ctypes.data_as(C.POINTER(C.c_double))) == 0
    scal = scal.reshape(COPIES, BLOCK)
    assert np.array_equal(scal, np.broadcast_to(scal[0], scal.shape))
    assert np.all(scal[0] < 0.0)                                        # 998 rescalings per pattern did happen

    # doubling the weights doubles logL bit for bit
    w2 = np.ascontiguousarray(2.0 * pr["weights"])
    E.check(L.B.beagleSetPatternWeights(L.inst, w2.ctypes.data_as(C.POINTER(C.c_double))), "set pattern weights")
    assert L.calc_log_likelihood(ea) == 2.0 * whole
    L.close()
    problem["whole"] = whole


def test_million_patterns_equal_the_sum_of_eight_slices(problem):
    from strom_b200.engine import EvalArgs, ShardedLikelihood
    from strom_b200.sharding import shard_bounds
    pr, t = problem, problem["tree"]
    if "whole" not in pr:
        pytest.skip("needs the whole-alignment value of the previous test")
    ea = EvalArgs(pr["model"], t.ops, t.pmatrix_index, t.edge_lengths, t.parent_buffer, 0)
    tips = np.tile(pr["block"], (1, COPIES))
    total = 0.0
    for r in range(8):
        lo, hi = shard_bounds(NPAT, r, 8)
        Ls = ShardedLikelihood(np.ascontiguousarray(tips[:, lo:hi]), pr["weights"][lo:hi], 4, device=0)
        total += Ls.calc_log_likelihood(ea)
        Ls.close()
    assert rel_err(total, pr["whole"]) < 1e-13, (total, pr["whole"])
