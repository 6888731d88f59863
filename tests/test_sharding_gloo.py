"""The N > 1 path on the CPU: two gloo ranks each evaluate their pattern slice (here with the CPU
oracle standing in for the per-GPU engine instance) and all-reduce one double; the sum must equal
the single-rank evaluation.  Exercises strom_b200.sharding exactly as ShardedLikelihood uses it."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import full_eval, load_golden, named_model, random_problem, rel_err


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, key, model_name, out_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    from strom_b200.beagle import BeagleAPI
    from strom_b200.build import ORACLE_LIB
    from strom_b200.sharding import shard_bounds, sum_over_ranks
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fx = load_golden(key) if key != "random" else random_problem(25, 101, 9)
    lo, hi = shard_bounds(fx["data"].shape[1], rank, world)
    part = dict(fx)
    part["data"] = fx["data"][:, lo:hi]
    part["counts"] = fx["counts"][lo:hi]
    ll = full_eval(BeagleAPI(ORACLE_LIB), part, named_model(model_name))
    t = torch.tensor([ll], dtype=torch.float64)
    sum_over_ranks(t)
    if rank == 0:
        np.save(out_path, t.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("key,model_name,world", [("rbcl10", "gtr_g4", 2), ("rbcl738", "gtr_i_g4", 2), ("random", "jc", 3)])
def test_two_rank_pattern_shards_sum_to_single_rank(tmp_path, oracle, key, model_name, world):
    out = str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(world, _free_port(), key, model_name, out), nprocs=world, join=True)
    total = float(np.load(out)[0])
    fx = load_golden(key) if key != "random" else random_problem(25, 101, 9)
    single = full_eval(oracle, fx, named_model(model_name))
    assert rel_err(total, single) < 1e-13


def test_shard_bounds_cover_every_pattern_once():
    from strom_b200.sharding import shard_bounds
    for P in (1, 7, 1235, 1000000):
        for R in (1, 2, 3, 8):
            edges = [shard_bounds(P, r, R) for r in range(R)]
            assert edges[0][0] == 0 and edges[-1][1] == P
            assert all(edges[i][1] == edges[i + 1][0] for i in range(R - 1))


def _mc3_worker(rank, world, port, out_path, kw):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    from strom_b200 import mc3
    from strom_b200.build import HOST_ORACLE_LIB
    from strom_b200.host import HostAPI
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H = HostAPI(HOST_ORACLE_LIB)
    fx = load_golden("rbcl10")
    cols = np.repeat(np.arange(fx["data"].shape[1]), fx["counts"].astype(int))
    h = H.data_from_codes(fx["data"].astype(np.int32)[:, cols], compress=True)
    rows, sec, its = mc3.run_one_chain_per_rank(H, h, str(fx["newick"]), **kw)
    trace = mc3.gather_trace(rows)
    if rank == 0:
        np.save(out_path, trace)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_one_chain_per_process_mc3_gives_the_side_by_side_trace(tmp_path, world):
    """BASELINE.json configs[4], the multi-process form: rank r steps heated chain r and the ranks exchange swap offers with a
    collective (gloo here, NCCL over NVLink on GPUs).  Same seed: bit for bit the trace of the single-process driver that steps the
    chains side by side on host threads (which tests/test_mcmc_cpu.py ties to the in-turn order)."""
    from strom_b200.build import build_host_oracle
    from strom_b200.host import HostAPI
    kw = dict(seed=4242, niter=60, burnin=40, samplefreq=5, heatfactor=0.5, ncateg=4, gammashape=0.5)
    out = str(tmp_path / "trace.npy")
    mp.spawn(_mc3_worker, args=(world, _free_port(), out, kw), nprocs=world, join=True)
    got = np.load(out)
    H = HostAPI(build_host_oracle())
    fx = load_golden("rbcl10")
    cols = np.repeat(np.arange(fx["data"].shape[1]), fx["counts"].astype(int))
    h = H.data_from_codes(fx["data"].astype(np.int32)[:, cols], compress=True)
    want, _ = H.mcmc(h, str(fx["newick"]), nchains=world, parallel_chains=True, **kw)
    H.data_free(h)
    assert got.shape == want.shape and got.shape[0] == 13
    assert np.array_equal(got, want)
