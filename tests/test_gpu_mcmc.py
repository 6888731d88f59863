"""Trace parity: the same host code and seed on the CUDA engine and on the reference's CPU call path give
the same MCMC trace up to tolerance (north-star).  A logL difference of 1e-13 can in principle flip one
accept decision; the test reports the first divergence instead of demanding bit-equality."""
import numpy as np
import pytest

from helpers import load_golden

pytestmark = pytest.mark.gpu


def _run(api, key, **kw):
    fx = load_golden(key)
    cols = np.repeat(np.arange(fx["data"].shape[1]), fx["counts"].astype(int))
    h = api.data_from_codes(fx["data"].astype(np.int32)[:, cols], compress=True)
    tr, sec = api.mcmc(h, str(fx["newick"]), **kw)
    api.data_free(h)
    return tr, sec


@pytest.mark.parametrize("key,kw", [("rbcl10", dict(niter=300, burnin=200, nchains=4, samplefreq=10, seed=13579)),
                                    ("rbcL", dict(niter=200, burnin=100, nchains=2, samplefreq=5, seed=7)),
                                    ("rbcl10", dict(niter=100, burnin=50, nchains=1, ncateg=1, samplefreq=5, seed=3))])
def test_trace_matches_cpu_call_path(key, kw):
    from strom_b200.build import HOST_ORACLE_LIB
    from strom_b200.host import HostAPI, host
    gpu, _ = _run(host(), key, **kw)
    cpu, _ = _run(HostAPI(HOST_ORACLE_LIB), key, **kw)
    assert gpu.shape == cpu.shape
    rel = np.abs(gpu - cpu) / np.maximum(np.abs(cpu), 1e-300)
    bad = np.argwhere(rel > 1e-9)
    assert bad.size == 0, "first divergence at sample row %d column %d: gpu %r cpu %r" % (
        bad[0][0], bad[0][1], gpu[bad[0][0], bad[0][1]], cpu[bad[0][0], bad[0][1]])


def test_chain_per_gpu_placement_runs():
    # config C5 shape: one Likelihood (engine instance) per heated chain; here all pinned to device 0
    from strom_b200.host import host
    tr, sec = _run(host(), "rbcl738", niter=10, burnin=10, nchains=2, samplefreq=5, seed=1, devices=(0, 0))
    assert tr.shape[0] == 3 and np.all(np.isfinite(tr)) and tr[-1, 1] > tr[0, 1]


def test_parallel_chains_match_cpu_call_path_and_turn_order():
    """Config C5's driver: heated chains stepped side by side on host threads, one engine instance per chain, spread
    over every visible GPU (all on device 0 when there is one).  Same trace as stepping them in turn, and as the
    CPU call path, whatever the thread timing."""
    import torch
    from strom_b200.build import HOST_ORACLE_LIB
    from strom_b200.host import HostAPI, host
    devices = tuple(range(torch.cuda.device_count()))
    kw = dict(niter=150, burnin=100, nchains=4, samplefreq=10, seed=99)
    side, _ = _run(host(), "rbcl10", parallel_chains=True, devices=devices, **kw)
    turn, _ = _run(host(), "rbcl10", per_chain_streams=True, devices=devices, **kw)
    cpu, _ = _run(HostAPI(HOST_ORACLE_LIB), "rbcl10", parallel_chains=True, **kw)
    assert np.array_equal(side, turn)
    rel = np.abs(side - cpu) / np.maximum(np.abs(cpu), 1e-300)
    assert rel.max() <= 1e-9


@pytest.mark.parametrize("key,extra", [("rbcl10", {}), ("rbcl738", dict(niter=12, burnin=8, samplefreq=4)), ("rbcl10", dict(incremental=True))])
def test_lockstep_chains_overlap_on_one_gpu_and_keep_the_trace(key, extra):
    """One host thread begins every update on all chains (proposal, asynchronous engine call) before finishing it on
    any: the chains' evaluations are in flight together on the GPU they share.  Bit for bit the trace of the chains
    stepped in turn with a random stream each, and the CPU call path's to tolerance."""
    from strom_b200.build import HOST_ORACLE_LIB
    from strom_b200.host import HostAPI, host
    kw = dict(niter=150, burnin=100, nchains=4, samplefreq=10, seed=99)
    kw.update(extra)
    lock, _ = _run(host(), key, lockstep_chains=True, **kw)
    turn, _ = _run(host(), key, per_chain_streams=True, **kw)
    assert np.array_equal(lock, turn)
    if key != "rbcl738":
        kw.pop("incremental", None)
        cpu, _ = _run(HostAPI(HOST_ORACLE_LIB), key, lockstep_chains=True, **kw)
        rel = np.abs(lock - cpu) / np.maximum(np.abs(cpu), 1e-300)
        assert rel.max() <= 1e-9


@pytest.mark.parametrize("key,kw", [("rbcl10", dict(niter=300, burnin=200, nchains=2, samplefreq=10, seed=13579)),
                                    ("rbcL", dict(niter=150, burnin=100, nchains=1, samplefreq=5, seed=7))])
def test_incremental_trace_matches_cpu_call_path(key, kw):
    """The whole MCMC with incremental likelihoods on the engine against the CPU call path that recomputes
    everything: same trace to the north-star tolerance."""
    from strom_b200.build import HOST_ORACLE_LIB
    from strom_b200.host import HostAPI, host
    gpu, _ = _run(host(), key, incremental=True, **kw)
    cpu, _ = _run(HostAPI(HOST_ORACLE_LIB), key, **kw)
    rel = np.abs(gpu - cpu) / np.maximum(np.abs(cpu), 1e-300)
    assert gpu.shape == cpu.shape and rel.max() <= 1e-9


def test_rbcl738_trace_matches_cpu_call_path():
    """BASELINE.json configs[2]: the real 738-taxon data set, two heated chains with swaps, all five updaters (reference call
    chain src/updater.hpp:184-225 -> src/likelihood.hpp:390-448): CUDA engine against the CPU call path, same seed."""
    from strom_b200.build import HOST_ORACLE_LIB
    from strom_b200.host import HostAPI, host
    kw = dict(niter=24, burnin=16, nchains=2, samplefreq=4, seed=2024, gammashape=0.5)
    gpu, _ = _run(host(), "rbcl738", **kw)
    cpu, _ = _run(HostAPI(HOST_ORACLE_LIB), "rbcl738", **kw)
    assert gpu.shape == cpu.shape and gpu.shape[0] == 7
    rel = np.abs(gpu - cpu) / np.maximum(np.abs(cpu), 1e-300)
    bad = np.argwhere(rel > 1e-9)
    assert bad.size == 0, "first divergence at sample row %d column %d: gpu %r cpu %r" % (
        bad[0][0], bad[0][1], gpu[bad[0][0], bad[0][1]], cpu[bad[0][0], bad[0][1]])


@pytest.mark.parametrize("key,kw", [("rbcl10", dict(niter=200, burnin=100, nchains=2, samplefreq=10, seed=5)),
                                    ("rbcl738", dict(niter=10, burnin=6, nchains=1, samplefreq=2, seed=11))])
def test_invariable_sites_mcmc_trace_matches_cpu_call_path(key, kw):
    """GTR+I+G4 (BASELINE.json configs[2]; +I is a north-star extension, the reference has no such parameter): an MCMC with a
    fixed proportion of invariable sites -- five categories on the engine -- against the CPU call path."""
    from strom_b200.build import HOST_ORACLE_LIB
    from strom_b200.host import HostAPI, host
    gpu, _ = _run(host(), key, pinvar=0.2, **kw)
    cpu, _ = _run(HostAPI(HOST_ORACLE_LIB), key, pinvar=0.2, **kw)
    plain, _ = _run(host(), key, **kw)
    assert gpu.shape == cpu.shape
    rel = np.abs(gpu - cpu) / np.maximum(np.abs(cpu), 1e-300)
    assert rel.max() <= 1e-9
    assert not np.allclose(gpu[:, 1], plain[:, 1])                    # (+I really is another model)


def test_eight_heated_chains_on_rbcl738_side_by_side_in_turn_and_cpu():
    """BASELINE.json configs[4] (reference loop src/strom.hpp:316-402): 8 heated chains on rbcl738.  Stepped side by side on host
    threads (one engine instance per chain, spread over the visible GPUs), in lockstep with their evaluations batched into one
    set of launches, and in turn: bit for bit the same trace; and the CPU call path's to tolerance."""
    import torch
    from strom_b200.build import HOST_ORACLE_LIB
    from strom_b200.host import HostAPI, host
    devices = tuple(range(torch.cuda.device_count()))
    kw = dict(niter=6, burnin=4, nchains=8, samplefreq=2, seed=31, gammashape=0.5)
    side, _ = _run(host(), "rbcl738", parallel_chains=True, devices=devices, **kw)
    turn, _ = _run(host(), "rbcl738", per_chain_streams=True, devices=devices, **kw)
    lock, _ = _run(host(), "rbcl738", lockstep_chains=True, **kw)
    lock_unbatched, _ = _run(host(), "rbcl738", lockstep_chains=True, batch_evaluations=False, **kw)
    assert np.array_equal(side, turn) and np.array_equal(lock, turn) and np.array_equal(lock_unbatched, turn)
    cpu, _ = _run(HostAPI(HOST_ORACLE_LIB), "rbcl738", per_chain_streams=True, **kw)
    rel = np.abs(turn - cpu) / np.maximum(np.abs(cpu), 1e-300)
    assert turn.shape == cpu.shape and rel.max() <= 1e-9
