"""Metropolis-coupled MCMC with ONE heated chain per process / GPU (BASELINE.json configs[4]; reference loop
src/strom.hpp:316-402, where one thread steps the chains in turn and a swap is an exchange of scalars on the host).

Here rank r of a ``torch.distributed`` job owns chain r: its tree, model and engine instance on its GPU, and its own random stream.
The swap schedule -- which two chains, and the deviate of the acceptance test -- depends only on the seed, so every rank computes
it for itself.  After every iteration the ranks all-gather one 16-double vector (NCCL over NVLink / NVSwitch, or gloo on CPUs): the
two chains of that iteration's proposal put their log kernel, heating power, index and tuning parameters in it, both take the same
decision, and an accepted swap is each of them adopting the other's heat.  Same chain set-up, streams, schedule and decisions as
the side-by-side driver of strom_b200/host/mcmc.hpp, so the trace is the same bit for bit (tests/test_sharding_gloo.py)."""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)
OFFER = 16          # doubles per chain in the exchange


def _bind(lib):
    if getattr(lib, "_mc3_bound", False):
        return
    lib.strom_mc3_chain_create.argtypes = [C.c_int, C.c_char_p, C.c_uint, C.c_uint, C.c_double, C.c_uint, C.c_double, C.c_double, _DP, _DP,
                                           C.c_int, C.c_uint, C.c_char_p, C.c_int]
    lib.strom_mc3_chain_step.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int]
    lib.strom_mc3_chain_stop_tuning.argtypes = [C.c_int]
    lib.strom_mc3_chain_offer.argtypes = [C.c_int, _DP, C.c_int, C.c_char_p, C.c_int]
    lib.strom_mc3_chain_take.argtypes = [C.c_int, _DP]
    lib.strom_mc3_chain_heat.argtypes = [C.c_int]
    lib.strom_mc3_chain_heat.restype = C.c_double
    lib.strom_mc3_chain_row.argtypes = [C.c_int, C.c_uint, _DP, C.c_char_p, C.c_int]
    lib.strom_mc3_chain_free.argtypes = [C.c_int]
    lib.strom_mc3_schedule.argtypes = [C.c_uint, C.c_uint, C.c_uint, _IP, _IP, _DP, C.c_uint]
    lib._mc3_bound = True


def run_one_chain_per_rank(H, data_handle: int, newick: str, *, seed=13579, niter=100, burnin=100, samplefreq=1, heatfactor=0.5, ncateg=4,
                           gammashape=0.5, pinvar=0.0, freqs=(0.25,) * 4, rmat=(1.0 / 6,) * 6, device=0, tensor_device="cpu", flags=0):
    """Runs chain ``rank`` of a ``world_size``-chain MC3 in this process (``torch.distributed`` must be initialised).
    Returns (rows sampled by THIS rank while its chain was the cold one [n][15], seconds, chain-iterations of the whole job)."""
    import torch
    import torch.distributed as dist
    lib = H.lib
    _bind(lib)
    rank, world = dist.get_rank(), dist.get_world_size()
    err = C.create_string_buffer(1024)

    def chk(rc):
        if rc < 0:
            raise RuntimeError(err.value.decode(errors="replace"))
        return rc
    f = np.ascontiguousarray(freqs, dtype=np.float64)
    r = np.ascontiguousarray(rmat, dtype=np.float64)
    h = chk(lib.strom_mc3_chain_create(data_handle, newick.encode(), seed, rank, heatfactor, ncateg, gammashape, pinvar, f.ctypes.data_as(_DP),
                                       r.ctypes.data_as(_DP), device, flags, err, 1024))
    total = burnin + niter
    ci, cj, logu = np.zeros(total, dtype=np.int32), np.zeros(total, dtype=np.int32), np.zeros(total)
    lib.strom_mc3_schedule(seed, world, total, ci.ctypes.data_as(_IP), cj.ctypes.data_as(_IP), logu.ctypes.data_as(_DP), flags)
    mine = torch.zeros(OFFER, dtype=torch.float64, device=tensor_device)
    everyone = torch.zeros(world * OFFER, dtype=torch.float64, device=tensor_device)
    offer = np.zeros(OFFER)
    rows = []
    row = np.zeros(15)
    if lib.strom_mc3_chain_heat(h) == 1.0:
        chk(lib.strom_mc3_chain_row(h, 0, row.ctypes.data_as(_DP), err, 1024))
        rows.append(row.copy())
    dist.barrier()
    t0 = time.perf_counter()

    def propose(t):
        if world < 2:
            return
        i, j = int(ci[t]), int(cj[t])
        offer[:] = 0.0
        if rank in (i, j):
            chk(lib.strom_mc3_chain_offer(h, offer.ctypes.data_as(_DP), OFFER, err, 1024))
        mine.copy_(torch.from_numpy(offer))
        dist.all_gather_into_tensor(everyone, mine)
        if rank in (i, j):
            both = everyone.cpu().numpy().reshape(world, OFFER)
            a, b = both[i], both[j]
            log_r = (a[1] - b[1]) * (b[0] - a[0])
            if logu[t] < log_r:
                other = np.ascontiguousarray(b if rank == i else a)
                lib.strom_mc3_chain_take(h, other.ctypes.data_as(_DP))
    for it in range(1, burnin + 1):
        chk(lib.strom_mc3_chain_step(h, it, err, 1024))
        propose(it - 1)
    lib.strom_mc3_chain_stop_tuning(h)
    for it in range(1, niter + 1):
        chk(lib.strom_mc3_chain_step(h, it, err, 1024))
        if it % samplefreq == 0 and lib.strom_mc3_chain_heat(h) == 1.0:
            chk(lib.strom_mc3_chain_row(h, it, row.ctypes.data_as(_DP), err, 1024))
            rows.append(row.copy())
        propose(burnin + it - 1)
    dist.barrier()
    seconds = time.perf_counter() - t0
    lib.strom_mc3_chain_free(h)
    return (np.array(rows) if rows else np.zeros((0, 15))), seconds, world * total


def gather_trace(rows: np.ndarray, tensor_device="cpu"):
    """All ranks' sampled rows put together in iteration order (every iteration is sampled by exactly one rank: the cold chain's)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=tensor_device)
    counts = [torch.zeros(1, dtype=torch.int64, device=tensor_device) for _ in range(world)]
    dist.all_gather(counts, n)
    m = int(max(int(c.item()) for c in counts))
    pad = torch.zeros((m, 15), dtype=torch.float64, device=tensor_device)
    if rows.shape[0]:
        pad[:rows.shape[0]] = torch.from_numpy(rows).to(tensor_device)
    allrows = [torch.zeros((m, 15), dtype=torch.float64, device=tensor_device) for _ in range(world)]
    dist.all_gather(allrows, pad)
    out = np.concatenate([a.cpu().numpy()[:int(c.item())] for a, c in zip(allrows, counts)], axis=0)
    return out[np.argsort(out[:, 0], kind="stable")]
