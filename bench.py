#!/usr/bin/env python
"""bench.py -- tree-likelihood throughput on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--patterns P] [--taxa n]

One step = one full Likelihood::calcLogLikelihood (reference src/likelihood.hpp:390-448): all
2n-3 x C transition matrices, all n-2 partials operations with per-node rescaling, the root-edge
reduction (and, for N > 1, the sum of one double per rank over NCCL).

Workload (BASELINE.json configs[3]): synthetic 1000-taxon alignment simulated under GTR+G4 on a
random tree, 1 000 000 site patterns (weight 1 each), patterns sharded over the N ranks: strong
scaling.  The working set (128 GB of FP64 partials at N=1) is far larger than L2, so no explicit
L2 flush is needed between steps.

Rank 0 prints ONE JSON line.  `value` = device-timed whole-job throughput with everything resident
in HBM; `e2e` = the same metric through the public call with host buffers (parameter block H2D +
scalar D2H inside the timed region); `roofline` = algorithmic bytes of the pruning launches
 / their CUDA-event time / measured HBM peak; `cpu_baseline` = the CPU oracle
on this box's host cores on a bounded sample; `rbcl738` = the latency-bound real-data case.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PAUP_PI = [0.309769, 0.163380, 0.121023, 1.0 - 0.309769 - 0.163380 - 0.121023]     # reference paup.nex:8
PAUP_R = [1.06923, 4.34568, 0.45898, 2.00465, 3.85915, 1.0]
PAUP_ALPHA = 0.517281
METRIC = "pattern_x_category_partial_updates_per_sec"
UNIT = "updates/s"


# ------------------------------------------------------------------------------------------------
# host-side model numbers (tiny; stays on the host in the reference too: src/model.hpp:212-313)
# ------------------------------------------------------------------------------------------------
def gtr_gamma_model(freqs, rmat, alpha, ncat):
    from strom_b200 import hostmodel
    evec, ivec, evals = hostmodel.gtr_eigen(freqs, rmat)
    rates, probs = hostmodel.gamma_rates(alpha, ncat)
    return dict(freqs=np.asarray(freqs, dtype=np.float64), evec=evec, ivec=ivec, evals=evals, rates=rates, probs=probs)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ------------------------------------------------------------------------------------------------
def oracle_api(threads: int):
    from strom_b200.beagle import BeagleAPI
    from strom_b200.build import ORACLE_LIB
    api = BeagleAPI(ORACLE_LIB)
    api.lib.oracleSetThreads.argtypes = [C.c_int]
    api.lib.oracleSetThreads(threads)
    return api


def time_oracle(tips_u8, weights, model, tree_ops, pidx, elen, parent, threads, steps, warmup, budget_s=25.0):
    """Evaluations/s of the CPU oracle on the given (sample) problem."""
    api = oracle_api(threads)
    n, P = tips_u8.shape
    inst = api.create(n, P, len(model["rates"]))
    api.set_data(inst, tips_u8.astype(np.int32), weights)
    ll = None
    for _ in range(warmup):
        ll = api.calc_log_likelihood(inst, model, tree_ops, pidx, elen, parent, 0)
    t0 = time.perf_counter()
    done = 0
    while done < steps:
        ll = api.calc_log_likelihood(inst, model, tree_ops, pidx, elen, parent, 0)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    api.finalize(inst)
    return done / dt, done, ll


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
def synthetic_inputs(args, rank, world, device):
    """Tree, model and this rank's pattern slice (uint8 tips on the host)."""
    from strom_b200 import synth
    tree = synth.RandomTree(args.taxa, seed=1, mean_edge=0.05)
    model = gtr_gamma_model(PAUP_PI, PAUP_R, PAUP_ALPHA, 4)
    from strom_b200.sharding import shard_bounds
    lo, hi = shard_bounds(args.patterns, rank, world)
    tips = synth.simulate_alignment(tree, model, lo, hi, seed=3, device=device)
    tips_h = tips.cpu().numpy()
    del tips
    return tree, model, tips_h, lo, hi


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (restated in oracle/; BeagleLib itself is not buildable
    here) on the host cores, bounded sample of the same workload."""
    if rank != 0:
        return None
    threads = host_threads()
    from strom_b200 import synth
    tree = synth.RandomTree(args.taxa, seed=1, mean_edge=0.05)
    model = gtr_gamma_model(PAUP_PI, PAUP_R, PAUP_ALPHA, 4)
    Ps = min(args.patterns, args.cpu_sample)
    tips = synth.simulate_alignment(tree, model, 0, Ps, seed=3, device="cpu", chunk=Ps).numpy()
    w = np.ones(Ps)
    evals_s, done, ll = time_oracle(tips, w, model, tree.ops, tree.pmatrix_index, tree.edge_lengths, tree.parent_buffer,
                                    threads, args.steps, args.warmup, budget_s=120.0)
    upd = (args.taxa - 2) * Ps * 4 * evals_s
    line = {
        "impl": "reference", "metric": METRIC, "value": upd, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": 1000.0 / evals_s, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, sample=Ps),
        "cpu_baseline": {"value": upd, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "first %d of %d patterns, same tree and model; one step = one full evaluation" % (Ps, args.patterns)},
        "e2e": {"value": upd, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "evals_per_sec_on_sample": evals_s, "logL_sample": ll,
    }
    return line


def workload_config(args, sample=None):
    cfg = {"workload": "synthetic %d-taxon alignment, %d site patterns, GTR+G4, random tree (seq. addition seed 1, "
                       "edges Exp(0.05)), simulated sites seed 3" % (args.taxa, args.patterns),
           "taxa": args.taxa, "patterns": args.patterns, "categories": 4, "operations_per_eval": args.taxa - 2,
           "l2": "inputs larger than L2 (no flush needed)", "parallelism": "patterns sharded over %d rank(s)" % args.gpus}
    if sample is not None:
        cfg["cpu_sample_patterns"] = sample
    return cfg


def rbcl738_block(E, device, steps=300, warmup=30, pinvar=0.0):
    """The real-data, latency-bound case of the north star: rbcl738 (738 taxa, 1235 patterns), GTR+G4 with the
    reference's paup.nex parameters (golden-checked) or, with pinvar > 0, GTR+I+G4 (BASELINE.json configs[2]: the
    invariable-sites class is a fifth category of rate 0).  Reads the committed fixture (derived from the
    reference's rbcl738.nex / rbcl738nj.tre)."""
    import torch
    from strom_b200 import hostmodel
    from strom_b200.engine import EvalArgs, ShardedLikelihood
    fx = np.load(os.path.join(ROOT, "tests", "golden", "rbcl738.npz"))
    model = gtr_gamma_model(PAUP_PI, PAUP_R, PAUP_ALPHA, 4)
    if pinvar > 0.0:
        model["rates"], model["probs"] = hostmodel.invariant_gamma_rates(PAUP_ALPHA, 4, pinvar)
    NC = len(model["rates"])
    data = np.minimum(fx["data"], 4).astype(np.uint8)
    n, P = data.shape
    L = ShardedLikelihood(data, fx["counts"], NC, device=device)
    ea = EvalArgs(model, fx["ops"], fx["pmatrix_index"], fx["edge_lengths"], int(fx["parent"]), 0)
    ll = L.calc_log_likelihood(ea)
    for _ in range(warmup):
        L.replay_async()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(L.stream):
        L.stream.synchronize()
        e0.record()
        for _ in range(steps):
            L.replay_async()
        e1.record()
        L.stream.synchronize()
    dev_ms = e0.elapsed_time(e1) / steps
    # e2e: the 13-call reference-facing path, host buffers in, scalar out, every step
    B = L.B
    for _ in range(warmup):
        B.calc_log_likelihood(L.inst, model, fx["ops"], fx["pmatrix_index"], fx["edge_lengths"], int(fx["parent"]), 0)
    t0 = time.perf_counter()
    for _ in range(steps):
        ll2 = B.calc_log_likelihood(L.inst, model, fx["ops"], fx["pmatrix_index"], fx["edge_lengths"], int(fx["parent"]), 0)
    e2e_ms = (time.perf_counter() - t0) * 1000.0 / steps
    # fused single call
    for _ in range(warmup):
        L.calc_log_likelihood(ea)
    t0 = time.perf_counter()
    for _ in range(steps):
        ll3 = L.calc_log_likelihood(ea)
    fused_ms = (time.perf_counter() - t0) * 1000.0 / steps
    alg, sch = C.c_double(), C.c_double()
    E.sb200LastTraffic(L.inst, C.byref(alg), C.byref(sch))
    L.close()
    # several independent evaluations at once on ONE GPU (each its own engine instance and stream), as the heated
    # chains of an MC3 run are: one rbcl738 evaluation leaves most SMs idle, so they overlap
    K = 4
    Ls = [ShardedLikelihood(data, fx["counts"], NC, device=device) for _ in range(K)]
    for Lk in Ls:
        Lk.calc_log_likelihood(ea)
    done = [torch.cuda.Event() for _ in range(K)]
    def rounds(count):
        for _ in range(count):
            for Lk in Ls:
                Lk.replay_async()
    rounds(warmup)
    torch.cuda.synchronize(device)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream(device)
    c0.record(main)
    for Lk in Ls:
        Lk.stream.wait_event(c0)
    rounds(steps)
    for Lk, ev in zip(Ls, done):
        ev.record(Lk.stream)
        main.wait_event(ev)
    c1.record(main)
    torch.cuda.synchronize(device)
    conc_ms = c0.elapsed_time(c1) / (steps * K)
    for Lk in Ls:
        Lk.close()
    # the C++ Likelihood facade (strom_b200/host/likelihood.hpp): what strom's Chain / Updaters call
    from strom_b200.host import host
    H = host()
    cols = np.repeat(np.arange(P), fx["counts"].astype(int))
    hd = H.data_from_codes(data.astype(np.int32)[:, cols], compress=True)
    llf, sec = H.loglik(hd, str(fx["newick"]), PAUP_PI, PAUP_R, PAUP_ALPHA, 4, pinvar=pinvar, repeats=steps + 1)
    H.data_free(hd)
    # ... and the same facade compiled as the reference has it: 13 BeagleLib calls per evaluation, library = this engine
    llf13, sec13 = None, None
    try:
        from strom_b200.build import build_host_13call
        from strom_b200.host import HostAPI
        H13 = HostAPI(build_host_13call())
        hd13 = H13.data_from_codes(data.astype(np.int32)[:, cols], compress=True)
        llf13, sec13 = H13.loglik(hd13, str(fx["newick"]), PAUP_PI, PAUP_R, PAUP_ALPHA, 4, pinvar=pinvar, repeats=steps + 1)
        H13.data_free(hd13)
    except Exception as e:                       # (this leg is extra evidence: the headline numbers do not depend on it)
        sys.stderr.write("13-call facade leg skipped: %r\n" % (e,))
    # CPU oracle on the same problem
    tips = data
    one, _, llc = time_oracle(tips, fx["counts"], model, fx["ops"], fx["pmatrix_index"], fx["edge_lengths"], int(fx["parent"]),
                              1, 40, 3, budget_s=8.0)
    thr = host_threads()
    allc, _, _ = time_oracle(tips, fx["counts"], model, fx["ops"], fx["pmatrix_index"], fx["edge_lengths"], int(fx["parent"]),
                             thr, 200, 5, budget_s=8.0)
    upd = (n - 2) * P * NC
    return {
        "workload": "rbcl738: 738 taxa, 1235 patterns, %s, rbcl738nj.tre (fixture tests/golden/rbcl738.npz)" %
                    ("GTR+I+G4 (pinvar %.2f, %d categories)" % (pinvar, NC) if pinvar > 0 else "GTR+G4"),
        "categories": NC,
        "logL": ll, "logL_13call": ll2, "logL_fused": ll3, "logL_cpu": llc, "golden": None if pinvar > 0 else -144730.8,
        "evals_per_sec_device": 1000.0 / dev_ms, "evals_per_sec_e2e_13call_python_ctypes": 1000.0 / e2e_ms,
        "evals_per_sec_e2e_fused_call": 1000.0 / fused_ms,
        "evals_per_sec_e2e_cpp_facade": 1.0 / sec, "logL_cpp_facade": llf,
        "evals_per_sec_e2e_cpp_facade_13_calls": (1.0 / sec13) if sec13 else None, "logL_cpp_facade_13_calls": llf13,
        "evals_per_sec_device_4_concurrent_chains": 1000.0 / conc_ms,
        "partial_updates_per_sec_device": upd * 1000.0 / dev_ms,
        "cpu_evals_per_sec_1thread": one, "cpu_evals_per_sec_all_threads": allc, "cpu_threads": thr,
        "speedup_e2e_vs_cpu_1thread": (1.0 / sec) / one, "speedup_e2e_vs_cpu_all_threads": (1.0 / sec) / allc,
        "algorithmic_bytes_per_eval": alg.value, "scheduled_bytes_per_eval": sch.value,
        "l2": "warm: the 116 MB working set is re-used by every MCMC step, as in the reference's usage",
    }


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from strom_b200.engine import EvalArgs, ShardedLikelihood, api
    E = api()                                    # raises if the CUDA library is missing: no fallback
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = "cuda:%d" % local_rank
    tree, model, tips_h, lo, hi = synthetic_inputs(args, rank, world, dev)
    torch.cuda.empty_cache()
    Pl = hi - lo
    L = ShardedLikelihood(tips_h, np.ones(Pl), 4, device=local_rank, distributed=distributed, single=(args.dtype == "f32"))
    ea = EvalArgs(model, tree.ops, tree.pmatrix_index, tree.edge_lengths, tree.parent_buffer, 0)
    n, K, W = args.taxa, args.steps, args.warmup

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    logl = L.calc_log_likelihood(ea)             # first call: uploads the plan
    # ---- device-only leg (`value`) -----------------------------------------------------------
    for _ in range(W):
        L.replay_async()
    E.sb200SetProfiling(L.inst, 1)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = L.launches()
    with torch.cuda.stream(L.stream):
        e0.record()
        for _ in range(K):
            L.replay_async()
        e1.record()
    barrier()
    launches = L.launches() - launches0
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    pms, pcalls = C.c_double(), C.c_long()
    E.sb200PartialsTime(L.inst, C.byref(pms), C.byref(pcalls))
    E.sb200SetProfiling(L.inst, 0)
    alg, sch = C.c_double(), C.c_double()
    E.sb200LastTraffic(L.inst, C.byref(alg), C.byref(sch))
    # ---- end-to-end leg (`e2e`): public call, host buffers in, scalar out, every step ---------
    for _ in range(W):
        L.calc_log_likelihood(ea)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(L.stream):
        e2.record()
    t0 = time.perf_counter()
    for _ in range(K):
        logl2 = L.calc_log_likelihood(ea)
    with torch.cuda.stream(L.stream):
        e3.record()
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1000.0
    e2e_ms = e2.elapsed_time(e3)
    # max over ranks
    t = torch.tensor([dev_ms, e2e_ms, e2e_wall_ms, pms.value / max(pcalls.value, 1)], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_wall_ms, prune_ms = [float(x) for x in t.tolist()]
    L.close()
    del L
    torch.cuda.empty_cache()

    if rank == 0:
        upd_per_eval = float(n - 2) * args.patterns * 4
        value = upd_per_eval * K / (dev_ms / 1000.0)
        e2e_val = upd_per_eval * K / (max(e2e_ms, e2e_wall_ms) / 1000.0)
        peak, peak_src = peaks()
        achieved = alg.value / (prune_ms / 1000.0) / 1e9            # this rank's slice / this rank's time
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("prune_chains_dram_bytes_per_eval")
                if traffic is not None:
                    traffic = traffic * (args.patterns / 1e6) / world      # captured at 1M patterns on one rank; linear in patterns
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": workload_config(args),
            "evals_per_sec": K / (dev_ms / 1000.0), "logL": logl, "logL_e2e": logl2,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": ea.h2d_bytes, "d2h_bytes_per_step": 8,
                    "evals_per_sec": K / (max(e2e_ms, e2e_wall_ms) / 1000.0), "ms_per_step_events": e2e_ms / K,
                    "ms_per_step_wall": e2e_wall_ms / K,
                    "call": "ShardedLikelihood.calc_log_likelihood -> sb200EvalAsync (+ NCCL sum) -> scalar on host"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "prune_chains_kernel (all level launches of one evaluation)",
                         "algorithmic_bytes_per_eval_per_rank": alg.value, "scheduled_bytes_per_eval_per_rank": sch.value,
                         "scheduled_frac": sch.value / (prune_ms / 1000.0) / 1e9 / peak,
                         "prune_ms_per_eval": prune_ms, "peak_source": peak_src},
        }
        if world == 1 and not args.no_cpu:
            thr = host_threads()
            Ps = min(args.patterns, args.cpu_sample)
            ev, done, _ = time_oracle(tips_h[:, :Ps], np.ones(Ps), model, tree.ops, tree.pmatrix_index, tree.edge_lengths,
                                      tree.parent_buffer, thr, 50, 1, budget_s=12.0)
            ev1, done1, _ = time_oracle(tips_h[:, :Ps], np.ones(Ps), model, tree.ops, tree.pmatrix_index, tree.edge_lengths,
                                        tree.parent_buffer, 1, 50, 1, budget_s=12.0)
            line["cpu_baseline"] = {"value": (n - 2) * Ps * 4 * ev, "unit": UNIT, "cores": thr, "kind": "port",
                                    "sample": "first %d of %d patterns, same tree/model, %d evaluations" % (Ps, args.patterns, done),
                                    "single_thread_value": (n - 2) * Ps * 4 * ev1}
            if not args.no_rbcl738:
                line["rbcl738"] = rbcl738_block(E, local_rank)
                line["rbcl738_i_g4"] = rbcl738_block(E, local_rank, pinvar=0.2)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return line if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--taxa", type=int, default=1000)
    ap.add_argument("--patterns", type=int, default=1000000)
    ap.add_argument("--cpu-sample", type=int, default=16384)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="f32 = the stated FP32 variant (partials/matrices float)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-rbcl738", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world
    # stdout carries exactly ONE line (the JSON, from rank 0): anything a library prints meanwhile (NCCL's version
    # banner, for one) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run_reference(args, rank, world) if args.impl == "reference" else run_ours(args, rank, world, local_rank)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
