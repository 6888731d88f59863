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


def mem_available_bytes():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) * 1024
    except Exception:
        pass
    return 8 << 30


def oracle_full_pass(args, tree, model, threads, slice_patterns, budget_s, max_passes):
    """The CPU path over the WHOLE alignment, `slice_patterns` at a time (the oracle keeps all n-2 partials buffers of a
    slice in host memory: 128 KB per pattern at 1000 taxa).  Returns (seconds per full evaluation, passes, summed logL)."""
    import torch
    from strom_b200 import synth
    api = oracle_api(threads)
    P = args.patterns
    bounds = [(lo, min(P, lo + slice_patterns)) for lo in range(0, P, slice_patterns)]
    # the SAME alignment as the GPU arm: its sites come from torch's CUDA generator (chunks of 125000 seeded by chunk
    # index), so the GPU of the box simulates the input here too; only the likelihood is the CPU's work
    gen_dev = "cuda:0" if torch.cuda.is_available() else "cpu"
    inst, inst_P = None, -1
    total, passes, logl = 0.0, 0, None
    t_start = time.perf_counter()
    while passes < max_passes:
        ll_sum, secs = 0.0, 0.0
        for lo, hi in bounds:
            tips = synth.simulate_alignment(tree, model, lo, hi, seed=3, device=gen_dev).cpu().numpy()
            if inst is None or inst_P != hi - lo:
                if inst is not None:
                    api.finalize(inst)
                inst, inst_P = api.create(args.taxa, hi - lo, 4), hi - lo
            api.set_data(inst, tips.astype(np.int32), np.ones(hi - lo))
            del tips
            if passes == 0 and lo == 0:      # untimed first touch of the partials arena
                api.calc_log_likelihood(inst, model, tree.ops, tree.pmatrix_index, tree.edge_lengths, tree.parent_buffer, 0)
            t0 = time.perf_counter()
            ll_sum += api.calc_log_likelihood(inst, model, tree.ops, tree.pmatrix_index, tree.edge_lengths, tree.parent_buffer, 0)
            secs += time.perf_counter() - t0
        total += secs
        passes += 1
        logl = ll_sum
        if time.perf_counter() - t_start + 1.3 * (time.perf_counter() - t_start) / passes > budget_s:
            break
    if inst is not None:
        api.finalize(inst)
    return total / passes, passes, logl, gen_dev


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (restated in oracle/; BeagleLib itself is not buildable
    here) on the host cores.  One step = one full evaluation of the SAME configuration as the GPU arm
    (all patterns, processed in slices that fit the host's memory); as many such steps as fit the time budget."""
    if rank != 0:
        return None
    threads = host_threads()
    from strom_b200 import synth
    tree = synth.RandomTree(args.taxa, seed=1, mean_edge=0.05)
    model = gtr_gamma_model(PAUP_PI, PAUP_R, PAUP_ALPHA, 4)
    per_pattern = (args.taxa - 2) * 4 * 4 * 8 + 64
    slice_patterns = int(max(1024, min(args.patterns, 125000, 0.4 * mem_available_bytes() / per_pattern)))
    sec, passes, ll, gen_dev = oracle_full_pass(args, tree, model, threads, slice_patterns, budget_s=args.cpu_budget,
                                                max_passes=max(1, args.steps))
    evals_s = 1.0 / sec
    upd = (args.taxa - 2) * args.patterns * 4 * evals_s
    nslices = (args.patterns + slice_patterns - 1) // slice_patterns
    line = {
        "impl": "reference", "metric": METRIC, "value": upd, "unit": UNIT, "n_gpus": args.gpus, "steps": passes,
        "warmup": 1, "ms_per_step": 1000.0 * sec, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "same_config": True,
        "cpu_baseline": {"value": upd, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "the whole alignment: %d patterns in %d slices of <= %d (host memory), same tree and model; "
                                   "one step = one full evaluation of all slices; %d step(s) within the %.0f s budget "
                                   "(alignment simulation and tip upload per slice not timed)"
                                   % (args.patterns, nslices, slice_patterns, passes, args.cpu_budget)},
        "e2e": {"value": upd, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "evals_per_sec": evals_s, "logL": ll,
        "data_generator": gen_dev + (" (same alignment as the GPU arm)" if gen_dev != "cpu" else
                                     " (no GPU: torch's CPU generator gives another alignment of the same distribution; logL not comparable)"),
        "note": "CPU restatement of the reference's BeagleLib path (oracle/beagle_cpu.c, plain C, %d threads); compare `logL` "
                "with the GPU arm's `logL`" % threads,
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
    # K independent evaluations (own tree / model each, as the heated chains of an MC3 run) as ONE set of kernel
    # launches on ONE GPU (sb200EvalBatchAsync): a single rbcl738 evaluation leaves most of the GPU idle
    K = 8
    from strom_b200.engine import StromEvalArgs
    rng = np.random.default_rng(5)
    Ls, eas = [], []
    for c in range(K):
        mk = gtr_gamma_model(PAUP_PI, PAUP_R, PAUP_ALPHA * (1.0 + 0.1 * c), 4)
        if pinvar > 0.0:
            mk["rates"], mk["probs"] = hostmodel.invariant_gamma_rates(PAUP_ALPHA * (1.0 + 0.1 * c), 4, pinvar)
        elen = fx["edge_lengths"] * np.exp(0.2 * rng.standard_normal(fx["edge_lengths"].shape)) if c else fx["edge_lengths"]
        Ls.append(ShardedLikelihood(data, fx["counts"], NC, device=device))
        eas.append(EvalArgs(mk, fx["ops"], fx["pmatrix_index"], elen, int(fx["parent"]), 0))
    single_vals = [Lk.calc_log_likelihood(ek) for Lk, ek in zip(Ls, eas)]
    insts = (C.c_int * K)(*[Lk.inst for Lk in Ls])
    bargs = (StromEvalArgs * K)(*[ek.c for ek in eas])
    res = C.c_double()

    def batch_eval():
        E.check(E.sb200EvalBatchAsync(K, insts, bargs), "batched evaluation")
        vals = []
        for Lk in Ls:
            E.check(E.sb200Finish(Lk.inst, C.byref(res)), "finish")
            vals.append(res.value)
        return vals
    batch_vals = batch_eval()
    batch_rel = max(abs(a - b) / abs(b) for a, b in zip(batch_vals, single_vals))
    for _ in range(warmup):
        E.check(E.sb200ReplayBatchAsync(K, insts), "replay")
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(Ls[0].stream):
        Ls[0].stream.synchronize()
        b0.record()
        for _ in range(steps):
            E.check(E.sb200ReplayBatchAsync(K, insts), "replay")
        b1.record()
        Ls[0].stream.synchronize()
    batch_dev_ms = b0.elapsed_time(b1) / steps
    t0 = time.perf_counter()
    for _ in range(steps):
        batch_eval()
    batch_e2e_ms = (time.perf_counter() - t0) * 1000.0 / steps
    # ... the same with transient partials (only the last node of every chain is written)
    for Lk in Ls:
        E.check(E.sb200SetTransientPartials(Lk.inst, 1), "transient partials")
    tr_vals = batch_eval()
    tr_rel = max(abs(a - b) / abs(b) for a, b in zip(tr_vals, single_vals))
    for _ in range(warmup):
        E.check(E.sb200ReplayBatchAsync(K, insts), "replay")
    with torch.cuda.stream(Ls[0].stream):
        Ls[0].stream.synchronize()
        b0.record()
        for _ in range(steps):
            E.check(E.sb200ReplayBatchAsync(K, insts), "replay")
        b1.record()
        Ls[0].stream.synchronize()
    tr_batch_dev_ms = b0.elapsed_time(b1) / steps
    t0 = time.perf_counter()
    for _ in range(steps):
        batch_eval()
    tr_batch_e2e_ms = (time.perf_counter() - t0) * 1000.0 / steps
    L1 = Ls[0]
    L1.calc_log_likelihood(eas[0])
    for _ in range(warmup):
        L1.replay_async()
    with torch.cuda.stream(L1.stream):
        L1.stream.synchronize()
        b0.record()
        for _ in range(steps):
            L1.replay_async()
        b1.record()
        L1.stream.synchronize()
    tr_dev_ms = b0.elapsed_time(b1) / steps
    for Lk in Ls:
        Lk.close()
    # ... and the stated FP32 variant (float partials / matrices / tables, FP64 scalers and root accumulation): 8 batched chains
    Lf = [ShardedLikelihood(data, fx["counts"], NC, device=device, single=True) for _ in range(K)]
    f32_single = [Lk.calc_log_likelihood(ek) for Lk, ek in zip(Lf, eas)]
    finsts = (C.c_int * K)(*[Lk.inst for Lk in Lf])
    E.check(E.sb200EvalBatchAsync(K, finsts, bargs), "batched evaluation (f32)")
    f32_vals = []
    for Lk in Lf:
        E.check(E.sb200Finish(Lk.inst, C.byref(res)), "finish")
        f32_vals.append(res.value)
    f32_rel = max(abs(a - b) / abs(b) for a, b in zip(f32_vals, single_vals))
    for _ in range(warmup):
        E.check(E.sb200ReplayBatchAsync(K, finsts), "replay")
    with torch.cuda.stream(Lf[0].stream):
        Lf[0].stream.synchronize()
        b0.record()
        for _ in range(steps):
            E.check(E.sb200ReplayBatchAsync(K, finsts), "replay")
        b1.record()
        Lf[0].stream.synchronize()
    f32_batch_dev_ms = b0.elapsed_time(b1) / steps
    for Lk in Lf:
        Lk.close()
    # the C++ Likelihood facade (strom_b200/host/likelihood.hpp): what strom's Chain / Updaters call
    from strom_b200.host import host
    H = host()
    cols = np.repeat(np.arange(P), fx["counts"].astype(int))
    hd = H.data_from_codes(data.astype(np.int32)[:, cols], compress=True)
    llf, sec = H.loglik(hd, str(fx["newick"]), PAUP_PI, PAUP_R, PAUP_ALPHA, 4, pinvar=pinvar, repeats=steps + 1)
    llfm, secm = H.loglik(hd, str(fx["newick"]), PAUP_PI, PAUP_R, PAUP_ALPHA, 4, pinvar=pinvar, repeats=steps + 1, transient_partials=False)
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
        "evals_per_sec_e2e_cpp_facade_every_buffer_written": 1.0 / secm, "logL_cpp_facade_every_buffer_written": llfm,
        "cpp_facade_note": "Likelihood::calcLogLikelihood (C++), transient partials as the facade has them by default; the second figure "
                           "with Likelihood::setTransientPartials(false)",
        "evals_per_sec_e2e_cpp_facade_13_calls": (1.0 / sec13) if sec13 else None, "logL_cpp_facade_13_calls": llf13,
        "batched_chains": {"chains": K, "call": "sb200EvalBatchAsync: one set of kernel launches for all chains (own tree and model each)",
                           "evals_per_sec_device": K * 1000.0 / batch_dev_ms, "evals_per_sec_e2e": K * 1000.0 / batch_e2e_ms,
                           "us_per_batch_device": batch_dev_ms * 1000.0, "max_rel_diff_vs_single_evaluations": batch_rel,
                           "speedup_e2e_vs_cpu_all_threads": (K * 1000.0 / batch_e2e_ms) / allc,
                           "speedup_e2e_vs_cpu_1thread": (K * 1000.0 / batch_e2e_ms) / one},
        "transient_partials": {"evals_per_sec_device": 1000.0 / tr_dev_ms, "evals_per_sec_device_8_batched_chains": K * 1000.0 / tr_batch_dev_ms,
                               "evals_per_sec_e2e_8_batched_chains": K * 1000.0 / tr_batch_e2e_ms,
                               "speedup_e2e_8_batched_vs_cpu_all_threads": (K * 1000.0 / tr_batch_e2e_ms) / allc,
                               "max_rel_diff_vs_materialised": tr_rel,
                               "what": "sb200SetTransientPartials: partials inside a fused chain are not written to HBM; OFF everywhere else in this block"},
        "f32_batched_chains": {"evals_per_sec_device": K * 1000.0 / f32_batch_dev_ms, "max_rel_diff_vs_f64": f32_rel, "stated_tolerance": 2e-5,
                               "what": "BEAGLE_FLAG_PRECISION_SINGLE instances, 8 chains as one launch set, every buffer written"},
        "partial_updates_per_sec_device": upd * 1000.0 / dev_ms,
        "roofline": {"bound": "hbm", "achieved": alg.value / (dev_ms / 1000.0) / 1e9, "peak": peaks()[0], "unit": "GB/s",
                     "frac": alg.value / (dev_ms / 1000.0) / 1e9 / peaks()[0],
                     "scheduled_frac": sch.value / (dev_ms / 1000.0) / 1e9 / peaks()[0],
                     "batched_frac": K * alg.value / (batch_dev_ms / 1000.0) / 1e9 / peaks()[0],
                     "batched_scheduled_frac": K * sch.value / (batch_dev_ms / 1000.0) / 1e9 / peaks()[0],
                     "traffic": None, "note": "whole evaluation (matrices, tables, pruning launches, root) per algorithmic / "
                                              "scheduled byte of the pruning launches; the working set of ONE evaluation fits L2"},
        "cpu_evals_per_sec_1thread": one, "cpu_evals_per_sec_all_threads": allc, "cpu_threads": thr,
        "speedup_e2e_vs_cpu_1thread": (1.0 / sec) / one, "speedup_e2e_vs_cpu_all_threads": (1.0 / sec) / allc,
        "algorithmic_bytes_per_eval": alg.value, "scheduled_bytes_per_eval": sch.value,
        "l2": "warm: the 116 MB working set is re-used by every MCMC step, as in the reference's usage",
    }


def timed_variant(E, tips_h, Pl, ea, local_rank, n, patterns, logl_ref, steps, **kw):
    """One more configuration of the engine on the benched alignment (FP32 storage, transient partials): device-timed
    evaluations and the roofline of its pruning launches."""
    import torch
    from strom_b200.engine import ShardedLikelihood
    Lf = ShardedLikelihood(tips_h, np.ones(Pl), 4, device=local_rank, **kw)
    lf = Lf.calc_log_likelihood(ea)
    for _ in range(3):
        Lf.replay_async()
    E.sb200SetProfiling(Lf.inst, 1)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(Lf.stream):
        Lf.stream.synchronize()
        f0.record()
        for _ in range(steps):
            Lf.replay_async()
        f1.record()
        Lf.stream.synchronize()
    fms = f0.elapsed_time(f1) / steps
    fp, fc = C.c_double(), C.c_long()
    E.sb200PartialsTime(Lf.inst, C.byref(fp), C.byref(fc))
    fa, fs = C.c_double(), C.c_double()
    E.sb200LastTraffic(Lf.inst, C.byref(fa), C.byref(fs))
    Lf.close()
    del Lf
    torch.cuda.empty_cache()
    fprune = fp.value / max(fc.value, 1)
    pk = peaks()[0]
    return {"value": float(n - 2) * patterns * 4 / (fms / 1000.0), "unit": UNIT, "ms_per_step": fms, "steps": steps,
            "logL": lf, "rel_diff_vs_f64": abs(lf - logl_ref) / abs(logl_ref),
            "roofline": {"bound": "hbm", "achieved": fa.value / (fprune / 1000.0) / 1e9, "peak": pk, "unit": "GB/s",
                         "frac": fa.value / (fprune / 1000.0) / 1e9 / pk, "scheduled_frac": fs.value / (fprune / 1000.0) / 1e9 / pk,
                         "algorithmic_bytes_per_eval": fa.value, "scheduled_bytes_per_eval": fs.value, "traffic": None}}


MC3_KW = dict(seed=13579, niter=400, burnin=200, samplefreq=10, heatfactor=0.5, ncateg=4, gammashape=PAUP_ALPHA, freqs=PAUP_PI,
              rmat=[x / sum(PAUP_R) for x in PAUP_R])


def mc3_data(H):
    fx = np.load(os.path.join(ROOT, "tests", "golden", "rbcl738.npz"))
    data = np.minimum(fx["data"], 4).astype(np.int32)
    cols = np.repeat(np.arange(data.shape[1]), fx["counts"].astype(int))
    return H.data_from_codes(data[:, cols], compress=True), str(fx["newick"])


def mc3_block(devices, nchains=8):
    """BASELINE.json configs[4]: Metropolis-coupled MCMC, 8 heated chains on rbcl738 (reference loop src/strom.hpp:316-402),
    through the C++ host driver in ONE process: chains stepped in turn (the reference's order of work) against side by side -- one
    chain per GPU on host threads when there are several GPUs, or, on ONE GPU, one host thread with the chains' evaluations batched
    into sets of launches.  Same seed: the two traces must be identical."""
    from strom_b200.host import host
    H = host()
    hd, newick = mc3_data(H)
    kw = dict(MC3_KW, nchains=nchains, devices=tuple(devices), time_iterations_only=True)      # (set-up of chains / instances not timed)
    H.mcmc(hd, newick, **dict(kw, niter=4, burnin=2), per_chain_streams=True)                  # warm-up
    tr_turn, sec_turn = H.mcmc(hd, newick, per_chain_streams=True, **kw)
    if len(devices) > 1:
        tr_side, sec_side = H.mcmc(hd, newick, parallel_chains=True, **kw)
        how = "one heated chain per GPU (%d GPUs), every chain on its own host thread of ONE process" % len(devices)
    else:
        tr_side, sec_side = H.mcmc(hd, newick, lockstep_chains=True, **kw)
        how = "ONE GPU, one host thread: the chains as two half-sets, each update of a half begun on all its chains and evaluated as ONE set of kernel launches while the host prepares the other half"
    H.data_free(hd)
    its = nchains * (kw["niter"] + kw["burnin"])
    same = tr_turn.shape == tr_side.shape and bool(np.array_equal(tr_turn, tr_side))
    return {"workload": "rbcl738 GTR+G4, %d heated chains, %d iterations each (5 updates + a swap proposal per iteration, a sample every 10)" % (nchains, kw["niter"] + kw["burnin"]),
            "side_by_side": how, "chain_iterations_per_sec_in_turn": its / sec_turn, "chain_iterations_per_sec_side_by_side": its / sec_side,
            "speedup": sec_turn / sec_side, "traces_identical": same, "gpus": len(devices)}, tr_turn


def mc3_one_chain_per_rank(local_rank, dev):
    """The same run with one chain per PROCESS (rank r = chain r on GPU r), swap offers exchanged by an NCCL all-gather over
    NVLink after every iteration (strom_b200/mc3.py).  Every rank calls this; returns (trace on every rank, seconds, chain-iterations)."""
    from strom_b200 import mc3
    from strom_b200.host import host
    H = host()
    hd, newick = mc3_data(H)
    kw = dict(MC3_KW, device=local_rank, tensor_device=dev)
    mc3.run_one_chain_per_rank(H, hd, newick, **dict(kw, niter=4, burnin=2))                   # warm-up (instances, plans, NCCL)
    rows, sec, its = mc3.run_one_chain_per_rank(H, hd, newick, **kw)
    trace = mc3.gather_trace(rows, tensor_device=dev)
    H.data_free(hd)
    return trace, sec, its


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
    L = ShardedLikelihood(tips_h, np.ones(Pl), 4, device=local_rank, distributed=distributed, single=(args.dtype == "f32"),
                          transient_partials=args.transient)
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
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]      # one event per step: p10 / p50 / p90
    barrier()
    launches0 = L.launches()
    with torch.cuda.stream(L.stream):
        evs[0].record()
        for i in range(K):
            L.replay_async()
            evs[i + 1].record()
    barrier()
    launches = L.launches() - launches0
    dev_ms = evs[0].elapsed_time(evs[K])
    step_ms = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(K)])
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
    t = torch.tensor([dev_ms, e2e_ms, e2e_wall_ms, pms.value / max(pcalls.value, 1)] + [float(x) for x in np.percentile(step_ms, [10, 50, 90])],
                     dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_wall_ms, prune_ms, p10, p50, p90 = [float(x) for x in t.tolist()]
    L.close()
    del L
    torch.cuda.empty_cache()

    # ---- self-check: the first cpu_sample patterns of THIS alignment on the GPU and on the CPU oracle ------------------
    verify = None
    if rank == 0 and world == 1 and not args.no_cpu:
        Ps = min(Pl, args.cpu_sample)
        Lv = ShardedLikelihood(np.ascontiguousarray(tips_h[:, :Ps]), np.ones(Ps), 4, device=local_rank, single=(args.dtype == "f32"))
        gpu_sample = Lv.calc_log_likelihood(ea)
        Lv.close()
        verify = {"sample_patterns": Ps, "logL_sample_gpu": gpu_sample}

    # ---- the stated FP32 variant on the same alignment (own roofline) --------------------------------------------------
    f32, transient = None, None
    if world == 1 and args.dtype == "f64" and not args.no_f32:
        f32 = timed_variant(E, tips_h, Pl, ea, local_rank, n, args.patterns, logl, max(3, min(K, 5)), single=True)
        f32["dtype"] = "f32 partials / matrices / tables, FP64 scalers and root accumulation (BEAGLE_FLAG_PRECISION_SINGLE)"
        f32["stated_tolerance"] = 2e-5
    if world == 1 and args.dtype == "f64" and not args.transient and not args.no_f32:
        transient = timed_variant(E, tips_h, Pl, ea, local_rank, n, args.patterns, logl, max(3, min(K, 5)), transient_partials=True)
        transient["what"] = "sb200SetTransientPartials / Likelihood::setTransientPartials: partials inside a fused chain of operations " \
                            "stay in registers and are not written to HBM (only what a later launch reads is); same operations, same " \
                            "bits; OFF in the headline numbers, where every node's buffer is written as BeagleLib does"

    if rank == 0:
        upd_per_eval = float(n - 2) * args.patterns * 4
        value = upd_per_eval * K / (dev_ms / 1000.0)
        e2e_val = upd_per_eval * K / (max(e2e_ms, e2e_wall_ms) / 1000.0)
        peak, peak_src = peaks()
        achieved = alg.value / (prune_ms / 1000.0) / 1e9            # this rank's slice / this rank's time
        # DRAM bytes of the pruning launches from an ncu capture (profiles/traffic.json): a STORED figure, scaled by the
        # pattern count, used only for the very shape it was captured on; otherwise null
        traffic, traffic_source = None, "none: no ncu capture stored for this dtype / taxon count (profiles/traffic.json)"
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if tj.get("dtype", "f64") == args.dtype and int(tj.get("taxa", 1000)) == args.taxa and tj.get("prune_chains_dram_bytes_per_eval"):
                    traffic = tj["prune_chains_dram_bytes_per_eval"] * (args.patterns / float(tj.get("patterns", 1e6))) / world
                    traffic_source = "stored: %s, scaled linearly to %d patterns / %d rank(s); not measured in this run" % (
                        tj.get("source", "profiles/traffic.json"), args.patterns, world)
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "ms_per_step_p10": p10, "ms_per_step_p50": p50, "ms_per_step_p90": p90,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic", "config": workload_config(args),
            "evals_per_sec": K / (dev_ms / 1000.0), "logL": logl, "logL_e2e": logl2,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": ea.h2d_bytes, "d2h_bytes_per_step": 8,
                    "evals_per_sec": K / (max(e2e_ms, e2e_wall_ms) / 1000.0), "ms_per_step_events": e2e_ms / K,
                    "ms_per_step_wall": e2e_wall_ms / K,
                    "call": "ShardedLikelihood.calc_log_likelihood -> sb200EvalAsync (+ NCCL sum) -> scalar on host"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_source,
                         "kernel": "prune_chains_kernel (all level launches of one evaluation)",
                         "algorithmic_bytes_per_eval_per_rank": alg.value, "scheduled_bytes_per_eval_per_rank": sch.value,
                         "scheduled_frac": sch.value / (prune_ms / 1000.0) / 1e9 / peak,
                         "prune_ms_per_eval": prune_ms, "peak_source": peak_src},
        }
        if f32 is not None:
            line["f32"] = f32
        if transient is not None:
            line["transient_partials"] = transient
        if world == 1 and not args.no_cpu:
            thr = host_threads()
            Ps = min(args.patterns, args.cpu_sample)
            ev, done, llc = time_oracle(tips_h[:, :Ps], np.ones(Ps), model, tree.ops, tree.pmatrix_index, tree.edge_lengths,
                                        tree.parent_buffer, thr, 50, 1, budget_s=12.0)
            ev1, done1, _ = time_oracle(tips_h[:, :Ps], np.ones(Ps), model, tree.ops, tree.pmatrix_index, tree.edge_lengths,
                                        tree.parent_buffer, 1, 50, 1, budget_s=12.0)
            line["cpu_baseline"] = {"value": (n - 2) * Ps * 4 * ev, "unit": UNIT, "cores": thr, "kind": "port",
                                    "sample": "first %d of %d patterns, same tree/model, %d evaluations" % (Ps, args.patterns, done),
                                    "single_thread_value": (n - 2) * Ps * 4 * ev1}
            if verify is not None:
                verify["logL_sample_cpu"] = llc
                verify["rel_diff"] = abs(verify["logL_sample_gpu"] - llc) / abs(llc)
                verify["ok"] = bool(verify["rel_diff"] < (2e-5 if args.dtype == "f32" else 1e-9))
                verify["note"] = "the timed kernels on the first patterns of the benched alignment against the CPU oracle; the " \
                                 "--impl reference arm prints the CPU logL of the WHOLE alignment to compare with `logL`"
                line["verify"] = verify
            if not args.no_rbcl738:
                line["rbcl738"] = rbcl738_block(E, local_rank)
                line["rbcl738_i_g4"] = rbcl738_block(E, local_rank, pinvar=0.2)
        if world == 1 and not args.no_mc3 and not args.no_rbcl738:
            line["mc3"] = mc3_block([local_rank])[0]
    if distributed:
        dist.barrier()                   # every rank has released its shard
        torch.cuda.synchronize()
        if world == 8 and not args.no_mc3:
            # BASELINE.json configs[4]: 8 heated chains on rbcl738, one per GPU.  First rank 0 alone drives all eight in ONE
            # process (in turn, then on host threads) while the other ranks wait on the rendezvous store (a socket, NOT an NCCL
            # barrier, whose kernels would spin on the very GPUs the chains use); then every rank runs ITS chain and the swap
            # offers travel by NCCL.
            import datetime
            store = dist.distributed_c10d._get_default_store()
            turn_trace = None
            if rank == 0:
                try:
                    line["mc3"], turn_trace = mc3_block(list(range(world)))
                except Exception as e:
                    line["mc3"] = {"error": repr(e)}
                store.set("sb200_mc3_done", "1")
            else:
                store.wait(["sb200_mc3_done"], datetime.timedelta(minutes=30))
            try:
                trace, sec, its = mc3_one_chain_per_rank(local_rank, dev)
                if rank == 0 and "error" not in line["mc3"]:
                    m = line["mc3"]
                    m["one_process_per_gpu"] = {
                        "how": "rank r = heated chain r on GPU r; after every iteration the ranks all-gather 16 doubles over NCCL "
                               "(NVLink): the two chains of the swap proposal exchange log kernel, heat and tuning parameters",
                        "chain_iterations_per_sec": its / sec, "speedup_vs_in_turn": (its / sec) / m["chain_iterations_per_sec_in_turn"],
                        "trace_identical_to_in_turn": bool(turn_trace is not None and trace.shape == turn_trace.shape and np.array_equal(trace, turn_trace))}
                    m["speedup_one_process_per_gpu"] = m["one_process_per_gpu"]["speedup_vs_in_turn"]
            except Exception as e:
                if rank == 0:
                    line.setdefault("mc3", {})["one_process_per_gpu"] = {"error": repr(e)}
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
    ap.add_argument("--cpu-budget", type=float, default=100.0, help="seconds the reference arm may spend on timed full evaluations")
    ap.add_argument("--no-f32", action="store_true")
    ap.add_argument("--transient", action="store_true", help="sb200SetTransientPartials: chain-interior partials are not written")
    ap.add_argument("--no-mc3", action="store_true")
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
