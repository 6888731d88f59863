#!/bin/bash
# End-of-round run on one B200: smoke, GPU tests, both bench arms, incremental-evaluation tools, ncu evidence.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python -m pytest tests -m gpu -q 2>&1 | tail -2
bash tools/gpu_bench_full.sh 2>&1 | tail -14
python tools/mcmc_incremental_run.py 400 2>&1 | tail -1 > gpurun_out/r2_mcmc_incremental.json; cat gpurun_out/r2_mcmc_incremental.json
python tools/incremental_run.py 2000 2>&1 | tail -1 > gpurun_out/r2_incremental_walk.json; cat gpurun_out/r2_incremental_walk.json
bash tools/gpu_evidence.sh 2>&1 | tail -3
