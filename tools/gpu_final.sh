#!/bin/bash
# round-end evidence on one B200: GPU test-suite, bench line (ours + reference arm), per-launch ncu lists of one evaluation
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/final_pytest_gpu.log 2>&1; tail -4 gpurun_out/final_pytest_gpu.log
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "ref rc $?"
python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err; echo "bench rc $?"
python bench.py --steps 10 --warmup 3 --dtype f32 --no-cpu --no-rbcl738 > gpurun_out/final_bench_1gpu_f32.json 2> gpurun_out/final_bench_1gpu_f32.err; echo "f32 rc $?"
bash profiles/levels.sh final2_levels_1000x250k.csv; echo "levels rc $?"
CMD="python tools/rbcl738_run.py 20"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"prune|tiptip|transition|root" -s 99 -c 9 --csv --log-file gpurun_out/final2_rbcl738_levels.csv $CMD > gpurun_out/ncu2.log 2>&1; echo "rbcl738 levels rc $?"
python tools/mc3_run.py 1 1000 > gpurun_out/final_mcmc_1chain.json 2>&1; tail -1 gpurun_out/final_mcmc_1chain.json
python tools/mc3_run.py 4 500 > gpurun_out/final_mcmc_4chains_1gpu.json 2>&1; tail -1 gpurun_out/final_mcmc_4chains_1gpu.json
