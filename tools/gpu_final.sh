#!/bin/bash
# round-end evidence on one B200: bench line (ours + reference arm), per-launch ncu list of one evaluation
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench_1gpu.json 2> gpurun_out/final_bench_1gpu.err
echo "bench rc $?"
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
echo "ref rc $?"
bash profiles/levels.sh final2_levels_1000x250k.csv
echo "levels rc $?"
CMD="python tools/rbcl738_run.py 20"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"prune|tiptip|transition|root" -s 99 -c 9 --csv --log-file gpurun_out/final2_rbcl738_levels.csv $CMD > gpurun_out/ncu2.log 2>&1
echo "rbcl738 levels rc $?"
head -c 600 gpurun_out/final_bench_1gpu.json
