#!/bin/bash
# Round-2 evidence: launch lists (time, DRAM bytes, issue, L1, occupancy) of one evaluation in each regime, and one
# --set full capture of the streaming pruning launches.  Every ncu run follows a plain run of the same command.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
B="python bench.py --taxa 1000 --patterns 250000 --steps 1 --warmup 1 --no-cpu --no-f32"
$B > gpurun_out/plain1.log 2>&1 && ncu --metrics $M -k regex:"prune|root|matrices|tiptip" --clock-control none -s 20 -c 8 --csv --log-file gpurun_out/r2_levels_1000x250k.csv $B > gpurun_out/ncu1.log 2>&1
R="python tools/rbcl738_run.py 3"
$R > gpurun_out/plain2.log 2>&1 && ncu --metrics $M --clock-control none -s 63 -c 9 --csv --log-file gpurun_out/r2_rbcl738_levels.csv $R > gpurun_out/ncu2.log 2>&1
K="python tools/rbcl738_batch.py 8 3"
$K > gpurun_out/plain3.log 2>&1 && ncu --metrics $M --clock-control none -s 153 -c 9 --csv --log-file gpurun_out/r2_rbcl738_batch8_levels.csv $K > gpurun_out/ncu3.log 2>&1
B2="python bench.py --taxa 1000 --patterns 250000 --steps 2 --warmup 3 --no-cpu --no-f32"
$B2 > gpurun_out/plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:prune_chains -s 10 -c 5 -f -o gpurun_out/r2_stream_prune $B2 > gpurun_out/ncu4.log 2>&1
$R > gpurun_out/plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:prune -s 18 -c 6 -f -o gpurun_out/r2_rbcl738_prune $R > gpurun_out/ncu5.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
