#!/bin/bash
# Launch lists (ncu metrics pass) of one rbcl738 evaluation and of eight as one launch set; each after a plain run.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
R="python tools/rbcl738_run.py 3"
$R > gpurun_out/plain2.log 2>&1 && ncu --metrics $M --clock-control none -s 63 -c 9 --csv --log-file gpurun_out/r2_rbcl738_levels.csv $R > gpurun_out/ncu2.log 2>&1
K="python tools/rbcl738_batch.py 8 3"
$K > gpurun_out/plain3.log 2>&1 && ncu --metrics $M --clock-control none -s 153 -c 9 --csv --log-file gpurun_out/r2_rbcl738_batch8_levels.csv $K > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/r2_rbcl738_levels.csv gpurun_out/r2_rbcl738_batch8_levels.csv
