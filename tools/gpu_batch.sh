#!/bin/bash
mkdir -p gpurun_out
{
for k in 1 2 4 8 16; do python tools/rbcl738_batch.py $k 200 2>&1 | tail -4; done
echo "== PINVAR"; PINVAR=0.2 python tools/rbcl738_batch.py 8 200 2>&1 | tail -3
for w in 1200 2400 4800 9600 20000 100000; do echo "== wide min warps $w"; SB200_WIDE_MIN_WARPS=$w python tools/rbcl738_batch.py 8 200 2>&1 | grep "device us"; done
for s in 888 3552 7104; do echo "== slots $s"; SB200_SLOTS=$s python tools/rbcl738_batch.py 8 200 2>&1 | grep "device us"; done
} > gpurun_out/batch1.log 2>&1
cat gpurun_out/batch1.log
python tools/rbcl738_batch.py 8 3 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -s 150 -c 12 --csv --log-file gpurun_out/r2_rbcl738_batch8_levels.csv python tools/rbcl738_batch.py 8 3 > gpurun_out/ncu1.log 2>&1
echo done
